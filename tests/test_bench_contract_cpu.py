"""bench.py contract pieces that can be checked without a GPU: the reference arm's JSON line (the unmodified reference when a
tree is present, else the oracle port, on a tiny workload) and that the product arm refuses to run without CUDA instead of
falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600, **extra_env):
    env = dict(os.environ, COUNCIL_CPU_THREADS='4', **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


import pytest


@pytest.mark.parametrize('disable_ref', ['0', '1'])
def test_reference_arm_prints_one_contract_line(disable_ref):
    r = run_bench('--impl', 'reference', '--workload', 'tiny_64_n2_b2', '--steps', '1', '--warmup', '0', COUNCIL_REF_DISABLE=disable_ref)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['metric'] == 'training images/sec (gen+dis step)' and d['steps'] == 1 and d['warmup'] == 0
    assert d['value'] > 0 and abs(d['value'] - 1e3 / d['ms_per_step']) < 1e-6 * d['value']
    assert d['config']['workload'] == 'tiny_64_n2_b2' and 'model' not in d['config']
    cb = d['cpu_baseline']
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_runner
    have_ref = disable_ref == '0' and ref_runner.find_reference() is not None
    assert cb['kind'] == ('reference' if have_ref else 'port') and cb['unmodified'] is have_ref
    assert cb['cores'] == 4 and cb['host_cores'] >= 1 and cb['value'] == d['value']
    assert cb['batch_ran'] == 1 and d['config']['batch_ran'] == 1 and d['config']['batch_per_gpu'] == 2  # the sample's batch is stated
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--workload', 'tiny_64_n2_b2',
                        '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('CUDA present: the product arm runs')
    r = run_bench('--workload', 'tiny_64_n2_b2', '--steps', '1', '--warmup', '0', '--no-cpu-baseline', timeout=300)
    assert r.returncode != 0, 'the product arm must fail loudly without a GPU, not fall back'
    assert r.stdout.strip() == ''
