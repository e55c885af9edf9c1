"""The oracle restatement (oracle/council_oracle.py) must reproduce what the unmodified reference
produced (tests/golden/*.json, written by oracle/make_golden.py).  CPU only."""
import pytest
import torch

import council_oracle as co
from common import close, load_golden, probe, run_oracle_iteration
from make_golden import PROBE_PARAMS

import os

CASES = ['glasses64_n2_b2_early', 'm2f64_n4_b2', 'anime64_n3_b2', 'glasses128_n2_b1', 'm2f256_n2_b1', 'glasses64_n2_b2_both']
# the BASELINE configurations at full council size / batch / resolution take minutes each on CPU: opt-in (COUNCIL_BIG=1); their
# fixtures are what the -m gpu parity tests and bench.py's pre-timing check compare the CUDA path against
BIG = ['anime256_n4_b4', 'm2f256_n4_b8', 'm2f512_n6_b2']
if os.environ.get('COUNCIL_BIG') == '1':
    CASES = CASES + BIG
RTOL = 2e-5  # both sides are torch-CPU fp32; the slack covers thread-count dependent summation order


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_golden(case):
    gold = load_golden(case)
    torch.set_num_threads(8)
    tr, hp, x_a, x_b = run_oracle_iteration(gold)
    d0 = tr.dirs[0]
    N = tr.N

    def chk(name, got, want, rtol=RTOL):
        assert len(got) == len(want), name
        for i, (g, w) in enumerate(zip(got, want)):
            assert close(float(g), w, rtol, 1e-7), '%s[%d]: oracle %r reference %r' % (name, i, float(g), w)

    chk('loss_dis_total', tr.loss_dis_total_s, gold['loss_dis_total'])
    assert tr.disc_ran == gold['dis_council_ran']
    if tr.disc_ran:
        chk('loss_dis_council_total', tr.loss_dis_council_total_s, gold['loss_dis_council_total'])
    chk('loss_gen_total', tr.loss_gen_total_s, gold['loss_gen_total'])
    chk('loss_gen_adv', tr.loss_gen_adv_s[d0], gold['loss_gen_adv'])
    if gold['council_loss'] and any(v != 0 for v in gold['council_loss']):
        chk('council_loss', tr.council_loss_s[d0], gold['council_loss'])
        assert close(float(tr.w_match[d0]), gold['w_match'], 1e-6)
    chk('mask01', tr.loss_gen_mask_zero_one_s[d0], gold['loss_gen_mask_zero_one'])
    if any(v != 0 for v in gold['loss_gen_mask_total']):
        chk('mask_total', tr.loss_gen_mask_total_s[d0], gold['loss_gen_mask_total'])
    if any(v != 0 for v in gold['loss_gen_mask_TV']):
        chk('mask_tv', tr.loss_gen_mask_TV_s[d0], gold['loss_gen_mask_TV'])

    if 'dirs' in gold:  # both directions: per-direction published losses
        for d in tr.dirs:
            chk('loss_gen_adv ' + d, tr.loss_gen_adv_s[d], gold['dirs'][d]['loss_gen_adv'])
            chk('council_loss ' + d, tr.council_loss_s[d], gold['dirs'][d]['council_loss'])

    # post-step parameters and generator grads
    for fam, dd in [(f, d) for d in tr.dirs for f in ('gen', 'dis', 'dis_council')]:
        name = '%s_%s' % (fam, dd)
        if name not in tr.P:
            continue
        for i in range(N):
            for key in PROBE_PARAMS[fam]:
                rec = gold['params'][('%s.%d.%s' % (fam, i, key)) if dd == d0 else ('%s_%s.%d.%s' % (fam, dd, i, key))]
                got = probe(tr.P[name][i][key])
                for f in ('mean', 'absmean', 'l2'):
                    assert close(got[f], rec['post'][f], 1e-4, 1e-9), (fam, i, key, f, got[f], rec['post'][f])
                for a, b in zip(got['samples'], rec['post']['samples']):
                    # Adam's first step moves every weight by ~lr*sign(g): compare with an lr-sized floor
                    assert abs(a - b) <= 2.1 * hp['lr'] * 1.0 + 1e-6, (fam, i, key, a, b)
                if 'grad' in rec:
                    gg = probe(tr.P[name][i][key].grad)
                    if key.endswith('.bias') and 'dec.model.9' not in key and 'mlp' not in key:
                        continue
                    assert close(gg['l2'], rec['grad']['l2'], 2e-3, 1e-9), (fam, i, key, gg['l2'], rec['grad']['l2'])

    # fresh forward of the updated member-0 generator
    with torch.no_grad():
        g0 = tr.state('gen_' + d0, 0)
        src = x_a if d0 == 'a2b' else x_b
        c = co.content_encode(g0, hp, src)
        s = torch.randn(gold['batch'], hp['gen']['style_dim'], 1, 1, generator=torch.Generator().manual_seed(5))
        xf, mask = co.decode(g0, hp, c, s, src)
    for got, want in ((probe(xf, 16), gold['post_x_fake0']), (probe(mask, 16), gold['post_mask0'])):
        assert close(got['absmean'], want['absmean'], 2e-3), (got['absmean'], want['absmean'])
        assert close(got['mean'], want['mean'], 2e-3, 2e-4), (got['mean'], want['mean'])


def test_oracle_three_iterations_match_reference():
    """Three consecutive iterations (council flip 2 on / 1 off, StepLR step_size 2): per-iteration losses, the loss-matching
    ratio and the final learning rate against the unmodified reference."""
    from test_trainer_host_cpu import run_oracle
    gold = load_golden('glasses64_n2_b2_iter3')
    torch.set_num_threads(8)
    log = []

    def grab(k, tr):
        log.append({'dis': [float(v) for v in tr.loss_dis_total_s], 'gen': [float(v) for v in tr.loss_gen_total_s],
                    'disc': [float(v) for v in tr.loss_dis_council_total_s] if tr.disc_ran else [], 'ran': tr.disc_ran,
                    'w': float(tr.w_match['a2b'])})
    tr, hp = run_oracle(gold, torch.float32, n_iters=3, on_iter=grab)
    # Tolerance per iteration: the first two Adam steps are lr*sign(g)-like and the focus loss 1/(|m-0.5|+eps) sits on masks
    # near 0.5, so fp32 summation-order noise is amplified from one iteration to the next -- the fp64 oracle differs from the
    # fp32 oracle by 3e-3 at the third iteration (and by 1e-6 at the first).  The exact check of the carried state is
    # tests/test_trainer_host_cpu.py::test_three_iterations_with_flips_and_lr_decay_exact_in_fp64.
    tol = [2e-5, 1e-4, 5e-3]
    for k, (got, want) in enumerate(zip(log, gold['iters'])):
        assert got['ran'] == want['dis_council_ran'], k
        assert close(got['w'], want['w_match'], 10 * tol[k]), (k, got['w'], want['w_match'])
        for name, key in (('dis', 'loss_dis_total'), ('gen', 'loss_gen_total'), ('disc', 'loss_dis_council_total')):
            assert len(got[name]) == len(want[key])
            for a, b in zip(got[name], want[key]):
                assert close(a, b, tol[k], 1e-7), (k, name, a, b)
    assert abs(tr.lr_now() - gold['lr_after']['gen'][0]) < 1e-15


def test_both_directions_fixture_has_both():
    gold = load_golden('glasses64_n2_b2_both')
    assert set(gold['dirs']) == {'a2b', 'b2a'} and any(k.startswith('gen_b2a.') for k in gold['params'])
