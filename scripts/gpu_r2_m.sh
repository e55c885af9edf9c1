#!/bin/bash
# Round 2, visit M: host-side launch cost profile on the small configuration; weight-gradient split plan for small problems.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2m_*
timeout 600 python scripts/prof_host.py glasses_128_n2_b1 > gpurun_out/r2m_host_profile_glasses.txt 2>&1
head -3 gpurun_out/r2m_host_profile_glasses.txt
timeout 600 python scripts/prof_host.py male2female_256_n4_b8 > gpurun_out/r2m_host_profile_m2f.txt 2>&1
head -3 gpurun_out/r2m_host_profile_m2f.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv_fwd_dgrad_wgrad or prod" > gpurun_out/r2m_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/r2m_summary.txt; tail -3 gpurun_out/r2m_pytest_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x > gpurun_out/r2m_pytest_trainer.log 2>&1
echo "pytest trainer rc=$?" >> gpurun_out/r2m_summary.txt; tail -3 gpurun_out/r2m_pytest_trainer.log
for i in 1 2; do
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2m_bench_glasses$i.json 2> gpurun_out/r2m_bench_glasses$i.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python - <<'PY'
import json
for f in ('glasses1', 'glasses2', ''):
    try:
        p = json.load(open('gpurun_out/r2m_bench%s.json' % (('_' + f) if f else '')))
        print(f, p['ms_per_step'], p['e2e']['ms_per_step'], p['clocks']['sm_mhz'], 'conv', sum(p['kernel_times_ms_per_step'].values()), p.get('parity_check'))
    except Exception as e:
        print(f, 'failed', e)
PY
cat gpurun_out/r2m_summary.txt
