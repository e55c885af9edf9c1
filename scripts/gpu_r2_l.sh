#!/bin/bash
# Round 2, visit L: narrower N tiles on small maps (mode bit 23 = old behaviour) -- kernel parity, small-configuration A/B, headline sanity.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2l_*
WIDE=$((7 | (1 << 23)))
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv_fwd_dgrad_wgrad or stats or prod" > gpurun_out/r2l_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/r2l_summary.txt; tail -3 gpurun_out/r2l_pytest_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x > gpurun_out/r2l_pytest_trainer.log 2>&1
echo "pytest trainer rc=$?" >> gpurun_out/r2l_summary.txt; tail -3 gpurun_out/r2l_pytest_trainer.log
for i in 1 2; do
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2l_bench_glasses_narrow$i.json 2> gpurun_out/r2l_bench_glasses_narrow$i.err
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $WIDE > gpurun_out/r2l_bench_glasses_wide$i.json 2> gpurun_out/r2l_bench_glasses_wide$i.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2l_bench_narrow.json 2> gpurun_out/r2l_bench_narrow.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $WIDE > gpurun_out/r2l_bench_wide.json 2> gpurun_out/r2l_bench_wide.err
python - <<'PY'
import json
for f in ('glasses_narrow1', 'glasses_wide1', 'glasses_narrow2', 'glasses_wide2', 'narrow', 'wide'):
    try:
        p = json.load(open('gpurun_out/r2l_bench_%s.json' % f))
        print(f, p['ms_per_step'], p['e2e']['ms_per_step'], p['clocks']['sm_mhz'], 'conv', sum(p['kernel_times_ms_per_step'].values()), p.get('parity_check'))
    except Exception as e:
        print(f, 'failed', e)
PY
cat gpurun_out/r2l_summary.txt
