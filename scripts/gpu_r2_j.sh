#!/bin/bash
# Round 2, visit J: coalescing epilogue v2 (loads prefetched, narrow tiles only) -- kernel parity, per-layer and whole-step A/B.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2j_*
OLD=$((7 | (1 << 20)))
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv_fwd_dgrad_wgrad or stats or prod" > gpurun_out/r2j_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/r2j_summary.txt; tail -3 gpurun_out/r2j_pytest_kernels.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2j_bench_new.json 2> gpurun_out/r2j_bench_new.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $OLD > gpurun_out/r2j_bench_old.json 2> gpurun_out/r2j_bench_old.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2j_bench_new2.json 2> gpurun_out/r2j_bench_new2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $OLD > gpurun_out/r2j_bench_old2.json 2> gpurun_out/r2j_bench_old2.err
for f in new old new2 old2; do python -c "
import json
p=json.load(open('gpurun_out/r2j_bench_$f.json'))
print('$f', p['ms_per_step'], p['clocks']['sm_mhz'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], 'conv', sum(p['kernel_times_ms_per_step'].values()), p.get('parity_check'))"; done
cat gpurun_out/r2j_summary.txt
