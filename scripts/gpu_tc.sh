#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scripts/grad_noise.py > gpurun_out/grad_noise.log 2>&1
grep -v Warn gpurun_out/grad_noise.log | tail -80
TC_ONLY=wgrad timeout 300 python scripts/tc_debug.py > gpurun_out/tc_debug_wgrad.log 2>&1
cut -c1-200 gpurun_out/tc_debug_wgrad.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -12 gpurun_out/test_kernels.log
timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?"; cut -c1-1500 gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
