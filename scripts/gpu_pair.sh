#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/test_kernels.log
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_tc.err
