#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/test_kernels.log
for tc in 1 262151; do
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tc $tc > gpurun_out/bench_tc$tc.json 2> gpurun_out/bench_tc$tc.err
echo "bench tc=$tc rc=$?"; tail -2 gpurun_out/bench_tc$tc.err
done
