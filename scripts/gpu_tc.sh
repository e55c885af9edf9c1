#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -6 gpurun_out/test_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --timeout 300 > gpurun_out/test_trainer.log 2>&1
echo "trainer rc=$?"; grep -E "passed|failed|Error" gpurun_out/test_trainer.log | head -30
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?"; cut -c1-400 gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
