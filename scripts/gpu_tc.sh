#!/bin/bash
set -x
mkdir -p gpurun_out
TC_ONLY=wgrad timeout 600 python scripts/tc_debug.py > gpurun_out/tc_debug_wgrad.log 2>&1
cat gpurun_out/tc_debug_wgrad.log | cut -c1-400
timeout 200 python bench.py --steps 2 --warmup 1 --workload tiny_64_n2_b2 --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
echo "bench tiny rc=$?"; cat gpurun_out/bench_tiny.json; tail -5 gpurun_out/bench_tiny.err
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?"; cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -15 gpurun_out/test_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --timeout 300 > gpurun_out/test_trainer.log 2>&1
echo "trainer rc=$?"; grep -E "tc=|passed|failed|Error" gpurun_out/test_trainer.log | head -30
