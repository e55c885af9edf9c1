#!/bin/bash
# tests, then A/B benches of tensor-core mode switches in ONE box visit (box-to-box power capping moves results by several %)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -6 gpurun_out/test_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --timeout 300 > gpurun_out/test_trainer.log 2>&1
echo "trainer rc=$?"; grep -E "passed|failed|Error" gpurun_out/test_trainer.log | head -30
for tc in ${AB_MODES:-1 55 15}; do
  timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --tc $tc > gpurun_out/bench_tc$tc.json 2> gpurun_out/bench_tc$tc.err
  echo "bench tc=$tc rc=$?"; cut -c1-330 gpurun_out/bench_tc$tc.json; tail -3 gpurun_out/bench_tc$tc.err
done
cp gpurun_out/bench_tc1.json gpurun_out/bench_tc.json
