#!/bin/bash
# Round-2 experiment (not yet run): weight gradients on a side stream (COUNCIL_WGRAD_STREAM=1) vs the default, same box.
set -x
mkdir -p gpurun_out
COUNCIL_WGRAD_STREAM=1 timeout 600 python -m pytest tests/test_trainer_gpu.py -x -q -m gpu > gpurun_out/test_trainer_wgrad_stream.log 2>&1
echo "trainer (side stream) rc=$?"; tail -3 gpurun_out/test_trainer_wgrad_stream.log
for v in 0 1; do
  COUNCIL_WGRAD_STREAM=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wgrad_stream$v.json 2> gpurun_out/bench_wgrad_stream$v.err
  echo "bench COUNCIL_WGRAD_STREAM=$v rc=$?"; cut -c1-200 gpurun_out/bench_wgrad_stream$v.json
done
