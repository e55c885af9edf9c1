#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -8 gpurun_out/test_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --timeout 300 > gpurun_out/test_trainer.log 2>&1
echo "trainer rc=$?"; grep -E "tc=|passed|failed|Error" gpurun_out/test_trainer.log | head -30
timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?"; cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
for t in 8 32 128; do
  COUNCIL_CPU_THREADS=$t timeout 200 python bench.py --impl reference --steps 1 --warmup 0 --workload glasses_128_n2_b1 > gpurun_out/cpu_t$t.json 2>&1
  echo "cpu threads $t rc=$?"; tail -1 gpurun_out/cpu_t$t.json | cut -c1-300
done
