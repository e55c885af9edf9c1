#!/bin/bash
# One GPU-box visit: kernel parity, trainer parity, smoke, a bench line, and an ncu launch list.
# Usage (from the build container): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick]'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/summary.txt
tail -25 gpurun_out/test_kernels.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
tail -5 gpurun_out/smoke.log
timeout 1200 python -m pytest tests/test_trainer_gpu.py -x -q -m gpu -s > gpurun_out/test_trainer.log 2>&1
echo "trainer rc=$?" >> gpurun_out/summary.txt
tail -25 gpurun_out/test_trainer.log
if [ "$1" != "quick" ]; then
  timeout 900 python bench.py --steps 3 --warmup 1 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?" >> gpurun_out/summary.txt
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 0 --workload male2female_256_n4_b8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
