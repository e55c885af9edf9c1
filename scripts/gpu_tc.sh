#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?"; tail -6 gpurun_out/test_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu -s --timeout 300 > gpurun_out/test_trainer.log 2>&1
echo "trainer rc=$?"; grep -E "passed|failed|Error" gpurun_out/test_trainer.log | head -30
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?"; cut -c1-400 gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
for wl in selfie2anime_256_n4_b4 male2female_512_n6_b2 glasses_128_n2_b1; do
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl rc=$?"; cut -c1-300 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"
