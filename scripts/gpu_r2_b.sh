#!/bin/bash
# Round 2, visit B: single-launch normalisation kernels -- tests, isolated timing, A/B inside the step, ncu DRAM bytes.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2b_*
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "norm or image_helpers or fused_losses" > gpurun_out/r2b_pytest_norm.log 2>&1
echo "pytest norm rc=$?" >> gpurun_out/r2b_summary.txt; tail -5 gpurun_out/r2b_pytest_norm.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x > gpurun_out/r2b_pytest_trainer.log 2>&1
echo "pytest trainer rc=$?" >> gpurun_out/r2b_summary.txt; tail -5 gpurun_out/r2b_pytest_trainer.log
for geo in "4 8 64 64 256" "4 8 128 128 128" "4 8 256 256 64"; do
  timeout 120 python scripts/prof_norm.py $geo >> gpurun_out/r2b_norm_timing.log 2>&1
done
for mb in 32 96; do
  COUNCIL_NORM_L2_MB=$mb timeout 120 python scripts/prof_norm.py 4 8 64 64 256 >> gpurun_out/r2b_norm_timing_budget$mb.log 2>&1
  COUNCIL_NORM_L2_MB=$mb timeout 120 python scripts/prof_norm.py 4 8 256 256 64 >> gpurun_out/r2b_norm_timing_budget$mb.log 2>&1
done
cat gpurun_out/r2b_norm_timing*.log
COUNCIL_COOP_NORM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2b_bench_coop0.json 2> gpurun_out/r2b_bench_coop0.err
COUNCIL_COOP_NORM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2b_bench_coop1.json 2> gpurun_out/r2b_bench_coop1.err
COUNCIL_COOP_NORM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2b_bench_coop0b.json 2> gpurun_out/r2b_bench_coop0.err
for f in coop0 coop1 coop0b; do python -c "
import json,sys
p=json.load(open('gpurun_out/r2b_bench_$f.json'))
print('$f', p['ms_per_step'], p['clocks'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], p.get('parity_check'))
print(p['hbm_kernel_times_ms_per_step'])
"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'norm_coop' -c 2 -o gpurun_out/r2b_ncu_norm_coop -f python scripts/prof_norm.py 4 8 64 64 256 1 > gpurun_out/r2b_ncu_norm_coop.log 2>&1
tail -3 gpurun_out/r2b_ncu_norm_coop.log
cat gpurun_out/r2b_summary.txt
