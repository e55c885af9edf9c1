#!/bin/bash
# Round 2, visit C: image-side shared-memory patch kernels (forward + weight gradient), L2-hinted cooperative norm, A/B in the step.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2c_*
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "img or disc0 or dis0 or prod_dc0 or prod_d0 or norm_single" > gpurun_out/r2c_pytest_img.log 2>&1
echo "pytest img rc=$?" >> gpurun_out/r2c_summary.txt; tail -15 gpurun_out/r2c_pytest_img.log
for geo in "4 32 256 256 8 64 3 1 1" "4 32 128 128 8 64 3 1 1" "4 16 256 256 4 64 4 2 1"; do
  timeout 120 python scripts/prof_layer.py all $geo >> gpurun_out/r2c_layers.log 2>&1
done
cat gpurun_out/r2c_layers.log
for geo in "4 8 64 64 256" "4 8 128 128 128" "4 8 256 256 64"; do
  timeout 120 python scripts/prof_norm.py $geo >> gpurun_out/r2c_norm_timing.log 2>&1
done
cat gpurun_out/r2c_norm_timing.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $((1 | 6 | (1<<19))) > gpurun_out/r2c_bench_noimg.json 2> gpurun_out/r2c_bench_noimg.err
COUNCIL_COOP_NORM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c_bench_coop1.json 2> gpurun_out/r2c_bench_coop1.err
for f in default noimg coop1; do python -c "
import json,sys
p=json.load(open('gpurun_out/r2c_bench_$f.json'))
print('$f', p['ms_per_step'], p['clocks'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], p.get('parity_check'))
kt=p['kernel_times_ms_per_step']
print({k:v for k,v in kt.items() if 'Cin8' in k or 'Cin4 ' in k})
"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'img_conv' -c 2 -o gpurun_out/r2c_ncu_img -f python scripts/prof_layer.py all 4 32 256 256 8 64 3 1 1 1 > gpurun_out/r2c_ncu_img.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'norm_coop' -c 2 -o gpurun_out/r2c_ncu_norm_coop -f python scripts/prof_norm.py 4 8 64 64 256 1 > gpurun_out/r2c_ncu_norm_coop.log 2>&1
cat gpurun_out/r2c_summary.txt
