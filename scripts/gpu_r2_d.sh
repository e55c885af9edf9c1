#!/bin/bash
# Round 2, visit D: batched-load image kernels, statistics epilogue in the CTA-pair kernel, A/B in the step, other BASELINE configs.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2d_*
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "img or disc0 or dis0 or prod_dc0 or prod_d0 or stats" > gpurun_out/r2d_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/r2d_summary.txt; tail -8 gpurun_out/r2d_pytest_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x > gpurun_out/r2d_pytest_trainer.log 2>&1
echo "pytest trainer rc=$?" >> gpurun_out/r2d_summary.txt; tail -8 gpurun_out/r2d_pytest_trainer.log
for geo in "4 32 256 256 8 64 3 1 1" "4 32 128 128 8 64 3 1 1" "4 16 256 256 4 64 4 2 1"; do
  timeout 120 python scripts/prof_layer.py all $geo >> gpurun_out/r2d_layers.log 2>&1
done
cat gpurun_out/r2d_layers.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err
COUNCIL_FUSE_STATS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2d_bench_nostats.json 2> gpurun_out/r2d_bench_nostats.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2d_bench_default2.json 2> gpurun_out/r2d_bench_default2.err
for f in default nostats default2; do python -c "
import json,sys
p=json.load(open('gpurun_out/r2d_bench_$f.json'))
print('$f', p['ms_per_step'], p['clocks'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], 'conv', sum(p['kernel_times_ms_per_step'].values()), p.get('parity_check'))
"; done
for wl in selfie2anime_256_n4_b4 male2female_512_n6_b2 glasses_128_n2_b1; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2d_bench_$wl.json 2> gpurun_out/r2d_bench_$wl.err
  python -c "
import json
p=json.load(open('gpurun_out/r2d_bench_$wl.json'))
print('$wl', round(p['ms_per_step'],2), 'img/s', round(p['value'],1), 'e2e', round(p['e2e']['value'],1), p['parity_check'], 'gpu baseline', {k:p['gpu_library_baseline'].get(k) for k in ('value','value_cudnn_benchmark','ours_over_baseline','kind','error')}, 'cpu', p['cpu_baseline'].get('value'))
"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'img_conv' -c 2 -o gpurun_out/r2d_ncu_img -f python scripts/prof_layer.py all 4 32 256 256 8 64 3 1 1 1 > gpurun_out/r2d_ncu_img.log 2>&1
cat gpurun_out/r2d_summary.txt
