#!/bin/bash
# Round 2, visit E: TMA-store epilogue of the image forward kernel, fused decoder tail, A/B in the step.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2e_*
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "img or disc0 or dis0 or prod_dc0 or prod_d0 or head_fused" > gpurun_out/r2e_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/r2e_summary.txt; tail -12 gpurun_out/r2e_pytest_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x > gpurun_out/r2e_pytest_trainer.log 2>&1
echo "pytest trainer rc=$?" >> gpurun_out/r2e_summary.txt; tail -8 gpurun_out/r2e_pytest_trainer.log
for geo in "4 32 256 256 8 64 3 1 1" "4 32 128 128 8 64 3 1 1" "4 16 256 256 4 64 4 2 1"; do
  timeout 120 python scripts/prof_layer.py fwd $geo >> gpurun_out/r2e_layers.log 2>&1
done
cat gpurun_out/r2e_layers.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_default.err
COUNCIL_FUSE_HEAD=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2e_bench_nohead.json 2> gpurun_out/r2e_bench_nohead.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2e_bench_default2.json 2> gpurun_out/r2e_bench_default2.err
for f in default nohead default2; do python -c "
import json,sys
p=json.load(open('gpurun_out/r2e_bench_$f.json'))
print('$f', p['ms_per_step'], p['clocks'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], 'conv', sum(p['kernel_times_ms_per_step'].values()), p.get('parity_check'))
print({k:v for k,v in p['hbm_kernel_times_ms_per_step'].items() if 'head' in k})
"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'head_fused|img_conv_fwd' -c 3 -o gpurun_out/r2e_ncu_head -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2e_ncu_head.log 2>&1
cat gpurun_out/r2e_summary.txt
