#!/bin/bash
# Round 2, visit G: vector-shaped layer kernels (heads, MLP gradient), full trainer parity, bench A/B (old SIMT path cannot be selected: compare with visit F).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2g_*
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "dis_out or mlp_linear or conv_fwd_dgrad" > gpurun_out/r2g_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/r2g_summary.txt; tail -6 gpurun_out/r2g_pytest_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x > gpurun_out/r2g_pytest_trainer.log 2>&1
echo "pytest trainer rc=$?" >> gpurun_out/r2g_summary.txt; tail -6 gpurun_out/r2g_pytest_trainer.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2g_bench_default2.json 2> gpurun_out/r2g_bench_default2.err
for f in default default2; do python -c "
import json,sys
p=json.load(open('gpurun_out/r2g_bench_$f.json'))
print('$f', p['ms_per_step'], p['clocks'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], 'conv', sum(p['kernel_times_ms_per_step'].values()), p.get('parity_check'), 'launches', p['gpu_launches'])
print({k:p['roofline'].get(k) for k in ('achieved','frac','tf32_cublas_tflops_measured_here','frac_of_tf32_cublas','tensor_pipe_pct_ncu')})
"; done
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2g_bench_glasses.json 2> gpurun_out/r2g_bench_glasses.err
python -c "
import json
p=json.load(open('gpurun_out/r2g_bench_glasses.json')); print('glasses', p['ms_per_step'], p['value'], p['parity_check'])"
cat gpurun_out/r2g_summary.txt
