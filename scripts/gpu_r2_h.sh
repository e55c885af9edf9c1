#!/bin/bash
# Round 2, visit H: input-pipeline kernels vs oracle, tensor-map cache / per-thread switches, full GPU suite, smoke launch list, bench (both arms).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2h_*
timeout 600 python -m pytest tests/test_augment_gpu.py -q -x > gpurun_out/r2h_pytest_augment.log 2>&1
echo "pytest augment rc=$?" >> gpurun_out/r2h_summary.txt; tail -15 gpurun_out/r2h_pytest_augment.log
timeout 300 python scripts/prof_augment.py > gpurun_out/r2h_augment_timing.json 2> gpurun_out/r2h_augment_timing.err; cat gpurun_out/r2h_augment_timing.json; tail -3 gpurun_out/r2h_augment_timing.err
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" >> gpurun_out/r2h_summary.txt; tail -8 gpurun_out/r2h_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2h_summary.txt
python -c "
import json
p=json.load(open('gpurun_out/r2h_bench_default.json'))
print(p['ms_per_step'], p['value'], p['clocks'], p['e2e'], p.get('parity_check'), 'launches', p['gpu_launches'], p.get('tensor_map_cache'))
print(p['roofline']); print(p['cpu_baseline']); print(p.get('gpu_library_baseline'))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
echo "bench ref rc=$?" >> gpurun_out/r2h_summary.txt; cat gpurun_out/r2h_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2h_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2h_summary.txt; tail -3 gpurun_out/r2h_smoke.log
cat gpurun_out/r2h_summary.txt
