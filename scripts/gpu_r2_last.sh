#!/bin/bash
# Round 2, last check of the final library build: GPU suite, smoke, one default bench line (no baselines).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2last_*
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2last_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" >> gpurun_out/r2last_summary.txt; tail -2 gpurun_out/r2last_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2last_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2last_summary.txt; tail -1 gpurun_out/r2last_smoke.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2last_bench.json 2> gpurun_out/r2last_bench.err
echo "bench rc=$?" >> gpurun_out/r2last_summary.txt
python -c "
import json
p=json.load(open('gpurun_out/r2last_bench.json'))
print(p['ms_per_step'], p['value'], p['e2e']['value'], p['clocks'], p['parity_check'], p['roofline']['tensor_pipe_pct_ncu'], p['gpu_launches'])"
cat gpurun_out/r2last_summary.txt
