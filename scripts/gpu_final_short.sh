#!/bin/bash
# Round-end style visit without the (unchanged) CPU reference arm: full `-m gpu` suite, smoke(), default bench.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/test_gpu_all.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?" >> gpurun_out/summary.txt; cut -c1-400 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
cat gpurun_out/summary.txt
