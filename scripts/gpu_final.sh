#!/bin/bash
# Round-end style visit: the full `-m gpu` suite the way the driver runs it, smoke(), both bench arms with default flags.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/test_gpu_all.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/test_gpu_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_reference.json; tail -4 gpurun_out/bench_reference.err
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?" >> gpurun_out/summary.txt; cut -c1-600 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
cat gpurun_out/summary.txt
