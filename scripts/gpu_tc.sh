#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1
cat gpurun_out/tc_debug.log
timeout 600 python bench.py --steps 3 --warmup 1 --tc 0 > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err
echo "bench simt rc=$?"; cat gpurun_out/bench_simt.json; tail -5 gpurun_out/bench_simt.err
