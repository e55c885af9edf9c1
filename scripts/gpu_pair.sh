#!/bin/bash
set -x
mkdir -p gpurun_out
COUNCIL_DEBUG=1 timeout 300 python scripts/prof_wgrad_pair.py > gpurun_out/prof_wgrad_pair.log 2>&1
cat gpurun_out/prof_wgrad_pair.log | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc2 -c 1 -o gpurun_out/ncu_wgrad2 -f python scripts/prof_wgrad_pair.py one > gpurun_out/ncu_wgrad2.log 2>&1
tail -2 gpurun_out/ncu_wgrad2.log
