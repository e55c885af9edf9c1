#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -c 1 -o gpurun_out/ncu_wgrad -f python scripts/prof_conv.py wgrad > gpurun_out/ncu_wgrad.log 2>&1
tail -2 gpurun_out/ncu_wgrad.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -c 1 -o gpurun_out/ncu_pair -f python scripts/prof_conv.py fwd > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log
