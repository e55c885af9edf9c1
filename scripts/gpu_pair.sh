#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/prof_pair.py > gpurun_out/prof_pair.log 2>&1
cat gpurun_out/prof_pair.log
bash scripts/gpu_tc.sh
