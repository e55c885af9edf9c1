#!/bin/bash
# Round 2, visit H2: the full GPU suite after the input-pipeline fix, then the ncu captures of visit I.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_gpu.log 2>&1
echo "pytest gpu (all) rc=$?" >> gpurun_out/r2h_summary.txt; tail -8 gpurun_out/r2h_pytest_gpu.log
bash scripts/gpu_r2_i.sh
