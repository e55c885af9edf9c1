#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/prof_wgrad_pair.py > gpurun_out/prof_wgrad_km.log 2>&1
tail -5 gpurun_out/prof_wgrad_km.log
