#!/bin/bash
set -x
mkdir -p gpurun_out
COUNCIL_DEBUG=1 TC_ONLY=wgrad timeout 600 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1
cut -c1-260 gpurun_out/tc_debug.log
