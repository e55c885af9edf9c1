#!/bin/bash
# Round 2, visit I: ncu --set full of the store-/shared-memory-bound conv_tc layers (one launch each) to decide what to fix next.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2i_*
cap() {  # name kernel-regex args...
  name=$1; shift; kr=$1; shift
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kr --launch-skip 1 -c 1 -f -o gpurun_out/r2i_$name python scripts/prof_layer.py "$@" 1 > gpurun_out/r2i_$name.log 2>&1
  ncu -i gpurun_out/r2i_$name.ncu-rep --page raw --csv > gpurun_out/r2i_$name.raw.csv 2>/dev/null
  python scripts/ncu_pick.py gpurun_out/r2i_$name.raw.csv | tee gpurun_out/r2i_$name.txt
}
cap dgrad_b32_c64 conv_tc_kernel dgrad 4 32 256 256 64 128 4 2 1
cap fwd_256_c64_k3 conv_tc_kernel fwd 4 8 256 256 64 64 3 1 1
cap fwd_256_c64_k1 conv_tc_kernel fwd 4 8 256 256 64 64 1 1 0
cap fwd_b32_c64_k4s2 "conv_tc" fwd 4 32 256 256 64 128 4 2 1
cap wgrad_b32_c64_k4s2 "wgrad_tc_kernel" wgrad 4 32 256 256 64 128 4 2 1
for a in "dgrad 4 32 256 256 64 128 4 2 1" "fwd 4 8 256 256 64 64 3 1 1" "fwd 4 8 256 256 64 64 1 1 0" "all 4 32 256 256 64 128 4 2 1"; do python scripts/prof_layer.py $a 10; done 2>&1 | tee gpurun_out/r2i_timing.log
