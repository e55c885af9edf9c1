#!/bin/bash
# Round 2, visit I: coalescing epilogue (mode bit 20 = old accumulator-layout stores) -- per-layer A/B, ncu --set full before/after, bench A/B.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2i_*
OLD=$((7 | (1 << 20)))
for a in "dgrad 4 32 256 256 64 128 4 2 1" "fwd 4 8 256 256 64 64 3 1 1" "fwd 4 8 256 256 64 64 1 1 0" "dgrad 4 8 256 256 64 64 1 1 0" "fwd 4 32 256 256 64 128 4 2 1" "fwd 4 8 128 128 128 128 3 1 1" "all 4 8 64 64 256 256 3 1 1" "dgrad 4 8 256 256 64 128 4 2 1" "fwd 4 8 256 256 128 64 3 1 1"; do
  python scripts/prof_layer.py $a 10
  COUNCIL_TC_MODE=$OLD python scripts/prof_layer.py $a 10
done 2>&1 | grep "ms per call" | tee gpurun_out/r2i_timing.log
cap() {  # name mode kernel-regex args...
  name=$1; shift; mode=$1; shift; kr=$1; shift
  COUNCIL_TC_MODE=$mode timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kr --launch-skip 1 -c 1 -f -o gpurun_out/r2i_$name python scripts/prof_layer.py "$@" 1 > gpurun_out/r2i_$name.log 2>&1
  ncu -i gpurun_out/r2i_$name.ncu-rep --page raw --csv > gpurun_out/r2i_$name.raw.csv 2>/dev/null
  python scripts/ncu_pick.py gpurun_out/r2i_$name.raw.csv | tee gpurun_out/r2i_$name.txt
}
cap dgrad_b32_c64_old $OLD conv_tc_kernel dgrad 4 32 256 256 64 128 4 2 1
cap dgrad_b32_c64_new 1 conv_tc_kernel dgrad 4 32 256 256 64 128 4 2 1
cap fwd_256_c64_k3_old $OLD conv_tc_kernel fwd 4 8 256 256 64 64 3 1 1
cap fwd_256_c64_k3_new 1 conv_tc_kernel fwd 4 8 256 256 64 64 3 1 1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2i_bench_new.json 2> gpurun_out/r2i_bench_new.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $OLD > gpurun_out/r2i_bench_old.json 2> gpurun_out/r2i_bench_old.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2i_bench_new2.json 2> gpurun_out/r2i_bench_new2.err
for f in new old new2; do python -c "
import json
p=json.load(open('gpurun_out/r2i_bench_$f.json'))
print('$f', p['ms_per_step'], p['clocks']['sm_mhz'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], 'conv', sum(p['kernel_times_ms_per_step'].values()), p.get('parity_check'))"; done
