#!/bin/bash
# Round 2, visit P: end-of-round ncu --set full captures of the kernels VERDICT names (weight gradient, normalise passes, single-CTA conv) and
# the PDL crossover on the mid-sized configuration.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2p_*
cap() {  # name kernel-regex script args...
  name=$1; shift; kr=$1; shift
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kr --launch-skip 2 -c 1 -f -o gpurun_out/r2p_$name python "$@" > gpurun_out/r2p_$name.log 2>&1
  ncu -i gpurun_out/r2p_$name.ncu-rep --page raw --csv > gpurun_out/r2p_$name.raw.csv 2>/dev/null
  python scripts/ncu_pick.py gpurun_out/r2p_$name.raw.csv > gpurun_out/r2p_$name.txt
  grep -E "gpu__time_duration.sum|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_(read|write).sum |dram__throughput.avg.pct" gpurun_out/r2p_$name.txt
}
cap wgrad_3x3_256 wgrad_tc_kernel scripts/prof_layer.py wgrad 4 8 64 64 256 256 3 1 1 3
cap dgrad_3x3_256 conv_tc2_kernel scripts/prof_layer.py dgrad 4 8 64 64 256 256 3 1 1 3
cap norm_act_fwd norm_act_fwd_kernel scripts/prof_norm.py 4 8 64 64 256 2
cap norm_bwd_apply norm_bwd_apply_kernel scripts/prof_norm.py 4 8 64 64 256 2
cap norm_bwd_partial norm_bwd_partial_kernel scripts/prof_norm.py 4 8 64 64 256 2
cap img_wgrad_dc0 img_conv_wgrad_kernel scripts/prof_layer.py wgrad 4 32 256 256 8 64 3 1 1 3
B="python bench.py --workload selfie2anime_256_n4_b4 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check"
for i in 1 2; do
COUNCIL_PDL=0 timeout 300 $B > gpurun_out/r2p_anime_pdl0_$i.json 2> gpurun_out/r2p_anime_pdl0_$i.err
COUNCIL_PDL=1 timeout 300 $B > gpurun_out/r2p_anime_pdl1_$i.json 2> gpurun_out/r2p_anime_pdl1_$i.err
done
B="python bench.py --workload tiny_64_n2_b2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity-check"
COUNCIL_PDL=0 timeout 300 $B > gpurun_out/r2p_tiny_pdl0.json 2> gpurun_out/r2p_tiny_pdl0.err
COUNCIL_PDL=1 timeout 300 $B > gpurun_out/r2p_tiny_pdl1.json 2> gpurun_out/r2p_tiny_pdl1.err
python - <<'PY'
import json
for f in ('anime_pdl0_1', 'anime_pdl1_1', 'anime_pdl0_2', 'anime_pdl1_2', 'tiny_pdl0', 'tiny_pdl1'):
    try:
        p = json.load(open('gpurun_out/r2p_%s.json' % f))
        print(f, 'ms', round(p['ms_per_step'], 3), 'e2e', round(p['e2e']['ms_per_step'], 3), p['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'failed', e)
PY
