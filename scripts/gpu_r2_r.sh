#!/bin/bash
# Round 2, visit R: PDL for helper kernels only (mode bit 24) vs off vs everywhere, headline configuration, COUNCIL_PDL=0 so that the mask decides.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2r_*
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check"
for i in 1 2 3; do
timeout 600 $B > gpurun_out/r2r_off$i.json 2> gpurun_out/r2r_off$i.err
timeout 600 $B --tc $((7 | (1 << 24))) > gpurun_out/r2r_helpers$i.json 2> gpurun_out/r2r_helpers$i.err
done
python - <<'PY'
import json
for f in ('off1', 'helpers1', 'off2', 'helpers2', 'off3', 'helpers3'):
    try:
        p = json.load(open('gpurun_out/r2r_%s.json' % f))
        print(f, 'ms', round(p['ms_per_step'], 2), 'e2e', round(p['e2e']['ms_per_step'], 2), p['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'failed', e)
PY
