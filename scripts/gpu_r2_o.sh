#!/bin/bash
# Round 2, visit O: knob sweep on the headline configuration with the final kernels (one box, alternating with the default).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2o_*
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check"
timeout 600 $B > gpurun_out/r2o_default1.json 2> gpurun_out/r2o_default1.err
timeout 600 $B --tc $((7 | 32)) > gpurun_out/r2o_pair64.json 2> gpurun_out/r2o_pair64.err
COUNCIL_FUSE_STATS=1 timeout 600 $B > gpurun_out/r2o_stats_all.json 2> gpurun_out/r2o_stats_all.err
timeout 600 $B > gpurun_out/r2o_default2.json 2> gpurun_out/r2o_default2.err
COUNCIL_FUSE_STATS=0 timeout 600 $B > gpurun_out/r2o_stats_none.json 2> gpurun_out/r2o_stats_none.err
COUNCIL_COOP_NORM=1 timeout 600 $B > gpurun_out/r2o_coop.json 2> gpurun_out/r2o_coop.err
timeout 600 $B --tc $((7 | (1 << 17))) > gpurun_out/r2o_one_cta.json 2> gpurun_out/r2o_one_cta.err
timeout 600 $B > gpurun_out/r2o_default3.json 2> gpurun_out/r2o_default3.err
python - <<'PY'
import json
for f in ('default1', 'pair64', 'stats_all', 'default2', 'stats_none', 'coop', 'one_cta', 'default3'):
    try:
        p = json.load(open('gpurun_out/r2o_%s.json' % f))
        print(f, 'ms', round(p['ms_per_step'], 2), 'e2e', round(p['e2e']['ms_per_step'], 2), p['clocks']['sm_mhz'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'], 'conv', round(sum(p['kernel_times_ms_per_step'].values()), 2))
    except Exception as e:
        print(f, 'failed', e)
PY
