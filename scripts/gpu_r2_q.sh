#!/bin/bash
# Round 2, visit Q: cooperative single-launch norm BACKWARD only (keeps the statistics epilogue of the forward) vs default.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2q_*
COUNCIL_COOP_NORM=bwd timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x -k "golden" > gpurun_out/r2q_pytest_trainer_coop_bwd.log 2>&1
echo "pytest trainer (coop bwd) rc=$?" >> gpurun_out/r2q_summary.txt; tail -2 gpurun_out/r2q_pytest_trainer_coop_bwd.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check"
for i in 1 2 3; do
timeout 600 $B > gpurun_out/r2q_default$i.json 2> gpurun_out/r2q_default$i.err
COUNCIL_COOP_NORM=bwd timeout 600 $B > gpurun_out/r2q_coop_bwd$i.json 2> gpurun_out/r2q_coop_bwd$i.err
done
python - <<'PY'
import json
for f in ('default1', 'coop_bwd1', 'default2', 'coop_bwd2', 'default3', 'coop_bwd3'):
    try:
        p = json.load(open('gpurun_out/r2q_%s.json' % f))
        print(f, 'ms', round(p['ms_per_step'], 2), 'e2e', round(p['e2e']['ms_per_step'], 2), p['clocks']['sm_mhz'], 'hbm', p['roofline_hbm']['all_hbm_kernels_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
cat gpurun_out/r2q_summary.txt
