#!/bin/bash
# Round 2, visit N: host launch path (stream pinned per update, cached parameter views), clean headline region; PDL A/B without per-kernel events.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2n_*
NOPDL=$((7 | (1 << 22)))
timeout 600 python scripts/prof_host.py glasses_128_n2_b1 > gpurun_out/r2n_host_profile_glasses.txt 2>&1
head -1 gpurun_out/r2n_host_profile_glasses.txt
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x > gpurun_out/r2n_pytest_trainer.log 2>&1
echo "pytest trainer rc=$?" >> gpurun_out/r2n_summary.txt; tail -3 gpurun_out/r2n_pytest_trainer.log
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2n_bench_pdl$i.json 2> gpurun_out/r2n_bench_pdl$i.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $NOPDL > gpurun_out/r2n_bench_nopdl$i.json 2> gpurun_out/r2n_bench_nopdl$i.err
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2n_bench_glasses_pdl$i.json 2> gpurun_out/r2n_bench_glasses_pdl$i.err
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $NOPDL > gpurun_out/r2n_bench_glasses_nopdl$i.json 2> gpurun_out/r2n_bench_glasses_nopdl$i.err
done
python - <<'PY'
import json
for f in ('pdl1', 'nopdl1', 'pdl2', 'nopdl2', 'glasses_pdl1', 'glasses_nopdl1', 'glasses_pdl2', 'glasses_nopdl2'):
    try:
        p = json.load(open('gpurun_out/r2n_bench_%s.json' % f))
        print(f, 'value-region', p['ms_per_step'], 'e2e', p['e2e']['ms_per_step'], 'profiled', p['roofline']['timed_region_ms_per_step'], p['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'failed', e)
PY
cat gpurun_out/r2n_summary.txt
