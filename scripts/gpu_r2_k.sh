#!/bin/bash
# Round 2, visit K: programmatic dependent launch between the library's kernels -- full GPU suite, whole-step A/B (mode bit 22 = plain launches).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2k_*
NOPDL=$((7 | (1 << 22)))
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" >> gpurun_out/r2k_summary.txt; tail -3 gpurun_out/r2k_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2k_bench_pdl.json 2> gpurun_out/r2k_bench_pdl.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $NOPDL > gpurun_out/r2k_bench_nopdl.json 2> gpurun_out/r2k_bench_nopdl.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2k_bench_pdl2.json 2> gpurun_out/r2k_bench_pdl2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $NOPDL > gpurun_out/r2k_bench_nopdl2.json 2> gpurun_out/r2k_bench_nopdl2.err
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2k_bench_glasses_pdl.json 2> gpurun_out/r2k_bench_glasses_pdl.err
timeout 300 python bench.py --workload glasses_128_n2_b1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity-check --tc $NOPDL > gpurun_out/r2k_bench_glasses_nopdl.json 2> gpurun_out/r2k_bench_glasses_nopdl.err
python - <<'PY'
import json
for f in ('pdl', 'nopdl', 'pdl2', 'nopdl2', 'glasses_pdl', 'glasses_nopdl'):
    try:
        p = json.load(open('gpurun_out/r2k_bench_%s.json' % f))
        print(f, p['ms_per_step'], p['e2e']['ms_per_step'], p['clocks']['sm_mhz'], p.get('parity_check'))
    except Exception as e:
        print(f, 'failed', e)
PY
cat gpurun_out/r2k_summary.txt
