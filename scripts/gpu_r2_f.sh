#!/bin/bash
# Round 2, visit F: full GPU suite + smoke + launch list of one bench step (where the non-conv time goes) + default bench with both baselines.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2f_*
timeout 1500 python -m pytest tests/ -q -m gpu --durations=8 > gpurun_out/r2f_pytest.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/r2f_summary.txt; tail -14 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2f_summary.txt; tail -2 gpurun_out/r2f_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?" >> gpurun_out/r2f_summary.txt; cut -c1-400 gpurun_out/r2f_bench.json; tail -4 gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2f_launches_run.log 2>&1
echo "launch list rc=$?" >> gpurun_out/r2f_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>&1
cat gpurun_out/r2f_summary.txt
