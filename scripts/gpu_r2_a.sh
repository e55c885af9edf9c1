#!/bin/bash
# Round 2, visit A: full GPU suite, smoke, bench with both comparison legs, launch list of smoke, council-D front-end timing + ncu.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2a_*
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
timeout 1500 python -m pytest tests/ -q -m gpu -x --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest -m gpu rc=$?" >> gpurun_out/r2a_summary.txt; tail -30 gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_summary.txt; tail -3 gpurun_out/r2a_smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_summary.txt; cut -c1-1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke_ncu.log 2>&1
echo "smoke launch list rc=$?" >> gpurun_out/r2a_summary.txt
for geo in "4 32 256 256 8 64 3 1 1" "4 32 256 256 64 128 4 2 1" "4 8 256 256 64 64 3 1 1" "4 8 256 256 128 64 3 1 1"; do
  timeout 120 python scripts/prof_layer.py all $geo >> gpurun_out/r2a_layers.log 2>&1
done
cat gpurun_out/r2a_layers.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'wgrad|im2col' -c 4 -o gpurun_out/r2a_ncu_dc0_wgrad -f python scripts/prof_layer.py wgrad 4 32 256 256 8 64 3 1 1 1 > gpurun_out/r2a_ncu_dc0_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_tc' -c 1 -o gpurun_out/r2a_ncu_dc0_fwd -f python scripts/prof_layer.py fwd 4 32 256 256 8 64 3 1 1 1 > gpurun_out/r2a_ncu_dc0_fwd.log 2>&1
ls -la gpurun_out/r2a_*
cat gpurun_out/r2a_summary.txt
