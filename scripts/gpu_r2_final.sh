#!/bin/bash
# Round 2, final visit: what the driver runs (GPU suite, smoke, both bench arms) plus the launch list and one ncu --set full of the dominant kernel.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2z_*
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" >> gpurun_out/r2z_summary.txt; tail -3 gpurun_out/r2z_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2z_summary.txt; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench_default_n1.json 2> gpurun_out/r2z_bench_default_n1.err
echo "bench rc=$?" >> gpurun_out/r2z_summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err
echo "bench ref rc=$?" >> gpurun_out/r2z_summary.txt
for w in selfie2anime_256_n4_b4 male2female_512_n6_b2 glasses_128_n2_b1; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_$w.json 2> gpurun_out/r2z_bench_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2z_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2z_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc2_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r2z_ncu_conv_pair_fwd_3x3_256 python scripts/prof_layer.py fwd 4 8 64 64 256 256 3 1 1 1 > gpurun_out/r2z_ncu_conv_pair.log 2>&1
ncu -i gpurun_out/r2z_ncu_conv_pair_fwd_3x3_256.ncu-rep --page raw --csv > gpurun_out/r2z_ncu_conv_pair.raw.csv 2>/dev/null
python scripts/ncu_pick.py gpurun_out/r2z_ncu_conv_pair.raw.csv > gpurun_out/r2z_ncu_conv_pair_fwd_3x3_256.txt
python - <<'PY'
import json
for f in ('default_n1', 'selfie2anime_256_n4_b4', 'male2female_512_n6_b2', 'glasses_128_n2_b1'):
    try:
        p = json.load(open('gpurun_out/r2z_bench_%s.json' % f))
        g = p.get('gpu_library_baseline') or {}
        print(f, 'ms', round(p['ms_per_step'], 2), 'img/s', round(p['value'], 1), 'e2e', round(p['e2e']['value'], 1), p['clocks']['sm_mhz'], p.get('parity_check'), 'ref-gpu', g.get('value'), g.get('value_cudnn_benchmark'))
    except Exception as e:
        print(f, 'failed', e)
print(open('gpurun_out/r2z_bench_reference_arm.json').read()[:400])
PY
cat gpurun_out/r2z_summary.txt
