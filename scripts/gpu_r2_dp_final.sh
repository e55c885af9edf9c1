#!/bin/bash
# Round 2, final multi-GPU check (gpurun --gpus 2): data-parallel correctness on NCCL with the final kernels, N=1 vs N=2 on the same box.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2zdp_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 scripts/dp_check.py > gpurun_out/r2zdp_dp_check.log 2>&1
echo "dp check rc=$?" >> gpurun_out/r2zdp_summary.txt; tail -6 gpurun_out/r2zdp_dp_check.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2zdp_bench_n1.json 2> gpurun_out/r2zdp_bench_n1.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2zdp_bench_n2.out 2> gpurun_out/r2zdp_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/r2zdp_summary.txt
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/r2zdp_bench_ref_n2.out 2> gpurun_out/r2zdp_bench_ref_n2.err
echo "bench ref n2 rc=$?" >> gpurun_out/r2zdp_summary.txt
python - <<'PY'
import json
def last_json(path):
    for l in reversed(open(path).read().splitlines()):
        if l.startswith('{'):
            return json.loads(l)
for f in ('gpurun_out/r2zdp_bench_n1.json', 'gpurun_out/r2zdp_bench_n2.out', 'gpurun_out/r2zdp_bench_ref_n2.out'):
    try:
        p = last_json(f)
        print(f, p.get('impl', 'ours'), 'n_gpus', p['n_gpus'], 'ms', round(p['ms_per_step'], 2), 'img/s', round(p['value'], 2), p.get('clocks'))
    except Exception as e:
        print(f, 'failed', e)
PY
cat gpurun_out/r2zdp_summary.txt
