#!/bin/bash
# Round 2, final 8-GPU run (gpurun --gpus 8): N=1 and N=8 on the same box with the final kernels.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2zdp8_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2zdp8_bench_n1.json 2> gpurun_out/r2zdp8_bench_n1.err
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2zdp8_bench_n8.out 2> gpurun_out/r2zdp8_bench_n8.err
echo "bench n8 rc=$?"
python - <<'PY'
import json
def last_json(path):
    for l in reversed(open(path).read().splitlines()):
        if l.startswith('{'):
            return json.loads(l)
for f in ('gpurun_out/r2zdp8_bench_n1.json', 'gpurun_out/r2zdp8_bench_n8.out'):
    try:
        p = last_json(f)
        print(f, 'n_gpus', p['n_gpus'], 'ms', round(p['ms_per_step'], 2), 'img/s', round(p['value'], 2), p.get('clocks'))
    except Exception as e:
        print(f, 'failed', e)
PY
