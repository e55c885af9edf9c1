#!/bin/bash
# Round 2, 8-GPU visit: overlapped vs joined gradient exchange at the BASELINE configs[3] shape (8 images per GPU, global batch 64).
N=8
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2dp8_*
nvidia-smi -L > gpurun_out/r2dp8_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2dp8_bench_overlap.json 2> gpurun_out/r2dp8_bench_overlap.err
COUNCIL_DP_SYNC=1 timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2dp8_bench_sync.json 2> gpurun_out/r2dp8_bench_sync.err
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2dp8_bench_n1.json 2> gpurun_out/r2dp8_bench_n1.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL,TUNING timeout 300 $TR --master-port 29516 bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "AllReduce: [0-9]+ Bytes" | sort | uniq -c > gpurun_out/r2dp8_nccl_algorithms.txt
cat gpurun_out/r2dp8_nccl_algorithms.txt | cut -c1-160
for f in n1 overlap sync; do python - <<PY
import json
txt=open('gpurun_out/r2dp8_bench_$f.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
p=json.loads(line)
print('$f', 'n_gpus', p['n_gpus'], 'ms', round(p['ms_per_step'],2), 'img/s', round(p['value'],1), p['clocks'])
PY
done
