#!/bin/bash
# Round 2, multi-GPU visit (run with gpurun --gpus N): data-parallel correctness on NCCL, overlapped vs joined gradient exchange.
N=${1:-2}
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2dp${N}_*
nvidia-smi -L > gpurun_out/r2dp${N}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_check.py > gpurun_out/r2dp${N}_dp_check.log 2>&1
echo "dp check rc=$?"; tail -8 gpurun_out/r2dp${N}_dp_check.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/r2dp${N}_bench_n1.json 2> gpurun_out/r2dp${N}_bench_n1.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2dp${N}_bench_overlap.json 2> gpurun_out/r2dp${N}_bench_overlap.err
COUNCIL_DP_SYNC=1 timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2dp${N}_bench_sync.json 2> gpurun_out/r2dp${N}_bench_sync.err
timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2dp${N}_bench_overlap2.json 2> gpurun_out/r2dp${N}_bench_overlap2.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 300 $TR --master-port 29516 bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2dp${N}_bench_nccl_debug.json 2> gpurun_out/r2dp${N}_nccl_debug.log
grep -E "NVLS|Channel|algo|proto|Using network|comm .* rank 0" gpurun_out/r2dp${N}_nccl_debug.log | head -40 > gpurun_out/r2dp${N}_nccl_summary.txt
for f in n1 overlap sync overlap2; do python -c "
import json
p=json.load(open('gpurun_out/r2dp${N}_bench_$f.json'))
print('$f', 'n_gpus', p['n_gpus'], 'ms', round(p['ms_per_step'],2), 'img/s', round(p['value'],1), p['clocks'])
"; done
tail -3 gpurun_out/r2dp${N}_bench_overlap.err
