#!/bin/bash
# BASELINE configs 4 and 5 at 2 / 4 / 8 GPUs of one box (one process per GPU, NCCL), plus the multi-GPU parity tests.
# usage (under gpurun --gpus 8): bash tools/multi_gpu_run.sh 2 4 8
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi_gpu or shard" 2>&1 | tail -3
port=29511
for n in "$@"; do
  for cfg in c4 c5; do
    port=$((port + 1))
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --config $cfg --steps 2 --warmup 1 > gpurun_out/${cfg}_n$n.json 2> gpurun_out/${cfg}_n$n.err
    echo "== $cfg N=$n rc=$?"; tail -c 1200 gpurun_out/${cfg}_n$n.json | cut -c1-400
  done
done
