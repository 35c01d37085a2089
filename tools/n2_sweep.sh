#!/bin/bash
# tuning sweep of the overlapped dW all-reduce at N GPUs (default 2): NCCL CTAs x ranges x SM limit
N=${1:-2}
port=29600
for cfg in "serial 16 6 0" "overlap 16 6 0" "overlap 16 6 1" "overlap 8 6 0" "overlap 32 6 0" "overlap 16 3 0" "overlap 16 9 0"; do
  set -- $cfg
  port=$((port+1))
  KD_BENCH_QUICK=1 KD_BENCH_SYNC=$1 KD_BENCH_NCCL_CTAS=$2 KD_BENCH_RANGES=$3 KD_BENCH_SM_LIMIT=$4 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 30 --warmup 5 2>/dev/null | tail -1 | sed "s/^/sync=$1 ctas=$2 ranges=$3 smlimit=$4 /"
done
