#!/bin/bash
# overlapped dW all-reduce at N GPUs: NCCL CTA cap x number of ranges (KD_BENCH_QUICK = device-resident step only)
N=${1:-8}
port=29900
for cfg in "overlap 16 6" "overlap 8 6" "overlap 32 12" "overlap 16 12" "overlap 24 9"; do
  set -- $cfg
  port=$((port+1))
  KD_BENCH_QUICK=1 KD_BENCH_SYNC=$1 KD_BENCH_NCCL_CTAS=$2 KD_BENCH_RANGES=$3 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 30 --warmup 5 2>/dev/null | tail -1 | sed "s/^/sync=$1 ctas=$2 ranges=$3 /"
done
