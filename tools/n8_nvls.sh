#!/bin/bash
# Token-shard dW all-reduce at N GPUs: default NCCL algorithm vs NVLS (in-switch reduction) at several CTA caps, and the
# same-box single-GPU step for the efficiency (KD_BENCH_QUICK = device-resident step only, 60 steps).
N=${1:-8}
port=29800
KD_BENCH_QUICK=1 python bench.py --gpus 1 --steps 60 --warmup 10 2>/dev/null | tail -1 | sed "s/^/N=1 /"
run() {  # name, extra env...
  name=$1; shift
  port=$((port+1))
  env KD_BENCH_QUICK=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 60 --warmup 10 2>/dev/null | tail -1 | sed "s/^/$name /"
}
run "default ctas=32 ranges=6" KD_BENCH_NCCL_CTAS=32 KD_BENCH_RANGES=6
run "nvls ctas=16 ranges=6" NCCL_ALGO=NVLS KD_BENCH_NCCL_CTAS=16 KD_BENCH_RANGES=6
run "nvls ctas=8 ranges=6" NCCL_ALGO=NVLS KD_BENCH_NCCL_CTAS=8 KD_BENCH_RANGES=6
run "nvls ctas=32 ranges=6" NCCL_ALGO=NVLS KD_BENCH_NCCL_CTAS=32 KD_BENCH_RANGES=6
run "after ctas=32" KD_BENCH_SYNC=after KD_BENCH_NCCL_CTAS=32
