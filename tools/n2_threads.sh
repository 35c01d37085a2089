#!/bin/bash
# Token-shard overlap at N GPUs vs NCCL's threads per CTA: a 256-thread NCCL CTA fits the registers a dW / dH GEMM CTA
# leaves on its SM (24 K), a default 512/640-thread one has to wait for a whole SM.  KD_BENCH_QUICK: step time only.
N=${1:-2}
port=29700
run() {
  name=$1; shift
  port=$((port+1))
  env KD_BENCH_QUICK=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 60 --warmup 10 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/$name /"
}
run "nocomm" KD_BENCH_SYNC=none_
run "default" KD_BENCH_NCCL_CTAS=32
run "nthreads256" NCCL_NTHREADS=256 KD_BENCH_NCCL_CTAS=32
run "nthreads128 ctas64" NCCL_NTHREADS=128 KD_BENCH_NCCL_CTAS=64
run "nthreads256 ranges12" NCCL_NTHREADS=256 KD_BENCH_NCCL_CTAS=32 KD_BENCH_RANGES=12
run "serial" KD_BENCH_SYNC=serial
