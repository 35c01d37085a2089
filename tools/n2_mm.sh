#!/bin/bash
# NCCL vs the NVLS multimem backend of GradSync at N GPUs (KD_BENCH_QUICK: step time only), then the full line with
# the parity block for the multimem backend.
N=${1:-2}
port=29650
run() {
  name=$1; shift
  port=$((port+1))
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 60 --warmup 10 2>gpurun_out/mm_err_$port.log | tail -1 | cut -c1-${CUT:-110} | sed "s/^/$name /"
}
run "nocomm" KD_BENCH_QUICK=1 KD_BENCH_SYNC=none_
run "nccl" KD_BENCH_QUICK=1
run "multimem ctas16" KD_BENCH_QUICK=1 KD_BENCH_BACKEND=multimem KD_BENCH_MM_CTAS=16
run "multimem ctas8" KD_BENCH_QUICK=1 KD_BENCH_BACKEND=multimem KD_BENCH_MM_CTAS=8
run "multimem ctas32" KD_BENCH_QUICK=1 KD_BENCH_BACKEND=multimem KD_BENCH_MM_CTAS=32
