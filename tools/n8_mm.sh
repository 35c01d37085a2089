#!/bin/bash
# NCCL vs the NVLS multimem backend of GradSync at 8 GPUs (KD_BENCH_QUICK: step time only)
N=${1:-8}
port=29550
run() {
  name=$1; shift
  port=$((port+1))
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 60 --warmup 10 2>gpurun_out/mm8_err_$port.log | tail -1 | cut -c1-110 | sed "s/^/$name /"
}
run "nccl" KD_BENCH_QUICK=1
run "multimem ctas16" KD_BENCH_QUICK=1 KD_BENCH_BACKEND=multimem KD_BENCH_MM_CTAS=16
run "multimem ctas8" KD_BENCH_QUICK=1 KD_BENCH_BACKEND=multimem KD_BENCH_MM_CTAS=8
run "multimem ctas16 ranges9" KD_BENCH_QUICK=1 KD_BENCH_BACKEND=multimem KD_BENCH_MM_CTAS=16 KD_BENCH_RANGES=9
