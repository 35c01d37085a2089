#!/bin/bash
# Does NCCL's CTA placement break the GEMMs' CTA pairs (cluster of 2 = both SMs of a TPC)?  NCCL_CGA_CLUSTER_SIZE groups
# NCCL's CTAs on neighbouring SMs; KD_UMMA_CTA_GROUP=1 removes the pairing constraint on our side.
N=${1:-2}
port=29750
run() {
  name=$1; shift
  port=$((port+1))
  env KD_BENCH_QUICK=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 60 --warmup 10 2>/dev/null | tail -1 | cut -c1-100 | sed "s/^/$name /"
}
run "default" KD_BENCH_NCCL_CTAS=32
run "cga2" NCCL_CGA_CLUSTER_SIZE=2 KD_BENCH_NCCL_CTAS=32
run "cga4" NCCL_CGA_CLUSTER_SIZE=4 KD_BENCH_NCCL_CTAS=32
run "cga0" NCCL_CGA_CLUSTER_SIZE=0 KD_BENCH_NCCL_CTAS=32
run "cg1 nocomm" KD_UMMA_CTA_GROUP=1 KD_BENCH_SYNC=none_
run "cg1 overlap" KD_UMMA_CTA_GROUP=1 KD_BENCH_NCCL_CTAS=32
