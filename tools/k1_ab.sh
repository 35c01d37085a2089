#!/bin/bash
# K1 step A/B matrix (one process per knob set); run on the GPU box: bash tools/k1_ab.sh > gpurun_out/k1_ab.log
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" python tools/k1_ab.py 2>&1 | tail -2; }
run AB_PARITY=1
run AB_PARITY=0 KD_GRAD_OVERLAP=0
run AB_PARITY=1 KD_LOGIT_CACHE_MB=0
run AB_PARITY=0 KD_LOGIT_CACHE_MB=0 KD_DW_ORDER=m
