#!/bin/bash
# Same-box A/B of library variants (python speech-distill_b200/build.py --variant NAME -D...): each variant runs
# tools/k1_ab.py (settled 2-s loop at configs[1] size) ROUNDS times, interleaved, so box-to-box and drift cancel.
#   bash tools/k1_abab.sh [ROUNDS] default nofwdmax nogccur ...
cd "$(dirname "$0")/.."
rounds=${1:-2}; shift
for r in $(seq 1 "$rounds"); do
  for v in "$@"; do
    lib="$PWD/speech-distill_b200/libkd_b200_$v.so"; [ "$v" = default ] && lib="$PWD/speech-distill_b200/libkd_b200.so"
    out=$(KD_B200_LIB="$lib" AB_PARITY=0 AB_SECONDS=${AB_SECONDS:-2} python tools/k1_ab.py 2>&1 | tail -1)
    echo "$v round $r: $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("settled %.3f ms  fwd %.3f  bwd %.3f" % (d["ms_per_step_settled"], d["fwd_ms"], d["bwd_ms"]))' 2>/dev/null || echo "$out" | tail -c 300)"
  done
done
