#!/bin/bash
# interleaved same-box A/B of two library builds for K2: tools/k2_ab.sh libA.so libB.so [rounds]
A=$1; B=$2; N=${3:-2}
for i in $(seq $N); do
  for lib in "$A" "$B"; do KD_B200_LIB="$lib" python tools/k2_ab.py 2>&1 | tail -1; done
done
