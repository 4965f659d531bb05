#!/bin/bash
# 1-GPU box: same-session A/B of library variants on per-class times: tools/gpu_ab_simple.sh "<variants>" "<modes>" [frames]
VARS=${1:-base}; MODES=${2:-"bf16 tf32"}; FR=${3:-1}
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
for rep in 1 2; do for v in base $VARS; do
  if [ $v = base ]; then cp /tmp/lib_orig.so codon_b200/libcodon_b200.so; else cp build/variants/lib_$v.so codon_b200/libcodon_b200.so; fi
  for m in $MODES; do
    echo "== $v $m"; timeout 200 python tools/gpu_class_time.py $m $FR 10 | grep "conv5x5\|pair\|conv3x3\|total"
  done
done; done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
timeout 300 python -m pytest tests -m gpu -x -q -k "forward_matches_reference" 2>&1 | tail -2
