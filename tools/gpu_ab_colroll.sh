#!/bin/bash
# 1-GPU box: same-session A/B of the column-rolled tap loops (product library) against the fully unrolled ones
# (build/variants/lib_old.so = -DCODON_TC_COLROLL=0), per-class CUDA-event times, then the parity tests.
VARS=${1:-old}; MODES=${2:-"bf16 tf32 f16x3"}
timeout 600 python -m pytest tests -m gpu -x -q -k "forward_matches_reference or ragged or cluster_kernels or random_shapes or batch_invariance" 2>&1 | tail -3
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
for rep in 1 2; do for v in base $VARS; do
  if [ $v = base ]; then cp /tmp/lib_orig.so codon_b200/libcodon_b200.so; else cp build/variants/lib_$v.so codon_b200/libcodon_b200.so; fi
  for m in $MODES; do
    echo "== $v $m"; timeout 200 python tools/gpu_class_time.py $m 1 10 | grep "conv5x5\|pair\|conv3x3\|total"
  done
done; done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
