#!/bin/bash
# 1-GPU box: same-session A/B of the CAC kernels (product library vs build/variants/lib_<v>.so): parity tests first, then
# per-class CUDA-event times at 1 and 8 frames.
VARS=${1:-oldcac}; MODES=${2:-"bf16 tf32 f16x3"}
timeout 600 python -m pytest tests -m gpu -x -q -k "forward_matches_reference or ragged or cluster_kernels or random_shapes or batch_invariance or narrow or cac" 2>&1 | tail -3
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
for rep in 1 2; do for v in base $VARS; do
  if [ $v = base ]; then cp /tmp/lib_orig.so codon_b200/libcodon_b200.so; else cp build/variants/lib_$v.so codon_b200/libcodon_b200.so; fi
  for m in $MODES; do
    echo "== $v $m"; timeout 200 python tools/gpu_class_time.py $m 1 10 | grep "cac_\|total"
  done
  echo "== $v bf16 8 frames"; timeout 200 python tools/gpu_class_time.py bf16 8 5 | grep "cac_\|total"
done; done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
