#!/bin/bash
# 1-GPU box: same-session A/B of an f16x3 kernel change (product library vs build/variants/lib_<v>.so): the f16x3 parity tests
# first (goldens, BASELINE shapes, 30-image three-decimal clause), then per-class CUDA-event times.
VARS=${1:-areuse0}
timeout 900 python -m pytest tests -m gpu -x -q -k "f16x3 or image_parity or cluster_kernels or random_shapes or batch_invariance" 2>&1 | tail -3
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
for rep in 1 2; do for v in base $VARS; do
  if [ $v = base ]; then cp /tmp/lib_orig.so codon_b200/libcodon_b200.so; else cp build/variants/lib_$v.so codon_b200/libcodon_b200.so; fi
  echo "== $v f16x3"; timeout 200 python tools/gpu_class_time.py f16x3 1 10 | grep "conv5x5\|pair\|conv3x3\|total"
done; done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
timeout 200 python tools/gpu_quick_time.py f16x3 1 20
