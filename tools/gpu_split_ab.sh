#!/bin/bash
# GPU box: A/B of f16x3 library variants (build/variants/lib_<name>.so): per-class time, accuracy vs the oracle,
# 30-image parity summary, issuer cycle account (CODON_TC_DEBUG=64 needs -DCODON_TC_EXPERIMENT builds).
VARS=${1:-"x_tpc1 x_tpc5"}
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
for v in $VARS; do
  cp build/variants/lib_$v.so codon_b200/libcodon_b200.so
  echo "=== $v"
  timeout 300 python tools/gpu_class_time.py f16x3 1 5
  timeout 300 python tools/gpu_quick_time.py f16x3 1 20
  timeout 300 python tests/checkers/debug_taps.py f16x3 1 160 240
  timeout 900 python tests/checkers/image_parity_table.py f16x3 | grep "^#"
  CODON_TC_DEBUG=64 timeout 300 python tools/gpu_quick_time.py f16x3 1 1 2>&1 | grep -a "issuer" | grep -a "cluster 0 \|cluster 24 " | cut -c1-200 | sort | uniq -c | sort -rn | head -8
done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
