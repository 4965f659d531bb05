#!/bin/bash
# GPU box: A/B of the f16x3 accumulation-chunk size (library variants x_tpc1 / x_tpc5): per-class time, accuracy,
# issuer cycle account (CODON_TC_DEBUG=64 needs -DCODON_TC_EXPERIMENT builds).
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
for v in x_tpc1 x_tpc5; do
  cp build/variants/lib_$v.so codon_b200/libcodon_b200.so
  echo "=== $v"
  timeout 300 python tools/gpu_class_time.py f16x3 1 5
  timeout 300 python tests/checkers/debug_taps.py f16x3 1 160 240
  CODON_TC_DEBUG=64 timeout 300 python tools/gpu_quick_time.py f16x3 1 1 2>&1 | grep "issuer" | sort | uniq -c | sort -rn | head -12
done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
