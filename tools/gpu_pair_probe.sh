#!/bin/bash
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
cp build/variants/lib_x_exp.so codon_b200/libcodon_b200.so
for m in bf16 tf32; do
  echo "=== $m"
  CODON_TC_DEBUG=64 timeout 300 python tools/gpu_quick_time.py $m 1 1 2>&1 | grep -a "issuer" | grep -a "cluster 24 " | cut -c1-200 | sort | uniq -c | sort -rn | head -6
done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
for m in bf16 tf32 f16x3; do timeout 200 python tools/gpu_class_time.py $m 1 10 | grep "conv5x5\|pair\|total"; done
