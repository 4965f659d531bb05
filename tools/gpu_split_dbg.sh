#!/bin/bash
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
cp build/variants/lib_x_tpc5i.so codon_b200/libcodon_b200.so
for d in 64 96 97 102 103; do
  echo "=== CODON_TC_DEBUG=$d"
  CODON_TC_DEBUG=$d timeout 300 python tools/gpu_quick_time.py f16x3 1 3 2>&1 | grep -a "issuer\|MP/s" | grep -a "cluster 24 \|MP/s" | cut -c1-200 | sort | uniq -c | sort -rn | head -4
done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
