#!/bin/bash
# issuer cycle account of the pair kernels (experiment build, CODON_TC_DEBUG=64), then class times with the product build
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
cp build/variants/lib_x_exp.so codon_b200/libcodon_b200.so
for m in bf16 tf32 f16x3; do
  echo "=== $m"
  CODON_TC_DEBUG=64 timeout 300 python tools/gpu_quick_time.py $m 1 1 2>&1 | grep -a "issuer" | grep -a "cluster 24 " | grep -a ",0,2>" | cut -c1-200 | sort | uniq -c | sort -rn | head -3
done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
timeout 200 python tools/gpu_class_time.py f16x3 1 10 | grep "conv5x5\|pair\|total"
timeout 300 python -m pytest tests -m gpu -x -q -k "forward_matches_reference and f16x3" 2>&1 | tail -2
