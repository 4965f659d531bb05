#!/bin/bash
# GPU box: does the tile geometry (NACC = accumulators per tile) explain the split kernel's MMA rate?
for m in fp16 bf16; do
  echo "== $m default"; timeout 300 python tools/gpu_class_time.py $m 1 10 | grep "conv5x5\|pair\|total"
  echo "== $m NACC=1"; CODON_TC_NACC_CONV=1 CODON_TC_NACC_PAIR=1 timeout 300 python tools/gpu_class_time.py $m 1 10 | grep "conv5x5\|pair\|total"
done
