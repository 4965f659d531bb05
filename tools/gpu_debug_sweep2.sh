#!/bin/bash
# 1-GPU box, experiment build (build/variants/lib_x_exp.so = -DCODON_TC_EXPERIMENT): which part of the cluster conv kernels
# costs what -- CODON_TC_DEBUG bits: 1 no epilogue stores, 32 no epilogue TMEM loads / Y staging, 2 no weight loads, 4 no
# patch loads (results are garbage for any non-zero value; only the per-class times mean something).
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
cp build/variants/lib_x_exp.so codon_b200/libcodon_b200.so
for MODE in ${MODES:-bf16 tf32}; do for FR in ${FRAMES:-1 8}; do for dbg in ${DBG_LIST:-0 1 33 6 39}; do
  echo "== $MODE frames $FR dbg $dbg: $(CODON_TC_DEBUG=$dbg timeout 200 python tools/gpu_class_time.py $MODE $FR 6 2>/dev/null | grep "conv5x5\|pair\|conv3x3" | awk '{printf "%s %.4f ms  ", $4, $5}')"
done; done; done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
