#!/bin/bash
# Perf experiment (GPU box): time the bench per-kernel with each library variant (build/variants/lib_<name>.so).
# usage: tools/gpu_variant_sweep.sh "<variants>" "<debug values>" [mode] [frames]
VARS=${1:-base}; DBGS=${2:-0}; MODE=${3:-bf16}; FR=${4:-8}
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
for v in $VARS; do
  cp build/variants/lib_$v.so codon_b200/libcodon_b200.so
  for dbg in $DBGS; do
    CODON_TC_DEBUG=$dbg python bench.py --mode $MODE --frames $FR --scale 8 --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['by_kernel_ms_per_step']
print('$v dbg=$dbg', 'ms/step %.2f'%d['ms_per_step'], 'conv5 %.3f pair %.3f conv3 %.3f'%(k['conv5x5_128to128'],k['pair_3x3_5x5_64to128'],k['conv3x3']), d['clocks']['sm_mhz'])"
  done
done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
