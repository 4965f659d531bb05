#!/bin/bash
# Needs an experiment build of the library (tools/build_variant.sh x -DCODON_TC_EXPERIMENT, copied over
# codon_b200/libcodon_b200.so): the knobs are compiled out of the product build.
# Perf experiment (GPU box): CODON_TC_DEBUG knobs of the cluster conv kernel; prints per-kernel ms per step.
# Results are numerically garbage for any non-zero knob; only the timing is meaningful.
MODE=${1:-bf16}; FR=${2:-8}
for dbg in ${DBG_LIST:-0 1 32 2 4 6 38}; do
  CODON_TC_DEBUG=$dbg python bench.py --mode $MODE --frames $FR --scale 8 --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['by_kernel_ms_per_step']
print('dbg=$dbg', 'ms/step %.2f'%d['ms_per_step'], 'conv5 %.3f pair %.3f conv3 %.3f'%(k['conv5x5_128to128'],k['pair_3x3_5x5_64to128'],k['conv3x3']), d['clocks']['sm_mhz'])"
done
