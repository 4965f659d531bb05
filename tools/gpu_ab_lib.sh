#!/bin/bash
# Same-session A/B of whole-library variants (GPU box): tools/gpu_ab_lib.sh "<variants>" [reps] ["<mode frames scale steps>;..."]
# (build/variants/lib_<name>.so); prints step time and the CAC apply kernel's time per workload.
VARS=${1:-base}; REPS=${2:-2}; CFGS=${3:-"tf32 1 4 30;bf16 1 8 30;bf16 8 8 15"}
cp codon_b200/libcodon_b200.so /tmp/lib_orig.so
IFS=';' read -ra CFG <<< "$CFGS"
for rep in $(seq $REPS); do for v in $VARS; do
  cp build/variants/lib_$v.so codon_b200/libcodon_b200.so
  for cfg in "${CFG[@]}"; do
    set -- $cfg
    python bench.py --mode $1 --frames $2 --scale $3 --steps $4 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['by_kernel_ms_per_step']; c=d['roofline']['cac_kernels']
print('$v $1 b$2', 'MP/s %.2f'%d['value'], 'e2e %.2f'%d['e2e']['value'], 'ms %.4f instrumented %.4f'%(d['ms_per_step'], d['roofline']['ms_per_step_instrumented']), 'apply ms %.4f frac %.3f'%(k['cac_apply'], c['frac']), d['clocks']['sm_mhz'])"
  done
done; done
cp /tmp/lib_orig.so codon_b200/libcodon_b200.so
