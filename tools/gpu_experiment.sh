#!/bin/bash
# perf experiments: per-kernel-class ms for a set of CODON_TC_DEBUG flag values
for f in "$@"; do
  echo "== CODON_TC_DEBUG=$f"
  CODON_TC_DEBUG=$f python bench.py --mode bf16 --frames 8 --steps 5 --warmup 3 --no-variants --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print({k:round(v,3) for k,v in r['by_kernel_ms_per_step'].items()}, 'conv5 TF/s', round(r['achieved'],1), 'total ms', round(d['ms_per_step'],2))
"
done
