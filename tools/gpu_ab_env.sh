#!/bin/bash
# Same-session A/B of an environment switch (GPU box): tools/gpu_ab_env.sh VAR [reps]; alternates VAR=1 / VAR=0.
VAR=$1; REPS=${2:-3}
for rep in $(seq $REPS); do for c in 1 0; do
env $VAR=$c python bench.py --no-cpu-baseline --no-variants --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tf32 b1 $VAR=$c', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
env $VAR=$c python bench.py --mode bf16 --frames 1 --scale 8 --no-cpu-baseline --no-variants --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 b1 $VAR=$c', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
env $VAR=$c python bench.py --mode bf16 --frames 8 --scale 8 --no-cpu-baseline --no-variants --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 b8 $VAR=$c', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
done; done
