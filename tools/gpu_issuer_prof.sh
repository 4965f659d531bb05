#!/bin/bash
# Needs an experiment build of the library (tools/build_variant.sh x -DCODON_TC_EXPERIMENT).
# Perf experiment (GPU box): cycle accounting of the cluster conv kernel's MMA issuer / weight producer warps
# (CODON_TC_DEBUG bit 64, printed by cluster 0 of every launch); one forward per mode is enough.
MODE=${1:-bf16}; FR=${2:-8}; DBG=${3:-64}
CODON_TC_DEBUG=$DBG python bench.py --mode $MODE --frames $FR --scale 8 --steps 1 --warmup 3 --no-cpu-baseline --no-variants 2>&1 | grep "conv_tc2<" | sort | uniq -c | sort -rn | head -${4:-40}
