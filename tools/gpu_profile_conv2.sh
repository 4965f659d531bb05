#!/bin/bash
# ncu --set full of the 2-CTA conv kernel launches (pair + 5x5-128) for one mode
TAG=${1:-r01b}; MODE=${2:-bf16}; OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-variants --no-cpu-baseline"
$CMD > $OUT/plain4_${TAG}_${MODE}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 6 -c 3 -f -o $OUT/prof_conv2_${TAG}_${MODE} $CMD > $OUT/ncu_conv2_${TAG}_${MODE}.log 2>&1
echo "conv2 capture rc=$?"
