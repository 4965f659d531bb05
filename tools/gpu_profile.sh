#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list, ncu --set full of the top kernels.
# usage: tools/gpu_profile.sh <tag> [mode]
set -u
TAG=${1:-r01}
MODE=${2:-tf32}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-variants --no-cpu-baseline"
$CMD > $OUT/plain_${TAG}_${MODE}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_${TAG}_${MODE}.csv $CMD > $OUT/ncu_launches_${TAG}_${MODE}.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/plain2_${TAG}_${MODE}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 10 -c 3 -f -o $OUT/prof_conv_${TAG}_${MODE} $CMD > $OUT/ncu_conv_${TAG}_${MODE}.log 2>&1
echo "conv capture rc=$?"
$CMD > $OUT/plain3_${TAG}_${MODE}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cac_ -s 3 -c 3 -f -o $OUT/prof_cac_${TAG}_${MODE} $CMD > $OUT/ncu_cac_${TAG}_${MODE}.log 2>&1
echo "cac capture rc=$?"
ls -la $OUT | tail -20
