#!/bin/bash
# Runs on the GPU box (under gpurun): for each mode a plain bench run, the ncu launch list of the same command and
# ncu --set full captures of the conv / CAC / edge kernels; summarises every report into text on the box and keeps
# only the cluster-conv report (gpurun_out is limited to 64 MiB).
# usage: tools/gpu_profile_all.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
for MODE in bf16 tf32; do
  CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-variants --no-cpu-baseline"
  $CMD > $OUT/plain_${TAG}_${MODE}.log 2>&1 || { echo "plain run failed ($MODE)"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_${TAG}_${MODE}.csv $CMD > /dev/null 2>&1
  python tools/summarize_ncu.py --launches $OUT/launches_${TAG}_${MODE}.csv $OUT/${TAG}_launches_${MODE}.txt
  for K in conv_tc2:conv_tc2_kernel:6:3 cac:cac_:3:3 edge:conv_first\|conv_last:2:2; do
    IFS=: read NAME RE SKIP CNT <<< "$K"
    ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o $OUT/prof_${NAME}_${TAG}_${MODE} $CMD > $OUT/ncu_${NAME}_${TAG}_${MODE}.log 2>&1
    python tools/summarize_ncu.py $OUT/prof_${NAME}_${TAG}_${MODE}.ncu-rep $OUT/${TAG}_${NAME}_${MODE}.txt
    [ "$NAME:$MODE" = "conv_tc2:bf16" ] || rm -f $OUT/prof_${NAME}_${TAG}_${MODE}.ncu-rep
  done
  rm -f $OUT/launches_${TAG}_${MODE}.csv
done
du -sh $OUT; ls $OUT
