#!/bin/bash
# Round 2 profiling pass on one B200 (under gpurun): for each mode a plain bench run (must exit 0), then the ncu launch
# list of the same command (this library's kernels only) and ncu --set full captures of the convolution and CAC kernels;
# every report is summarised into text on the box (gpurun_out is limited to 64 MiB; only one .ncu-rep is kept).
# usage: tools/gpu_profile_r02.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for MODE in tf32 f16x3 bf16; do
  CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-variants --no-cpu-baseline"
  $CMD > $OUT/plain_${TAG}_${MODE}.log 2>&1 || { echo "plain run failed ($MODE)"; tail -5 $OUT/plain_${TAG}_${MODE}.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|cac_" -c 360 --csv --log-file $OUT/launches_${TAG}_${MODE}.csv $CMD > /dev/null 2>&1
  python tools/summarize_ncu.py --launches $OUT/launches_${TAG}_${MODE}.csv $OUT/${TAG}_launches_${MODE}.txt; rm -f $OUT/launches_${TAG}_${MODE}.csv
  for K in conv_tc2:conv_tc2_kernel:38:4 cac:cac_:18:3; do
    IFS=: read NAME RE SKIP CNT <<< "$K"
    ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o $OUT/prof_${NAME}_${TAG}_${MODE} $CMD > $OUT/ncu_${NAME}_${TAG}_${MODE}.log 2>&1
    python tools/summarize_ncu.py $OUT/prof_${NAME}_${TAG}_${MODE}.ncu-rep $OUT/${TAG}_${NAME}_${MODE}.txt
    [ "$NAME:$MODE" = "conv_tc2:f16x3" ] || rm -f $OUT/prof_${NAME}_${TAG}_${MODE}.ncu-rep
  done
done
du -sh $OUT; ls $OUT | grep $TAG
