#!/bin/bash
# 1-GPU box: BASELINE configs 2-5 at the final round-2 kernels, and the CAC kernels of config 5's shape (4 x 1280x720, bf16)
# under ncu --set full (HBM roofline of the attention kernels at 720p).
OUT=gpurun_out
timeout 300 python tools/run_configs.py > $OUT/configs_r02c_1gpu.jsonl 2> $OUT/configs_r02c.err; echo "configs rc=$?"; cut -c1-200 $OUT/configs_r02c_1gpu.jsonl
CMD="python bench.py --steps 1 --warmup 3 --mode bf16 --frames 4 --height 720 --width 1280 --scale 16 --no-variants --no-cpu-baseline"
timeout 200 $CMD > $OUT/plain_r02c_720p.log 2>&1 || { echo "plain 720p run failed"; tail -3 $OUT/plain_r02c_720p.log; exit 0; }
python -c "
import json; d=json.loads(open('$OUT/plain_r02c_720p.log').read().strip().splitlines()[-1]); r=d['roofline']; print('720p x16 bf16 4 frames:', round(d['value'],1), 'MP/s, apply frac', round(r['cac_kernels']['frac'],3), r['by_kernel_ms_per_step'])"
timeout 300 ncu --set full --clock-control none -k regex:cac_ -s 18 -c 3 -f -o $OUT/prof_cac_r02c_720p $CMD > $OUT/ncu_cac_r02c_720p.log 2>&1
python tools/summarize_ncu.py $OUT/prof_cac_r02c_720p.ncu-rep $OUT/r02c_cac_bf16_720p.txt; rm -f $OUT/prof_cac_r02c_720p.ncu-rep
grep "^== \|time_duration\|dram_throughput\|traffic = " $OUT/r02c_cac_bf16_720p.txt | cut -c1-150
