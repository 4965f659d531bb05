#!/bin/bash
# Round-end run on one B200 (under gpurun): GPU tests, smoke, the two bench workloads, the reference arm, then (after
# those exited) the ncu launch lists and the --set full capture of the CAC kernels (the conv / edge kernels are
# unchanged since r01h: profiles/r01h_conv_tc2_*.txt, r01h_edge_*.txt).
set -u
TAG=${1:-r01i}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_${TAG}.log
python __graft_entry__.py --smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"
python bench.py > $OUT/bench_${TAG}_tf32.json 2> $OUT/bench_${TAG}_tf32.err; echo "bench rc=$?"
python bench.py --mode bf16 --frames 8 --scale 8 --no-variants --no-cpu-baseline > $OUT/bench_${TAG}_bf16_x8_b8.json 2> $OUT/bench_${TAG}_bf16.err; echo "bench bf16 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>/dev/null; echo "ref rc=$?"
for MODE in bf16 tf32; do
  CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-variants --no-cpu-baseline"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 240 --csv --log-file $OUT/launches_${TAG}_${MODE}.csv $CMD > /dev/null 2>&1
  python tools/summarize_ncu.py --launches $OUT/launches_${TAG}_${MODE}.csv $OUT/${TAG}_launches_${MODE}.txt; rm -f $OUT/launches_${TAG}_${MODE}.csv
  ncu --set full --clock-control none --import-source on -k regex:cac_ -s 3 -c 3 -f -o $OUT/prof_cac_${TAG}_${MODE} $CMD > $OUT/ncu_cac_${TAG}_${MODE}.log 2>&1
  python tools/summarize_ncu.py $OUT/prof_cac_${TAG}_${MODE}.ncu-rep $OUT/${TAG}_cac_${MODE}.txt; rm -f $OUT/prof_cac_${TAG}_${MODE}.ncu-rep
done
python - $TAG <<'PYEOF'
import json, sys
tag = sys.argv[1]
for f in [f"bench_{tag}_tf32.json", f"bench_{tag}_bf16_x8_b8.json"]:
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    print(f, round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "blocking", round(d["e2e"]["blocking_call"]["value"], 2),
          "5x5 frac", round(d["roofline"]["frac"], 3), "apply frac", round(d["roofline"]["cac_kernels"]["frac"], 3), d["clocks"],
          d.get("parity"), d.get("cpu_baseline", {}).get("value"), d.get("variants"))
PYEOF
cat $OUT/bench_${TAG}_reference.json | cut -c1-300
