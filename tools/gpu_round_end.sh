set -u
TAG=${1:-r01h}
OUT=gpurun_out
python -m pytest tests -m gpu -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_${TAG}.log
python __graft_entry__.py --smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"
python bench.py > $OUT/bench_${TAG}_tf32.json 2> $OUT/bench_${TAG}_tf32.err; echo "bench rc=$?"
python bench.py --mode bf16 --frames 8 --scale 8 --no-variants > $OUT/bench_${TAG}_bf16_x8_b8.json 2> $OUT/bench_${TAG}_bf16.err; echo "bench bf16 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>/dev/null; echo "ref rc=$?"
python tools/run_configs.py > $OUT/configs_${TAG}_1gpu.jsonl 2> $OUT/configs_${TAG}_1gpu.err; echo "configs rc=$?"
for MODE in bf16 tf32; do
  CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-variants --no-cpu-baseline"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_${TAG}_${MODE}.csv $CMD > /dev/null 2>&1
  python tools/summarize_ncu.py --launches $OUT/launches_${TAG}_${MODE}.csv $OUT/${TAG}_launches_${MODE}.txt; rm -f $OUT/launches_${TAG}_${MODE}.csv
  ncu --set full --clock-control none --import-source on -k regex:conv_first\|conv_last -s 2 -c 2 -f -o $OUT/prof_edge_${TAG}_${MODE} $CMD > /dev/null 2>&1
  python tools/summarize_ncu.py $OUT/prof_edge_${TAG}_${MODE}.ncu-rep $OUT/${TAG}_edge_${MODE}.txt; rm -f $OUT/prof_edge_${TAG}_${MODE}.ncu-rep
done
python - $TAG <<'EOF'
import json, sys
tag = sys.argv[1]
for f in [f"bench_{tag}_tf32.json", f"bench_{tag}_bf16_x8_b8.json"]:
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    print(f, round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["roofline"]["by_kernel_ms_per_step"], d["clocks"],
          d.get("parity"), d.get("cpu_baseline", {}).get("value"))
EOF
cat $OUT/configs_${TAG}_1gpu.jsonl | cut -c1-200
