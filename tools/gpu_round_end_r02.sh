#!/bin/bash
# Round-2 end-of-round run on one B200 (under gpurun): GPU tests, smoke, the bench workloads, the reference arm, the
# 30-image parity table of every mode, a stress run.  ncu evidence: tools/gpu_profile_r02.sh.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_${TAG}.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $OUT/bench_${TAG}_tf32.json 2> $OUT/bench_${TAG}_tf32.err; echo "bench rc=$?"
timeout 600 python bench.py --mode f16x3 --no-variants > $OUT/bench_${TAG}_f16x3.json 2> $OUT/bench_${TAG}_f16x3.err; echo "bench f16x3 rc=$?"
timeout 600 python bench.py --mode bf16 --frames 8 --scale 8 --no-variants --no-cpu-baseline > $OUT/bench_${TAG}_bf16_x8_b8.json 2> $OUT/bench_${TAG}_bf16.err; echo "bench bf16 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 900 python tests/checkers/image_parity_table.py > $OUT/${TAG}_image_parity.txt 2>&1; grep "^# " $OUT/${TAG}_image_parity.txt | tail -6
timeout 200 python tools/gpu_stress.py 90 > $OUT/${TAG}_stress.txt 2>&1; cat $OUT/${TAG}_stress.txt | tail -2
python - $TAG <<'PYEOF'
import json, sys
tag = sys.argv[1]
for f in [f"bench_{tag}_tf32.json", f"bench_{tag}_f16x3.json", f"bench_{tag}_bf16_x8_b8.json"]:
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f, round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "stream", round(d["e2e"]["streaming"]["value"], 2),
          "5x5 TF/s", round(r["achieved"], 1), "frac", round(r["frac"], 3), "mma frac of dtype peak", round(r["frac_of_dtype_peak"] or 0, 3),
          "apply frac", round(r["cac_kernels"]["frac"], 3), d["clocks"], d["config"].get("parity_max_abs_vs_cpu_reference"),
          d.get("cpu_baseline", {}).get("value"), {k: round(v.get("value", 0), 2) for k, v in r.get("variants", {}).items()})
PYEOF
cat $OUT/bench_${TAG}_reference.json | cut -c1-200
