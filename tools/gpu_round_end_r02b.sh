#!/bin/bash
# Round-2 HEAD validation on one B200 (under gpurun): the full GPU test suite, smoke, the default bench line, the
# reference arm as the driver launches it, BASELINE configs 2-5 on one GPU.
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_${TAG}.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_${TAG}.log
timeout 600 python bench.py > $OUT/bench_${TAG}_tf32.json 2> $OUT/bench_${TAG}_tf32.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python tools/run_configs.py > $OUT/configs_${TAG}_1gpu.jsonl 2> $OUT/configs_${TAG}.err; echo "configs rc=$?"
cut -c1-260 $OUT/configs_${TAG}_1gpu.jsonl
python - $TAG <<'PYEOF'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_{tag}_tf32.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "5x5 TF/s", round(r["achieved"], 1), "frac", round(r["frac"], 3),
      "apply frac", round(r["cac_kernels"]["frac"], 3), d["clocks"], d["config"].get("parity_max_abs_vs_cpu_reference"),
      d.get("cpu_baseline", {}).get("value"), {k: round(v.get("value", 0), 2) for k, v in r.get("variants", {}).items()})
print({k: round(v, 3) for k, v in r["by_kernel_ms_per_step"].items()})
PYEOF
cut -c1-300 $OUT/bench_${TAG}_reference.json
