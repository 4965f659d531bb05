#!/bin/bash
# Round-end multi-GPU run (gpurun --gpus 8): the multi-GPU tests, the weak-scaling bench, all BASELINE configs and the
# single-frame row-band mode.
OUT=gpurun_out; TAG=${1:-r01h}
python -m pytest tests -m gpu -q -k "gpus or two_gpus or over_gpus" > $OUT/pytest_${TAG}_multigpu.log 2>&1; echo "pytest multi rc=$?"; tail -2 $OUT/pytest_${TAG}_multigpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --no-cpu-baseline --no-variants > $OUT/bench_${TAG}_8gpu_tf32.json 2> $OUT/bench_${TAG}_8gpu_tf32.err; echo "bench8 tf32 rc=$?"
$TR bench.py --gpus 8 --mode bf16 --frames 8 --scale 8 --no-cpu-baseline --no-variants > $OUT/bench_${TAG}_8gpu_bf16_x8_b8.json 2> $OUT/bench_${TAG}_8gpu_bf16.err; echo "bench8 bf16 rc=$?"
$TR tools/run_configs.py > $OUT/configs_${TAG}_8gpu.jsonl 2> $OUT/configs_${TAG}_8gpu.err; echo "configs8 rc=$?"
python tools/run_group.py > $OUT/group_${TAG}_8gpu.jsonl 2> $OUT/group_${TAG}_8gpu.err; echo "group rc=$?"
python - <<EOF
import json
for f in ["bench_${TAG}_8gpu_tf32.json","bench_${TAG}_8gpu_bf16_x8_b8.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["n_gpus"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
EOF
cut -c1-170 $OUT/configs_${TAG}_8gpu.jsonl; cut -c1-200 $OUT/group_${TAG}_8gpu.jsonl
