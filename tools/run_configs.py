#!/usr/bin/env python
"""Runs the BASELINE.json configs 2-5 on the local GPU(s) of this rank and prints one JSON line each.

  python tools/run_configs.py [--quick]            (1 GPU)
  torchrun --nproc-per-node N tools/run_configs.py (N GPUs, frames sharded, no data-path collective)

config 2: x4, 1 x 640x480, fp32 parity mode (tf32)       config 3: x8, 64 x 640x480 bf16 over the ranks
config 4: x16, 1920x1080 frames, bf16 (in-kernel halo tiling)   config 5: 256 x 1280x720 cycling x4/x8/x16
Throughput = frames*H*W / device time (CUDA events, max over ranks); inputs resident in HBM.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codon_b200 import engine as E, synthetic as syn  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def sync_max(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    engines = {}

    def eng(scale, mode):
        if (scale, mode) not in engines:
            e = E.Engine(scale, mode, local)
            e.load_state_dict(syn.synthetic_state_dict(scale, 0))
            engines[(scale, mode)] = e
        return engines[(scale, mode)]

    def timed(jobs, reps):
        """jobs: list of (engine, x, y, out); one pass = all jobs once."""
        for e, x, y, o in jobs[:3]:
            e.forward(x, y, o)
        torch.cuda.synchronize()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            for e, x, y, o in jobs:
                e.forward(x, y, o)
        b.record()
        torch.cuda.synchronize()
        barrier()
        return sync_max(a.elapsed_time(b))

    def frames(n, h, w, seed):
        x, y = syn.synthetic_frames(1, h, w, seed)
        return x.to(dev).expand(n, 1, h, w).contiguous(), y.to(dev).expand(n, 1, h, w).contiguous()

    def report(name, total_frames, h, w, ms, extra):
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "frames": total_frames, "height": h, "width": w,
                              "ms": ms, "MP_per_s": total_frames * h * w / 1e6 / (ms / 1e3), **extra}), flush=True)

    # config 2
    x, y = frames(1, 480, 640, 1234 + rank)
    reps = 5 if quick else 30
    ms = timed([(eng(4, "tf32"), x, y, torch.empty_like(x))], reps)
    report("2: x4 1x640x480 fp32-parity(tf32) per GPU", reps * world, 480, 640, ms, {"mode": "tf32"})
    # config 3: 64 frames over the ranks (8 per GPU on 8 GPUs); on fewer GPUs each rank takes 64/world in chunks of 8
    per_rank = 64 // world
    x, y = frames(8, 480, 640, 99 + rank)
    jobs = [(eng(8, "bf16"), x, y, torch.empty_like(x))] * max(1, per_rank // 8)
    reps = 1 if quick else 3
    ms = timed(jobs, reps)
    report("3: x8 64x640x480 bf16 sharded", reps * len(jobs) * 8 * world, 480, 640, ms, {"mode": "bf16", "frames_per_call": 8})
    # config 4: 1080p, x16
    x, y = frames(2, 1080, 1920, 7 + rank)
    reps = 1 if quick else 4
    ms = timed([(eng(16, "bf16"), x, y, torch.empty_like(x))], reps)
    report("4: x16 1920x1080 bf16, 2 frames per call per GPU", reps * 2 * world, 1080, 1920, ms,
           {"mode": "bf16", "tiling": "in-kernel 16x16-px tiles, (k-1)-px TMA halo, fixed-order CAC reduction",
            "workspace_GB": eng(16, "bf16").workspace_bytes(2, 1080, 1920) / 1e9})
    # config 5: 256 frames 1280x720 cycling x4/x8/x16, sharded over ranks, 4 frames per call
    total = 24 if quick else 256
    per_rank = total // world
    x, y = frames(4, 720, 1280, 5 + rank)
    o = torch.empty_like(x)
    jobs = [(eng((4, 8, 16)[k % 3], "bf16"), x, y, o) for k in range(max(1, per_rank // 4))]
    ms = timed(jobs, 1)
    report("5: mixed x4/x8/x16 sweep, 1280x720 bf16, 4 frames per call", len(jobs) * 4 * world, 720, 1280, ms, {"mode": "bf16"})
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
