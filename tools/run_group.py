#!/usr/bin/env python
"""Single-frame latency mode: one frame split into row bands over 1/2/4/8 GPUs of this process
(codon_group_*: NVLink peer-memory halo exchange after every layer, no NCCL).  Prints one JSON line per
(shape, mode, GPU count): wall-clock ms per frame through the host entry point (H2D + forward + D2H)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codon_b200 import engine as E, synthetic as syn  # noqa: E402


def main():
    ngpu = torch.cuda.device_count()
    shapes = [(480, 640, 4), (1080, 1920, 16)]
    reps = 10
    for (h, w, scale) in shapes:
        sd = syn.synthetic_state_dict(scale, 0)
        x, y = syn.synthetic_frames(1, h, w, 1234)
        xn, yn = x.numpy()[0, 0], y.numpy()[0, 0]
        for mode in ("tf32", "bf16"):
            base = None
            n = 1
            while n <= ngpu:
                grp = E.FrameGroup(scale, mode, list(range(n)), sd)
                for _ in range(3):
                    out = grp.forward_host(xn, yn)
                t0 = time.perf_counter()
                for _ in range(reps):
                    out = grp.forward_host(xn, yn)
                ms = (time.perf_counter() - t0) / reps * 1e3
                grp.close()
                if base is None:
                    base, ref = ms, out
                print(json.dumps({"workload": f"single frame {w}x{h} x{scale} {mode}", "n_gpus": n, "ms_per_frame": ms,
                                  "MP_per_s": h * w / 1e6 / (ms / 1e3), "speedup_vs_1gpu": base / ms,
                                  "max_abs_diff_vs_1gpu": float(np.abs(out - ref).max())}), flush=True)
                n *= 2


if __name__ == "__main__":
    main()
