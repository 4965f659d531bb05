#!/usr/bin/env python
"""Multi-GPU box: one small frame through the row-band single-frame mode on all visible GPUs vs one GPU (run under a
short `timeout`: a protocol bug in the cross-GPU ordering would hang, not fail)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codon_b200 import engine as E, synthetic as syn  # noqa: E402

n = torch.cuda.device_count()
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "bf16", "tf32", "f16x3"]
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (203, 176)
sd = syn.synthetic_state_dict(4, 0)
x, y = syn.synthetic_frames(1, h, w, 17)
xn, yn = x.numpy()[0, 0], y.numpy()[0, 0]
for mode in modes:
    one = E.FrameGroup(4, mode, [0], sd)
    ref = one.forward_host(xn, yn)
    one.close()
    print(f"{mode}: single GPU done", flush=True)
    grp = E.FrameGroup(4, mode, list(range(n)), sd)
    got = grp.forward_host(xn, yn)
    print(f"{mode}: first {n}-GPU forward done", flush=True)
    t0 = time.perf_counter()
    for _ in range(5):
        again = grp.forward_host(xn, yn)
    ms = (time.perf_counter() - t0) / 5 * 1e3
    grp.close()
    print(f"{mode} {w}x{h} over {n} GPUs: max |band - single| = {np.abs(got - ref).max():.3e}, repeatable {np.array_equal(got, again)}, "
          f"{ms:.3f} ms per frame", flush=True)
