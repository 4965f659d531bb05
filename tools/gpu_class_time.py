#!/usr/bin/env python
"""Per-kernel-class device time of one forward (CUDA events inside the library), GPU box:
tools/gpu_class_time.py <mode> <frames> [steps] [H W]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codon_b200 import engine as E, synthetic as syn  # noqa: E402

mode, B = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
H, W = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (480, 640)
eng = E.Engine(4, mode, 0)
eng.load_state_dict(syn.synthetic_state_dict(4, 0))
xh, yh = syn.synthetic_frames(B, H, W, 1234)
x, y = xh.cuda(), yh.cuda()
out = torch.empty_like(x)
for _ in range(3):
    eng.forward(x, y, out)
torch.cuda.synchronize()
eng.profile_reset()
eng.profile_enable(True)
for _ in range(steps):
    eng.forward(x, y, out)
torch.cuda.synchronize()
eng.profile_enable(False)
prof = eng.profile_read()
tot = 0.0
for k, v in prof.items():
    ms = v["ms"] / steps
    tot += ms
    tf = v["work"] / steps / (ms * 1e-3) / 1e12 if ms > 0 and v["work"] else 0.0
    print(f"{mode} b{B} {W}x{H}  {k:22s} {ms:8.4f} ms/step  {v['launches'] // steps:3d} launches  {tf:8.1f} T(FLOP|B)/s")
print(f"{mode} b{B} {W}x{H}  total {tot:.4f} ms/step -> {B * H * W / 1e3 / tot:.2f} MP/s (instrumented)")
