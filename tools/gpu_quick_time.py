#!/usr/bin/env python
"""Device time of the forward alone (no instrumentation, L2 flushed between steps), for same-session A/B runs of
environment switches or library variants:  tools/gpu_quick_time.py <mode> <frames> [steps] [H W]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codon_b200 import engine as E, synthetic as syn  # noqa: E402

mode, B = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
H, W = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (480, 640)
eng = E.Engine(4, mode, 0)
eng.load_state_dict(syn.synthetic_state_dict(4, 0))
xh, yh = syn.synthetic_frames(B, H, W, 1234)
x, y = xh.cuda(), yh.cuda()
out = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    eng.forward(x, y, out)
torch.cuda.synchronize()
time.sleep(1.0)
for _ in range(3):
    eng.forward(x, y, out)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for a, b in ev:
    flush.zero_()
    a.record()
    eng.forward(x, y, out)
    b.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
print(f"{mode} b{B} {W}x{H}: mean {sum(ts) / len(ts):.4f} ms  median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f} ms  "
      f"-> {B * H * W / 1e3 / (sum(ts) / len(ts)):.2f} MP/s  launches {eng.last_launch_count}  checksum {float(out.double().sum()):.6f}")
