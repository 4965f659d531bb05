#!/usr/bin/env python
"""Per-mode, per-tap error report of the CUDA forward against the oracle (debug aid; GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import codon_oracle as orc  # noqa: E402
from codon_b200.CODON_x4 import CODONNet  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    B, H, W = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (2, 45, 70)))
    sd = orc.synthetic_state_dict(4, 0)
    x, y = orc.synthetic_frames(B, H, W, 7)
    with torch.no_grad():
        ref, taps = orc.forward(sd, x, y, return_taps=True)
    net = CODONNet().eval().set_mode(mode)
    net.load_state_dict(sd)
    with torch.no_grad():
        out = net(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    eng = net.engine(torch.device("cuda", 0))
    want = {"enc": torch.cat((taps["enc_d"], taps["enc_c"]), 1),
            "feat": torch.cat((taps["out_d4"], taps["out_c4"]), 1),
            "fuse": taps["fuse"], "out_fuse": taps["out_fuse"]}
    for name, r in want.items():
        got = eng.debug_tap(name, B, H, W).cpu()
        d = (got - r).abs()
        print(f"[{mode}] tap {name:9s} max-abs {float(d.max()):.3e}  mean-abs {float(d.mean()):.3e}  ref max {float(r.abs().max()):.3f}")
    d = (out.cpu() - ref).abs()
    print(f"[{mode}] output       max-abs {float(d.max()):.3e}  mean-abs {float(d.mean()):.3e}  launches {eng.last_launch_count}")


if __name__ == "__main__":
    main()
