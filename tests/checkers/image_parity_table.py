#!/usr/bin/env python
"""Per-image parity table on the 30 bundled Middlebury inputs (north_star: "RMSE/SSIM identical to 3
decimals on the shipped input_color / input_depth / input_label images").

For every scale (x4, x8, x16) and image: the CPU oracle's fp32 forward (the reference algorithm, synthetic seed-0
weights -- the .pth files are absent from the reference checkout) is quantised with the driver rule and scored with
EvaluationResults / ssim_exact against the label; the GPU engine does the same in each arithmetic mode with its own
quantise / RMSE / SSIM kernels.  Prints one line per (scale, image, mode) and a summary per mode.

  python tests/checkers/image_parity_table.py --make-ref      # CPU only: oracle outputs -> build/parity_ref/*.npy (travels with gpurun)
  python tests/checkers/image_parity_table.py                 # GPU box: table (uses build/parity_ref if present, else computes)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))   # this tool is a checker (tests/-style use of the oracle)
import codon_oracle as orc  # noqa: E402

IMG = os.path.join(ROOT, "tests", "golden", "images")
REF = os.path.join(ROOT, "build", "parity_ref")


def imread(p):
    import cv2
    return cv2.imread(p, 0)


def oracle_out(scale, name, sd):
    p = os.path.join(REF, f"x{scale}_{name}.npy")
    if os.path.exists(p):
        return np.load(p)
    d = imread(os.path.join(IMG, f"depth_x{scale}", name))
    g = imread(os.path.join(IMG, "gray", name))
    x = torch.from_numpy(d / 255).float()[None, None]
    y = torch.from_numpy(g / 255).float()[None, None]
    with torch.no_grad():
        out = orc.forward(sd, x, y)[0, 0].numpy()
    os.makedirs(REF, exist_ok=True)
    np.save(p, out)
    return out


def main():
    names = sorted(os.listdir(os.path.join(IMG, "gray")))
    torch.set_num_threads(os.cpu_count() or 1)
    if "--make-ref" in sys.argv:
        for scale in (4, 8, 16):
            sd = orc.synthetic_state_dict(scale, 0)
            for n in names:
                oracle_out(scale, n, sd)
                print("ref", scale, n, flush=True)
        return
    from codon_b200 import engine
    modes = tuple(a for a in sys.argv[1:] if not a.startswith("-")) or ("f16x3", "fp32", "tf32", "fp16", "bf16")
    # reference metrics: the REAL reference's fp32 forward, committed by oracle/make_golden.py images
    committed = json.load(open(os.path.join(ROOT, "tests", "golden", "image_parity.json")))
    summary = {m: dict(n=0, rmse_eq3=0, ssim_eq3=0, max_d_rmse=0.0, max_d_ssim=0.0, max_abs=0.0, px_changed=0.0) for m in modes}
    print("# scale image mode  rmse_ref ssim_ref | rmse ssim | d_rmse d_ssim | max_abs_err px_changed")
    for scale in (4, 8, 16):
        sd = orc.synthetic_state_dict(scale, 0)
        engs = {}
        for m in modes:
            engs[m] = engine.Engine(scale, m, 0)
            engs[m].load_state_dict(sd)
        for n in names:
            d = imread(os.path.join(IMG, f"depth_x{scale}", n))
            g = imread(os.path.join(IMG, "gray", n))
            lab = imread(os.path.join(IMG, "label", n))
            ref = oracle_out(scale, n, sd)
            q_ref = orc.quantise_output(ref)
            r_ref, s_ref = committed[f"x{scale}/{n}"]["rmse"], committed[f"x{scale}/{n}"]["ssim"]
            x = torch.from_numpy(d / 255).float()[None, None].cuda()
            y = torch.from_numpy(g / 255).float()[None, None].cuda()
            labg = torch.from_numpy(lab).cuda()[None]
            for m in modes:
                out = engs[m].forward(x, y)
                q = engine.quantise_u8(out[0, 0])
                r = float(engine.masked_rmse(labg, q[None])[0])
                s = float(engine.ssim_gauss(labg, q[None])[0])
                err = float(np.abs(out.cpu().numpy()[0, 0] - ref).max())
                ch = float((q.cpu().numpy() != q_ref).mean())
                print(f"x{scale} {n:13s} {m:5s} {r_ref:.4f} {s_ref:.5f} | {r:.4f} {s:.5f} | {r - r_ref:+.4f} {s - s_ref:+.5f} | {err:.2e} {ch:.4f}")
                a = summary[m]
                a["n"] += 1
                a["rmse_eq3"] += round(r, 3) == round(r_ref, 3)
                a["ssim_eq3"] += round(s, 3) == round(s_ref, 3)
                a["max_d_rmse"] = max(a["max_d_rmse"], abs(r - r_ref))
                a["max_d_ssim"] = max(a["max_d_ssim"], abs(s - s_ref))
                a["max_abs"] = max(a["max_abs"], err)
                a["px_changed"] = max(a["px_changed"], ch)
        for e in engs.values():
            e.close()
    print("# summary over 30 images (3 scales x 10): images whose 3-decimal RMSE / SSIM equal the fp32 reference's")
    for m in modes:
        print("#", m, json.dumps(summary[m]))


if __name__ == "__main__":
    main()
