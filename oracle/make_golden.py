#!/usr/bin/env python
"""Generate tests/golden/ by running the REAL reference (read-only at /root/reference).

TEST INFRASTRUCTURE ONLY; runs in the build container (CPU).  The reference is Python and
cannot travel to the GPU box, so its outputs on fixed inputs are committed as fixtures:

  tests/golden/fwd_*.npz      CODONNet.forward of the reference classes (fp32 CPU, and an
                              fp64 run as the rounding-free truth) on the synthetic
                              weights/frames of oracle/codon_oracle.py
  tests/golden/cac_*.npz      CAC_channel / CAC_spatial / ChannelGate / SpatialGate / ResCBAM
  tests/golden/images/        the bundled Middlebury inputs as read by test.py:116-118
                              (cv2.imread(path, 0) -> single-channel uint8 PNG), labels and the
                              authors' x4/x8/x16 outputs (data, not code)
  tests/golden/preproc.npz    cv2 gray conversion of a colour crop; cv2.resize(INTER_CUBIC) of LR depth crops
  tests/golden/metrics.json   EvaluationResults (test.py:148-164, exec'd from source lines) and
                              ssim_2.ssim_exact on those images
  tests/golden/images/ref_fp32_x{4,8,16}/  the reference's fp32 CPU forward (synthetic seed-0 weights) on the 30
                              bundled input pairs, quantised with the driver rule (test.py:130-132) -> uint8 PNG,
                              and tests/golden/image_parity.json = its EvaluationResults / ssim_exact vs the label:
                              the north_star's "RMSE/SSIM identical to 3 decimals" is asserted against these
  tests/golden/big_*.npz      the reference forward at the BASELINE.json shapes (640x480, 8 x 640x480, 1280x720,
                              1920x1080; fp32 and fp64): row sums, column sums, a stride-4 sample and full-resolution
                              crops of the output (the whole frames would be tens of MB)

Usage:  python oracle/make_golden.py [all|images|big|small]   (needs /root/reference)

"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CODON_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
import codon_oracle as orc  # noqa: E402


def load_reference(scale):
    """Import the reference model module for a scale; the three directories reuse module names."""
    for m in ("CAC_module", "attention", "attention.ResCBAM", "CODON_x4", "CODON_x8", "CODON_x16", "ssim_2"):
        sys.modules.pop(m, None)
    d = os.path.join(REF, f"CODON_X{scale}")
    sys.path.insert(0, d)
    try:
        mod = importlib.import_module(f"CODON_x{scale}")
        cac = importlib.import_module("CAC_module")
        ssim2 = importlib.import_module("ssim_2")
        rescbam = importlib.import_module("attention.ResCBAM") if scale in (4, 8) else None
    finally:
        sys.path.remove(d)
    return mod, cac, ssim2, rescbam


def reference_rmse_fn():
    """EvaluationResults cannot be imported (test.py does not import here); exec its source lines."""
    import math  # noqa: F401  (used by the exec'd source)
    src = open(os.path.join(REF, "CODON_X4", "test.py"), encoding="utf-8").read().split("\n")
    body = "\n".join(src[147:164])
    ns = {"np": np, "math": math}
    exec(body, ns)
    return ns["EvaluationResults"]


def run_forward(scale, seed, shape, name, frame_seed):
    mod, _, _, _ = load_reference(scale)
    sd = orc.synthetic_state_dict(scale, seed)
    b, h, w = shape
    x, y = orc.synthetic_frames(b, h, w, frame_seed)
    torch.manual_seed(0)
    net = mod.CODONNet().eval()
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out32 = net(x, y)
        net64 = net.double()
        out64 = net64(x.double(), y.double())
    np.savez_compressed(os.path.join(GOLD, f"fwd_{name}.npz"),
                        scale=scale, seed=seed, frame_seed=frame_seed,
                        x=x.numpy(), y=y.numpy(), out_fp32=out32.numpy(),
                        out_fp64=out64.numpy())
    print(f"fwd_{name}: x{scale} seed {seed} {shape}  residual max "
          f"{float((out64 - x.double()).abs().max()):.4f}  fp32-vs-fp64 {float((out32.double() - out64).abs().max()):.2e}")


def run_image_forward(scale, seed, image, name):
    import cv2
    mod, _, _, _ = load_reference(scale)
    sd = orc.synthetic_state_dict(scale, seed)
    d = cv2.imread(os.path.join(GOLD, "images", f"depth_x{scale}", image), 0)
    g = cv2.imread(os.path.join(GOLD, "images", "gray", image), 0)
    x = torch.from_numpy(d / 255).float()[None, None]
    y = torch.from_numpy(g / 255).float()[None, None]
    net = mod.CODONNet().eval()
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out32 = net(x, y)
    np.savez_compressed(os.path.join(GOLD, f"fwd_{name}.npz"), scale=scale, seed=seed, image=image,
                        out_fp32=out32.numpy())
    print(f"fwd_{name}: x{scale} seed {seed} {image} {tuple(x.shape)}")


def run_cac():
    _, cac, _, rescbam = load_reference(4)
    g = torch.Generator().manual_seed(7)
    sd = orc.synthetic_state_dict(4, 3)
    x = torch.randn(2, 128, 21, 27, generator=g)
    ch = cac.CAC_channel(128).eval()
    ch.load_state_dict({k[len("attention_c2."):]: v for k, v in sd.items() if k.startswith("attention_c2.")})
    sp = cac.CAC_spatial().eval()
    sp.load_state_dict({k[len("attention_s2."):]: v for k, v in sd.items() if k.startswith("attention_s2.")})
    x64 = torch.randn(2, 64, 19, 23, generator=g)
    cg = rescbam.ChannelGate(64).eval()
    cg.load_state_dict({k[len("attention_c5."):]: v for k, v in sd.items() if k.startswith("attention_c5.")})
    sg = rescbam.SpatialGate().eval()
    sg.load_state_dict({k[len("attention_s5."):]: v for k, v in sd.items() if k.startswith("attention_s5.")})
    with torch.no_grad():
        np.savez_compressed(os.path.join(GOLD, "cac_modules.npz"),
                            x=x.numpy(), channel=ch(x)[:, :, 0, 0].numpy(), spatial=sp(x).numpy(),
                            pool=cac.ChannelPool()(x).numpy(),
                            x64=x64.numpy(), channel_gate=cg(x64).numpy(), spatial_gate=sg(x64).numpy())
    print("cac_modules done")


def copy_images_and_metrics():
    import cv2
    rmse_fn = reference_rmse_fn()
    _, _, ssim2, _ = load_reference(4)
    names = sorted(os.listdir(os.path.join(REF, "CODON_X4", "input_color")))
    img_dir = os.path.join(GOLD, "images")
    for sub in ("gray", "label", "depth_x4", "depth_x8", "depth_x16", "ref_out_x4", "ref_out_x8", "ref_out_x16"):
        os.makedirs(os.path.join(img_dir, sub), exist_ok=True)
    metrics = {}
    for n in names:
        gray = cv2.imread(os.path.join(REF, "CODON_X4", "input_color", n), 0)      # test.py:118
        label = cv2.imread(os.path.join(REF, "CODON_X4", "input_label", n), 0)     # test.py:117
        cv2.imwrite(os.path.join(img_dir, "gray", n), gray, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        cv2.imwrite(os.path.join(img_dir, "label", n), label, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        for s in (4, 8, 16):
            dep = cv2.imread(os.path.join(REF, f"CODON_X{s}", "input_depth", n), 0)   # test.py:116
            out = cv2.imread(os.path.join(REF, f"CODON_X{s}", "output", n), 0)
            cv2.imwrite(os.path.join(img_dir, f"depth_x{s}", n), dep, [cv2.IMWRITE_PNG_COMPRESSION, 9])
            cv2.imwrite(os.path.join(img_dir, f"ref_out_x{s}", n), out, [cv2.IMWRITE_PNG_COMPRESSION, 9])
            metrics[f"x{s}/{n}"] = {
                "rmse_out": rmse_fn(label, out), "ssim_out": float(ssim2.ssim_exact(label / 255, out / 255)),
                "rmse_in": rmse_fn(label, dep), "ssim_in": float(ssim2.ssim_exact(label / 255, dep / 255)),
            }
    json.dump(metrics, open(os.path.join(GOLD, "metrics.json"), "w"), indent=1, sort_keys=True)
    for s in (4, 8, 16):
        r = np.mean([metrics[f"x{s}/{n}"]["rmse_out"] for n in names])
        q = np.mean([metrics[f"x{s}/{n}"]["ssim_out"] for n in names])
        print(f"x{s}: mean RMSE {r:.4f} SSIM {q:.4f}")


def run_preproc():
    """Fixtures for the GPU pre-processing kernels: cv2's own gray conversion of a decoded colour image
    (what cv2.imread(path, 0) returns, test.py:118) and cv2.resize(INTER_CUBIC) on float32."""
    import cv2
    col = cv2.imread(os.path.join(REF, "CODON_X4", "input_color", "Art.png"), cv2.IMREAD_COLOR)
    gray = cv2.imread(os.path.join(REF, "CODON_X4", "input_color", "Art.png"), 0)
    r, g, b = (col[..., i].astype(np.uint32) for i in (2, 1, 0))
    libpng = ((r * 9797 + g * 19234 + b * 3737) >> 15).astype(np.uint8)       # png_set_rgb_to_gray(0.299, 0.587)
    cvt = cv2.cvtColor(col, cv2.COLOR_BGR2GRAY)
    print("imread(path,0) == libpng formula on imread(path,1):", bool((libpng == gray).all()))
    y0, x0 = 100, 150
    bgr_crop = np.ascontiguousarray(col[y0:y0 + 64, x0:x0 + 80])
    gray_crop = np.ascontiguousarray(gray[y0:y0 + 64, x0:x0 + 80])
    dep = cv2.imread(os.path.join(REF, "CODON_X4", "input_label", "Art.png"), 0)
    out = {"bgr": bgr_crop, "gray": gray_crop, "gray_cvtcolor": np.ascontiguousarray(cvt[y0:y0 + 64, x0:x0 + 80])}
    for s, (H, W) in ((4, (92, 116)), (8, (120, 88)), (16, (97, 131))):
        hr = dep[40:40 + H, 60:60 + W].astype(np.float32) / 255
        h, w = max(1, H // s), max(1, W // s)
        lr = cv2.resize(hr, (w, h), interpolation=cv2.INTER_AREA)
        up = cv2.resize(lr, (W, H), interpolation=cv2.INTER_CUBIC)
        out[f"lr_x{s}"] = lr
        out[f"up_x{s}"] = up
    np.savez_compressed(os.path.join(GOLD, "preproc.npz"), **out)
    print("preproc fixtures written")


def run_image_parity_refs():
    """The reference's own fp32 result on every bundled input pair, as the uint8 image its driver would write
    (test.py:122-132: /255 -> float -> forward -> clip -> *255 -> astype(uint8)) and the two metrics of that image."""
    import cv2
    rmse_fn = reference_rmse_fn()
    img_dir = os.path.join(GOLD, "images")
    names = sorted(os.listdir(os.path.join(img_dir, "gray")))
    table = {}
    for scale in (4, 8, 16):
        mod, _, ssim2, _ = load_reference(scale)
        net = mod.CODONNet().eval()
        net.load_state_dict(orc.synthetic_state_dict(scale, 0), strict=True)
        os.makedirs(os.path.join(img_dir, f"ref_fp32_x{scale}"), exist_ok=True)
        for n in names:
            d = cv2.imread(os.path.join(img_dir, f"depth_x{scale}", n), 0)
            g = cv2.imread(os.path.join(img_dir, "gray", n), 0)
            lab = cv2.imread(os.path.join(img_dir, "label", n), 0)
            x = torch.from_numpy(d / 255).float()[None, None]
            y = torch.from_numpy(g / 255).float()[None, None]
            with torch.no_grad():
                out = net(x, y).squeeze().numpy()
            out = np.clip(out, 0, 1)
            q = (out * 255).astype(np.uint8)
            cv2.imwrite(os.path.join(img_dir, f"ref_fp32_x{scale}", n), q, [cv2.IMWRITE_PNG_COMPRESSION, 9])
            table[f"x{scale}/{n}"] = {"rmse": rmse_fn(lab, q), "ssim": float(ssim2.ssim_exact(lab / 255, q / 255))}
            print(f"image parity ref x{scale} {n}: rmse {table[f'x{scale}/{n}']['rmse']:.4f} ssim {table[f'x{scale}/{n}']['ssim']:.5f}", flush=True)
    json.dump(table, open(os.path.join(GOLD, "image_parity.json"), "w"), indent=1, sort_keys=True)


BIG_CROP = 96


def big_summary(out):
    """Size-bounded fingerprint of a [H, W] output: row / column sums (fp64), a stride-4 sample, five crops."""
    H, W = out.shape
    o = out.astype(np.float64)
    c = BIG_CROP
    ys = (0, H - c, (H - c) // 2)
    xs = (0, W - c, (W - c) // 2)
    crops = {"tl": out[:c, :c], "tr": out[:c, W - c:], "bl": out[H - c:, :c], "br": out[H - c:, W - c:],
             "ce": out[ys[2]:ys[2] + c, xs[2]:xs[2] + c]}
    return {"row_sum": o.sum(1), "col_sum": o.sum(0), "sample4": out[1::4, 2::4].copy(), **{f"crop_{k}": v.copy() for k, v in crops.items()}}


def run_big(scale, seed, shape, name, frame_seed, frames=(0,), with_fp64=True):
    """Reference forward at a BASELINE.json shape; stores fingerprints of the frames listed in `frames`."""
    mod, _, _, _ = load_reference(scale)
    sd = orc.synthetic_state_dict(scale, seed)
    b, h, w = shape
    x, y = orc.synthetic_frames(b, h, w, frame_seed)
    net = mod.CODONNet().eval()
    net.load_state_dict(sd, strict=True)
    rec = {"scale": scale, "seed": seed, "frame_seed": frame_seed, "shape": np.array(shape), "frames": np.array(frames),
           "x_sum": x.double().sum(dim=(1, 2, 3)).numpy(), "y_sum": y.double().sum(dim=(1, 2, 3)).numpy()}
    with torch.no_grad():
        out32 = net(x, y).numpy()
        for f in frames:
            for k, v in big_summary(out32[f, 0]).items():
                rec[f"f{f}_fp32_{k}"] = v
        if with_fp64:
            # frame by frame (frames are independent: CAC pools per sample): a whole fp64 batch does not fit in memory
            net64 = net.double()
            for f in frames:
                out64 = net64(x[f:f + 1].double(), y[f:f + 1].double()).numpy()
                for k, v in big_summary(out64[0, 0]).items():
                    rec[f"f{f}_fp64_{k}"] = v
                print(f"big_{name}: frame {f} fp32-vs-fp64 max-abs {np.abs(out32[f:f + 1] - out64).max():.2e}", flush=True)
    np.savez_compressed(os.path.join(GOLD, f"big_{name}.npz"), **rec)
    print(f"big_{name}: x{scale} seed {seed} {shape} done", flush=True)


def write_c_case(npz_name, out_name):
    """A committed forward golden (real reference output) in the raw layout examples/c_consumer.c reads:
    "CODONC1\\0" | int32 B, H, W | x | y | out_fp32   (float32, little endian)."""
    import struct
    g = np.load(os.path.join(GOLD, npz_name))
    x, y, out = g["x"], g["y"], g["out_fp32"]
    b, _, h, w = x.shape
    with open(os.path.join(GOLD, out_name), "wb") as f:
        f.write(b"CODONC1\0")
        f.write(struct.pack("<3i", b, h, w))
        for a in (x, y, out):
            f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())
    print(f"{out_name}: {b}x{h}x{w}")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "c_case":
        write_c_case("fwd_x4_s0_b2_48x64.npz", "c_case_x4_s0_b2_48x64.bin")
    if what in ("all", "small"):
        copy_images_and_metrics()
        run_cac()
        run_forward(4, 0, (2, 48, 64), "x4_s0_b2_48x64", 1234)
        run_forward(4, 1, (1, 37, 53), "x4_s1_b1_37x53", 77)
        run_forward(8, 2, (1, 40, 72), "x8_s2_b1_40x72", 5)
        run_forward(16, 2, (1, 64, 80), "x16_s2_b1_64x80", 99)
        run_forward(4, 0, (1, 120, 160), "x4_s0_b1_120x160", 4321)
        run_image_forward(4, 0, "Tsukuba.png", "x4_s0_tsukuba")
        run_preproc()
        write_c_case("fwd_x4_s0_b2_48x64.npz", "c_case_x4_s0_b2_48x64.bin")
    if what in ("all", "images"):
        run_image_parity_refs()
    if what.startswith("big:"):       # one shape: big:x8 | big:720 | big:1080
        which = what.split(":")[1]
        if which == "x8":
            run_big(8, 1, (8, 480, 640), "x8_b8_640x480", 2000, frames=(0, 7))
        elif which == "720":
            run_big(4, 2, (1, 720, 1280), "x4_1280x720", 3000)
        elif which == "1080":
            run_big(16, 0, (1, 1080, 1920), "x16_1920x1080", 4000, with_fp64=False)
    if what in ("all", "big"):
        # BASELINE.json configs[1..4]: x4 640x480; x8 batch of 8 x 640x480 (one GPU's shard of the 64);
        # x16 1920x1080; 1280x720 (x4 here; configs[4] cycles the scales)
        run_big(4, 0, (1, 480, 640), "x4_640x480", 1234)
        run_big(8, 1, (8, 480, 640), "x8_b8_640x480", 2000, frames=(0, 7))
        run_big(4, 2, (1, 720, 1280), "x4_1280x720", 3000)
        run_big(16, 0, (1, 1080, 1920), "x16_1920x1080", 4000, with_fp64=False)   # fp64 at 1080p needs > 60 GB
