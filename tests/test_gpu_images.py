"""GPU: the test.py-compatible driver on the bundled Middlebury images (tests/golden/images), and
RMSE/SSIM parity of the reduced-precision modes against the reference fp32 forward (north_star:
"RMSE/SSIM identical to 3 decimals ... in bf16 mode").

The reference's own result for Tsukuba (fp32 CPU forward of the real reference class with synthetic
weights seed 0) is committed as tests/golden/fwd_x4_s0_tsukuba.npz.  SURVEY.md 7.3(4) measured that
truncation to uint8 makes 3-decimal RMSE equality unreachable for bf16 operands; the achieved
deltas are printed and bounded here instead of being assumed.
"""
import json
import os

import numpy as np
import pytest
import torch

import codon_oracle as orc
from codon_b200 import engine, test as driver
from codon_b200.CODON_x4 import CODONNet

pytestmark = pytest.mark.gpu


def _imread(p):
    import cv2
    return cv2.imread(p, 0)


def _metrics_cpu(out_f32, label_u8):
    q = orc.quantise_output(out_f32)
    return orc.masked_rmse(label_u8, q), orc.ssim_gauss(label_u8 / 255, q / 255), q


def test_tsukuba_modes_vs_reference_fp32(golden_dir):
    g = np.load(os.path.join(golden_dir, "fwd_x4_s0_tsukuba.npz"))
    img = os.path.join(golden_dir, "images")
    d = _imread(os.path.join(img, "depth_x4", "Tsukuba.png"))
    c = _imread(os.path.join(img, "gray", "Tsukuba.png"))
    lab = _imread(os.path.join(img, "label", "Tsukuba.png"))
    ref = g["out_fp32"][0, 0]
    r_ref, s_ref, q_ref = _metrics_cpu(ref, lab)
    x = torch.from_numpy(d / 255).float()[None, None].cuda()
    y = torch.from_numpy(c / 255).float()[None, None].cuda()
    sd = orc.synthetic_state_dict(4, 0)
    report = {}
    for mode in ("fp32", "tf32", "fp16", "bf16"):
        net = CODONNet().eval().set_mode(mode)
        net.load_state_dict(sd)
        with torch.no_grad():
            out = net(x, y)
        err = float(np.abs(out.cpu().numpy()[0, 0] - ref).max())
        q = engine.quantise_u8(out[0, 0])
        r = float(engine.masked_rmse(torch.from_numpy(lab).cuda()[None], q[None])[0])
        s = float(engine.ssim_gauss(torch.from_numpy(lab).cuda()[None], q[None])[0])
        flipped = float((q.cpu().numpy() != q_ref).mean())
        report[mode] = dict(max_abs=err, rmse=r, ssim=s, d_rmse=r - r_ref, d_ssim=s - s_ref, px_changed=flipped)
    print(json.dumps({"reference_fp32": dict(rmse=r_ref, ssim=s_ref), **report}, indent=1))
    assert report["fp32"]["max_abs"] <= 1e-4 and report["tf32"]["max_abs"] <= 1e-3 and report["fp16"]["max_abs"] <= 1e-3
    assert round(report["fp32"]["rmse"], 3) == round(r_ref, 3) and round(report["fp32"]["ssim"], 3) == round(s_ref, 3)
    for mode, (dr, ds) in {"tf32": (5e-3, 5e-4), "fp16": (5e-3, 5e-4), "bf16": (5e-2, 1e-3)}.items():
        assert abs(report[mode]["d_rmse"]) <= dr, (mode, report[mode])
        assert abs(report[mode]["d_ssim"]) <= ds, (mode, report[mode])


# ---- north_star: "RMSE/SSIM (ssim_2.py) identical to 3 decimals on the shipped input_color / input_depth / input_label
# images".  Reference = the REAL reference classes' fp32 CPU forward (synthetic seed-0 weights; the .pth files are not
# shipped), quantised by the driver rule and scored by the reference's own EvaluationResults / ssim_exact:
# tests/golden/images/ref_fp32_x*/ + tests/golden/image_parity.json (oracle/make_golden.py images).
_PARITY_MODES = ("f16x3", "fp32", "tf32", "fp16", "bf16")


def _image_parity(golden_dir, scale, modes):
    table = json.load(open(os.path.join(golden_dir, "image_parity.json")))
    img = os.path.join(golden_dir, "images")
    names = sorted(os.listdir(os.path.join(img, "gray")))
    sd = orc.synthetic_state_dict(scale, 0)
    rows = []
    for mode in modes:
        eng = engine.Engine(scale, mode, 0)
        eng.load_state_dict(sd)
        for n in names:
            d = _imread(os.path.join(img, f"depth_x{scale}", n))
            g = _imread(os.path.join(img, "gray", n))
            lab = _imread(os.path.join(img, "label", n))
            q_ref = _imread(os.path.join(img, f"ref_fp32_x{scale}", n))
            ref = table[f"x{scale}/{n}"]
            x = torch.from_numpy(d / 255).float()[None, None].cuda()      # test.py:122-123
            y = torch.from_numpy(g / 255).float()[None, None].cuda()
            out = eng.forward(x, y)
            q = engine.quantise_u8(out[0, 0])                              # test.py:130-132 on the GPU
            labg = torch.from_numpy(lab).cuda()[None]
            r = float(engine.masked_rmse(labg, q[None])[0])                # test.py:148-164
            s = float(engine.ssim_gauss(labg, q[None])[0])                 # ssim_2.py:36-52
            rows.append(dict(scale=scale, image=n, mode=mode, rmse=r, ssim=s, rmse_ref=ref["rmse"], ssim_ref=ref["ssim"],
                             px_changed=int((q.cpu().numpy() != q_ref).sum()), px=int(q_ref.size)))
        eng.close()
    return rows


@pytest.mark.parametrize("scale", [4, 8, 16])
def test_image_parity_three_decimals(golden_dir, scale):
    """f16x3 (the fp32-accurate tensor-core mode) and fp32 (FFMA) must reproduce the reference's RMSE and SSIM to three
    decimals on all ten images of the scale; the 16-bit / tf32 modes are reported with their achieved deltas."""
    rows = _image_parity(golden_dir, scale, _PARITY_MODES)
    summary = {}
    for r in rows:
        a = summary.setdefault(r["mode"], dict(n=0, rmse_eq3=0, ssim_eq3=0, max_d_rmse=0.0, max_d_ssim=0.0, max_px_changed=0))
        a["n"] += 1
        a["rmse_eq3"] += round(r["rmse"], 3) == round(r["rmse_ref"], 3)
        a["ssim_eq3"] += round(r["ssim"], 3) == round(r["ssim_ref"], 3)
        a["max_d_rmse"] = max(a["max_d_rmse"], abs(r["rmse"] - r["rmse_ref"]))
        a["max_d_ssim"] = max(a["max_d_ssim"], abs(r["ssim"] - r["ssim_ref"]))
        a["max_px_changed"] = max(a["max_px_changed"], r["px_changed"])
    print(json.dumps({f"x{scale}": summary}, indent=1))
    for mode in ("f16x3", "fp32"):
        bad = [(r["image"], r["rmse"], r["rmse_ref"], r["ssim"], r["ssim_ref"]) for r in rows if r["mode"] == mode and
               (round(r["rmse"], 3) != round(r["rmse_ref"], 3) or round(r["ssim"], 3) != round(r["ssim_ref"], 3))]
        assert not bad, (mode, bad)
    # the plain tensor-core modes: SSIM to three decimals within 1e-3, RMSE bounded (uint8 truncation flips pixels that
    # sit on a grey-level boundary)
    for mode, (dr, ds) in {"tf32": (5e-3, 5e-4), "fp16": (5e-3, 5e-4), "bf16": (6e-2, 1e-3)}.items():
        assert summary[mode]["max_d_rmse"] <= dr and summary[mode]["max_d_ssim"] <= ds, (mode, summary[mode])


def test_driver_end_to_end_on_bundled_images(golden_dir, tmp_path, capsys):
    """codon_b200.test.main with the reference's flags + data-path flags; checks the printed
    per-image lines and means against the CPU oracle metrics of the written PNGs."""
    img = os.path.join(golden_dir, "images")
    out_dir = tmp_path / "CODON_result_save"
    res = driver.main(["--gpus", "0", "--scale", "4", "--mode", "fp16", "--input-depth", os.path.join(img, "depth_x4"),
                       "--input-color", os.path.join(img, "gray"), "--label", os.path.join(img, "label"),
                       "--out", str(out_dir) + "/", "--logfile", str(tmp_path / "log.txt"), "--seed", "1"])
    import sys
    sys.stdout = sys.__stdout__
    mean_rmse, mean_ssim, n = res
    assert n == 10
    names = sorted(os.listdir(os.path.join(img, "gray")))
    rm, ss = [], []
    for nme in names:
        out = _imread(str(out_dir / nme))
        lab = _imread(os.path.join(img, "label", nme))
        assert out is not None and out.shape == lab.shape
        rm.append(orc.masked_rmse(lab, out))
        ss.append(orc.ssim_gauss(lab / 255, out / 255))
    assert abs(np.mean(rm) - mean_rmse) < 1e-9 and abs(np.mean(ss) - mean_ssim) < 1e-9
    log = open(tmp_path / "log.txt").read()
    assert "Tsukuba.png" in log and log.strip().splitlines()[-2] == "10"


def test_driver_with_true_low_resolution_depth(golden_dir, tmp_path):
    """--lr-depth: the driver uploads the LR depth as uint8 (s^2 fewer bytes) and up-samples it on the GPU with
    cv2.INTER_CUBIC semantics; the result must match running the driver on depth maps pre-upsampled by cv2 itself
    (the reference's offline "Bicubic/X4" step, test.py:77) to within the rounding of that offline uint8 image."""
    import cv2
    import sys
    img = os.path.join(golden_dir, "images")
    names = sorted(os.listdir(os.path.join(img, "gray")))[:3]
    lr_dir, up_dir, col_dir, lab_dir = (tmp_path / d for d in ("lr", "up", "col", "lab"))
    for d in (lr_dir, up_dir, col_dir, lab_dir):
        d.mkdir()
    for n in names:
        lab = _imread(os.path.join(img, "label", n))
        H, W = lab.shape
        lr = cv2.resize(lab, (W // 4, H // 4), interpolation=cv2.INTER_AREA)
        up = np.clip(cv2.resize(lr.astype(np.float32) / 255, (W, H), interpolation=cv2.INTER_CUBIC), 0, 1)
        cv2.imwrite(str(lr_dir / n), lr)
        cv2.imwrite(str(up_dir / n), np.round(up * 255).astype(np.uint8))
        cv2.imwrite(str(col_dir / n), _imread(os.path.join(img, "gray", n)))
        cv2.imwrite(str(lab_dir / n), lab)
    common = ["--gpus", "0", "--scale", "4", "--mode", "tf32", "--input-color", str(col_dir), "--label", str(lab_dir),
              "--logfile", "", "--seed", "1"]
    a = driver.main(common + ["--input-depth", str(up_dir), "--lr-depth", str(lr_dir), "--out", str(tmp_path / "o1") + "/"])
    b = driver.main(common + ["--input-depth", str(up_dir), "--out", str(tmp_path / "o2") + "/"])
    sys.stdout = sys.__stdout__
    assert a[2] == b[2] == 3
    print(f"--lr-depth: mean RMSE {a[0]:.4f} SSIM {a[1]:.5f}; pre-upsampled uint8 input: {b[0]:.4f} {b[1]:.5f}")
    assert abs(a[0] - b[0]) < 0.1 and abs(a[1] - b[1]) < 2e-3        # the offline path rounds the input to grey levels


def test_evaluationresults_and_ssim_exact_dropins(golden_dir):
    metrics = json.load(open(os.path.join(golden_dir, "metrics.json")))
    img = os.path.join(golden_dir, "images")
    lab = _imread(os.path.join(img, "label", "Art.png"))
    out = _imread(os.path.join(img, "ref_out_x8", "Art.png"))
    from codon_b200.ssim_2 import ssim_exact
    assert abs(driver.EvaluationResults(lab, out) - metrics["x8/Art.png"]["rmse_out"]) < 1e-12
    assert abs(ssim_exact(lab / 255, out / 255) - metrics["x8/Art.png"]["ssim_out"]) < 1e-10
