"""CPU: the oracle (oracle/codon_oracle.py) against the fixtures generated from the REAL reference
(oracle/make_golden.py, run in the build container where /root/reference exists)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import codon_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


FWD = ["fwd_x4_s0_b2_48x64.npz", "fwd_x4_s1_b1_37x53.npz", "fwd_x8_s2_b1_40x72.npz", "fwd_x16_s2_b1_64x80.npz"]


@pytest.mark.parametrize("name", FWD)
def test_forward_matches_reference_fp64(golden_dir, name):
    g = _load(golden_dir, name)
    sd = orc.synthetic_state_dict(int(g["scale"]), int(g["seed"]))
    x, y = torch.from_numpy(g["x"]).double(), torch.from_numpy(g["y"]).double()
    with torch.no_grad():
        out = orc.forward(sd, x, y).numpy()
    assert out.shape == g["out_fp64"].shape
    np.testing.assert_allclose(out, g["out_fp64"], rtol=0, atol=1e-10)


def test_forward_fp32_close_to_reference_fp32(golden_dir):
    g = _load(golden_dir, FWD[0])
    sd = orc.synthetic_state_dict(4, 0)
    with torch.no_grad():
        out = orc.forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["y"])).numpy()
    np.testing.assert_allclose(out, g["out_fp32"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out, g["out_fp64"], rtol=0, atol=2e-5)


def test_synthetic_frames_are_reproducible(golden_dir):
    g = _load(golden_dir, FWD[1])
    x, y = orc.synthetic_frames(1, 37, 53, int(g["frame_seed"]))
    np.testing.assert_array_equal(x.numpy(), g["x"])
    np.testing.assert_array_equal(y.numpy(), g["y"])


def test_param_inventory():
    assert len(orc.param_shapes(4)) == 49 and len(orc.param_shapes(8)) == 49 and len(orc.param_shapes(16)) == 44
    assert sum(int(np.prod(s)) for s in orc.param_shapes(4).values()) == 1866136
    assert sum(int(np.prod(s)) for s in orc.param_shapes(16).values()) == 1865506
    assert orc.flops_per_pixel() == 14855552


def test_cac_modules_match_reference(golden_dir):
    g = _load(golden_dir, "cac_modules.npz")
    sd = orc.synthetic_state_dict(4, 3)
    x = torch.from_numpy(g["x"])
    np.testing.assert_allclose(orc.cac_channel_scale(sd, "attention_c2", x).numpy(), g["channel"], atol=1e-6)
    np.testing.assert_allclose(orc.cac_spatial_scale(sd, "attention_s2", x).numpy(), g["spatial"], atol=1e-6)
    np.testing.assert_allclose(orc.channel_pool(x).numpy(), g["pool"], atol=1e-6)
    x64 = torch.from_numpy(g["x64"])
    np.testing.assert_allclose(orc.channel_gate(sd, "attention_c5", x64).numpy(), g["channel_gate"], atol=1e-6)
    np.testing.assert_allclose(orc.spatial_gate(sd, "attention_s5", x64).numpy(), g["spatial_gate"], atol=1e-6)


def _imread(path):
    import cv2
    return cv2.imread(path, 0)


def test_metrics_match_reference_on_bundled_images(golden_dir):
    """EvaluationResults / ssim_exact of the reference on the authors' outputs."""
    metrics = json.load(open(os.path.join(golden_dir, "metrics.json")))
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(golden_dir, "images", "label", "*.png")))
    assert len(names) == 10
    means = {}
    for s in (4, 8, 16):
        rm, ss = [], []
        for n in names:
            lab = _imread(os.path.join(golden_dir, "images", "label", n))
            out = _imread(os.path.join(golden_dir, "images", f"ref_out_x{s}", n))
            m = metrics[f"x{s}/{n}"]
            r = orc.masked_rmse(lab, out)
            q = orc.ssim_gauss(lab / 255, out / 255)
            assert abs(r - m["rmse_out"]) < 1e-9
            assert abs(q - m["ssim_out"]) < 1e-9
            rm.append(r)
            ss.append(q)
        means[s] = (np.mean(rm), np.mean(ss))
    # BASELINE.md section 2
    assert round(means[4][0], 4) == 1.7779 and round(means[4][1], 4) == 0.9542
    assert round(means[8][0], 4) == 3.4789 and round(means[8][1], 4) == 0.9370
    assert round(means[16][0], 4) == 5.8032 and round(means[16][1], 4) == 0.9097


def test_quantise_truncates():
    a = np.array([-0.2, 0.0, 0.5, 0.999, 1.0, 1.7], np.float32)
    np.testing.assert_array_equal(orc.quantise_output(a), np.array([0, 0, 127, 254, 255, 255], np.uint8))
