"""GPU parity of the stand-alone attention modules and metrics against the reference fixtures."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import codon_oracle as orc
from codon_b200 import engine
from codon_b200.CAC_module import BasicConv, CAC_channel, CAC_spatial, ChannelPool, logsumexp_2d
from codon_b200.attention.ResCBAM import ChannelGate, ResCBAM, ResCBAM_c, ResCBAM_d, SpatialGate

pytestmark = pytest.mark.gpu


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def test_cac_modules_match_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "cac_modules.npz"))
    sd = orc.synthetic_state_dict(4, 3)
    x = torch.from_numpy(g["x"]).cuda()
    ch = CAC_channel(128).eval()
    ch.load_state_dict(_sub(sd, "attention_c2."))
    sp = CAC_spatial().eval()
    sp.load_state_dict(_sub(sd, "attention_s2."))
    with torch.no_grad():
        c = ch.cuda()(x)
        s = sp.cuda()(x)
        p = ChannelPool()(x)
    assert c.shape == (2, 64, 21, 27) and s.shape == (2, 1, 21, 27)
    np.testing.assert_allclose(c[:, :, 0, 0].cpu().numpy(), g["channel"], atol=2e-6)
    assert torch.equal(c[:, :, 0, 0], c[:, :, 5, 7])
    np.testing.assert_allclose(s.cpu().numpy(), g["spatial"], atol=2e-6)
    np.testing.assert_allclose(p.cpu().numpy(), g["pool"], atol=2e-6)
    x64 = torch.from_numpy(g["x64"]).cuda()
    cg = ChannelGate(64).eval()
    cg.load_state_dict(_sub(sd, "attention_c5."))
    sg = SpatialGate().eval()
    sg.load_state_dict(_sub(sd, "attention_s5."))
    with torch.no_grad():
        np.testing.assert_allclose(cg.cuda()(x64).cpu().numpy(), g["channel_gate"], atol=5e-6)
        np.testing.assert_allclose(sg.cuda()(x64).cpu().numpy(), g["spatial_gate"], atol=5e-6)


@pytest.mark.parametrize("cls,pools", [(ResCBAM, ["avg", "max"]), (ResCBAM_c, ["avg"]), (ResCBAM_d, ["max"])])
def test_rescbam_wrappers(cls, pools):
    torch.manual_seed(0)
    m = cls(32).eval()
    x = torch.randn(2, 32, 13, 9)
    sd = {("attention_c." + k[len("ChannelGate."):] if k.startswith("ChannelGate.") else "attention_s." + k[len("SpatialGate."):]): v
          for k, v in m.state_dict().items()}
    # oracle: channel gate with the selected pools, then spatial gate, + x  (ResCBAM.py:94-134)
    h = {"avg": x.mean((2, 3)), "max": x.amax((2, 3))}
    z = sum(orc._mlp(sd, "attention_c", h[p]) for p in pools)
    xc = x * torch.sigmoid(z)[:, :, None, None]
    ref = orc.spatial_gate(sd, "attention_s", xc) + x
    with torch.no_grad():
        out = m.cuda()(x.cuda()).cpu()
    np.testing.assert_allclose(out.numpy(), ref.detach().numpy(), atol=5e-6)


def test_pool_types_lp_lse_and_basicconv():
    torch.manual_seed(1)
    x = torch.randn(2, 16, 11, 14)
    st = engine.channel_stats(x.cuda()).cpu()
    np.testing.assert_allclose(st[0].numpy(), x.mean((2, 3)).numpy(), atol=1e-6)
    np.testing.assert_allclose(st[1].numpy(), x.amax((2, 3)).numpy(), atol=0)
    np.testing.assert_allclose(st[2].numpy(), x.pow(2).sum((2, 3)).sqrt().numpy(), rtol=1e-6)
    np.testing.assert_allclose(st[3].numpy(), torch.logsumexp(x.flatten(2), 2).numpy(), rtol=1e-6)
    assert logsumexp_2d(x.cuda()).shape == (2, 16, 1)
    conv = BasicConv(16, 8, 3, stride=2, padding=1, dilation=1, groups=2, relu=True, bn=True, bias=True).eval()
    conv.bn.running_mean.normal_()
    conv.bn.running_var.uniform_(0.5, 2.0)
    import torch.nn.functional as F
    c = conv.conv
    ref = F.relu(F.batch_norm(F.conv2d(x, c.weight, c.bias, 2, 1, 1, 2), conv.bn.running_mean, conv.bn.running_var,
                              conv.bn.weight, conv.bn.bias, False, 0.0, conv.bn.eps))
    with torch.no_grad():
        out = conv.cuda()(x.cuda()).cpu()
    np.testing.assert_allclose(out.numpy(), ref.detach().numpy(), atol=2e-5)


def _imread(path):
    import cv2
    return cv2.imread(path, 0)


def test_gpu_metrics_match_reference_table(golden_dir):
    """codon_masked_rmse / codon_ssim_gauss vs EvaluationResults / ssim_exact of the reference
    (tests/golden/metrics.json) on the bundled label / output PNGs."""
    metrics = json.load(open(os.path.join(golden_dir, "metrics.json")))
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(golden_dir, "images", "label", "*.png")))
    for s in (4, 16):
        for n in names:
            lab = torch.from_numpy(_imread(os.path.join(golden_dir, "images", "label", n))).cuda()
            out = torch.from_numpy(_imread(os.path.join(golden_dir, "images", f"ref_out_x{s}", n))).cuda()
            m = metrics[f"x{s}/{n}"]
            r = float(engine.masked_rmse(lab, out)[0])
            q = float(engine.ssim_gauss(lab, out)[0])
            q64 = float(engine.ssim_gauss(lab.double() / 255, out.double() / 255)[0])
            assert abs(r - m["rmse_out"]) < 1e-12, (n, r, m["rmse_out"])
            assert abs(q - m["ssim_out"]) < 1e-10, (n, q, m["ssim_out"])
            assert abs(q64 - m["ssim_out"]) < 1e-10


def test_quantise_matches_driver_rule():
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-0.2, 1.2, 100000), np.arange(256) / 255.0, [0.0, 1.0, 0.5]]).astype(np.float32)
    got = engine.quantise_u8(torch.from_numpy(a).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, orc.quantise_output(a))
    h = a.astype(np.float16)
    got_h = engine.quantise_u8(torch.from_numpy(a).cuda(), via_half=True).cpu().numpy()
    np.testing.assert_array_equal(got_h, orc.quantise_output(h))


def test_gpu_preprocessing_matches_cv2_fixtures(golden_dir):
    """BGR->gray (cv2.imread(path, 0) arithmetic), /255 normalisation, bicubic pre-upsampling
    (cv2.resize INTER_CUBIC on float32) against fixtures generated with cv2 (oracle/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "preproc.npz"))
    gray = engine.bgr_to_gray_u8(torch.from_numpy(g["bgr"]).cuda())
    np.testing.assert_array_equal(gray.cpu().numpy(), g["gray"])
    np.testing.assert_array_equal(engine.bgr_to_gray_u8(torch.from_numpy(g["bgr"]).cuda(), "cvtcolor").cpu().numpy(),
                                  g["gray_cvtcolor"])
    unit = engine.u8_to_unit_f32(gray)
    np.testing.assert_array_equal(unit.cpu().numpy(), (g["gray"] / 255).astype(np.float32))
    for s in (4, 8, 16):
        lr, up = g[f"lr_x{s}"], g[f"up_x{s}"]
        got = engine.bicubic_upsample(torch.from_numpy(lr).cuda()[None], up.shape[0], up.shape[1])[0].cpu().numpy()
        err = np.abs(got - up).max()
        print(f"bicubic x{s}: max err vs cv2 {err:.2e}")
        assert err <= 2e-6
