"""CPU oracle for the CODON guided depth super-resolution forward pass.

TEST INFRASTRUCTURE ONLY.  This module is the checker for the CUDA engine in
``codon_b200``; it is imported by ``tests/``, by ``__graft_entry__.smoke()`` and by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and by nothing else.  The
product path never routes through it (see ``codon_b200/engine.py``: a missing CUDA
library is a hard error).

It is a *functional restatement* of the reference forward, written against a plain
``state_dict`` (name -> tensor) instead of ``nn.Module`` objects, with the reference
file:line each block follows.  All ``path:line`` citations are relative to
``/root/reference``.

Parity pin: ``oracle/make_golden.py`` imports the real reference classes
(``CODON_X4/CODON_x4.py``, ``CODON_X16/CODON_x16.py``, ``CAC_module.py``,
``attention/ResCBAM.py``, ``ssim_2.py``, ``test.py:148-164``) in the build container,
runs them on the synthetic weights/frames defined here and on the bundled Middlebury
images, and commits their outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this restatement against those outputs, so the
oracle is pinned to the reference on synthetic weights.  Real-weight parity (X4.pth,
X8.pth, X16.pth are absent from the reference checkout, ``.MISSING_LARGE_BLOBS:1-3``)
is UNPINNED.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------
# Parameter inventory (CODON_X4/CODON_x4.py:24-47, 54-65; CODON_X16/CODON_x16.py:98-134)
# --------------------------------------------------------------------------------------

#: trunk convolutions: name -> (Cout, Cin, k); all bias-free (CODON_x4.py:24-47)
TRUNK_CONVS = {
    "input": (64, 1, 3), "conv_input": (64, 64, 3),
    "conv1": (64, 64, 3), "conv2": (64, 64, 5), "conv3": (128, 128, 5), "confuse": (64, 128, 1),
    "input_c": (64, 1, 3), "conv_input_c": (64, 64, 3),
    "conv4": (64, 64, 5), "conv5": (64, 64, 3), "conv6": (128, 128, 5), "confuse_c": (64, 128, 1),
    "conv7": (64, 128, 3),
    "conv8": (64, 64, 5), "conv9": (64, 64, 3), "conv10": (128, 128, 5), "confuse_fuse": (64, 128, 1),
    "conv11": (64, 64, 3),
    "output": (1, 64, 3),
}
N_CAC_STAGES = 5       # CODON_x4.py:74
N_FUSE_STAGES = 3      # CODON_x4.py:122


def param_shapes(scale: int) -> Dict[str, Tuple[int, ...]]:
    """Every state_dict key and shape of ``CODONNet`` for x4/x8 (49 keys) or x16 (44 keys).

    x4/x8 carry the never-called ``attention_c5`` (``ChannelGate(64)``, reduction 16 ->
    hidden 4, ResCBAM.py:26-36) and ``attention_s5`` (CODON_x4.py:64-65); x16 does not
    (CODON_x16.py:124-134).
    """
    shapes: Dict[str, Tuple[int, ...]] = {}
    for name, (co, ci, k) in TRUNK_CONVS.items():
        shapes[f"{name}.weight"] = (co, ci, k, k)
    for s in range(N_CAC_STAGES):
        shapes[f"attention_c{s}.mlp.1.weight"] = (8, 128)     # CAC_module.py:31 (128 // 16)
        shapes[f"attention_c{s}.mlp.1.bias"] = (8,)
        shapes[f"attention_c{s}.mlp.3.weight"] = (64, 8)      # CAC_module.py:33 (128 // 2)
        shapes[f"attention_c{s}.mlp.3.bias"] = (64,)
        shapes[f"attention_s{s}.spatial.conv.weight"] = (1, 2, 5, 5)   # CAC_module.py:88
    if scale in (4, 8):
        shapes["attention_c5.mlp.1.weight"] = (4, 64)
        shapes["attention_c5.mlp.1.bias"] = (4,)
        shapes["attention_c5.mlp.3.weight"] = (64, 4)
        shapes["attention_c5.mlp.3.bias"] = (64,)
        shapes["attention_s5.spatial.conv.weight"] = (1, 2, 5, 5)
    elif scale != 16:
        raise ValueError(f"scale must be 4, 8 or 16, got {scale}")
    return shapes


def synthetic_state_dict(scale: int, seed: int, output_gain: float = 0.002) -> StateDict:
    """Seeded synthetic weights (the real .pth files are not available).

    Distributions follow the reference initialisation -- trunk convs N(0, sqrt(2/(k*k*Cout)))
    (CODON_x4.py:50-53); CAC Linear / 2->1 conv U(+-1/sqrt(fan_in)) (torch defaults, the CAC
    modules are created after the init loop, CODON_x4.py:54-65) -- but are drawn from a
    numpy ``default_rng`` in sorted-key order so that they are identical on every machine
    and independent of torch's RNG stream.  ``output.weight`` is scaled by ``output_gain``
    so that the predicted residual stays inside [0, 1] depth.
    """
    rng = np.random.default_rng(1000003 * scale + seed)
    sd: StateDict = {}
    shapes = param_shapes(scale)
    for key in sorted(shapes):
        shp = shapes[key]
        base = key.split(".")[0]
        if base in TRUNK_CONVS:
            co, _, k = TRUNK_CONVS[base]
            arr = rng.normal(0.0, math.sqrt(2.0 / (k * k * co)), size=shp)
            if base == "output":
                arr = arr * output_gain
        else:
            if key.endswith("spatial.conv.weight"):
                fan_in = 2 * 5 * 5
            elif key.endswith("mlp.1.weight") or key.endswith("mlp.1.bias"):
                fan_in = shapes[key.rsplit(".", 1)[0] + ".weight"][1]
            else:  # mlp.3.*
                fan_in = shapes[key.rsplit(".", 1)[0] + ".weight"][1]
            bound = 1.0 / math.sqrt(fan_in)
            arr = rng.uniform(-bound, bound, size=shp)
        sd[key] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return sd


def synthetic_frames(batch: int, height: int, width: int, seed: int = 1234
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Synthetic RGB-D pair in the reference's input domain.

    Both tensors are [B,1,H,W] float32 with values k/255 (test.py:116,122 reads uint8 and
    divides by 255).  Depth is a blurred piece-wise smooth scene (stands for the bicubic
    pre-upsampled LR depth, test.py:77), guide is an edge-aligned gray texture.
    """
    depth = np.empty((batch, 1, height, width), np.float32)
    guide = np.empty((batch, 1, height, width), np.float32)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    for b in range(batch):
        rng = np.random.default_rng(seed + b)
        gt = 0.1 + 0.3 * (xx / max(width - 1, 1)) * rng.uniform(0.3, 1.0) \
                 + 0.2 * (yy / max(height - 1, 1)) * rng.uniform(0.3, 1.0)
        for _ in range(6):
            cy, cx = rng.uniform(0, height), rng.uniform(0, width)
            ry, rx = rng.uniform(0.08, 0.35) * height, rng.uniform(0.08, 0.35) * width
            inside = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
            gt = np.where(inside, rng.uniform(0.15, 0.9), gt)
        gt = np.clip(gt, 0.0, 1.0)
        # low-pass (5-tap box, twice) = the loss of detail of an upsampled LR depth map
        lr = gt.copy()
        for _ in range(2):
            p = np.pad(lr, 2, mode="edge")
            lr = sum(p[i:i + height, 2:2 + width] for i in range(5)) / 5.0
            p = np.pad(lr, 2, mode="edge")
            lr = sum(p[2:2 + height, i:i + width] for i in range(5)) / 5.0
        tex = rng.uniform(0.0, 1.0, size=(height, width))
        p = np.pad(tex, 1, mode="edge")
        tex = sum(p[i:i + height, j:j + width] for i in range(3) for j in range(3)) / 9.0
        g = 0.5 * gt + 0.5 * tex
        depth[b, 0] = np.floor(np.clip(lr, 0, 1) * 255.0) / 255.0
        guide[b, 0] = np.floor(np.clip(g, 0, 1) * 255.0) / 255.0
    return torch.from_numpy(depth), torch.from_numpy(guide)


# --------------------------------------------------------------------------------------
# CAC blocks (CODON_X4/CAC_module.py)
# --------------------------------------------------------------------------------------

def _mlp(sd: StateDict, prefix: str, v: torch.Tensor) -> torch.Tensor:
    """Flatten -> Linear -> ReLU -> Linear (CAC_module.py:29-35)."""
    h = F.relu(F.linear(v.flatten(1), sd[f"{prefix}.mlp.1.weight"], sd[f"{prefix}.mlp.1.bias"]))
    return F.linear(h, sd[f"{prefix}.mlp.3.weight"], sd[f"{prefix}.mlp.3.bias"])


def cac_channel_scale(sd: StateDict, prefix: str, fcat: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,C/2] channel scale (CAC_module.py:38-63).

    Global average and global max pool per channel (:43, :47), the shared MLP applied to
    each and summed -- so the second-layer bias enters twice (:44, :48, :59-61) -- sigmoid
    (:62).  The reference then expands the result to [B,C/2,H,W]; callers here broadcast.
    """
    avg = fcat.mean(dim=(2, 3))
    mx = fcat.amax(dim=(2, 3))
    return torch.sigmoid(_mlp(sd, prefix, avg) + _mlp(sd, prefix, mx))


def channel_pool(fcat: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,2,H,W]: channel 0 = max over C, channel 1 = mean over C (CAC_module.py:78-81)."""
    return torch.stack((fcat.amax(dim=1), fcat.mean(dim=1)), dim=1)


def cac_spatial_scale(sd: StateDict, prefix: str, fcat: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,1,H,W] spatial scale (CAC_module.py:83-94): pooled map -> 5x5 conv
    2->1, zero padding 2, no bias, no ReLU (:88) -> sigmoid (:93)."""
    q = F.conv2d(channel_pool(fcat), sd[f"{prefix}.spatial.conv.weight"], padding=2)
    return torch.sigmoid(q)


def channel_gate(sd: StateDict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Classic CBAM channel gate, returns x*scale (attention/ResCBAM.py:38-61)."""
    avg = x.mean(dim=(2, 3))
    mx = x.amax(dim=(2, 3))
    s = torch.sigmoid(_mlp(sd, prefix, avg) + _mlp(sd, prefix, mx))
    return x * s[:, :, None, None]


def spatial_gate(sd: StateDict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Classic CBAM spatial gate, returns x*scale (attention/ResCBAM.py:75-87)."""
    return x * cac_spatial_scale(sd, prefix, x)


# --------------------------------------------------------------------------------------
# The network (CODON_X4/CODON_x4.py:66-132 == CODON_X8; CODON_X16/CODON_x16.py:136-202)
# --------------------------------------------------------------------------------------

def _conv(sd: StateDict, name: str, x: torch.Tensor, relu: bool = True) -> torch.Tensor:
    w = sd[f"{name}.weight"]
    y = F.conv2d(x, w, padding=w.shape[-1] // 2)
    return F.relu(y) if relu else y


def forward(sd: StateDict, x: torch.Tensor, y: torch.Tensor, return_taps: bool = False):
    """``CODONNet.forward(x_depth, y_gray)`` restated (CODON_x4.py:66-132).

    x, y: [B,1,H,W]; dtype is taken from ``x`` (float32 or float64); ``sd`` tensors are
    cast to it.  With ``return_taps`` also returns a dict of intermediate tensors used by
    the per-layer parity tests.
    """
    dt = x.dtype
    sd = {k: v.to(dt) for k, v in sd.items()}
    taps = {}
    # encoders (CODON_x4.py:68-73): their outputs are the residual carriers of all 5 stages
    enc_d = _conv(sd, "conv_input", _conv(sd, "input", x))
    enc_c = _conv(sd, "conv_input_c", _conv(sd, "input_c", y))
    out_d, out_c = enc_d, enc_c
    taps["enc_d"], taps["enc_c"] = enc_d, enc_c
    for s in range(N_CAC_STAGES):                                     # CODON_x4.py:74
        # depth branch: [3x3 | 5x5] (:75,77,79); colour branch: [5x5 | 3x3] (:76,78,80)
        ms_d = torch.cat((_conv(sd, "conv1", out_d), _conv(sd, "conv2", out_d)), 1)
        ms_c = torch.cat((_conv(sd, "conv4", out_c), _conv(sd, "conv5", out_c)), 1)
        r2_d = _conv(sd, "conv3", ms_d)                               # :81
        r2_c = _conv(sd, "conv6", ms_c)                               # :82
        out_c = _conv(sd, "confuse_c", r2_c, relu=False)              # :83
        out_d = _conv(sd, "confuse", r2_d, relu=False)                # :84
        fcat = torch.cat((out_c, out_d), 1)                           # :85  colour first
        s_c = cac_channel_scale(sd, f"attention_c{s}", fcat)          # :86-115
        s_s = cac_spatial_scale(sd, f"attention_s{s}", fcat)
        gate = s_c[:, :, None, None] * s_s                            # :88 etc.  [B,64,H,W]
        if s == 0:
            taps["ms_d0"], taps["ms_c0"] = ms_d, ms_c
            taps["fcat0"], taps["s_c0"], taps["s_s0"] = fcat, s_c, s_s
        out_d = out_d * gate + enc_d                                  # :89/:118
        out_c = out_c * gate + enc_c                                  # :91/:117
        taps[f"out_d{s}"], taps[f"out_c{s}"] = out_d, out_c
    fuse = _conv(sd, "conv7", torch.cat((out_d, out_c), 1))           # :119-120 depth first
    taps["fuse"] = fuse
    of = fuse
    for _ in range(N_FUSE_STAGES):                                    # :122
        ms = torch.cat((_conv(sd, "conv8", of), _conv(sd, "conv9", of)), 1)       # :123-125 [5x5 | 3x3]
        of = _conv(sd, "confuse_fuse", _conv(sd, "conv10", ms), relu=False) + fuse  # :126-128
    taps["out_fuse"] = of
    res = _conv(sd, "output", _conv(sd, "conv11", of), relu=False) + x            # :129-131
    return (res, taps) if return_taps else res


# --------------------------------------------------------------------------------------
# Driver semantics and metrics (CODON_X4/test.py:116-164, CODON_X4/ssim_2.py:36-52)
# --------------------------------------------------------------------------------------

def quantise_output(out: np.ndarray) -> np.ndarray:
    """clip to [0,1], *255, truncate to uint8 (test.py:130,132), evaluated in out's dtype."""
    return (np.clip(out, 0, 1) * 255).astype(np.uint8)


def masked_rmse(label_u8: np.ndarray, out_u8: np.ndarray) -> float:
    """``EvaluationResults`` (test.py:148-164): RMSE in 0..255 units over pixels whose label
    is non-zero, label cropped to the output's shape, float64."""
    lab = label_u8.astype(np.float64)[:out_u8.shape[0], :out_u8.shape[1]]
    o = out_u8.astype(np.float64)
    valid = lab != 0
    n = int(valid.sum())
    e = np.where(valid, lab - o, 0.0)
    return math.sqrt(float((e ** 2).sum()) / n)


def ssim_gauss(img1: np.ndarray, img2: np.ndarray, sd: float = 1.5,
               c1: float = 0.01 ** 2, c2: float = 0.03 ** 2) -> float:
    """``ssim_exact`` (ssim_2.py:36-52): Gaussian-window SSIM, sigma 1.5, scipy defaults
    (truncate 4.0 -> 13 taps, 'reflect' boundary), mean of the SSIM map."""
    from scipy.ndimage import gaussian_filter
    a = np.asarray(img1, np.float64)
    b = np.asarray(img2, np.float64)
    mu1, mu2 = gaussian_filter(a, sd), gaussian_filter(b, sd)
    s11 = gaussian_filter(a * a, sd) - mu1 * mu1
    s22 = gaussian_filter(b * b, sd) - mu2 * mu2
    s12 = gaussian_filter(a * b, sd) - mu1 * mu2
    num = (2 * mu1 * mu2 + c1) * (2 * s12 + c2)
    den = (mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2)
    return float(np.mean(num / den))


def flops_per_pixel() -> int:
    """2*MAC per output pixel of the trunk convolutions: 14,855,552."""
    calls = {"input": 1, "conv_input": 1, "input_c": 1, "conv_input_c": 1,
             "conv1": 5, "conv2": 5, "conv3": 5, "confuse": 5,
             "conv4": 5, "conv5": 5, "conv6": 5, "confuse_c": 5,
             "conv7": 1, "conv8": 3, "conv9": 3, "conv10": 3, "confuse_fuse": 3,
             "conv11": 1, "output": 1}
    return 2 * sum(co * ci * k * k * calls[n] for n, (co, ci, k) in TRUNK_CONVS.items())
