"""Synthetic workloads for benchmarks and tests: seeded weights with the reference's parameter
inventory and seeded RGB-D frame pairs in the reference's input domain.

Part of the product package (bench.py and the drivers must not import oracle/); the oracle keeps
its own identical generators and tests/test_synthetic.py asserts that the two agree bit for bit.
No arithmetic of the forward pass lives here -- numpy only builds inputs.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

StateDict = Dict[str, torch.Tensor]

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


FLOPS_PER_PIXEL = 14855552   # 2 * MAC of the 59 trunk convolutions per output pixel
CAC_BYTES_PER_PIXEL_PER_STAGE_PER_ELEM_BYTE = 512   # stats 128 + apply 384
