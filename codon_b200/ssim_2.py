"""Drop-in for the reference ``ssim_2`` (CODON_X4/ssim_2.py, identical in X8/X16).

``ssim_exact(img1, img2, sd=1.5, C1=0.01**2, C2=0.03**2)`` is the function the driver calls
(test.py:139): Gaussian-window SSIM with scipy.ndimage.gaussian_filter semantics, float64.  Here it
runs in the libcodon_b200 kernels (codon_ssim_gauss); inputs may be numpy arrays (copied to the
current CUDA device) or CUDA tensors.  The reference's ``ssim`` / ``block_view`` are broken on
Python 3 (a float lands in an array shape, ssim_2.py:15) and unused; they are not provided.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _eng


def ssim_exact(img1, img2, sd=1.5, C1=0.01 ** 2, C2=0.03 ** 2):
    a = torch.as_tensor(np.ascontiguousarray(img1) if isinstance(img1, np.ndarray) else img1)
    b = torch.as_tensor(np.ascontiguousarray(img2) if isinstance(img2, np.ndarray) else img2)
    if not torch.cuda.is_available():
        raise _eng.CodonError("ssim_exact runs on the GPU; no CUDA device is available (no CPU path)")
    a, b = a.cuda(), b.cuda()
    if a.dim() != 2 or a.shape != b.shape:
        raise _eng.CodonError(f"ssim_exact needs two 2-D images of one shape, got {tuple(a.shape)} and {tuple(b.shape)}")
    return float(_eng.ssim_gauss(a[None], b[None], sd, C1, C2)[0])
