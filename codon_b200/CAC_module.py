"""Drop-in for the reference ``CAC_module`` (CODON_X4/CAC_module.py == CODON_X8; CODON_X16/CAC_module.py).

Same class names, constructor arguments, sub-module names (hence state_dict keys) and return
conventions as the reference -- ``CAC_channel`` and ``CAC_spatial`` return a *scale*, not a gated
tensor (CAC_module.py:62-63, 93-94) -- but ``forward`` runs the sm_100a kernels of libcodon_b200
through the C ABI.  The modules only hold parameters; there is no PyTorch arithmetic and no CPU
path (a CPU tensor raises ``CodonError``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine as _eng


def _fold_bn(weight, bias, bn):
    """Eval-mode BatchNorm folded into the convolution's weight/bias (parameter preparation)."""
    if bn is None:
        return weight, bias
    if bn.training:
        raise _eng.CodonError("BasicConv(bn=True) is inference-only here: call .eval() first")
    inv = (bn.running_var + bn.eps).rsqrt()
    g = bn.weight if bn.weight is not None else torch.ones_like(inv)
    b = bn.bias if bn.bias is not None else torch.zeros_like(inv)
    w = weight * (g * inv).view(-1, 1, 1, 1)
    base = bias if bias is not None else torch.zeros_like(inv)
    return w, (base - bn.running_mean) * g * inv + b


class BasicConv(nn.Module):
    """Conv2d (+ optional eval-mode BatchNorm) (+ optional ReLU): CAC_module.py:6-20."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 relu=True, bn=False, bias=False):
        super().__init__()
        self.out_channels = out_planes
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_planes, eps=1e-5, momentum=0.01, affine=True) if bn else None
        self.relu = nn.ReLU() if relu else None

    def forward(self, x):
        c = self.conv
        w, b = _fold_bn(c.weight, c.bias, self.bn)
        y = _eng.conv2d_nchw(x, w, b, stride=c.stride, padding=c.padding, dilation=c.dilation, groups=c.groups,
                             relu=self.relu is not None)
        return y.to(x.dtype)


class Flatten(nn.Module):
    """[B, ...] -> [B, -1] (CAC_module.py:22-24); a view, no arithmetic."""

    def forward(self, x):
        return x.reshape(x.size(0), -1)


def _gate_params(mlp):
    return mlp[1].weight, mlp[1].bias, mlp[3].weight, mlp[3].bias


class CAC_channel(nn.Module):
    """Cross-domain channel gate: [B,C,H,W] -> sigmoid(sum_pool mlp(pool(x))) expanded to
    [B,C//2,H,W] (CAC_module.py:26-63)."""

    def __init__(self, gate_channels, reduction_ratio=16, pool_types=['avg', 'max']):
        super().__init__()
        self.gate_channels = gate_channels
        self.mlp = nn.Sequential(
            Flatten(),
            nn.Linear(gate_channels, gate_channels // reduction_ratio),
            nn.ReLU(),
            nn.Linear(gate_channels // reduction_ratio, gate_channels // 2),
        )
        self.pool_types = pool_types

    def forward(self, x):
        s = _eng.cac_channel_scale(x, *_gate_params(self.mlp), pool_types=self.pool_types)
        return s.to(x.dtype)[:, :, None, None].expand(x.shape[0], x.shape[1] // 2, x.shape[2], x.shape[3])


def logsumexp_2d(tensor):
    """[B,C,H,W] -> [B,C,1] log-sum-exp over the plane (CAC_module.py:71-76), on the GPU kernels."""
    st = _eng.channel_stats(tensor)
    return st[3].to(tensor.dtype).reshape(tensor.size(0), tensor.size(1), 1)


class ChannelPool(nn.Module):
    """[B,C,H,W] -> [B,2,H,W]: channel 0 = max over C, channel 1 = mean over C (CAC_module.py:78-81)."""

    def forward(self, x):
        return _eng.channel_pool(x).to(x.dtype)


class CAC_spatial(nn.Module):
    """Cross-domain spatial gate: sigmoid(conv5x5(ChannelPool(x))) -> [B,1,H,W] (CAC_module.py:83-94)."""

    def __init__(self):
        super().__init__()
        kernel_size = 5
        self.compress = ChannelPool()
        self.spatial = BasicConv(2, 1, kernel_size, stride=1, padding=(kernel_size - 1) // 2, relu=False)

    def forward(self, x):
        return _eng.cac_spatial_scale(x, self.spatial.conv.weight).to(x.dtype)
