"""Drop-in for the reference ``attention.ResCBAM`` (CODON_X4/attention/ResCBAM.py == CODON_X8).

CODONNet only instantiates ``ChannelGate(64)`` as ``attention_c5`` and never calls it
(CODON_X4/CODON_x4.py:64); the classes are provided with the reference's names, arguments,
state_dict keys and semantics (gates return ``x * scale``, ResCBAM.py:61,87; the ResCBAM wrappers
add the input back, :102-106), computed by the libcodon_b200 kernels.
"""
from __future__ import annotations

import torch.nn as nn

from .. import engine as _eng
from ..CAC_module import BasicConv, ChannelPool, Flatten, _gate_params, logsumexp_2d  # noqa: F401


class ChannelGate(nn.Module):
    """x * sigmoid(sum_pool mlp(pool(x))), MLP C -> C/r -> C (ResCBAM.py:26-61)."""

    def __init__(self, gate_channels, reduction_ratio=16, pool_types=['avg', 'max']):
        super().__init__()
        self.gate_channels = gate_channels
        self.mlp = nn.Sequential(
            Flatten(),
            nn.Linear(gate_channels, gate_channels // reduction_ratio),
            nn.ReLU(),
            nn.Linear(gate_channels // reduction_ratio, gate_channels),
        )
        self.pool_types = pool_types

    def forward(self, x):
        s = _eng.cac_channel_scale(x, *_gate_params(self.mlp), pool_types=self.pool_types)
        return _eng.cac_apply(x, sc=s).to(x.dtype)


class SpatialGate(nn.Module):
    """x * sigmoid(conv5x5(ChannelPool(x))) (ResCBAM.py:75-87)."""

    def __init__(self):
        super().__init__()
        kernel_size = 5
        self.compress = ChannelPool()
        self.spatial = BasicConv(2, 1, kernel_size, stride=1, padding=(kernel_size - 1) // 2, relu=False)

    def forward(self, x):
        s = _eng.cac_spatial_scale(x, self.spatial.conv.weight)
        return _eng.cac_apply(x, ss=s).to(x.dtype)


class _ResCBAMBase(nn.Module):
    def __init__(self, gate_channels, reduction_ratio, pool_types, no_spatial):
        super().__init__()
        self.ChannelGate = ChannelGate(gate_channels, reduction_ratio, pool_types)
        self.no_spatial = no_spatial
        if not no_spatial:
            self.SpatialGate = SpatialGate()

    def forward(self, x):
        sc = _eng.cac_channel_scale(x, *_gate_params(self.ChannelGate.mlp), pool_types=self.ChannelGate.pool_types)
        if self.no_spatial:
            return _eng.cac_apply(x, sc=sc, res=x).to(x.dtype)
        xc = _eng.cac_apply(x, sc=sc)
        ss = _eng.cac_spatial_scale(xc, self.SpatialGate.spatial.conv.weight)
        return _eng.cac_apply(xc, ss=ss, res=x).to(x.dtype)


class ResCBAM(_ResCBAMBase):
    """channel gate -> spatial gate -> + x (ResCBAM.py:94-106)."""

    def __init__(self, gate_channels, reduction_ratio=8, pool_types=['avg', 'max'], no_spatial=False):
        super().__init__(gate_channels, reduction_ratio, pool_types, no_spatial)


class ResCBAM_c(_ResCBAMBase):
    """avg-pool-only variant (ResCBAM.py:108-120)."""

    def __init__(self, gate_channels, reduction_ratio=8, pool_types=['avg'], no_spatial=False):
        super().__init__(gate_channels, reduction_ratio, pool_types, no_spatial)


class ResCBAM_d(_ResCBAMBase):
    """max-pool-only variant (ResCBAM.py:122-134)."""

    def __init__(self, gate_channels, reduction_ratio=8, pool_types=['max'], no_spatial=False):
        super().__init__(gate_channels, reduction_ratio, pool_types, no_spatial)
