"""``CODONNet`` with the reference's module surface, computed by libcodon_b200.

Mirrors CODON_X4/CODON_x4.py:18-132 (== CODON_X8/CODON_x8.py) and CODON_X16/CODON_x16.py:92-202:
the constructor takes no arguments; the sub-modules carry the reference's attribute names so that
``state_dict()`` / ``load_state_dict()`` exchange the same 49 (x4/x8) or 44 (x16) keys; the
reference idioms ``.cuda()``, ``.half()``, ``.eval()``, ``torch.nn.DataParallel(model)`` all work.
The ``nn.Conv2d`` / ``nn.Linear`` objects are *parameter holders only*: ``forward`` hands the two
frames to the engine (one C-ABI call), and no PyTorch operator touches the data.

Arithmetic mode follows the parameter dtype, as in the reference (test.py:52 ``.half()``):
  float32 parameters  -> "fp32"  (fp32 FFMA trunk; parity mode)
  float16 parameters  -> "fp16"  (tcgen05 kind::f16, fp16 operands, fp32 accumulation)
  bfloat16 parameters -> "bf16"  (tcgen05 kind::f16, bf16 operands, fp32 accumulation)
``set_mode("tf32" | ...)`` overrides it.
"""
from __future__ import annotations

import threading
from math import sqrt
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import engine as _eng
from .CAC_module import CAC_channel, CAC_spatial
from .attention.ResCBAM import ChannelGate

# name -> (in_channels, out_channels, kernel): CODON_X4/CODON_x4.py:24-47, declaration order
_TRUNK: Tuple[Tuple[str, int, int, int], ...] = (
    ("input", 1, 64, 3), ("conv_input", 64, 64, 3), ("conv1", 64, 64, 3), ("conv2", 64, 64, 5),
    ("conv3", 128, 128, 5), ("confuse", 128, 64, 1),
    ("input_c", 1, 64, 3), ("conv_input_c", 64, 64, 3), ("conv4", 64, 64, 5), ("conv5", 64, 64, 3),
    ("conv6", 128, 128, 5), ("confuse_c", 128, 64, 1),
    ("conv7", 128, 64, 3), ("conv8", 64, 64, 5), ("conv9", 64, 64, 3), ("conv10", 128, 128, 5),
    ("confuse_fuse", 128, 64, 1), ("conv11", 64, 64, 3), ("output", 64, 1, 3),
)
_DTYPE_MODE = {torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}


class CODONNetBase(nn.Module):
    """Shared implementation; ``SCALE`` selects the parameter set (x4/x8 carry attention_c5/s5)."""

    SCALE = 4

    def __init__(self):
        super().__init__()
        for name, cin, cout, k in _TRUNK:
            setattr(self, name, nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=k // 2, bias=False))
        self.relu = nn.ReLU()
        # He-normal with fan = k*k*Cout for the trunk convs only (CODON_x4.py:50-53); the attention
        # modules are created afterwards and keep torch's default init (:54-65)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, sqrt(2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
        for k in range(5):
            setattr(self, f"attention_c{k}", CAC_channel(128))
        for k in range(5):
            setattr(self, f"attention_s{k}", CAC_spatial())
        if self.SCALE in (4, 8):
            self.attention_c5 = ChannelGate(64)      # declared, never called (CODON_x4.py:64-65)
            self.attention_s5 = CAC_spatial()
        self._mode_override: Optional[str] = None
        self._engines: Dict[Tuple[int, str], Tuple[_eng.Engine, tuple]] = {}
        self._engine_lock = threading.Lock()          # DataParallel replicas share _engines across threads
        self._wver = 0
        self._plist = None

    # ---- engine management --------------------------------------------------------------------
    def set_mode(self, mode: Optional[str]) -> "CODONNetBase":
        if mode is not None and mode not in _eng.MODES:
            raise ValueError(f"mode must be one of {sorted(_eng.MODES)} or None")
        self._mode_override = mode
        return self

    @property
    def mode(self) -> str:
        if self._mode_override:
            return self._mode_override
        return _DTYPE_MODE.get(self.input.weight.dtype, "fp32")

    # Weight changes are tracked by a version counter that every nn.Module-level mutation bumps (load_state_dict, and
    # _apply = .cuda() / .half() / .to() / .float()), plus the autograd version counters of the parameters (in-place
    # ops through the Parameter itself: p.copy_(), p.mul_()).  In-place edits through ``.data`` (p.data.normal_(), the
    # reference's own init idiom, CODON_x4.py:53) bump neither: call ``refresh_weights()`` after such an edit.
    def load_state_dict(self, *args, **kwargs):
        res = super().load_state_dict(*args, **kwargs)
        self._wver += 1
        return res

    def _apply(self, fn, *args, **kwargs):
        res = super()._apply(fn, *args, **kwargs)
        self._wver += 1
        self._plist = None
        return res

    def refresh_weights(self) -> "CODONNetBase":
        """Forces the engines to re-read the parameters on the next forward (needed only after in-place edits through
        ``param.data``, which PyTorch does not version)."""
        self._wver += 1
        return self

    invalidate = refresh_weights

    def _weights_key(self) -> tuple:
        if getattr(self, "_is_replica", False):
            # torch.nn.DataParallel replica: its parameters are fresh broadcast copies on every forward; the master's
            # counter (copied at replication time) is the only meaningful version
            return (self._wver, None)
        if self._plist is None:
            self._plist = list(self.parameters())
        return (self._wver, tuple((p.data_ptr(), p._version) for p in self._plist))

    def _current_weights(self) -> Dict[str, torch.Tensor]:
        """name -> tensor of this module tree.  A DataParallel replica keeps its broadcast copies in
        ``_former_parameters`` (its ``_parameters`` are empty, so ``state_dict()`` would be too)."""
        if not getattr(self, "_is_replica", False):
            return dict(self.state_dict())
        out: Dict[str, torch.Tensor] = {}
        for prefix, m in self.named_modules():
            held = dict(getattr(m, "_former_parameters", None) or {})
            held.update({k: v for k, v in m._parameters.items() if v is not None})
            for k, v in held.items():
                out[f"{prefix}.{k}" if prefix else k] = v
        return out

    def engine(self, device: torch.device) -> _eng.Engine:
        """The (cached) engine for this device/mode, with the current parameter values uploaded."""
        mode = self.mode
        key = (device.index if device.index is not None else torch.cuda.current_device(), mode)
        wkey = self._weights_key()
        hit = self._engines.get(key)
        if hit is not None and hit[1][0] == wkey[0] and (wkey[1] is None or hit[1][1] is None or hit[1][1] == wkey[1]):
            return hit[0]
        with self._engine_lock:
            hit = self._engines.get(key)
            eng = hit[0] if hit is not None else _eng.Engine(self.SCALE, mode, key[0])
            eng.load_state_dict(self._current_weights())
            if wkey[1] is None and hit is not None:
                wkey = (wkey[0], hit[1][1])
            self._engines[key] = (eng, wkey)
        return eng

    # ---- the hot path -------------------------------------------------------------------------
    def forward(self, x, y):
        """x: depth [B,1,H,W], y: gray guide [B,1,H,W] -> [B,1,H,W] in x's dtype (CODON_x4.py:66-132)."""
        if not x.is_cuda:
            raise _eng.CodonError("CODONNet.forward needs CUDA tensors: codon_b200 has no CPU path "
                                  "(the reference's CPU forward lives in oracle/ as a test checker only)")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise _eng.CodonError("codon_b200 is an inference engine: call model.eval() / torch.no_grad() "
                                  "(the reference test.py does, CODON_X4/test.py:67)")
        return self.engine(x.device).forward(x, y.to(x.dtype))

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engines"] = {}
        st["_engine_lock"] = None
        st["_plist"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._engine_lock = threading.Lock()
