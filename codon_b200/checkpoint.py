"""Checkpoint import for the reference's ``X4.pth`` / ``X8.pth`` / ``X16.pth``.

The reference saves ``{"epoch": e, "model": <pickled nn.Module>}`` and reads it back with
``checkpoint["model"].state_dict()`` (CODON_X4/test.py:56-59; X16 through ``DataParallel``, so the
keys carry ``module.``, CODON_X16/test.py:52,58-60).  Unpickling a whole module needs its defining
classes importable under the module path recorded at training time (``base_net_withoutBN``,
``CODON_x4``, ``model.CODONet`` ... -- some of which cannot even be imported in the reference
checkout).  ``load_checkpoint`` therefore unpickles with a ``find_class`` that maps every class it
cannot resolve to a stub ``nn.Module`` and then walks ``_modules`` / ``_parameters`` / ``_buffers`` to
rebuild the flat state_dict.  Plain state_dict files and ``{"state_dict": ...}`` wrappers are
accepted too.  The .pth files are absent from the reference checkout (.MISSING_LARGE_BLOBS), so
this path is exercised by tests with synthetic checkpoints saved the reference's way.
"""
from __future__ import annotations

import collections
import importlib
import pickle
from typing import Dict, Tuple

import torch
import torch.nn as nn


class _StubModule(nn.Module):
    """Stands in for any class the checkpoint names that is not importable here."""

    def __init__(self, *a, **k):
        super().__init__()

    def __setstate__(self, state):
        self.__dict__.update(state)
        for k in ("_parameters", "_buffers", "_modules"):
            self.__dict__.setdefault(k, collections.OrderedDict())


# Modules a checkpoint may import from while being unpickled: torch's own tensor / parameter / container machinery
# and the numpy reconstructors torch.save can emit.  Every other class the pickle names -- the training-time model
# classes (``base_net_withoutBN...``, ``CODON_x4.CODONNet``, ``model.CODONet...``), and anything an attacker put there --
# is mapped to a stub nn.Module WITHOUT being imported, so that loading never executes code from the working
# directory or from arbitrary installed packages.  (A whole-module .pth is still a pickle: only load files you trust.)
_ALLOWED_PREFIXES = ("torch", "collections", "numpy", "_codecs", "builtins", "copyreg")
_ALLOWED_BUILTINS = {"set", "frozenset", "dict", "list", "tuple", "int", "float", "bool", "str", "bytes", "bytearray",
                     "complex", "slice", "range", "object", "getattr"}


def _import_allowed(module: str, name: str) -> bool:
    root = module.split(".", 1)[0]
    if root not in _ALLOWED_PREFIXES:
        return False
    if root == "builtins":
        return name in _ALLOWED_BUILTINS and name != "getattr"
    return True


class _LenientUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if _import_allowed(module, name):
            try:
                mod = importlib.import_module(module)
                obj = mod
                for part in name.split("."):
                    obj = getattr(obj, part)
                return obj
            except Exception:
                pass
        return type(name.rsplit(".", 1)[-1], (_StubModule,), {"__module__": module})


class _LenientPickle:
    """`pickle_module` shim for torch.load."""
    __name__ = "pickle"
    Unpickler = _LenientUnpickler
    load = staticmethod(pickle.load)
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError


def _walk(module, prefix: str, out: Dict[str, torch.Tensor]) -> None:
    for k, p in getattr(module, "_parameters", {}).items():
        if p is not None:
            out[prefix + k] = p.data if hasattr(p, "data") else p
    for k, b in getattr(module, "_buffers", {}).items():
        if b is not None:
            out[prefix + k] = b
    for k, m in getattr(module, "_modules", {}).items():
        if m is not None:
            _walk(m, prefix + k + ".", out)


def extract_state_dict(obj) -> Dict[str, torch.Tensor]:
    """nn.Module (real or stub) / state_dict / wrapper dict -> flat state_dict without ``module.``."""
    if isinstance(obj, dict) and "model" in obj and not torch.is_tensor(obj["model"]):
        obj = obj["model"]
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        obj = obj["state_dict"]
    if isinstance(obj, dict):
        sd = {k: v for k, v in obj.items() if torch.is_tensor(v)}
    else:
        sd = {}
        _walk(obj, "", sd)
    if not sd:
        raise ValueError("no tensors found in the checkpoint")
    out = {}
    for k, v in sd.items():
        while k.startswith("module."):
            k = k[len("module."):]
        out[k] = v.detach().to("cpu", torch.float32)
    return out


def load_checkpoint(path: str) -> Tuple[Dict[str, torch.Tensor], dict]:
    """Returns (state_dict, meta) where meta carries e.g. ``epoch`` (test.py:58)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_LenientPickle)
    meta = {k: v for k, v in ckpt.items() if isinstance(v, (int, float, str))} if isinstance(ckpt, dict) else {}
    return extract_state_dict(ckpt), meta


def infer_scale(sd: Dict[str, torch.Tensor]) -> str:
    """'x4/x8' checkpoints carry attention_c5 / attention_s5 (CODON_x4.py:64-65); x16 ones do not."""
    return "x4/x8" if any(k.startswith("attention_c5.") for k in sd) else "x16"


# ---- flat weight file: the C-host route --------------------------------------------------------
_FLAT_MAGIC = b"CODONW1\0"


def export_flat(sd: Dict[str, torch.Tensor], path: str) -> int:
    """Writes a state_dict as the flat little-endian file ``codon_load_weights_file`` (include/codon_b200.h) reads:
    ``"CODONW1\\0" | uint32 n | n x { uint16 name_len | name | uint8 ndim | ndim x int64 dims | float32 data }``.
    A leading ``module.`` is stripped.  Returns the number of tensors written."""
    import struct
    items = []
    for k, v in sd.items():
        while k.startswith("module."):
            k = k[len("module."):]
        items.append((k, v.detach().to("cpu", torch.float32).contiguous()))
    with open(path, "wb") as f:
        f.write(_FLAT_MAGIC)
        f.write(struct.pack("<I", len(items)))
        for k, t in items:
            name = k.encode("utf-8")
            if not (0 < len(name) < 256) or not (1 <= t.dim() <= 4):
                raise ValueError(f"cannot export {k!r} with shape {tuple(t.shape)}")
            f.write(struct.pack("<H", len(name)))
            f.write(name)
            f.write(struct.pack("<B", t.dim()))
            f.write(struct.pack(f"<{t.dim()}q", *t.shape))
            f.write(t.numpy().astype("<f4", copy=False).tobytes())
    return len(items)


def load_flat(path: str) -> Dict[str, torch.Tensor]:
    """Reads a file written by ``export_flat`` back into a state_dict (the Python twin of the C loader)."""
    import struct
    import numpy as np
    out: Dict[str, torch.Tensor] = {}
    with open(path, "rb") as f:
        if f.read(8) != _FLAT_MAGIC:
            raise ValueError(f"{path} is not a CODONW1 weight file")
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (ln,) = struct.unpack("<H", f.read(2))
            name = f.read(ln).decode("utf-8")
            (nd,) = struct.unpack("<B", f.read(1))
            dims = struct.unpack(f"<{nd}q", f.read(8 * nd))
            cnt = 1
            for d in dims:
                cnt *= d
            buf = f.read(4 * cnt)
            if len(buf) != 4 * cnt:
                raise ValueError(f"{path}: truncated data of {name!r}")
            out[name] = torch.from_numpy(np.frombuffer(buf, dtype="<f4").reshape(dims).copy())
    return out
