"""ctypes binding of libcodon_b200.so (include/codon_b200.h) and the Engine wrapper.

PyTorch is plumbing here: it owns device memory (inputs, outputs, workspace) and streams; every
arithmetic operation of the forward runs in the hand-written sm_100a kernels behind the C ABI.
There is NO CPU path and NO fallback: if the library is missing or no CUDA device is present,
calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, Optional

import torch

from . import build as _build

MODES = {"fp32": 0, "bf16": 1, "fp16": 2, "tf32": 3, "f16x3": 4}
_IO_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}

# every symbol include/codon_b200.h declares (tests/test_abi.py checks the list against the header)
ABI_SYMBOLS = [
    "codon_create", "codon_destroy", "codon_last_error", "codon_version", "codon_selftest", "codon_set_weight",
    "codon_finalize_weights", "codon_weights_generation", "codon_load_weights_file", "codon_workspace_bytes", "codon_forward", "codon_forward_host",
    "codon_forward_host_submit", "codon_forward_host_wait",
    "codon_last_launch_count", "codon_debug_tap", "codon_profile_enable", "codon_profile_read",
    "codon_profile_reset", "codon_profile_category_name", "codon_cac_channel", "codon_cac_spatial",
    "codon_cac_apply", "codon_channel_stats", "codon_channel_pool", "codon_conv2d_nchw",
    "codon_masked_rmse", "codon_ssim_gauss", "codon_quantise_u8",
    "codon_bgr_to_gray_u8", "codon_u8_to_unit_f32", "codon_bicubic_upsample_f32",
    "codon_group_create", "codon_group_destroy", "codon_group_forward_host", "codon_group_last_error",
    "codon_group_last_ms",
]

_lib = None
_lib_lock = threading.Lock()


class CodonError(RuntimeError):
    pass


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Loads (once) the in-tree shared library and declares the argument types."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or _build.LIB_PATH
        if not os.path.exists(p):
            raise CodonError(
                f"{p} not found: build it with `python -m codon_b200.build` (or __graft_entry__.build()). "
                "codon_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(p)
        c = ctypes
        vp, ip, fp = c.c_void_p, c.c_int, c.POINTER(c.c_float)
        lib.codon_create.argtypes = [c.POINTER(vp), ip, ip, ip]
        lib.codon_destroy.argtypes = [vp]
        lib.codon_destroy.restype = None
        lib.codon_last_error.argtypes = [vp]
        lib.codon_last_error.restype = c.c_char_p
        lib.codon_selftest.argtypes = []
        lib.codon_selftest.restype = c.c_int
        lib.codon_version.argtypes = []
        lib.codon_version.restype = c.c_char_p
        lib.codon_set_weight.argtypes = [vp, c.c_char_p, fp, c.POINTER(c.c_int64), ip]
        lib.codon_finalize_weights.argtypes = [vp]
        lib.codon_workspace_bytes.argtypes = [vp, ip, ip, ip]
        lib.codon_workspace_bytes.restype = c.c_size_t
        lib.codon_weights_generation.argtypes = [vp]
        lib.codon_load_weights_file.argtypes = [vp, c.c_char_p]
        lib.codon_weights_generation.restype = c.c_ulonglong
        lib.codon_forward.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, vp, c.c_size_t, vp]
        lib.codon_forward_host.argtypes = [vp, vp, vp, vp, ip, ip, ip]
        lib.codon_forward_host_submit.argtypes = [vp, vp, vp, vp, ip, ip, ip]
        lib.codon_forward_host_wait.argtypes = [vp]
        lib.codon_last_launch_count.argtypes = [vp]
        lib.codon_debug_tap.argtypes = [vp, c.c_char_p, vp, c.POINTER(ip), vp]
        lib.codon_profile_enable.argtypes = [vp, ip]
        lib.codon_profile_read.argtypes = [vp, ip, c.POINTER(c.c_double), c.POINTER(c.c_double), c.POINTER(c.c_longlong)]
        lib.codon_profile_reset.argtypes = [vp]
        lib.codon_profile_category_name.argtypes = [ip]
        lib.codon_profile_category_name.restype = c.c_char_p
        lib.codon_cac_channel.argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, vp, ip, ip, ip, vp, vp]
        lib.codon_cac_spatial.argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, vp]
        lib.codon_cac_apply.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, ip, vp, vp]
        lib.codon_channel_stats.argtypes = [vp, ip, ip, ip, ip, vp, vp]
        lib.codon_channel_pool.argtypes = [vp, ip, ip, ip, ip, vp, vp]
        lib.codon_conv2d_nchw.argtypes = [vp, vp, vp] + [ip] * 15 + [vp, vp]
        lib.codon_masked_rmse.argtypes = [vp, vp, ip, ip, ip, vp, vp]
        lib.codon_ssim_gauss.argtypes = [vp, vp, ip, ip, ip, ip, c.c_double, c.c_double, c.c_double, vp, vp, c.c_size_t, vp]
        lib.codon_quantise_u8.argtypes = [vp, vp, c.c_size_t, ip, vp]
        lib.codon_bgr_to_gray_u8.argtypes = [vp, vp, c.c_size_t, ip, vp]
        lib.codon_u8_to_unit_f32.argtypes = [vp, vp, c.c_size_t, vp]
        lib.codon_bicubic_upsample_f32.argtypes = [vp, vp, ip, ip, ip, ip, ip, vp]
        lib.codon_group_create.argtypes = [c.POINTER(vp), c.POINTER(vp), ip]
        lib.codon_group_destroy.argtypes = [vp]
        lib.codon_group_destroy.restype = None
        lib.codon_group_forward_host.argtypes = [vp, vp, vp, vp, ip, ip]
        lib.codon_group_last_error.argtypes = [vp]
        lib.codon_group_last_error.restype = c.c_char_p
        lib.codon_group_last_ms.argtypes = [vp]
        lib.codon_group_last_ms.restype = c.c_double
        for name in ABI_SYMBOLS:
            fn = getattr(lib, name)
            if name not in ("codon_destroy", "codon_last_error", "codon_version", "codon_selftest", "codon_workspace_bytes",
                            "codon_weights_generation",
                            "codon_profile_category_name", "codon_group_destroy", "codon_group_last_error",
                            "codon_group_last_ms"):
                fn.restype = c.c_int
        if path is None:
            _lib = lib
        return lib


def check(rc: int, ctx=None) -> None:
    if rc != 0:
        lib = load_library()
        msg = lib.codon_last_error(ctx).decode("utf-8", "replace") if ctx else lib.codon_last_error(None).decode()
        raise CodonError(f"libcodon_b200 error {rc}: {msg}")


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise CodonError(f"{what} must be a CUDA tensor: codon_b200 has no CPU path (got device {t.device})")


def current_stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Engine:
    """One `codon_ctx`: the weights of one CODONNet on one GPU, in one arithmetic mode.

    Replaces ``CODONNet().cuda().half()`` + ``load_state_dict`` + ``model(depth, gray)``
    (CODON_X4/test.py:48-59,125).
    """

    def __init__(self, scale: int = 4, mode: str = "bf16", device: int | torch.device = 0):
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}, got {mode!r}")
        self.lib = load_library()
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        if dev.type != "cuda":
            raise CodonError("codon_b200.Engine needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.scale, self.mode = scale, mode
        self._ctx = ctypes.c_void_p()
        check(self.lib.codon_create(ctypes.byref(self._ctx), self.device.index, scale, MODES[mode]))
        self._ws: Optional[torch.Tensor] = None
        self._lock = threading.Lock()

    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.codon_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ---------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Accepts a reference state_dict (optionally with DataParallel's ``module.`` prefix)."""
        for name, t in sd.items():
            a = t.detach().to("cpu", torch.float32).contiguous()
            shape = (ctypes.c_int64 * a.dim())(*a.shape)
            check(self.lib.codon_set_weight(self._ctx, name.encode(), ctypes.cast(a.data_ptr(), ctypes.POINTER(ctypes.c_float)),
                                            shape, a.dim()), self._ctx)
        check(self.lib.codon_finalize_weights(self._ctx), self._ctx)

    def load_weights_file(self, path: str) -> None:
        """Loads a flat weight file written by ``codon_b200.checkpoint.export_flat`` (the C-host route)."""
        check(self.lib.codon_load_weights_file(self._ctx, os.fsencode(path)), self._ctx)

    @property
    def weights_generation(self) -> int:
        """Bumped by every load_state_dict; captured graphs re-capture when it moves."""
        return int(self.lib.codon_weights_generation(self._ctx))

    # ---- forward ---------------------------------------------------------------------------------
    def workspace_bytes(self, B: int, H: int, W: int) -> int:
        return int(self.lib.codon_workspace_bytes(self._ctx, B, H, W))

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, depth: torch.Tensor, guide: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """depth, guide: [B,1,H,W] (or [B,H,W]) CUDA tensors of one dtype (fp32 / fp16 / bf16);
        returns [B,1,H,W] of that dtype.  Asynchronous on the current stream."""
        _require_cuda(depth, "depth")
        _require_cuda(guide, "guide")
        if depth.shape != guide.shape or depth.dtype != guide.dtype:
            raise CodonError(f"depth {tuple(depth.shape)}/{depth.dtype} and guide {tuple(guide.shape)}/{guide.dtype} differ")
        if depth.dtype not in _IO_DTYPES:
            raise CodonError(f"unsupported frame dtype {depth.dtype}")
        if depth.dim() == 4:
            if depth.shape[1] != 1:
                raise CodonError(f"frames must be single-channel, got {tuple(depth.shape)}")
            B, _, H, W = depth.shape
        elif depth.dim() == 3:
            B, H, W = depth.shape
        else:
            raise CodonError(f"frames must be [B,1,H,W] or [B,H,W], got {tuple(depth.shape)}")
        depth, guide = depth.contiguous(), guide.contiguous()
        if out is None:
            out = torch.empty_like(depth)
        with self._lock, torch.cuda.device(self.device):
            ws = self._workspace(self.workspace_bytes(B, H, W))
            check(self.lib.codon_forward(self._ctx, depth.data_ptr(), guide.data_ptr(), out.data_ptr(), B, H, W,
                                         _IO_DTYPES[depth.dtype], ws.data_ptr(), ws.numel(),
                                         current_stream_ptr(self.device)), self._ctx)
        return out

    __call__ = forward

    @staticmethod
    def pinned_frames(*shape):
        """A page-locked float32 numpy array (torch pin_memory): forward_host copies such buffers to / from the
        device directly instead of staging them through the context's own pinned buffers."""
        return torch.empty(*shape, dtype=torch.float32, pin_memory=True).numpy()

    def forward_host(self, depth, guide, out=None):
        """HOST float32 numpy arrays [B,H,W] (or [B,1,H,W]) in, numpy out; copies included
        (the reference's H2D / D2H around the model call, CODON_X4/test.py:122-128).  `out` may be a
        preallocated (ideally page-locked, see pinned_frames) array of the input shape."""
        import numpy as np
        d = np.ascontiguousarray(depth, dtype=np.float32)
        g = np.ascontiguousarray(guide, dtype=np.float32)
        if d.shape != g.shape:
            raise CodonError("depth and guide shapes differ")
        shp = d.shape
        H, W = shp[-2], shp[-1]
        B = int(d.size // (H * W))
        if out is None:
            out = np.empty_like(d)
        elif out.dtype != np.float32 or out.shape != d.shape or not out.flags.c_contiguous:
            raise CodonError("out must be a C-contiguous float32 array of the input shape")
        with self._lock:
            check(self.lib.codon_forward_host(self._ctx, d.ctypes.data, g.ctypes.data, out.ctypes.data, B, H, W), self._ctx)
        return out

    def submit_host(self, depth, guide, out) -> None:
        """Non-blocking forward_host: enqueues H2D copy -> forward -> D2H copy of one call and returns.  All three
        arrays must be page-locked C-contiguous float32 (see pinned_frames) and must not be touched until the
        matching wait_host() returns.  At most two calls may be in flight; keeping one submitted ahead hides the
        copies of neighbouring calls under the kernels of the current one (stream_host does that)."""
        import numpy as np
        for a, what in ((depth, "depth"), (guide, "guide"), (out, "out")):
            if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags.c_contiguous:
                raise CodonError(f"{what} must be a C-contiguous float32 numpy array (page-locked, see pinned_frames)")
        if depth.shape != guide.shape or out.shape != depth.shape:
            raise CodonError("depth, guide and out shapes differ")
        H, W = depth.shape[-2], depth.shape[-1]
        B = int(depth.size // (H * W))
        with self._lock:
            check(self.lib.codon_forward_host_submit(self._ctx, depth.ctypes.data, guide.ctypes.data, out.ctypes.data,
                                                     B, H, W), self._ctx)

    def wait_host(self) -> None:
        """Blocks until the oldest submit_host() call has written its `out`."""
        check(self.lib.codon_forward_host_wait(self._ctx), self._ctx)

    def stream_host(self, calls):
        """Runs an iterable of (depth, guide, out) page-locked triples through submit_host / wait_host with one call
        submitted ahead, and yields each `out` as it completes (in order).  Throughput form of the reference's
        per-image loop (CODON_X4/test.py:109-145): the H2D copy of call i+1 and the D2H copy of call i-1 run while
        the kernels of call i do."""
        pending = []
        for d, g, o in calls:
            self.submit_host(d, g, o)
            pending.append(o)
            if len(pending) == 2:
                self.wait_host()
                yield pending.pop(0)
        while pending:
            self.wait_host()
            yield pending.pop(0)

    def capture_graph(self, B: int, H: int, W: int, dtype: torch.dtype = torch.float32) -> "GraphedForward":
        """Captures the forward for a fixed shape into a CUDA graph: one
        cudaGraphLaunch replaces the ~36 kernel launches of a frame."""
        return GraphedForward(self, B, H, W, dtype)

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.codon_last_launch_count(self._ctx))

    # ---- per-kernel-class profiling (CUDA events on the forward's stream) ---------------------------
    def profile_enable(self, on: bool = True) -> None:
        check(self.lib.codon_profile_enable(self._ctx, int(on)), self._ctx)

    def profile_reset(self) -> None:
        check(self.lib.codon_profile_reset(self._ctx), self._ctx)

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        """{category: {"ms": total device ms, "work": FLOP or bytes, "launches": n}} since the last reset."""
        out = {}
        for cat in range(8):
            ms, work, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_longlong(0)
            check(self.lib.codon_profile_read(self._ctx, cat, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n)), self._ctx)
            out[self.lib.codon_profile_category_name(cat).decode()] = {"ms": ms.value, "work": work.value, "launches": int(n.value)}
        return out

    def debug_tap(self, name: str, B: int, H: int, W: int) -> torch.Tensor:
        chans = {"enc": 128, "feat": 128, "ms": 128, "fuse": 64, "out_fuse": 64}[name]
        dst = torch.empty(B, chans, H, W, dtype=torch.float32, device=self.device)
        c = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            check(self.lib.codon_debug_tap(self._ctx, name.encode(), dst.data_ptr(), ctypes.byref(c),
                                           current_stream_ptr(self.device)), self._ctx)
        return dst


class GraphedForward:
    """A CUDA-graph replay of ``Engine.forward`` for one (B, H, W, dtype).  The library's launches are
    stream-ordered and allocation-free (the workspace is caller-owned), so the whole forward is
    capturable; inputs are copied into static buffers before each replay."""

    def __init__(self, eng: "Engine", B: int, H: int, W: int, dtype: torch.dtype = torch.float32):
        self.eng = eng
        dev = eng.device
        self.x = torch.zeros(B, 1, H, W, dtype=dtype, device=dev)
        self.y = torch.zeros(B, 1, H, W, dtype=dtype, device=dev)
        self.out = torch.empty(B, 1, H, W, dtype=dtype, device=dev)
        self._capture()

    def _capture(self) -> None:
        """(Re-)captures the forward.  The graph bakes in device pointers (weights, workspace) and per-layer constants
        derived from the weights; ``Engine.load_state_dict`` keeps the pointers valid (in-place re-upload) and bumps
        ``weights_generation``, which ``replay`` checks."""
        eng, dev = self.eng, self.eng.device
        self.generation = eng.weights_generation
        eng.profile_enable(False)
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):                      # warm-up: function attributes, tensor maps, workspace
                    eng.forward(self.x, self.y, self.out)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                eng.forward(self.x, self.y, self.out)

    def __call__(self, depth: torch.Tensor, guide: torch.Tensor) -> torch.Tensor:
        self.x.copy_(depth.reshape(self.x.shape))
        self.y.copy_(guide.reshape(self.y.shape))
        return self.replay()

    def replay(self) -> torch.Tensor:
        """Replays on the data already in ``self.x`` / ``self.y`` (after re-capturing if the engine's weights were
        reloaded since the capture)."""
        if self.eng.weights_generation != self.generation:
            self._capture()
        self.graph.replay()
        return self.out


# ---- stand-alone attention pieces (module-level API of CAC_module / attention.ResCBAM) -------------

POOL_BITS = {"avg": 1, "max": 2, "lp": 4, "lse": 8}


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    _require_cuda(t, what)
    return t.detach().to(torch.float32).contiguous()


def cac_channel_scale(x, w1, b1, w2, b2, pool_types=("avg", "max")) -> torch.Tensor:
    """[B,C,H,W] -> [B,C_out] sigmoid gate (CAC_module.py:38-63 / ResCBAM.py:38-61)."""
    lib = load_library()
    xf = _f32c(x, "x")
    B, C, H, W = xf.shape
    w1f, b1f, w2f, b2f = (_f32c(t, "weight").to(xf.device) for t in (w1, b1, w2, b2))
    hidden, c_out = w1f.shape[0], w2f.shape[0]
    mask = 0
    for p in pool_types:
        mask |= POOL_BITS[p]
    out = torch.empty(B, c_out, dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        check(lib.codon_cac_channel(xf.data_ptr(), B, C, H, W, w1f.data_ptr(), b1f.data_ptr(), w2f.data_ptr(),
                                    b2f.data_ptr(), hidden, c_out, mask, out.data_ptr(), current_stream_ptr(xf.device)))
    return out


def cac_spatial_scale(x, w, return_pooled: bool = False):
    """[B,C,H,W] -> [B,1,H,W] sigmoid(conv5x5(ChannelPool(x))) (CAC_module.py:78-94)."""
    lib = load_library()
    xf = _f32c(x, "x")
    B, C, H, W = xf.shape
    wf = _f32c(w, "weight").to(xf.device)
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=xf.device)
    pooled = torch.empty(B, 2, H, W, dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        check(lib.codon_cac_spatial(xf.data_ptr(), B, C, H, W, wf.data_ptr(), out.data_ptr(), pooled.data_ptr(),
                                    current_stream_ptr(xf.device)))
    return (out, pooled) if return_pooled else out


def cac_apply(x, sc=None, ss=None, res=None) -> torch.Tensor:
    """y = x * sc[b, c % c_gate] * ss[b,h,w] (+ res)   (CODON_x4.py:88-91,117-118; ResCBAM.py:61,87)."""
    lib = load_library()
    xf = _f32c(x, "x")
    B, C, H, W = xf.shape
    scf = _f32c(sc, "sc") if sc is not None else None
    ssf = _f32c(ss, "ss") if ss is not None else None
    rf = _f32c(res, "res") if res is not None else None
    y = torch.empty_like(xf)
    with torch.cuda.device(xf.device):
        check(lib.codon_cac_apply(xf.data_ptr(), scf.data_ptr() if scf is not None else None,
                                  ssf.data_ptr() if ssf is not None else None,
                                  rf.data_ptr() if rf is not None else None, B, C, H, W,
                                  scf.shape[1] if scf is not None else 1, y.data_ptr(), current_stream_ptr(xf.device)))
    return y


def channel_stats(x) -> torch.Tensor:
    """[B,C,H,W] -> [4,B,C]: mean, max, lp(2), lse per plane (CAC_module.py:43,47,50-55,71-76)."""
    lib = load_library()
    xf = _f32c(x, "x")
    B, C, H, W = xf.shape
    out = torch.empty(4, B, C, dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        check(lib.codon_channel_stats(xf.data_ptr(), B, C, H, W, out.data_ptr(), current_stream_ptr(xf.device)))
    return out


def channel_pool(x) -> torch.Tensor:
    """ChannelPool (CAC_module.py:78-81): [B,C,H,W] -> [B,2,H,W] (max, mean)."""
    lib = load_library()
    xf = _f32c(x, "x")
    B, C, H, W = xf.shape
    out = torch.empty(B, 2, H, W, dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        check(lib.codon_channel_pool(xf.data_ptr(), B, C, H, W, out.data_ptr(), current_stream_ptr(xf.device)))
    return out


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def conv2d_nchw(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1, relu=False) -> torch.Tensor:
    """Generic small NCHW fp32 convolution (BasicConv, CAC_module.py:6-20)."""
    lib = load_library()
    xf = _f32c(x, "x")
    B, Cin, H, W = xf.shape
    wf = _f32c(w, "weight").to(xf.device)
    bf = _f32c(bias, "bias").to(xf.device) if bias is not None else None
    Cout, _, kh, kw = wf.shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    y = torch.empty(B, Cout, OH, OW, dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        check(lib.codon_conv2d_nchw(xf.data_ptr(), wf.data_ptr(), bf.data_ptr() if bf is not None else None,
                                    B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, int(relu),
                                    y.data_ptr(), current_stream_ptr(xf.device)))
    return y


# ---- driver post-processing and metrics on the GPU -------------------------------------------------

def quantise_u8(out: torch.Tensor, via_half: bool = False) -> torch.Tensor:
    """clip(0,1) * 255 truncated to uint8 (CODON_X4/test.py:130,132)."""
    lib = load_library()
    of = _f32c(out, "out")
    dst = torch.empty(of.shape, dtype=torch.uint8, device=of.device)
    with torch.cuda.device(of.device):
        check(lib.codon_quantise_u8(of.data_ptr(), dst.data_ptr(), of.numel(), int(via_half), current_stream_ptr(of.device)))
    return dst


def masked_rmse(label_u8: torch.Tensor, out_u8: torch.Tensor) -> torch.Tensor:
    """EvaluationResults (CODON_X4/test.py:148-164) per frame: uint8 [B,H,W] CUDA tensors -> float64 [B]."""
    lib = load_library()
    _require_cuda(label_u8, "label")
    _require_cuda(out_u8, "out")
    o = out_u8.reshape(-1, out_u8.shape[-2], out_u8.shape[-1]).contiguous()
    lab = label_u8.reshape(-1, label_u8.shape[-2], label_u8.shape[-1])[:, :o.shape[1], :o.shape[2]].contiguous()
    if lab.dtype != torch.uint8 or o.dtype != torch.uint8 or lab.shape != o.shape:
        raise CodonError("masked_rmse needs uint8 tensors of one shape")
    B, H, W = o.shape
    r = torch.empty(B, dtype=torch.float64, device=o.device)
    with torch.cuda.device(o.device):
        check(lib.codon_masked_rmse(lab.data_ptr(), o.data_ptr(), B, H, W, r.data_ptr(), current_stream_ptr(o.device)))
    return r


def ssim_gauss(img1: torch.Tensor, img2: torch.Tensor, sd: float = 1.5, c1: float = 0.01 ** 2,
               c2: float = 0.03 ** 2) -> torch.Tensor:
    """ssim_exact (CODON_X4/ssim_2.py:36-52) per frame.  uint8 inputs are scaled by 1/255 (the
    driver's call, test.py:139); float inputs are used as float64 as given.  Returns float64 [B]."""
    lib = load_library()
    _require_cuda(img1, "img1")
    _require_cuda(img2, "img2")
    a = img1.reshape(-1, img1.shape[-2], img1.shape[-1])
    b = img2.reshape(-1, img2.shape[-2], img2.shape[-1])
    if a.shape != b.shape:
        raise CodonError("ssim_gauss needs images of one shape")
    if a.dtype == torch.uint8 and b.dtype == torch.uint8:
        dt = 0
        a, b = a.contiguous(), b.contiguous()
    else:
        dt = 1
        a, b = a.to(torch.float64).contiguous(), b.to(torch.float64).contiguous()
    B, H, W = a.shape
    ws = torch.empty((5 * H * W + H) * B, dtype=torch.float64, device=a.device)
    r = torch.empty(B, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.codon_ssim_gauss(a.data_ptr(), b.data_ptr(), dt, B, H, W, sd, c1, c2, r.data_ptr(), ws.data_ptr(),
                                   ws.numel() * 8, current_stream_ptr(a.device)))
    return r


# ---- driver pre-processing on the GPU ----------------------------------------------------------------

def bgr_to_gray_u8(bgr: torch.Tensor, method: str = "imread") -> torch.Tensor:
    """uint8 [..., H, W, 3] (BGR, as cv2 decodes) -> uint8 [..., H, W].  method "imread": what
    cv2.imread(png, 0) returns (libpng conversion, test.py:118); "cvtcolor": cv2.cvtColor(BGR2GRAY)."""
    lib = load_library()
    _require_cuda(bgr, "bgr")
    if bgr.dtype != torch.uint8 or bgr.shape[-1] != 3:
        raise CodonError("bgr_to_gray_u8 needs a uint8 [..., 3] tensor")
    src = bgr.contiguous()
    dst = torch.empty(src.shape[:-1], dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        check(lib.codon_bgr_to_gray_u8(src.data_ptr(), dst.data_ptr(), dst.numel(), {"imread": 0, "cvtcolor": 1}[method],
                                       current_stream_ptr(src.device)))
    return dst


def u8_to_unit_f32(img: torch.Tensor) -> torch.Tensor:
    """uint8 -> float32(v / 255) (CODON_X4/test.py:122-123)."""
    lib = load_library()
    _require_cuda(img, "img")
    if img.dtype != torch.uint8:
        raise CodonError("u8_to_unit_f32 needs a uint8 tensor")
    src = img.contiguous()
    dst = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(lib.codon_u8_to_unit_f32(src.data_ptr(), dst.data_ptr(), src.numel(), current_stream_ptr(src.device)))
    return dst


def bicubic_upsample(lr: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """float32 [B,h,w] (or [B,1,h,w]) -> [B,H,W] with cv2.resize(INTER_CUBIC) semantics."""
    lib = load_library()
    src = _f32c(lr, "lr")
    shp = src.shape
    src = src.reshape(-1, shp[-2], shp[-1])
    B, h, w = src.shape
    dst = torch.empty(B, H, W, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(lib.codon_bicubic_upsample_f32(src.data_ptr(), dst.data_ptr(), B, h, w, H, W, current_stream_ptr(src.device)))
    return dst.reshape(*shp[:-2], H, W)


# ---- one frame over several GPUs -----------------------------------------------------------------------

class FrameGroup:
    """Single-frame latency mode: one frame is split into horizontal bands, one per GPU of
    this process; halo rows travel over NVLink peer memory after every layer and the CAC channel
    statistics are all-gathered (codon_group_* in include/codon_b200.h).  No NCCL, no torch.distributed."""

    def __init__(self, scale: int, mode: str, devices, state_dict: Dict[str, torch.Tensor]):
        self.lib = load_library()
        self.engines = [Engine(scale, mode, int(d)) for d in devices]
        for e in self.engines:
            e.load_state_dict(state_dict)
        arr = (ctypes.c_void_p * len(self.engines))(*[e._ctx for e in self.engines])
        self._grp = ctypes.c_void_p()
        check(self.lib.codon_group_create(ctypes.byref(self._grp), arr, len(self.engines)))

    def forward_host(self, depth, guide):
        """HOST float32 [H,W] arrays in, [H,W] out (copies and the cross-GPU exchange included)."""
        import numpy as np
        d = np.ascontiguousarray(depth, dtype=np.float32)
        g = np.ascontiguousarray(guide, dtype=np.float32)
        if d.ndim != 2 or d.shape != g.shape:
            raise CodonError("FrameGroup.forward_host needs two [H,W] arrays of one shape")
        out = np.empty_like(d)
        rc = self.lib.codon_group_forward_host(self._grp, d.ctypes.data, g.ctypes.data, out.ctypes.data, d.shape[0], d.shape[1])
        if rc != 0:
            raise CodonError(f"libcodon_b200 error {rc}: {self.lib.codon_group_last_error(None).decode()}")
        return out

    @property
    def last_ms(self) -> float:
        return float(self.lib.codon_group_last_ms(self._grp))

    def close(self):
        if getattr(self, "_grp", None) and self._grp.value:
            self.lib.codon_group_destroy(self._grp)
            self._grp = ctypes.c_void_p()
        for e in getattr(self, "engines", []):
            e.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
