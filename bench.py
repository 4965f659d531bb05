#!/usr/bin/env python
"""Benchmark of the CODON forward pass on B200 (metric: HR depth megapixels per second).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--mode tf32|f16x3|fp32|fp16|bf16] [--scale 4|8|16] [--frames F] [--height H] [--width W]

A step is one forward pass over one batch of F synthetic frames per GPU.  The default workload is
BASELINE.json configs[1]: CODON x4, a single 640x480 frame, "fp32 parity mode" (mode tf32: fp32
activations in HBM, TF32 tensor-core products, fp32 accumulation; max-abs error vs the reference
fp32 forward <= 1e-3, measured in this run and printed under "parity").  With --gpus N>1 (launched
under torchrun, one rank per GPU) every rank processes its own F frames per step: independent
frames, no collective on the data path, weak scaling; NCCL is used for the barrier and the
max-over-ranks timing only.

One JSON line is printed by rank 0 (see the driver contract in the task description):
  value        device-timed throughput, frames resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through the drop-in call the reference's driver makes (CODON_X4/test.py:122-128):
               pinned host frames -> .cuda() -> CODONNet.forward (codon_b200.CODON_x4.CODONNet, one codon_forward
               call) -> .cpu(), every step inside the timed region.  e2e.streaming = the library's streaming
               host entry points (codon_forward_host_submit / _wait, one call submitted ahead: copies overlap the
               neighbouring call's kernels); e2e.blocking_call = one codon_forward_host per step
  roofline     dominant kernel (5x5 128->128 tcgen05 implicit GEMM with the fused 1x1): algorithmic FLOP
               of the 5x5 alone / CUDA-event time of its launches, against MEASURED_PEAKS.json.  The
               per-launch events are recorded over a second pass of the same K steps: events between the
               launches serialise them, and `value` times the forward as a caller runs it.
  cpu_baseline the reference's own CPU forward (its CODONNet class from oracle/_ref, torch fp32, all host cores;
               kind "reference") -- or the oracle's restatement of it (kind "port") when oracle/_ref was never
               built -- on a bounded sample
--impl reference times that CPU forward alone, as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The GPU sits at its power cap inside the tensor-core kernels even on a single 640x480 frame, so a measurement pass
# that starts right behind another one starts with lower clocks.  Every pass (device-timed, instrumented, blocking
# host calls, streaming host calls) therefore starts after the same short idle, like separate runs would.
IDLE_S = 1.5
METRIC = "hr_depth_megapixels_per_second"
UNIT = "MP/s"
DTYPE_NAME = {"fp32": "f32", "tf32": "f32 (tf32 tensor-core products, f32 accumulate)",
              "f16x3": "f32-accurate (split f16 hi+lo operands, 3 tensor-core MMAs per K step, f32 accumulate)",
              "fp16": "f16 operands, f32 accumulate", "bf16": "bf16 operands, f32 accumulate"}
# tensor-core MMA work per algorithmic FLOP (f16x3 issues hi*hi + lo*hi + hi*lo)
MMA_FACTOR = {"fp32": 1.0, "tf32": 1.0, "f16x3": 3.0, "fp16": 1.0, "bf16": 1.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="codon_b200", choices=["codon_b200", "reference"])
    ap.add_argument("--mode", default="tf32", choices=["fp32", "tf32", "f16x3", "fp16", "bf16"])
    ap.add_argument("--scale", type=int, default=4, choices=[4, 8, 16])
    ap.add_argument("--frames", type=int, default=1, help="frames per GPU per step")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the other arithmetic modes")
    ap.add_argument("--bundled", action="store_true",
                    help="with --impl reference: BASELINE configs[0] -- the CPU forward over the 10 bundled Middlebury "
                         "images (tests/golden/images) with the test.py loop semantics, printing name rmse ssim per image")
    return ap.parse_args()


def workload_name(a):
    return (f"CODON x{a.scale}, {a.frames} x {a.width}x{a.height} synthetic RGB-D frame(s) per GPU per step, "
            f"mode {a.mode}")


def config_dict(a):
    """The `config` object of BOTH arms (the driver compares them key by key): what is measured, not how."""
    return {"workload": workload_name(a), "scale": a.scale, "frames_per_gpu": a.frames, "height": a.height, "width": a.width,
            "mode": a.mode, "weights": "synthetic seed 0 (reference init, output.weight x0.002)",
            "sharding": "independent frames per rank, no data-path collective"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.paused = False
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            if self.paused:
                self._stop.wait(self.period)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def idle(self, seconds: float):
        """Sleeps without sampling (the clocks of an idle GPU are not the clocks under load)."""
        self.paused = True
        time.sleep(seconds)
        self.paused = False

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# CPU leg (oracle; the only place bench.py touches oracle/)

def cpu_forward_fn(scale):
    """The CPU forward the CPU legs time: the reference's own `CODONNet` (oracle/_ref, built from /root/reference by
    oracle/build_ref.py; kind "reference"), else the oracle's restatement (kind "port").  Returns (fn(x, y), kind, orc)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import codon_oracle as orc
    import build_ref
    sd = orc.synthetic_state_dict(scale, 0)
    cls = build_ref.load_model_class(scale)
    if cls is not None:
        net = cls().eval()
        net.load_state_dict(sd, strict=True)
        return (lambda x, y: net(x, y)), "reference", orc
    return (lambda x, y: orc.forward(sd, x, y)), "port", orc


def cpu_forward_sample(scale, height, width, budget_s, steps, warmup):
    """Times the reference's fp32 CPU forward (see cpu_forward_fn) on a bounded sample: the top `rows` rows of one
    frame of the workload, rows chosen so that (steps + warmup) forwards fit `budget_s`.
    Returns (MP/s, description, cores, out, rows, mean seconds, kind)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind, orc = cpu_forward_fn(scale)
    x, y = orc.synthetic_frames(1, height, width, 1234)
    probe_rows = min(height, 96)
    with torch.no_grad():
        fwd(x[:, :, :probe_rows], y[:, :, :probe_rows])          # oneDNN primitive warm-up
        t0 = time.perf_counter()
        fwd(x[:, :, :probe_rows], y[:, :, :probe_rows])
        rate = probe_rows * width / (time.perf_counter() - t0)               # px/s
    rows = int(min(height, max(32, budget_s * rate / max(1, steps + warmup) / width)))
    xs, ys = x[:, :, :rows].contiguous(), y[:, :, :rows].contiguous()
    times, out = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = fwd(xs, ys)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    mean_s = sum(times) / len(times)
    what = "the reference's CODONNet class (oracle/_ref)" if kind == "reference" else "the oracle's restatement of the reference forward"
    desc = (f"{what}: {steps} timed forward(s) after {warmup} warm-up on the top {rows} of {height} rows of one "
            f"{width}x{height} frame (x{scale} synthetic weights seed 0, torch {torch.__version__} fp32, {cores} threads)")
    return rows * width / 1e6 / mean_s, desc, cores, out, rows, mean_s, kind


def run_reference_bundled(a):
    """BASELINE configs[0]: the reference's CPU path on the bundled images.  The .pth files are absent from the
    reference checkout, so the weights are the synthetic seed-0 set (image quality is meaningless; the timing
    and the pipeline are what is measured).  Loop semantics of CODON_X4/test.py:109-145."""
    import torch
    import cv2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind, orc = cpu_forward_fn(a.scale)
    img = os.path.join(ROOT, "tests", "golden", "images")
    names = sorted(os.listdir(os.path.join(img, "gray")))
    rmse_sum = ssim_sum = 0.0
    px = 0
    t_fwd = 0.0
    for n in names:
        d = cv2.imread(os.path.join(img, f"depth_x{a.scale}", n), 0)
        g = cv2.imread(os.path.join(img, "gray", n), 0)
        lab = cv2.imread(os.path.join(img, "label", n), 0)
        x = torch.from_numpy(d / 255).float()[None, None]
        y = torch.from_numpy(g / 255).float()[None, None]
        t0 = time.perf_counter()
        with torch.no_grad():
            out = fwd(x, y)[0, 0].numpy()
        t_fwd += time.perf_counter() - t0
        q = orc.quantise_output(out)
        r, s_ = orc.masked_rmse(lab, q), orc.ssim_gauss(lab / 255, q / 255)
        rmse_sum += r
        ssim_sum += s_
        px += d.size
        print(n, r, s_, file=sys.stderr)
    print(len(names), file=sys.stderr)
    print(rmse_sum / len(names), ssim_sum / len(names), file=sys.stderr)
    mps = px / 1e6 / t_fwd
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": mps, "unit": UNIT, "n_gpus": a.gpus,
                      "steps": len(names), "warmup": 0, "ms_per_step": t_fwd / len(names) * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "bundled Middlebury images, synthetic weights",
                      "config": {"workload": f"BASELINE configs[0]: CODON x{a.scale} CPU forward over the 10 bundled images"},
                      "cpu_baseline": {"value": mps, "unit": UNIT, "cores": cores, "kind": kind,
                                       "sample": f"{len(names)} bundled images, forward only, {cores} threads"},
                      "mean_rmse": rmse_sum / len(names), "mean_ssim": ssim_sum / len(names),
                      "e2e": {"value": mps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}), flush=True)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.bundled:
        return run_reference_bundled(a)
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    mps, desc, cores, _, rows, mean_s, kind = cpu_forward_sample(a.scale, a.height, a.width, 150.0, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": mps, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a), "timing": "host wall clock around the CPU forward (fp32 whatever --mode says)",
        "cpu_baseline": {"value": mps, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": mps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm

def time_mode(eng, x, y, out, steps, warmup, flush, dist_barrier, profile=False):
    """Device time of `steps` forwards (CUDA events on the current stream, L2 flushed between steps)."""
    import torch
    for _ in range(warmup):
        eng.forward(x, y, out)
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    if profile:
        eng.profile_reset()
        eng.profile_enable(True)
    dist_barrier()
    torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()                      # evict L2 (buffer larger than the 126 MB L2), outside the timed window
        starts[i].record()
        eng.forward(x, y, out)
        ends[i].record()
    torch.cuda.synchronize()
    dist_barrier()
    prof = None
    if profile:
        eng.profile_enable(False)
        prof = eng.profile_read()
    return sum(s.elapsed_time(e) for s, e in zip(starts, ends)), prof


def measure_cublas_peak(dev, kind: str, n: int = 8192, reps: int = 20):
    """Dense GEMM rate of cuBLAS on this GPU, now: bf16, or fp32 inputs with TF32 tensor-core products."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        dt = torch.bfloat16 if kind == "bf16" else torch.float32
        torch.backends.cuda.matmul.allow_tf32 = True
        a_ = torch.randn(n, n, device=dev, dtype=dt)
        b_ = torch.randn(n, n, device=dev, dtype=dt)
        for _ in range(3):
            torch.matmul(a_, b_)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a_, b_)
        e1.record()
        torch.cuda.synchronize(dev)
        tf = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) / 1e3) / 1e12
        return {"tflops": tf, "kind": kind, "how": f"torch.matmul {n}^3 x {reps} (cuBLAS), CUDA events, this run"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def run_gpu(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from codon_b200 import engine as E, synthetic as syn, build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; codon_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    build.build_library()
    peaks = load_peaks()
    B, H, W = a.frames, a.height, a.width
    P = B * H * W
    sd = syn.synthetic_state_dict(a.scale, 0)
    # every rank gets its own frames (seed offset by rank): independent shards, no data-path collective
    xh, yh = syn.synthetic_frames(B, H, W, 1234 + rank * B)
    x, y = xh.to(dev), yh.to(dev)
    out = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps, warmup = max(1, a.steps), max(3, a.warmup)

    # the drop-in module (codon_b200.CODON_x4 / _x8 / _x16 . CODONNet) owns the engine: every number below is measured on
    # the engine a user of the reference's module surface gets
    from importlib import import_module
    net = import_module(f"codon_b200.CODON_x{a.scale}").CODONNet().eval().set_mode(a.mode)
    net.load_state_dict(sd)
    eng = net.engine(dev)

    # ---- device-timed throughput + per-kernel-class profile ---------------------------------------
    with ClockSampler(local) as clk:
        # pass 1: the headline -- K forwards exactly as a caller runs them (no instrumentation between the launches,
        # so programmatic dependent launch overlaps each kernel's ramp-up with its predecessor's tail)
        total_ms, _ = time_mode(eng, x, y, out, steps, warmup, flush, barrier)
        launches = eng.last_launch_count * steps
        # pass 2: the same K forwards with a CUDA event pair around every launch (recorded inside the library on the
        # forward's stream) -> per-kernel times for the roofline; the events serialise the launches
        clk.idle(IDLE_S)
        prof_ms, prof = time_mode(eng, x, y, out, steps, 3, flush, barrier, profile=True)
        # ---- end to end through the host entry point (pinned H2D + forward + D2H per step) -----------
        # the step's inputs live in PINNED host memory and the result is read back into pinned host memory
        xn, yn, res = (E.Engine.pinned_frames(*xh.shape) for _ in range(3))
        xn[...] = xh.numpy()
        yn[...] = yh.numpy()
        for _ in range(2):
            eng.forward_host(xn, yn, out=res)
        # (a) one blocking call per step: copy in, forward, copy out, synchronise
        clk.idle(IDLE_S)
        eng.forward_host(xn, yn, out=res)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.forward_host(xn, yn, out=res)
        e2e_blocking_s = time.perf_counter() - t0
        # (b) the streaming form of the same call (Engine.stream_host: submit / wait with one call submitted ahead,
        # two sets of pinned buffers): every step still copies its own inputs in and its own result out inside the
        # timed region; the copies of neighbouring steps run under the kernels of the current one
        xn2, yn2, res2 = (E.Engine.pinned_frames(*xh.shape) for _ in range(3))
        xn2[...] = xn
        yn2[...] = yn
        bufs = [(xn, yn, res), (xn2, yn2, res2)]
        clk.idle(IDLE_S)
        for _ in eng.stream_host(bufs[i & 1] for i in range(3)):
            pass
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in eng.stream_host(bufs[i & 1] for i in range(steps)):
            pass
        e2e_s = time.perf_counter() - t0
        barrier()
        assert np.array_equal(res, res2)
        # (c) THE HEADLINE e2e: the reference driver's own sequence around the model call (CODON_X4/test.py:122-128) with
        # the drop-in module -- pinned host frames -> .cuda() -> model(x, y) -> .cpu() -- every step
        xp, yp = torch.from_numpy(xn), torch.from_numpy(yn)           # views of the pinned buffers
        rp = torch.from_numpy(res)

        def dropin_step():
            with torch.no_grad():
                o = net(xp.to(dev, non_blocking=True), yp.to(dev, non_blocking=True))       # test.py:122-125
            rp.copy_(o, non_blocking=True)                                                  # test.py:127-128 (.cpu())
            torch.cuda.synchronize(dev)
        clk.idle(IDLE_S)
        for _ in range(3):
            dropin_step()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            dropin_step()
        e2e_dropin_s = time.perf_counter() - t0
        barrier()
        assert np.array_equal(res, res2)
        # host overhead of the module surface alone: CODONNet.forward vs Engine.forward on resident frames
        with torch.no_grad():
            t0 = time.perf_counter()
            for _ in range(200):
                net.engine(dev)
            key_us = (time.perf_counter() - t0) / 200 * 1e6
    total_ms = max_over_ranks(total_ms)
    prof_ms = max_over_ranks(prof_ms)
    e2e_s = max_over_ranks(e2e_s)
    e2e_blocking_s = max_over_ranks(e2e_blocking_s)
    e2e_dropin_s = max_over_ranks(e2e_dropin_s)
    value = world * P * steps / 1e6 / (total_ms / 1e3)
    e2e_value = world * P * steps / 1e6 / e2e_dropin_s
    # CODON_TC_DEBUG knobs (kernel perf experiments) produce garbage on purpose; never set for a bench line
    assert os.environ.get("CODON_TC_DEBUG", "0") != "0" or np.isfinite(res).all()

    # ---- roofline of the dominant kernel (measured live above) --------------------------------------
    dom = prof["conv5x5_128to128"]
    tfs = dom["work"] / (dom["ms"] / 1e3) / 1e12 if dom["ms"] > 0 else 0.0
    conv_ms = sum(prof[k]["ms"] for k in ("conv5x5_128to128", "pair_3x3_5x5_64to128", "conv3x3", "conv1x1_128to64"))
    cac_ms = prof["cac_stats"]["ms"] + prof["cac_apply"]["ms"]
    cac_bytes = prof["cac_stats"]["work"] + prof["cac_apply"]["work"]
    apply_ms, apply_bytes = prof["cac_apply"]["ms"], prof["cac_apply"]["work"]
    all_ms = sum(v["ms"] for v in prof.values())
    peak_tf = peaks["bf16_tflops"]
    # the tensor-core rate of this mode's MMA kind, measured in this run with cuBLAS (torch.matmul): kind::tf32 for the
    # tf32 mode, 16-bit otherwise (f16x3 issues three 16-bit MMAs per algorithmic multiply-add: MMA_FACTOR)
    measured_peak = measure_cublas_peak(dev, "tf32" if a.mode == "tf32" else "bf16") if a.mode != "fp32" else None
    mma_tfs = tfs * MMA_FACTOR[a.mode]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{a.mode}|{B}|{H}x{W}")
    roofline = {
        "kernel": "conv_tc2_kernel<FUSE> (5x5 128->128 implicit GEMM + ReLU + fused 1x1 128->64, tcgen05 cta_group::2)",
        "bound": "tensor",
        "achieved": tfs, "peak": peak_tf, "unit": "TFLOP/s", "frac": tfs / peak_tf,
        "traffic": traffic["bytes"] if traffic else None,
        "traffic_detail": traffic if traffic else "no ncu --set full capture of this workload is committed (profiles/traffic.json "
                                                  "holds the captured ones, keyed mode|frames|HxW)",
        "peak_source": peaks["source"] + ", dense bf16 burst" +
                       ("; tf32 runs at half the bf16 rate, so frac <= 0.5 in this mode" if a.mode == "tf32" else "") +
                       ("; f16x3 spends 3 MMAs per multiply-add, so frac <= 1/3 in this mode" if a.mode == "f16x3" else ""),
        # the same kernel against the rate of ITS instruction kind, measured in this run (cuBLAS through torch.matmul)
        "mma_tflops": mma_tfs, "mma_per_flop": MMA_FACTOR[a.mode],
        "peak_for_this_dtype": measured_peak,
        "frac_of_dtype_peak": (mma_tfs / measured_peak["tflops"]) if measured_peak else None,
        "frac_of_sustained_peak": (mma_tfs / peaks["bf16_tflops_sustained"] / (0.5 if a.mode == "tf32" else 1.0))
                                  if peaks.get("bf16_tflops_sustained") else None,
        "flop_per_launch": dom["work"] / max(1, dom["launches"]), "ms_per_launch": dom["ms"] / max(1, dom["launches"]),
        "launches": dom["launches"], "share_of_step": dom["ms"] / all_ms if all_ms else None,
        "timed_with": "CUDA event pair around every launch, over a second pass of the same K steps "
                      "(the events serialise the launches; `value` is the un-instrumented first pass)",
        "ms_per_step_instrumented": prof_ms / steps,
        "trunk_all_convs": {"achieved": syn.FLOPS_PER_PIXEL * P * steps / (conv_ms / 1e3) / 1e12 if conv_ms else None,
                            "unit": "TFLOP/s", "ms_per_step": conv_ms / steps},
        # CAC: the HBM-bound kernel is cac_apply (reads F and E, writes F: 384*e B/px/stage).  In the tensor-core modes the
        # statistics come out of the fused conv epilogue ("the stats read vanishes"); what is left of that
        # pass is a fold of the epilogue's 32 B/px partials, listed separately (fp32 mode: one 128*e B/px read of F).
        "cac_kernels": {"bound": "hbm", "kernel": "cac_apply_kernel",
                        "achieved": apply_bytes / (apply_ms / 1e3) / 1e9 if apply_ms else None,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": apply_bytes / (apply_ms / 1e3) / 1e9 / peaks["hbm_gbs"] if apply_ms else None,
                        "ms_per_step": apply_ms / steps,
                        "stats_pass": {"ms_per_step": prof["cac_stats"]["ms"] / steps,
                                       "bytes_per_step": prof["cac_stats"]["work"] / steps,
                                       "GB/s": prof["cac_stats"]["work"] / (prof["cac_stats"]["ms"] / 1e3) / 1e9
                                       if prof["cac_stats"]["ms"] else None},
                        "stats_plus_apply": {"GB/s": cac_bytes / (cac_ms / 1e3) / 1e9 if cac_ms else None,
                                             "ms_per_step": cac_ms / steps}},
        "by_kernel_ms_per_step": {k: v["ms"] / steps for k, v in prof.items()},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE_NAME[a.mode], "data": "synthetic",
        "config": config_dict(a),
        "timing": {"l2": "256 MiB buffer written between timed steps (L2 flush, outside the event window)",
                   "passes": f"device-timed, instrumented, blocking host calls, streaming host calls, drop-in calls; "
                             f"{IDLE_S} s idle before each"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * P * 4, "d2h_bytes_per_step": P * 4,
                "api": f"codon_b200.CODON_x{a.scale}.CODONNet.forward, called as the reference driver calls its model "
                       "(CODON_X4/test.py:122-128): pinned host fp32 frames -> .to(cuda) -> model(x, y) -> copy back to pinned "
                       "host memory -> synchronise, every step",
                "module_overhead_us": key_us,
                "streaming": {"value": world * P * steps / 1e6 / e2e_s, "unit": UNIT,
                              "api": "Engine.stream_host -> codon_forward_host_submit / _wait (one call submitted ahead: its H2D "
                                     "copy and the previous call's D2H copy run under the current call's kernels)"},
                "blocking_call": {"value": world * P * steps / 1e6 / e2e_blocking_s, "unit": UNIT,
                                  "api": "Engine.forward_host -> codon_forward_host (copy in, forward, copy out, synchronise; "
                                         "nothing overlapped)"}},
        "gpu_launches": launches, "roofline": roofline, "clocks": clk.summary(),
    }

    if rank == 0 and world == 1:
        # ---- other arithmetic modes on the same workload (reported, not the headline) ---------------
        if not a.no_variants:
            variants = {}
            outs = {a.mode: out.clone()}
            for m in ("fp32", "tf32", "f16x3", "fp16", "bf16"):
                if m == a.mode:
                    continue
                e2 = E.Engine(a.scale, m, local)
                e2.load_state_dict(sd)
                o2 = torch.empty_like(x)
                k = steps if m != "fp32" else max(2, steps // 5)
                ms, _ = time_mode(e2, x, y, o2, k, 3, flush, barrier)
                variants[m] = {"value": P * k / 1e6 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / k}
                outs[m] = o2.clone()
                e2.close()
            # the same workload replayed from a CUDA graph (one graph launch instead of ~36 kernel launches)
            try:
                g = eng.capture_graph(B, H, W)
                g.x.copy_(x); g.y.copy_(y)
                for _ in range(3):
                    g.replay()
                torch.cuda.synchronize()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                for s_, e_ in ev:
                    flush.zero_()
                    s_.record(); g.replay(); e_.record()
                torch.cuda.synchronize()
                gms = sum(s_.elapsed_time(e_) for s_, e_ in ev)
                variants[a.mode + "+cuda_graph"] = {"value": P * steps / 1e6 / (gms / 1e3), "unit": UNIT, "ms_per_step": gms / steps}
            except Exception as exc:   # noqa: BLE001 - a variant, never the headline
                variants[a.mode + "+cuda_graph"] = {"error": str(exc)[:200]}
            line["roofline"]["variants"] = variants      # (under roofline: the driver's record keeps this object whole)
        else:
            outs = {a.mode: out.clone()}
        # ---- CPU baseline (oracle) on a bounded sample + parity of the GPU result against it --------
        if not a.no_cpu_baseline:
            mps, desc, cores, ref, rows, _, kind = cpu_forward_sample(a.scale, H, W, 20.0, 1, 1)
            line["cpu_baseline"] = {"value": mps, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
            # The sample is the top `rows` rows; the receptive field is 47 px and the CAC
            # pooling is global, so compare the GPU forward of the SAME cropped input.
            xs, ys = x[:1, :, :rows].contiguous(), y[:1, :, :rows].contiguous()
            par = {}
            for m in outs:
                e2 = E.Engine(a.scale, m, local)
                e2.load_state_dict(sd)
                o = e2.forward(xs, ys)
                torch.cuda.synchronize()
                par[m] = float((o.cpu() - ref).abs().max())
                e2.close()
            # (under config: the driver's record keeps this object whole)
            line["config"]["parity_max_abs_vs_cpu_reference"] = par
            line["config"]["parity_note"] = (f"max |GPU - CPU {kind} fp32 forward| on the top {rows} rows of frame 0; north_star "
                                             "tolerance for the fp32 parity mode: 1e-3")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()
