"""GPU parity: the CUDA forward (through the C ABI, via the CODONNet drop-in) against the committed
reference outputs (tests/golden/fwd_*.npz, produced by the real reference) and the oracle."""
import os

import numpy as np
import pytest
import torch

import codon_oracle as orc
from codon_b200 import engine
from codon_b200 import CODON_x4, CODON_x8, CODON_x16

pytestmark = pytest.mark.gpu

MODS = {4: CODON_x4, 8: CODON_x8, 16: CODON_x16}
FWD = ["fwd_x4_s0_b2_48x64.npz", "fwd_x4_s1_b1_37x53.npz", "fwd_x8_s2_b1_40x72.npz", "fwd_x16_s2_b1_64x80.npz",
       "fwd_x4_s0_b1_120x160.npz"]
# max-abs tolerance in normalised depth vs the reference's fp64 forward.  north_star: fp32 mode <= 1e-3.
# fp16 / tf32 operands (10-bit mantissa) also hold 1e-3; bf16 (7-bit) is judged on RMSE/SSIM instead
# (tests/test_gpu_images.py) and only sanity-bounded here.
# f16x3 (split-fp16 operands, chunked accumulation) is the fp32-accurate tensor-core mode: <= 2e-6.
TOL = {"fp32": 1e-4, "tf32": 1e-3, "fp16": 1e-3, "bf16": 2e-2, "f16x3": 2e-6}


def _net(scale, seed, mode):
    net = MODS[scale].CODONNet().eval().set_mode(mode)
    net.load_state_dict(orc.synthetic_state_dict(scale, seed))
    return net


@pytest.mark.parametrize("mode", ["fp32", "tf32", "fp16", "bf16", "f16x3"])
@pytest.mark.parametrize("name", FWD)
def test_forward_matches_reference(golden_dir, name, mode):
    g = np.load(os.path.join(golden_dir, name))
    net = _net(int(g["scale"]), int(g["seed"]), mode)
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    with torch.no_grad():
        out = net(x, y)
    torch.cuda.synchronize()
    assert out.shape == x.shape and out.dtype == x.dtype
    err = np.abs(out.cpu().numpy().astype(np.float64) - g["out_fp64"]).max()
    print(f"{name} {mode}: max-abs {err:.3e}")
    assert err <= TOL[mode]


# ---- BASELINE.json shapes: the real reference's forward at 640x480, 8 x 640x480, 1280x720 and 1920x1080, committed as
# fingerprints (oracle/make_golden.py big): row sums, column sums, a stride-4 sample and five full-resolution crops.
BIG = ["big_x4_640x480.npz", "big_x8_b8_640x480.npz", "big_x4_1280x720.npz", "big_x16_1920x1080.npz"]


def _fingerprint(out):
    H, W = out.shape
    c = 96
    o = out.astype(np.float64)
    return {"row_sum": o.sum(1), "col_sum": o.sum(0), "sample4": out[1::4, 2::4],
            "crop_tl": out[:c, :c], "crop_tr": out[:c, W - c:], "crop_bl": out[H - c:, :c], "crop_br": out[H - c:, W - c:],
            "crop_ce": out[(H - c) // 2:(H - c) // 2 + c, (W - c) // 2:(W - c) // 2 + c]}


@pytest.mark.parametrize("mode", ["f16x3", "tf32", "fp16", "bf16", "fp32"])
@pytest.mark.parametrize("name", BIG)
def test_baseline_shapes_match_reference(golden_dir, name, mode):
    """Every arithmetic mode against the reference forward at the BASELINE shapes (thousands of halo-overlapped
    tiles per launch, the 2-CTA cluster kernels, tail splitting): max-abs over the sampled pixels and crops within
    the mode's tolerance, and row / column sums within tolerance x pixels-per-line."""
    if not os.path.exists(os.path.join(golden_dir, name)):
        pytest.skip(f"{name} not generated (oracle/make_golden.py big)")
    g = np.load(os.path.join(golden_dir, name))
    scale, seed, fseed = int(g["scale"]), int(g["seed"]), int(g["frame_seed"])
    B, H, W = (int(v) for v in g["shape"])
    x, y = orc.synthetic_frames(B, H, W, fseed)
    np.testing.assert_allclose(x.double().sum(dim=(1, 2, 3)).numpy(), g["x_sum"], rtol=0, atol=1e-9)   # same inputs
    net = _net(scale, seed, mode)
    with torch.no_grad():
        out = net(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    ref_kind = "fp64" if f"f{int(g['frames'][0])}_fp64_sample4" in g.files else "fp32"
    worst = 0.0
    for f in (int(v) for v in g["frames"]):
        fp = _fingerprint(out[f, 0])
        for k, v in fp.items():
            ref = g[f"f{f}_{ref_kind}_{k}"]
            err = float(np.abs(v.astype(np.float64) - ref).max())
            lim = TOL[mode] * (W if k == "row_sum" else H if k == "col_sum" else 1)
            worst = max(worst, err / (W if k == "row_sum" else H if k == "col_sum" else 1))
            assert err <= lim, (name, mode, f, k, err, lim)
    print(f"{name} {mode}: worst per-pixel error vs reference {ref_kind} {worst:.3e}")


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16", "tf32", "f16x3"])
def test_intermediate_taps_match_oracle(mode):
    """Layer-level attribution: encoder, stage, fusion tensors vs the oracle's taps."""
    sd = orc.synthetic_state_dict(4, 0)
    x, y = orc.synthetic_frames(2, 45, 70, 7)
    with torch.no_grad():
        _, taps = orc.forward(sd, x, y, return_taps=True)
    net = _net(4, 0, mode)
    with torch.no_grad():
        net(x.cuda(), y.cuda())
    eng = net.engine(torch.device("cuda", 0))
    rel = {"fp32": 1e-5, "tf32": 6e-3, "fp16": 3e-3, "bf16": 3e-2, "f16x3": 1e-5}[mode]
    want = {"enc": torch.cat((taps["enc_d"], taps["enc_c"]), 1),
            "feat": torch.cat((taps["out_d4"], taps["out_c4"]), 1),
            "fuse": taps["fuse"], "out_fuse": taps["out_fuse"]}
    for name, ref in want.items():
        got = eng.debug_tap(name, 2, 45, 70).cpu()
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        print(f"{mode} tap {name}: max-abs {err:.3e} (tensor max {scale:.3f})")
        assert err <= rel * scale, (name, err, scale)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_io_like_reference_driver(dtype):
    """model.cuda().half() with half frames (CODON_X4/test.py:52,122-123): mode follows the dtype."""
    sd = orc.synthetic_state_dict(4, 1)
    x, y = orc.synthetic_frames(1, 40, 56, 3)
    net = CODON_x4.CODONNet().eval()
    net.load_state_dict(sd)
    net = net.cuda().to(dtype)
    assert net.mode == {torch.float16: "fp16", torch.bfloat16: "bf16"}[dtype]
    with torch.no_grad():
        out = net(x.cuda().to(dtype), y.cuda().to(dtype))
        ref = orc.forward(sd, x.to(dtype).float(), y.to(dtype).float())
    assert out.dtype == dtype
    tol = 4e-3 if dtype == torch.float16 else 3e-2
    assert float((out.float().cpu() - ref).abs().max()) <= tol


def test_batch_invariance_and_determinism():
    """Frames are independent (CAC pools per sample): a frame's result does not depend on its batch
    neighbours or on the run -- bit-exact."""
    sd = orc.synthetic_state_dict(8, 2)
    x, y = orc.synthetic_frames(3, 50, 66, 11)
    for mode in ("fp32", "bf16"):
        net = _net(8, 2, mode)
        with torch.no_grad():
            full = net(x.cuda(), y.cuda()).clone()
            again = net(x.cuda(), y.cuda()).clone()
            single = net(x[1:2].cuda(), y[1:2].cuda()).clone()
        assert torch.equal(full, again)
        assert torch.equal(full[1:2], single)


def test_ragged_and_tiny_shapes():
    sd = orc.synthetic_state_dict(4, 0)
    net = _net(4, 0, "fp32")
    netb = _net(4, 0, "fp16")
    nets = _net(4, 0, "f16x3")
    for (h, w) in [(1, 1), (3, 5), (17, 16), (16, 33), (65, 31), (300, 1)]:
        x, y = orc.synthetic_frames(1, h, w, h * 100 + w)
        with torch.no_grad():
            ref = orc.forward(sd, x.double(), y.double()).float()
            out = net(x.cuda(), y.cuda()).cpu()
            outb = netb(x.cuda(), y.cuda()).cpu()
            outs = nets(x.cuda(), y.cuda()).cpu()
        assert float((out - ref).abs().max()) <= 1e-4, (h, w)
        assert float((outb - ref).abs().max()) <= 1e-3, (h, w)
        assert float((outs - ref).abs().max()) <= 2e-6, (h, w)


def test_narrow_frames_in_a_large_batch_use_the_cell_statistics_buffer_safely():
    """One-pixel-wide, tall frames in a batch large enough for the fused cluster kernels: the per-cell statistics chunks
    (one per eight 8 x 16-pixel cells) outnumber the 256-pixel chunks the partial buffer used to be sized for."""
    sd = orc.synthetic_state_dict(4, 0)
    x, y = orc.synthetic_frames(17, 300, 1, 31)
    with torch.no_grad():
        ref = orc.forward(sd, x[16:17].double(), y[16:17].double()).float()
    for mode, tol in (("bf16", 2e-2), ("tf32", 1e-3), ("f16x3", 2e-6)):
        net = _net(4, 0, mode)
        with torch.no_grad():
            out = net(x.cuda(), y.cuda())
        assert float((out[16:17].cpu() - ref).abs().max()) <= tol, mode


def test_cluster_kernels_ragged_frame_odd_tile_count():
    """203 x 331: large enough for the 2-CTA cluster kernels (>= 148 tiles per job), ragged in both directions
    (partial sub-tiles, TMA zero fill on every border) and with an ODD tile count per job (the peer CTA of the last
    pair recomputes a tile and stores nothing); the fusion stages fall to one accumulator per tile (NACC = 1)."""
    sd = orc.synthetic_state_dict(4, 1)
    x, y = orc.synthetic_frames(3, 203, 331, 77)
    with torch.no_grad():
        ref = orc.forward(sd, x[1:2].double(), y[1:2].double()).float()
    for mode, tol in (("fp16", 1e-3), ("tf32", 1e-3), ("bf16", 2e-2), ("f16x3", 2e-6)):
        net = _net(4, 1, mode)
        with torch.no_grad():
            single = net(x[1:2].cuda(), y[1:2].cuda()).clone()
            full = net(x.cuda(), y.cuda()).clone()
        err = float((single.cpu() - ref).abs().max())
        print(f"203x331 {mode}: {err:.3e}")
        assert err <= tol, (mode, err)
        # a different batch size changes the tile -> CTA schedule and the tail split, never the bits of a frame
        assert torch.equal(full[1:2], single), mode


def test_random_shapes_deterministic_batch_invariant_and_close_to_fp32_mode():
    """A short version of tools/gpu_stress.py: random frame shapes and batch sizes through the cluster kernels; every
    case is finite, bit-reproducible, batch-invariant and within tolerance of the fp32 FFMA mode."""
    rng = np.random.default_rng(7)
    nets = {m: _net(4, 1, m) for m in ("fp32", "fp16", "bf16", "tf32", "f16x3")}
    tol = {"fp16": 1e-3, "tf32": 1e-3, "bf16": 2e-2, "f16x3": 2e-6}
    for _ in range(12):
        B, H, W = int(rng.integers(1, 4)), int(rng.integers(150, 420)), int(rng.integers(150, 560))
        x, y = orc.synthetic_frames(B, H, W, int(rng.integers(1 << 30)))
        x, y = x.cuda(), y.cuda()
        with torch.no_grad():
            ref = nets["fp32"](x[:1], y[:1]).clone()
            for mode, t in tol.items():
                a = nets[mode](x, y).clone()
                b = nets[mode](x, y).clone()
                s = nets[mode](x[:1], y[:1]).clone()
                assert torch.isfinite(a).all(), (mode, B, H, W)
                assert torch.equal(a, b), ("nondeterministic", mode, B, H, W)
                assert torch.equal(a[:1], s), ("batch variance", mode, B, H, W)
                assert float((s - ref).abs().max()) <= t, (mode, B, H, W)


def test_host_entry_point_matches_device_entry_point():
    sd = orc.synthetic_state_dict(4, 0)
    x, y = orc.synthetic_frames(2, 33, 47, 5)
    net = _net(4, 0, "fp16")
    eng = net.engine(torch.device("cuda", 0))
    with torch.no_grad():
        dev = net(x.cuda(), y.cuda()).cpu().numpy()
    host = eng.forward_host(x.numpy(), y.numpy())
    np.testing.assert_array_equal(host, dev)
    assert eng.last_launch_count >= 30
    # page-locked caller buffers are copied directly (no staging): same bits
    xp, yp, op = (engine.Engine.pinned_frames(*x.shape) for _ in range(3))
    xp[...] = x.numpy()
    yp[...] = y.numpy()
    assert eng.forward_host(xp, yp, out=op) is op
    np.testing.assert_array_equal(op, dev)
    with pytest.raises(engine.CodonError):
        eng.forward_host(xp, yp, out=np.empty((1, 2, 3), np.float32))


def test_streaming_host_calls_match_blocking_calls():
    """submit_host / wait_host (copies of neighbouring calls overlap the kernels): same bits as forward_host, in
    order, for calls of different shapes (slot and workspace re-allocation in flight); call-order errors are loud."""
    net = _net(4, 0, "bf16")
    eng = net.engine(torch.device("cuda", 0))
    shapes = [(1, 40, 56), (2, 33, 47), (1, 96, 130), (1, 40, 56), (3, 17, 23), (1, 96, 130), (1, 64, 64)]
    calls, want = [], []
    for i, (B, H, W) in enumerate(shapes):
        x, y = orc.synthetic_frames(B, H, W, 100 + i)
        want.append(eng.forward_host(x.numpy(), y.numpy()).copy())
        xp, yp, op = (engine.Engine.pinned_frames(B, 1, H, W) for _ in range(3))
        xp[...] = x.numpy()
        yp[...] = y.numpy()
        op[...] = -1.0
        calls.append((xp, yp, op))
    got = list(eng.stream_host(calls))
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g is calls[i][2]
        np.testing.assert_array_equal(g, w, err_msg=f"call {i} {shapes[i]}")
    # call order
    with pytest.raises(engine.CodonError):
        eng.wait_host()                                              # nothing submitted
    eng.submit_host(*calls[0])
    eng.submit_host(*calls[1])
    with pytest.raises(engine.CodonError):
        eng.submit_host(*calls[2])                                   # two already in flight
    with pytest.raises(engine.CodonError):
        eng.forward_host(calls[0][0], calls[0][1])                   # blocking call while submits are outstanding
    eng.wait_host()
    eng.wait_host()
    np.testing.assert_array_equal(calls[1][2], want[1])
    with pytest.raises(engine.CodonError):                           # pageable buffers cannot be copied asynchronously
        eng.submit_host(np.zeros((1, 8, 8), np.float32), np.zeros((1, 8, 8), np.float32), np.zeros((1, 8, 8), np.float32))
    np.testing.assert_array_equal(eng.forward_host(calls[0][0], calls[0][1]), want[0])


def test_errors_are_loud():
    net = _net(4, 0, "fp32")
    x = torch.zeros(1, 1, 8, 8)
    with pytest.raises(engine.CodonError):
        net(x, x)                                   # CPU tensor
    with pytest.raises(engine.CodonError):
        net(x.cuda().expand(1, 2, 8, 8), x.cuda().expand(1, 2, 8, 8))   # not single-channel
    eng = engine.Engine(4, "fp32", 0)
    with pytest.raises(engine.CodonError):
        eng.forward(x.cuda(), x.cuda())             # weights not loaded
    with pytest.raises(engine.CodonError):
        eng.load_state_dict({"conv1.weight": torch.zeros(3, 3)})        # wrong shape


def test_full_size_linearity_property_free_checks():
    """640x480 (BASELINE config 2): finite output, zero-weight output conv gives identity (global
    residual, CODON_x4.py:131), and fp16/bf16 agree with fp32 mode."""
    sd = orc.synthetic_state_dict(4, 0)
    x, y = orc.synthetic_frames(1, 480, 640, 1234)
    xc, yc = x.cuda(), y.cuda()
    outs = {}
    for mode in ("fp32", "fp16", "bf16", "tf32"):
        net = _net(4, 0, mode)
        with torch.no_grad():
            outs[mode] = net(xc, yc)
    assert torch.isfinite(outs["fp32"]).all()
    for mode, tol in (("fp16", 1e-3), ("tf32", 1e-3), ("bf16", 2e-2)):
        err = float((outs[mode] - outs["fp32"]).abs().max())
        print(f"640x480 {mode} vs fp32 mode: {err:.3e}")
        assert err <= tol
    sd0 = dict(sd)
    sd0["output.weight"] = torch.zeros_like(sd["output.weight"])
    net = CODON_x4.CODONNet().eval().set_mode("bf16")
    net.load_state_dict(sd0)
    with torch.no_grad():
        assert torch.equal(net(xc, yc), xc)


def test_huge_activations_saturate_instead_of_overflowing():
    """The fused 5x5 -> ReLU -> 1x1 kernel stages the 128-channel intermediate in fp16 (tf32 / fp16 / f16x3 modes).  With
    weights scaled so that post-ReLU activations exceed the fp16 range (65504) the staging saturates: the output stays
    finite (the reference's own .half() GPU path would produce inf / nan here).  fp32-storage modes keep tracking the
    fp32 FFMA mode up to the saturation."""
    sd = dict(orc.synthetic_state_dict(4, 0))
    sd["conv_input.weight"] = sd["conv_input.weight"] * 1.0e6          # encoder outputs ~2e5 > 65504
    sd["conv_input_c.weight"] = sd["conv_input_c.weight"] * 1.0e6
    sd["output.weight"] = sd["output.weight"] * 1e-9
    x, y = orc.synthetic_frames(1, 160, 240, 5)
    outs = {}
    for mode in ("fp32", "tf32", "f16x3"):
        net = CODON_x4.CODONNet().eval().set_mode(mode)
        net.load_state_dict(sd)
        with torch.no_grad():
            outs[mode] = net(x.cuda(), y.cuda())
        eng = net.engine(torch.device("cuda", 0))
        peak = float(eng.debug_tap("enc", 1, 160, 240).abs().max())
        print(f"{mode}: encoder peak {peak:.3e}, output finite {bool(torch.isfinite(outs[mode]).all())}")
        assert torch.isfinite(outs[mode]).all(), mode
        if mode == "fp32":
            assert peak > 65504.0            # the regime the test is about
    assert float(eng.debug_tap("enc", 1, 160, 240).abs().max()) <= 2 * 65504.0      # split storage saturates (hi + lo)


def test_dataparallel_wrapper_keeps_engines_and_weights():
    """torch.nn.DataParallel(model) as in CODON_X16/test.py:52,132: the forward works through the wrapper on one GPU and,
    when two are visible, through its replicas -- which share the master's engines: weights are uploaded once per device,
    not once per forward (the weights generation of each engine stays put)."""
    sd = orc.synthetic_state_dict(16, 2)
    net = CODON_x16.CODONNet().eval().set_mode("bf16")
    net.load_state_dict(sd)
    ndev = min(torch.cuda.device_count(), 2)
    x, y = orc.synthetic_frames(2 * ndev, 64, 80, 99)
    with torch.no_grad():
        # reference: the same per-GPU batches on one GPU (DataParallel scatters the batch in equal chunks; a frame's bits
        # are batch-invariant as long as the batch takes the same kernel variants, which depend on the tile count)
        per = x.shape[0] // ndev
        ref = torch.cat([net(x[i:i + per].cuda(), y[i:i + per].cuda()).clone() for i in range(0, x.shape[0], per)])
        dp = torch.nn.DataParallel(net, device_ids=list(range(ndev))).cuda()
        assert set(dp.state_dict()) == {"module." + k for k in sd}
        first = dp(x.cuda(), y.cuda()).clone()      # (.cuda() above went through _apply: one legitimate re-upload)
        gens = {k: e.weights_generation for k, (e, _) in net._engines.items()}
        outs = [dp(x.cuda(), y.cuda()).clone() for _ in range(3)]
    for o in [first] + outs:
        assert torch.equal(o.cpu(), ref.cpu())                     # frames are independent: same bits on any GPU
    assert len(gens) == ndev, gens
    assert {k: e.weights_generation for k, (e, _) in net._engines.items()} == gens     # no upload per forward


def test_cuda_graph_replay_is_bit_exact():
    """The whole forward is capturable into a CUDA graph (no allocation, no sync inside codon_forward)."""
    sd = orc.synthetic_state_dict(4, 0)
    x, y = orc.synthetic_frames(1, 96, 176, 21)
    x2, y2 = orc.synthetic_frames(1, 96, 176, 22)
    for mode in ("bf16", "tf32", "f16x3"):
        net = _net(4, 0, mode)
        eng = net.engine(torch.device("cuda", 0))
        with torch.no_grad():
            direct = eng.forward(x.cuda(), y.cuda()).clone()
            direct2 = eng.forward(x2.cuda(), y2.cuda()).clone()
        g = eng.capture_graph(1, 96, 176)
        assert torch.equal(g(x.cuda(), y.cuda()), direct)
        assert torch.equal(g(x2.cuda(), y2.cuda()), direct2)
        assert torch.equal(g(x.cuda(), y.cuda()), direct)
        # reloading weights keeps the graph usable: same device allocations, and the replay re-captures (per-layer
        # constants such as the f16x3 weight scales live in the captured kernel parameters)
        sd2 = orc.synthetic_state_dict(4, 1)
        eng.load_state_dict(sd2)
        with torch.no_grad():
            want = eng.forward(x.cuda(), y.cuda()).clone()
        assert not torch.equal(want, direct)
        assert torch.equal(g(x.cuda(), y.cuda()), want)
        eng.load_state_dict(sd)


def test_1080p_x16_in_kernel_tiling_properties():
    """BASELINE config 4 shape (x16 class, 1920x1080): the in-kernel halo tiling covers ~8000 tiles per launch.
    Size-independent checks: finite, deterministic, batch-invariant (a frame alone == the same frame inside a
    batch, bit-exact: the CAC reductions are per frame and fixed-order), and the 16-bit / tf32 modes agree
    within the fp32-parity tolerance."""
    sd = orc.synthetic_state_dict(16, 2)
    x, y = orc.synthetic_frames(2, 1080, 1920, 3)
    xc, yc = x.cuda(), y.cuda()
    outs = {}
    for mode in ("tf32", "fp16"):
        net = CODON_x16.CODONNet().eval().set_mode(mode)
        net.load_state_dict(sd)
        with torch.no_grad():
            both = net(xc, yc).clone()
            single = net(xc[1:2], yc[1:2]).clone()
            again = net(xc, yc).clone()
        assert torch.isfinite(both).all()
        assert torch.equal(both, again)
        assert torch.equal(both[1:2], single)
        outs[mode] = both
    err = float((outs["tf32"] - outs["fp16"]).abs().max())
    print(f"1080p x16 tf32 vs fp16: {err:.3e}")
    assert err <= 1e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_one_process_same_bits():
    """Frame sharding over GPUs in one process (one host thread per GPU): every GPU produces
    the same bits for the same frame, concurrently."""
    from codon_b200 import scheduler as sch
    sd = orc.synthetic_state_dict(4, 0)
    x, y = orc.synthetic_frames(4, 130, 170, 9)
    net = _net(4, 0, "bf16")
    with torch.no_grad():
        ref = net(x.cuda(0), y.cuda(0)).cpu()

    def run(gpu, i):
        dev = torch.device("cuda", gpu)
        with torch.no_grad():
            o = net.engine(dev).forward(x[i:i + 1].to(dev), y[i:i + 1].to(dev))
        torch.cuda.synchronize(dev)
        return o.cpu()

    outs = sch.MultiGpuExecutor([0, 1]).map(run, list(range(4)))
    for i, o in enumerate(outs):
        assert torch.equal(o, ref[i:i + 1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_single_frame_over_gpus_matches_single_gpu(mode):
    """codon_group_*: one frame split into row bands over the GPUs (NVLink peer halo exchange after every layer,
    all-gathered CAC statistics).  Equals the single-GPU forward up to the summation order of the average pool."""
    n = min(torch.cuda.device_count(), 4)
    sd = orc.synthetic_state_dict(4, 0)
    for (h, w) in ((203, 176), (480, 640)):
        x, y = orc.synthetic_frames(1, h, w, 17)
        net = _net(4, 0, mode)
        with torch.no_grad():
            ref = net(x.cuda(0), y.cuda(0)).cpu().numpy()[0, 0]
        grp = engine.FrameGroup(4, mode, list(range(n)), sd)
        got = grp.forward_host(x.numpy()[0, 0], y.numpy()[0, 0])
        again = grp.forward_host(x.numpy()[0, 0], y.numpy()[0, 0])
        grp.close()
        err = float(np.abs(got - ref).max())
        print(f"{mode} {h}x{w} over {n} GPUs: max |band - single| = {err:.3e}")
        assert np.array_equal(got, again)
        # the gate differs by ~1e-7 (summation order); 16-bit / tf32 rounding then amplifies that to the mode's own rounding noise
        assert err <= (2e-6 if mode == "fp32" else 5e-4 if mode != "bf16" else 4e-3)
