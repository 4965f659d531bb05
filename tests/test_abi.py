"""CPU: the C-ABI library builds, loads, and exports every symbol include/codon_b200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from codon_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "codon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(codon_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(engine.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.codon_version()


def test_host_selftest_weight_packing_and_tap_schedules(lib):
    # host code only: the packed K-major SWIZZLE_128B weight image and the compile-time tap schedule of every
    # cluster-kernel kind agree with the plans the engine builds
    assert lib.codon_selftest() == 0


def test_no_torch_types_in_signatures():
    src = open(os.path.join(ROOT, "include", "codon_b200.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S).lower()
    assert "at::" not in src and "c10::" not in src


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    ctx = ctypes.c_void_p()
    rc = lib.codon_create(ctypes.byref(ctx), 0, 4, 0)
    assert rc == -3 and not ctx.value
    assert b"no CPU fallback" in lib.codon_last_error(None)
    from codon_b200.CODON_x4 import CODONNet
    net = CODONNet().eval()
    x = torch.zeros(1, 1, 8, 8)
    with pytest.raises(engine.CodonError):
        net(x, x)
    with pytest.raises(engine.CodonError):
        net.attention_c0(torch.zeros(1, 128, 4, 4))


def test_argument_validation(lib):
    assert lib.codon_create(None, 0, 4, 0) == -1
    ctx = ctypes.c_void_p()
    assert lib.codon_create(ctypes.byref(ctx), 0, 5, 0) == -1
    assert lib.codon_create(ctypes.byref(ctx), 0, 4, 9) == -1
    assert lib.codon_workspace_bytes(None, 1, 8, 8) == 0


def test_module_surface_matches_reference_inventory():
    import codon_oracle as orc
    from codon_b200 import CODON_x4, CODON_x8, CODON_x16
    for mod, scale in ((CODON_x4, 4), (CODON_x8, 8), (CODON_x16, 16)):
        net = mod.CODONNet()
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert got == orc.param_shapes(scale)
        # a reference state_dict (and one saved through DataParallel) loads strictly
        sd = orc.synthetic_state_dict(scale, 0)
        net.load_state_dict(sd, strict=True)
        dp = torch.nn.DataParallel(mod.CODONNet())
        dp.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=True)
    assert CODON_x4.CODONNet().half().mode == "fp16"
    assert CODON_x4.CODONNet().bfloat16().mode == "bf16"
    assert CODON_x4.CODONNet().mode == "fp32"


def _build_c_consumer(tmp_path, args=()):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    engine.load_library()                                            # builds the library if needed
    exe = str(tmp_path / "c_consumer")
    libdir = os.path.join(ROOT, "codon_b200")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_consumer.c"), "-o", exe, "-L" + libdir, "-lcodon_b200", "-Wl,-rpath," + libdir,
           "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr                               # the header is plain C99, no C++-isms
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_plain_c_consumer_links_and_fails_loudly_without_gpu(tmp_path):
    r = _build_c_consumer(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no device: rc=-3" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_plain_c_consumer_on_gpu(tmp_path):
    r = _build_c_consumer(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "refused as documented" in r.stdout and r.stdout.strip().endswith("ok")


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [(4, 2e-6), (0, 1e-4), (2, 1e-3), (1, 2e-2)])
def test_forward_from_plain_c_matches_reference(tmp_path, mode, tol):
    """A non-Python host end to end: the flat weight file exported by checkpoint.export_flat is
    loaded by codon_load_weights_file, codon_forward_host runs on the committed frames, and the C program itself
    compares with the real reference's output (tests/golden/c_case_*.bin, oracle/make_golden.py)."""
    import codon_oracle as orc
    from codon_b200 import checkpoint as ck
    wpath = str(tmp_path / "x4_s0.codonw")
    assert ck.export_flat(orc.synthetic_state_dict(4, 0), wpath) == 49
    case = os.path.join(ROOT, "tests", "golden", "c_case_x4_s0_b2_48x64.bin")
    r = _build_c_consumer(tmp_path, (wpath, case, str(mode), repr(tol)))
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "forward from C: 2 x 48 x 64" in r.stdout and r.stdout.strip().endswith("ok")
