"""CPU: host-side logic -- frame sharding / batching, the gloo metric reduction (world_size 2),
checkpoint import, Logger."""
import os
import sys
import tempfile
import types

import pytest
import torch
import torch.multiprocessing as mp

import codon_oracle as orc
from codon_b200 import checkpoint as ck
from codon_b200 import scheduler as sch
from codon_b200.Loger import Logger


def test_shard_indices_partition():
    for n in (0, 1, 7, 10, 64, 257):
        for world in (1, 2, 3, 8):
            parts = [sch.shard_indices(n, world, r) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sch.shard_indices(4, 2, 2)


def _reduce_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = list(range(10))
    mine = sch.shard_indices(len(items), world, rank)
    rm = sum(1.5 * i for i in mine)
    ss = sum(0.01 * i for i in mine)
    out = sch.reduce_metric_sums(rm, ss, len(mine))
    q.put((rank, out))
    dist.destroy_process_group()


def test_metric_reduction_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (sum(1.5 * i for i in range(10)), sum(0.01 * i for i in range(10)), 10)
    for _, out in got:
        assert out[2] == want[2] and abs(out[0] - want[0]) < 1e-12 and abs(out[1] - want[1]) < 1e-12


def test_checkpoint_import_reference_style():
    """{"epoch", "model": pickled DataParallel(module)} whose defining module is gone at load time."""
    from codon_b200.CODON_x16 import CODONNet
    sd = orc.synthetic_state_dict(16, 1)
    mod = types.ModuleType("model_zoo_gone")
    sys.modules["model_zoo_gone"] = mod
    src = ("import torch.nn as nn\n"
           "class Holder(nn.Module):\n"
           "    def __init__(self, sd):\n"
           "        super().__init__()\n"
           "        import torch\n"
           "        for k, v in sd.items():\n"
           "            parts = k.split('.')\n"
           "            m = self\n"
           "            for p in parts[:-1]:\n"
           "                if p not in m._modules:\n"
           "                    m.add_module(p, nn.Module())\n"
           "                m = m._modules[p]\n"
           "            m.register_parameter(parts[-1], nn.Parameter(v.clone()))\n")
    exec(src, mod.__dict__)
    holder = torch.nn.DataParallel(mod.Holder(sd))
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "X16.pth")
        torch.save({"epoch": 93, "model": holder}, p)
        del sys.modules["model_zoo_gone"]
        got, meta = ck.load_checkpoint(p)
        assert meta["epoch"] == 93 and ck.infer_scale(got) == "x16"
        assert got.keys() == sd.keys() and all(torch.equal(got[k], sd[k]) for k in sd)
        CODONNet().load_state_dict(got, strict=True)
        p2 = os.path.join(d, "plain.pth")
        torch.save({"module." + k: v for k, v in sd.items()}, p2)
        got2, _ = ck.load_checkpoint(p2)
        assert got2.keys() == sd.keys()
    assert ck.infer_scale(orc.synthetic_state_dict(4, 0)) == "x4/x8"


def test_flat_weight_file_roundtrip(tmp_path):
    """checkpoint.export_flat / load_flat: the file codon_load_weights_file reads (layout pinned byte for byte)."""
    import struct
    sd = {"module." + k: v for k, v in orc.synthetic_state_dict(16, 3).items()}
    path = str(tmp_path / "w.codonw")
    assert ck.export_flat(sd, path) == 44
    back = ck.load_flat(path)
    assert list(back) == [k[len("module."):] for k in sd]
    assert all(torch.equal(back[k[len("module."):]], v) for k, v in sd.items())
    raw = open(path, "rb").read()
    assert raw[:8] == b"CODONW1\0" and struct.unpack("<I", raw[8:12])[0] == 44
    first = next(iter(back))
    (ln,) = struct.unpack("<H", raw[12:14])
    assert raw[14:14 + ln].decode() == first and raw[14 + ln] == back[first].dim()
    with pytest.raises(ValueError):
        open(path, "wb").write(b"garbage!" + raw[8:])
        ck.load_flat(path)


def test_unpickler_does_not_import_foreign_modules(tmp_path):
    """A checkpoint naming a class of an importable non-torch module must get a stub, not an import (the module is
    never executed), and the tensors must still come out."""
    marker = tmp_path / "imported.flag"
    pkg = tmp_path / "evil_zoo.py"
    pkg.write_text("import pathlib\n"
                   f"pathlib.Path({str(marker)!r}).write_text('x')\n"
                   "import torch.nn as nn\n"
                   "class Net(nn.Module):\n"
                   "    def __init__(self):\n"
                   "        super().__init__()\n"
                   "        self.conv1 = nn.Conv2d(64, 64, 3, bias=False)\n")
    sys.path.insert(0, str(tmp_path))
    try:
        import importlib
        zoo = importlib.import_module("evil_zoo")
        net = zoo.Net()
        p = str(tmp_path / "m.pth")
        torch.save({"epoch": 1, "model": net}, p)
        del sys.modules["evil_zoo"]
        marker.unlink()
        got, meta = ck.load_checkpoint(p)                 # evil_zoo is importable from sys.path -- and must not be
        assert not marker.exists() and "evil_zoo" not in sys.modules
        assert list(got) == ["conv1.weight"] and torch.equal(got["conv1.weight"], net.conv1.weight.data)
    finally:
        sys.path.remove(str(tmp_path))


def test_model_weight_versioning_without_gpu():
    """CODONNet tracks weight changes by version (load_state_dict, .half()/.to(), in-place ops on the parameters,
    refresh_weights() for edits through .data); DataParallel replicas trust the master's counter."""
    from codon_b200.CODON_x4 import CODONNet
    net = CODONNet().eval()
    k0 = net._weights_key()
    assert net._weights_key() == k0
    net.load_state_dict(orc.synthetic_state_dict(4, 0))
    k1 = net._weights_key()
    assert k1 != k0
    with torch.no_grad():
        net.conv3.weight.mul_(2.0)
    k2 = net._weights_key()
    assert k2 != k1
    net.conv3.weight.data.mul_(0.5)                      # not versioned by PyTorch ...
    assert net._weights_key() == k2
    net.refresh_weights()                                # ... hence the explicit call
    k3 = net._weights_key()
    assert k3 != k2
    net.half()
    assert net._weights_key() != k3 and net.mode == "fp16"
    rep = net._replicate_for_data_parallel()
    rep._is_replica = True
    assert rep._weights_key() == (net._wver, None) and rep._engines is net._engines
    import copy
    import pickle
    assert copy.deepcopy(net)._weights_key()[0] == net._wver
    assert pickle.loads(pickle.dumps(net)).state_dict().keys() == net.state_dict().keys()


def test_logger_tees(capsys):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sub", "log.txt")
        lg = Logger(path)
        lg.write("hello\n")
        lg.flush()
        lg.close()
        assert open(path).read() == "hello\n"
    assert "hello" in capsys.readouterr().out
