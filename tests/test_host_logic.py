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


def test_plan_batches_groups_by_shape_and_budget():
    shapes = [(480, 640)] * 5 + [(370, 463)] * 2 + [(1080, 1920)]
    b = sch.plan_batches(shapes, range(len(shapes)), max_pixels=2 * 480 * 640)
    assert [x.indices for x in b] == [[0, 1], [2, 3], [4], [5, 6], [7]]
    assert b[-1].shape == (1080, 1920)
    assert sch.max_pixels_for_budget(1800.0, 18_000_000) == 10_000


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


def test_logger_tees(capsys):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sub", "log.txt")
        lg = Logger(path)
        lg.write("hello\n")
        lg.flush()
        lg.close()
        assert open(path).read() == "hello\n"
    assert "hello" in capsys.readouterr().out
