"""CPU: the driver-facing contract of bench.py's reference arm (the CPU forward of the oracle, the one bench leg that
runs without a GPU) and the no-GPU behaviour of the product arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--steps", "1", "--warmup", "0", "--height", "32", "--width", "48"]


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", *SMALL])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hr_depth_megapixels_per_second" and d["unit"] == "MP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    # the reference's own classes when oracle/_ref has been built (oracle/build_ref.py), the oracle port otherwise
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "CODON_X4.zip"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # both arms print the same config object for the same command line (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    import argparse
    a = argparse.Namespace(scale=4, frames=1, height=32, width=48, mode="tf32")
    assert d["config"] == bench.config_dict(a)


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", *SMALL], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return   # covered by the GPU runs
    r = _run(SMALL)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_only_checkers_import_the_oracle():
    """oracle/ is test infrastructure: tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it; the product
    package and the tools never do."""
    import re
    allowed = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    offenders = []
    for d, dirs, files in os.walk(ROOT):
        dirs[:] = [x for x in dirs if x not in (".git", "gpurun_out", "build", "__pycache__", "tests", "oracle", "baseline")]
        for f in files:
            p = os.path.join(d, f)
            if f.endswith((".py", ".sh", ".cu", ".cuh", ".h", ".c")) and p not in allowed:
                if re.search(r"import\s+codon_oracle|from\s+codon_oracle|oracle/codon_oracle", open(p, errors="replace").read()):
                    offenders.append(os.path.relpath(p, ROOT))
    assert offenders == []
    # and inside bench.py the oracle is reached only from the CPU legs
    src = open(os.path.join(ROOT, "bench.py")).read()
    gpu_arm = src[src.index("def run_gpu("):src.index("def main(")]
    assert "codon_oracle" not in gpu_arm.replace("cpu_forward_sample", "")
