import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """GPU tests never silently pass on a box without a GPU: they are skipped, loudly."""
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device: codon_b200 has no CPU path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def lib():
    """The in-tree shared library, built on demand (nvcc cross-compiles without a GPU)."""
    from codon_b200 import build, engine
    build.build_library()
    return engine.load_library()
