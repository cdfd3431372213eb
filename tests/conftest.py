"""Shared fixtures.  Marker tiers follow the reference's layout (its pyproject.toml:94-101 has
slow/integration/external); here the split is `gpu` (needs a B200, run through gpurun) vs the
default CPU suite."""

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on a B200")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests skip (rather than fail inside lxg_init) on a box without a CUDA device, so a plain
    `pytest tests` is green on the CPU-only build container too."""
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (sm_100a); run with -m gpu on a B200")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_corpus(n, d, seed=0, dtype=np.float16):
    """BASELINE.md synthetic corpus: N(0,1) rows, L2-normalised in fp32, cast to `dtype`."""
    c = np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    return c.astype(dtype)


def make_queries(nq, d, seed=1):
    """BASELINE.md synthetic queries: N(0,1), left un-normalised."""
    return np.random.default_rng(seed).standard_normal((nq, d), dtype=np.float32)


@pytest.fixture(scope="session")
def lxg():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lean_explore_b200 import _lib

    return _lib.init(0)
