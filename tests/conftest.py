import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _paths  # noqa: E402,F401

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def ref_pinned():
    return np.load(os.path.join(GOLDEN, "reference_pinned.npz"))


@pytest.fixture(scope="session")
def oracle_frozen():
    return np.load(os.path.join(GOLDEN, "oracle_frozen.npz"))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def small_cloud(seed, n, spread=1.2, q=0.05, batch=None):
    """Unique voxel coords [M,3] (or [M,4] with a batch column) of a random blob."""
    from oracle import quantize as oq
    rng = np.random.default_rng(seed)
    p = rng.normal(0, spread, (n, 3)).astype(np.float32)
    c = oq.sparse_quantize_me(p, q)[0]
    if batch is not None:
        c = np.concatenate([np.full((c.shape[0], 1), batch, np.int32), c], 1)
    return c
