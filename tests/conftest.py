import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def random_in_csr(V, nnz, seed, with_eids=True, hub=False):
    """Random directed multigraph-free graph as an in-CSR (indptr, indices, eids) in COO edge-id order."""
    import scipy.sparse as spsp
    rng = np.random.default_rng(seed)
    src = rng.integers(0, V, nnz)
    dst = rng.integers(0, V, nnz) if not hub else np.minimum(rng.geometric(0.02, nnz) - 1, V - 1)
    key = np.unique(src * V + dst)
    rng.shuffle(key)
    src, dst = key // V, key % V
    order = np.argsort(dst, kind="stable")
    indptr = np.zeros(V + 1, np.int64)
    np.cumsum(np.bincount(dst, minlength=V), out=indptr[1:])
    coo = spsp.coo_matrix((np.ones(len(src), np.int64), (src, dst)), shape=(V, V))
    return indptr, src[order].astype(np.int64), (order.astype(np.int64) if with_eids else None), coo


@pytest.fixture
def golden_dir():
    return GOLDEN
