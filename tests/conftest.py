import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def golden_cis_files():
    g = load_golden("cis_pipeline.npz")
    files = {}
    for k in g["order"].tolist():
        files[k] = g[k.replace("-", "_")]
    return g, files


def golden_trans_files():
    g = load_golden("trans_pipeline.npz")
    files = {}
    for k in g["order"].tolist():
        files[k] = g[k.replace("-", "_")]
    return g, files


def txt_rows(txt):
    return [line.split("\t") for line in str(txt).rstrip("\n").split("\n")][1:]


def random_case(rng, it, nmax=600):
    """Small adversarial point sets: blobs, sparse, duplicate-heavy, unordered rows."""
    n = int(rng.integers(1, nmax))
    span = int(rng.choice([20, 50, 200, 1000, 5000]))
    eps = int(rng.integers(1, 60))
    minPts = int(rng.integers(2, 14))
    kind = it % 4
    if kind == 0:
        k = int(rng.integers(1, 7))
        cx = rng.integers(0, span, size=k)
        cy = cx + rng.integers(0, span, size=k)
        r = rng.integers(0, k, size=n)
        s = eps * rng.uniform(0.3, 2.5)
        x = np.abs(np.rint(rng.normal(cx[r], s))).astype(np.int64)
        y = np.abs(np.rint(rng.normal(cy[r], s))).astype(np.int64)
        xy = np.stack([np.minimum(x, y), np.maximum(x, y)], 1)
    elif kind == 1:
        x = rng.integers(0, span, size=n)
        xy = np.stack([x, x + rng.integers(0, span, size=n)], 1)
    elif kind == 2:
        x = rng.integers(0, max(2, span // 10), size=n) * 3
        xy = np.stack([x, x + rng.integers(0, max(2, span // 10), size=n) * 2], 1)
    else:
        xy = np.stack([rng.integers(0, span, size=n), rng.integers(0, span, size=n)], 1)
    return np.ascontiguousarray(xy.astype(np.int64)), eps, minPts


@pytest.fixture(scope="session")
def gpu_ctx():
    from cloops2_b200 import _lib
    return _lib.default_context(0)
