"""Live differential test of the oracle against the reference itself; runs only where /root/reference exists
(the build container).  Skipped on the GPU box."""
import numpy as np
import pytest

import oracle
from oracle import ref_shim
from conftest import random_case

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


def _ref_labels(cls, xy, eps, minPts, as_float=False):
    n = len(xy)
    mat = np.zeros((n, 3))
    mat[:, 0] = np.arange(n)
    mat[:, 1:] = xy
    if not as_float:
        mat = mat.astype("int")
    db = cls(mat, eps, minPts)
    lab = -np.ones(n, dtype=np.int32)
    for k, v in db.labels.items():
        lab[int(k)] = v
    return lab


def test_random_cases_both_algorithms():
    ref = ref_shim.load()
    rng = np.random.default_rng(99)
    for it in range(300):
        xy, eps, minPts = random_case(rng, it, nmax=300)
        assert np.array_equal(_ref_labels(ref.blockDBSCAN, xy, eps, minPts), oracle.block_dbscan(xy, eps, minPts)[0])
        assert np.array_equal(_ref_labels(ref.cDBSCAN, xy, eps, minPts), oracle.cdbscan2(xy, eps, minPts)[0])


def test_block_float_matrix_as_trans_path():
    """callTransLoops.py:43-47 passes a float64 matrix; labels must not change."""
    ref = ref_shim.load()
    rng = np.random.default_rng(5)
    for it in range(60):
        xy, eps, minPts = random_case(rng, it, nmax=200)
        assert np.array_equal(_ref_labels(ref.blockDBSCAN, xy, eps, minPts, as_float=True),
                              oracle.block_dbscan(xy, eps, minPts)[0])
