"""`filterPETs -knn` (second "next" row): blockDBSCAN's noise removal as a PET filter."""
import json

import numpy as np
import pytest

import oracle
from conftest import golden_cis_files, load_golden


def test_oracle_noise_filter_matches_reference_rows():
    """The oracle's keep flags select exactly the rows the reference writes (as a set; order is checked on the GPU)."""
    k = load_golden("filter_knn.npz")
    _, files = golden_cis_files()
    for key, xy in files.items():
        keep = oracle.block_noise_keep(xy, int(k["eps"]), int(k["minPts"]))
        want = k["knn_" + key.replace("-", "_")]
        assert keep.sum() == len(want)
        assert np.array_equal(np.unique(xy[keep], axis=0), np.unique(want, axis=0))


@pytest.mark.gpu
def test_gpu_filter_knn_files_identical(gpu_ctx, tmp_path):
    import joblib
    from cloops2_b200 import synth
    from cloops2_b200.filter import filterPETsByKNNs
    k = load_golden("filter_knn.npz")
    _, files = golden_cis_files()
    d, o = str(tmp_path / "d"), str(tmp_path / "o")
    synth.write_ixy_dir(files, d)
    (tmp_path / "o").mkdir()
    filterPETsByKNNs(d, o, eps=int(k["eps"]), minPts=int(k["minPts"]))
    tot = 0
    for key in files:
        got = joblib.load(o + "/%s.ixy" % key)
        assert np.array_equal(got, k["knn_" + key.replace("-", "_")]), key      # rows AND row order
        tot += len(got)
    meta = json.load(open(o + "/petMeta.json"))
    assert meta["Unique PETs"] == tot and sorted(meta["data"]["cis"]) == sorted(files)
