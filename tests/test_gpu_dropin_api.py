"""The reference-named work-unit functions (what joblib dispatches in the reference: callCisLoops.py:134-137, 561-572;
callTransLoops.py:95-97, 237-243) called the way the reference calls them, on real .ixy files."""
import numpy as np
import pytest

import oracle
from cloops2_b200 import synth
from cloops2_b200 import callCisLoops as cis
from cloops2_b200 import callTransLoops as trans

pytestmark = pytest.mark.gpu

FIELDS = ["chromX", "chromY", "x_start", "x_end", "x_center", "ra", "y_start", "y_end", "y_center", "rb", "rab", "cis",
          "distance", "density", "ES", "P2LL", "FDR", "hypergeometric_p_value", "poisson_p_value", "binomial_p_value",
          "x_peak_poisson_p_value", "x_peak_es", "y_peak_poisson_p_value", "y_peak_es"]


def _same(a, b, fields):
    assert len(a) == len(b)
    for la, lb in zip(a, b):
        for f in fields:
            assert getattr(la, f) == getattr(lb, f), f


def test_cis_work_units(gpu_ctx, tmp_path):
    files = {"chr21-chr21": synth.cis_chrom(6_000_000, 60_000, 3, "hitrac")}
    synth.write_ixy_dir(files, str(tmp_path))
    fixy = str(tmp_path / "chr21-chr21.ixy")
    xy = files["chr21-chr21"]
    for cut, mcut in ((0, -1), (1500, 1_000_000)):
        key, loops = cis.runCisDBSCANLoops(fixy, 500, 5, cut=cut, mcut=mcut)
        xyf = oracle.parse_ixy_filter(xy, cut, mcut)
        wkey, want = oracle.cis_candidates(("chr21", "chr21"), xyf, 500, 5, cut=cut)
        assert key == wkey == "chr21-chr21"
        _same(loops, want, ["chromX", "chromY", "x_start", "x_end", "x_center", "y_start", "y_end", "y_center", "rab",
                            "cis", "distance"])
        key2, est = cis.estLoopSig(key, loops, fixy, len(xy), minPts=5, cut=cut, mcut=mcut)
        _, west = oracle.est_loop_sig_cis(key, want, xyf, len(xy), minPts=5)
        _same(est, west, FIELDS)
        assert not hasattr(est[0], "significant")            # set by markSigLoops, as in the reference
        _, marked = cis.markSigLoops(key, est)
        wm = oracle.mark_sig_cis(west)
        assert [lp.significant for lp in marked] == [lp.significant for lp in wm]
        _, sel = cis.selSigLoops(key, marked)
        assert [(lp.x_start, lp.y_end) for lp in sel] == [(lp.x_start, lp.y_end) for lp in oracle.sel_sig_loops(wm)]
    # a trans file name makes the cis work unit return None (callCisLoops.py:73-74)
    synth.write_ixy_dir({"chr21-chr22": xy[:100]}, str(tmp_path / "t"))
    assert cis.runCisDBSCANLoops(str(tmp_path / "t" / "chr21-chr22.ixy"), 500, 5) is None
    # nothing left after the distance cut -> None (callCisLoops.py:80-83)
    assert cis.runCisDBSCANLoops(fixy, 500, 5, cut=10**9) is None


def test_trans_work_units(gpu_ctx, tmp_path):
    files = {"chr20-chr22": synth.trans_pair(1_500_000, 2_400_000, 40_000, 12, sigma=800, per=1500)}
    synth.write_ixy_dir(files, str(tmp_path))
    fixy = str(tmp_path / "chr20-chr22.ixy")
    xy = files["chr20-chr22"]
    key, loops = trans.runTransDBSCANLoops(fixy, 3000, 8)
    wkey, want = oracle.trans_candidates(("chr20", "chr22"), xy, 3000, 8)
    assert key == wkey
    _same(loops, want, ["chromX", "chromY", "x_start", "x_end", "y_start", "y_end", "rab", "cis", "distance"])
    _, est = trans.estLoopSig(key, loops, fixy, minPts=8)
    _, west = oracle.est_loop_sig_trans(key, want, xy, minPts=8)
    _same(est, west, FIELDS)
    _, marked = trans.markSigLoops(key, est)
    assert [lp.significant for lp in marked] == [lp.significant for lp in oracle.mark_sig_trans(west)]


def test_cdbscan_class_protocol(gpu_ctx):
    from cloops2_b200.cDBSCAN2 import cDBSCAN
    xy = synth.cis_chrom(2_000_000, 20_000, 9, "hic")
    mat = np.zeros((len(xy), 3), dtype=np.int64)
    mat[:, 0] = np.arange(len(xy))
    mat[:, 1:] = xy
    db = cDBSCAN(mat, 2000, 10)
    want, _ = oracle.cdbscan2(xy, 2000, 10)
    assert db.labels == {i: int(v) for i, v in enumerate(want.tolist()) if v >= 0}
