"""`callDiffLoops.quantDloops` and the clustering of `callPeaks` (SURVEY.md 8f "next" rows): oracle restatements vs
the reference's golden output (CPU), the interval form of stichPeaks vs the position replay (CPU), and the GPU
implementation vs the golden output."""
import numpy as np
import pytest

import oracle
from conftest import golden_cis_files, load_golden

DL_CASES = ((1, 0), (2, 3000))
PEAK_CASES = (("e200", dict(eps=200, minPts=3, mcut=1000)), ("e300", dict(eps=300, minPts=4, mcut=2000)),
              ("split", dict(eps=150, minPts=5, mcut=1000, split=True, splitExt=50)))
DL_FIELDS = ["raw_trt_ra", "raw_trt_rb", "raw_con_ra", "raw_con_rb", "raw_trt_rab", "raw_con_rab", "raw_trt_mrab",
             "raw_con_mrab", "trt_es", "con_es", "distance", "size"]


@pytest.mark.parametrize("win,cut", DL_CASES)
def test_oracle_quant_dloops_matches_reference(win, cut):
    d = load_golden("dloops.npz")
    _, files = golden_cis_files()
    a, b = files["chr21-chr21"], d["control"]
    assert d["fields"].tolist() == DL_FIELDS
    ts, cs, rows = oracle.quant_dloops(d["loops"], oracle.parse_ixy_filter(a, cut), oracle.parse_ixy_filter(b, cut), win=win)
    assert np.array_equal(np.array(ts), d["ts_w%d" % win])
    assert np.array_equal(np.array(cs), d["cs_w%d" % win])
    assert np.array_equal(np.array(rows, dtype=np.float64), d["rows_w%d" % win][:, :10])


@pytest.mark.parametrize("name,kw", PEAK_CASES)
def test_oracle_dbscan_peaks_match_reference(name, kw):
    p = load_golden("peaks.npz")
    _, files = golden_cis_files()
    kw = dict(kw)
    xy = oracle.parse_ixy_filter(files["chr21-chr21"], 0, kw.pop("mcut"))
    _, st, _ = oracle.dbscan_peaks(xy, **kw)
    got = np.array([[s, e, e - s + 1] for s, e in st], dtype=np.int64).reshape(-1, 3)
    assert len(got) > 0 and np.array_equal(got, p[name])


def test_stich_peaks_interval_form_equals_position_replay():
    from cloops2_b200.callPeaks import stichPeaks
    from cloops2_b200.ds import Peak
    rng = np.random.default_rng(5)
    for it in range(300):
        k = int(rng.integers(0, 12))
        s = rng.integers(0, 200, size=k)
        e = s + rng.integers(0, 25, size=k) * (rng.random(k) < 0.85)      # some single-position intervals
        margin = int(rng.choice([0, 1, 2, 5, 30]))
        peaks = []
        for a, b in zip(s.tolist(), e.tolist()):
            p = Peak()
            p.chrom, p.start, p.end = "chr1", int(a), int(b)
            peaks.append(p)
        want = oracle.stich_peaks([(p.start, p.end) for p in peaks], margin=margin)
        got = stichPeaks(peaks, margin=margin)
        assert [(p.start, p.end) for p in got] == want, (it, margin, list(zip(s.tolist(), e.tolist())))
        assert all(p.length == p.end - p.start + 1 and p.chrom == "chr1" for p in got)


@pytest.mark.gpu
@pytest.mark.parametrize("win,cut", DL_CASES)
def test_gpu_quant_dloops(gpu_ctx, tmp_path, win, cut):
    from cloops2_b200 import synth
    from cloops2_b200.callDiffLoops import quantDloops
    from cloops2_b200.quant import parseTxt2Loops
    d = load_golden("dloops.npz")
    _, files = golden_cis_files()
    synth.write_ixy_dir({"chr21-chr21": files["chr21-chr21"]}, str(tmp_path / "a"))
    synth.write_ixy_dir({"chr21-chr21": d["control"]}, str(tmp_path / "b"))
    (tmp_path / "loops.txt").write_text(str(d["loops_txt"]))
    loops = parseTxt2Loops(str(tmp_path / "loops.txt"))["chr21-chr21"]
    assert np.array_equal(np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops]), d["loops"])
    key, ts, cs, dl = quantDloops("chr21-chr21", loops, str(tmp_path / "a" / "chr21-chr21.ixy"),
                                  str(tmp_path / "b" / "chr21-chr21.ixy"), cut=cut, win=win, ctx=gpu_ctx)
    assert key == "chr21-chr21"
    assert np.array_equal(np.array(ts, dtype=np.float64), d["ts_w%d" % win])
    assert np.array_equal(np.array(cs, dtype=np.float64), d["cs_w%d" % win])
    got = np.array([[float(getattr(x, f)) for f in DL_FIELDS] for x in dl], dtype=np.float64)
    assert np.array_equal(got, d["rows_w%d" % win])
    assert [x.id for x in dl] == [lp.id for lp in loops]
    assert quantDloops("chr21-chr21", [], "unused", "unused", ctx=gpu_ctx) is None


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", PEAK_CASES)
def test_gpu_dbscan_peaks(gpu_ctx, tmp_path, name, kw):
    from cloops2_b200 import synth
    from cloops2_b200.callPeaks import runDBSCANPeaks
    p = load_golden("peaks.npz")
    _, files = golden_cis_files()
    synth.write_ixy_dir({"chr21-chr21": files["chr21-chr21"], "chr21-chr22": files["chr21-chr21"][:100]}, str(tmp_path / "d"))
    kw = dict(kw)
    peaks = runDBSCANPeaks(str(tmp_path / "d" / "chr21-chr21.ixy"), kw.pop("eps"), kw.pop("minPts"), ctx=gpu_ctx, **kw)
    got = np.array([[q.start, q.end, q.length] for q in peaks], dtype=np.int64).reshape(-1, 3)
    assert np.array_equal(got, p[name])
    assert all(q.chrom == "chr21" for q in peaks)
    assert runDBSCANPeaks(str(tmp_path / "d" / "chr21-chr22.ixy"), 200, 3, ctx=gpu_ctx) is None     # trans file
    assert runDBSCANPeaks(str(tmp_path / "d" / "chr21-chr21.ixy"), 200, 3, cut=10**9, ctx=gpu_ctx) is None   # no PETs
