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


def _ref_xy(ref, xy):
    return ref.ds.XY(xy[:, 0], xy[:, 1])


def _loop(ref, c):
    lp = ref.ds.Loop()
    lp.x_start, lp.x_end, lp.y_start, lp.y_end = (int(v) for v in c)
    lp.x_center, lp.y_center = (lp.x_start + lp.x_end) / 2, (lp.y_start + lp.y_end) / 2
    return lp


def test_counts_and_nearby_windows_random():
    """queryLoop / queryPeak / getLoopNearbyPETs of the reference (ds.py:176-190, callCisLoops.py:173-223) against the
    oracle's sorted-array counting, on random points and random boxes (overlapping anchors, boxes at the origin, empty
    regions), for several window counts."""
    ref = ref_shim.load()
    rng = np.random.default_rng(17)
    for it in range(25):
        n = int(rng.integers(200, 3000))
        span = int(rng.choice([3000, 20000, 200000]))
        x = rng.integers(0, span, n)
        y = x + rng.integers(0, span // 3 + 1, n)
        if it % 3 == 0:                                        # unordered rows, as a trans file has
            y = rng.integers(0, span, n)
        xy = np.stack([x, y], axis=1).astype(np.int64)
        L = 12
        xs = rng.integers(0, span, L)
        ys = xs + rng.integers(-span // 10, span // 2, L)
        ys = np.maximum(ys, 0)
        loops = np.stack([xs, xs + rng.integers(1, span // 20 + 2, L), ys, ys + rng.integers(1, span // 20 + 2, L)], 1)
        loops[0, 0] = 0                                        # windows that reach below zero are clamped (:194-195)
        rxy = _ref_xy(ref, xy)
        idx = oracle.XYIndex(xy)
        for win in (1, 3, 5):
            core, na, nb, nab, anc = idx.loop_counts(loops, win=win, mode=0)
            for i, c in enumerate(loops.tolist()):
                ra, rb, rab = rxy.queryLoop(c[0], c[1], c[2], c[3])
                assert (len(ra), len(rb), len(rab)) == tuple(core[i, :3].tolist()), (it, i)
                assert len(rxy.queryPeak(c[0], c[1])) == anc[i, 0] and len(rxy.queryPeak(c[2], c[3])) == anc[i, 2]
                rabs, nbps = ref.cis.getLoopNearbyPETs(_loop(ref, c), rxy, win)
                assert np.array_equal(np.asarray(rabs, dtype=np.int64), nab[i]), (it, i, win)


def test_stich_peaks_random():
    ref = ref_shim.load()
    rng = np.random.default_rng(3)
    for it in range(200):
        k = int(rng.integers(1, 10))
        s = rng.integers(0, 300, size=k)
        e = s + rng.integers(1, 40, size=k)
        margin = int(rng.choice([1, 2, 10, 50]))
        peaks = []
        for a, b in zip(s.tolist(), e.tolist()):
            p = ref.ds.Peak()
            p.chrom, p.start, p.end = "chr1", a, b
            peaks.append(p)
        want = [(p.start, p.end) for p in ref.geo.stichPeaks(peaks, margin=margin)]
        assert oracle.stich_peaks(list(zip(s.tolist(), e.tolist())), margin=margin) == want, it


def test_pre_line_parsers_random():
    """PET(line) / Pair(line) (ds.py:42-137) against the oracle's line handling on random fields."""
    ref = ref_shim.load()
    rng = np.random.default_rng(8)
    junk = ["", "*", "1.5", "+7", " 12", "12 ", "12a", "-5", "1e3"]
    chroms = ["chr1", "chr10", "chr2", "chrX"]

    def num():
        return junk[int(rng.integers(0, len(junk)))] if rng.random() < 0.1 else str(int(rng.integers(0, 10**6)))

    bed, prs = [], []
    for _ in range(1500):
        a, b = chroms[int(rng.integers(0, 4))], chroms[int(rng.integers(0, 4))]
        f = [a, num(), num(), b, num(), num(), "id", num(), "+", "-"]
        bed.append("\t".join(f[:int(rng.integers(6, 11))] if rng.random() < 0.1 else f))
        g = ["id", a, num(), b, num(), "+", "-", ["UU", "UR", "NN", "MM"][int(rng.integers(0, 4))]]
        prs.append("\t".join(g[:int(rng.integers(5, 9))] if rng.random() < 0.3 else g))
    rows, meta = oracle.pre_bedpe(bed, mapq=0)
    want = {}
    for line in bed:
        try:
            p = ref.ds.PET(line.split("\t"))
        except Exception:
            continue
        if p.mapq < 0:
            continue
        want.setdefault("%s-%s" % (p.chromA, p.chromB), set()).add((p.cA, p.cB))
    assert rows == {k: sorted(v) for k, v in want.items()}
    rows, meta = oracle.pre_pairs(prs)
    want = {}
    for line in prs:
        d = line.split("\t")
        if len(d) == 8 and d[7] not in ["UU", "MR", "MU", "RU", "UR"]:
            continue
        if len(d) not in (7, 8):
            continue
        try:
            p = ref.ds.Pair(d)
        except Exception:
            continue
        want.setdefault("%s-%s" % (p.chromA, p.chromB), set()).add((p.cA, p.cB))
    assert rows == {k: sorted(v) for k, v in want.items()}
