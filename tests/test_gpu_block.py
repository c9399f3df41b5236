"""blockDBSCAN on the GPU through the C ABI vs the oracle / the reference's golden labels.  Bit-exact, including
cluster numbering."""
import numpy as np
import pytest

import oracle
from cloops2_b200 import _lib, synth
from conftest import load_golden, random_case

pytestmark = pytest.mark.gpu


def gpu_block(ctx, xy, eps, minPts):
    pts = _lib.Points(ctx, xy)
    try:
        if pts.n == 0:
            return np.zeros(0, np.int32), 0, np.zeros((0, 4), np.int64), np.zeros(0, np.int64)
        lab, ncl = pts.block_dbscan(eps, minPts)
        bbox, size = pts.cluster_bbox()
        return lab, ncl, bbox, size
    finally:
        pts.close()


def test_golden_labels(gpu_ctx):
    g = load_golden("dbscan_small.npz")
    offs = g["offsets"]
    for i, (eps, minPts) in enumerate(g["params"].tolist()):
        xy = g["xy"][offs[i]:offs[i + 1]]
        lab, ncl, _, _ = gpu_block(gpu_ctx, xy, eps, minPts)
        assert np.array_equal(lab, g["labels_block"][offs[i]:offs[i + 1]]), (i, eps, minPts)


def test_random_small_cases_labels_bbox(gpu_ctx):
    rng = np.random.default_rng(2026)
    for it in range(400):
        xy, eps, minPts = random_case(rng, it)
        want, wn = oracle.block_dbscan(xy, eps, minPts)
        lab, ncl, bbox, size = gpu_block(gpu_ctx, xy, eps, minPts)
        assert ncl == wn and np.array_equal(lab, want), (it, len(xy), eps, minPts)
        wb, ws = oracle.cluster_bbox(xy, want, wn)
        assert np.array_equal(bbox, wb) and np.array_equal(size, ws), it


def test_noise_removal_only(gpu_ctx):
    rng = np.random.default_rng(7)
    for it in range(60):
        xy, eps, minPts = random_case(rng, it)
        pts = _lib.Points(gpu_ctx, xy)
        try:
            keep = pts.block_noise_keep(eps, minPts)
        finally:
            pts.close()
        assert np.array_equal(keep, oracle.block_noise_keep(xy, eps, minPts)), it


@pytest.mark.parametrize("profile,n,length,eps,minPts", [
    ("hitrac", 300_000, 46_709_983, 200, 5), ("hitrac", 300_000, 46_709_983, 1000, 5),
    ("hitrac", 300_000, 20_000_000, 500, 10), ("hic", 200_000, 3_000_000, 5000, 20),
    ("hic", 200_000, 3_000_000, 10000, 50), ("chiapet", 250_000, 30_000_000, 2000, 3)])
def test_synthetic_chromosomes(gpu_ctx, profile, n, length, eps, minPts):
    xy = synth.cis_chrom(length, n, 1234 + eps, profile)
    want, wn = oracle.block_dbscan(xy, eps, minPts)
    lab, ncl, bbox, size = gpu_block(gpu_ctx, xy, eps, minPts)
    assert ncl == wn
    assert np.array_equal(lab, want)
    wb, ws = oracle.cluster_bbox(xy, want, wn)
    assert np.array_equal(bbox, wb) and np.array_equal(size, ws)


def test_dense_5m_labels(gpu_ctx):
    """5 M Hi-C-like PETs at the density of BASELINE configs[2] (324 k PETs/Mb; eps 5000 / 10000, minPts 20 / 50): cells
    of thousands of points, the heavy-pair tiers of the adjacency kernel, a diagonal component holding most cells.
    Labels (numbering included) and boxes identical to the oracle."""
    xy = synth.cis_chrom(15_432_098, 5_000_000, 33, "hic")
    for eps, minPts in ((5000, 20), (10000, 50)):
        want, wn = oracle.block_dbscan(xy, eps, minPts)
        lab, ncl, bbox, size = gpu_block(gpu_ctx, xy, eps, minPts)
        assert ncl == wn and np.array_equal(lab, want), (eps, minPts)
        wb, ws = oracle.cluster_bbox(xy, want, wn)
        assert np.array_equal(bbox, wb) and np.array_equal(size, ws), (eps, minPts)


def test_minpts_sweep_reuses_grid(gpu_ctx):
    """eps-major sweep: the cached grid of one eps must give the same labels as a fresh build."""
    xy = synth.cis_chrom(10_000_000, 120_000, 77, "hitrac")
    pts = _lib.Points(gpu_ctx, xy)
    try:
        for eps in (300, 900):
            for minPts in (12, 6, 3):
                lab, ncl = pts.block_dbscan(eps, minPts)
                want, wn = oracle.block_dbscan(xy, eps, minPts)
                assert ncl == wn and np.array_equal(lab, want), (eps, minPts)
    finally:
        pts.close()


def test_cut_mcut_filter_on_device(gpu_ctx):
    xy = synth.cis_chrom(5_000_000, 50_000, 5, "hitrac")
    for cut, mcut in ((0, -1), (3000, -1), (0, 500_000), (2000, 800_000)):
        pts = _lib.Points(gpu_ctx, xy, cut=cut, mcut=mcut)
        try:
            want = oracle.parse_ixy_filter(xy, cut, mcut)
            assert pts.n == len(want)
            assert np.array_equal(pts.download(), want)
            if pts.n:
                ext = pts.extent()
                assert ext.tolist() == [want[:, 0].min(), want[:, 0].max(), want[:, 1].min(), want[:, 1].max()]
        finally:
            pts.close()


def test_edge_inputs(gpu_ctx):
    for xy in (np.zeros((0, 2), np.int64), np.array([[5, 9]], np.int64), np.array([[5, 9]] * 40, np.int64),
               np.array([[0, 0], [2**31 - 2, 2**31 - 2]], np.int64)):
        for eps, minPts in ((10, 1), (10, 2), (1, 50), (2**31, 2)):
            want, wn = oracle.block_dbscan(xy, eps, minPts)
            lab, ncl, _, _ = gpu_block(gpu_ctx, xy, eps, minPts)
            assert ncl == wn and np.array_equal(lab, want), (xy.shape, eps, minPts)
    with pytest.raises(_lib.Cl2Error):
        _lib.Points(gpu_ctx, np.array([[-1, 5]], np.int64))
    with pytest.raises(_lib.Cl2Error):
        _lib.Points(gpu_ctx, np.array([[1, 2**31]], np.int64))
    pts = _lib.Points(gpu_ctx, np.array([[1, 2]], np.int64))
    with pytest.raises(_lib.Cl2Error):
        pts.block_dbscan(0, 5)
    pts.close()


def test_dropin_class_protocol(gpu_ctx):
    from cloops2_b200.blockDBSCAN import blockDBSCAN
    xy = synth.cis_chrom(2_000_000, 20_000, 9, "hitrac")
    mat = np.zeros((len(xy), 3))
    mat[:, 0] = np.arange(len(xy))
    mat[:, 1:] = xy
    db = blockDBSCAN(mat, 500, 5)          # float matrix, as callTransLoops.py:43-47 passes it
    want, _ = oracle.block_dbscan(xy, 500, 5)
    assert db.labels == {i: int(v) for i, v in enumerate(want.tolist()) if v >= 0}


def test_full_size_chr1_of_config2(gpu_ctx):
    """BASELINE configs[1] at full size for its largest file: chr1 of the 100 M-PET genome (8.06 M PETs), eps 200 and
    1000, minPts 5 -- labels and boxes equal to the oracle, plus size-independent invariants."""
    n = int(round(100e6 * synth.HG38[0][1] / synth.HG38_TOTAL))
    xy = synth.cis_chrom(synth.HG38[0][1], n, 20261019 + 2000, "hitrac")
    pts = _lib.Points(gpu_ctx, xy)
    try:
        for eps in (200, 1000):
            lab, ncl = pts.block_dbscan(eps, 5)
            bbox, size = pts.cluster_bbox()
            want, wn = oracle.block_dbscan(xy, eps, 5)
            assert ncl == wn and np.array_equal(lab, want)
            # invariants that do not need the oracle
            assert lab.min() >= -1 and lab.max() == ncl - 1
            cnt = np.bincount(lab[lab >= 0], minlength=ncl)
            assert np.array_equal(cnt, size) and (cnt > 0).all()
            order = np.argsort(lab, kind="stable")
            order = order[lab[order] >= 0]
            starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
            assert np.array_equal(np.minimum.reduceat(xy[order, 0], starts), bbox[:, 0])
            assert np.array_equal(np.maximum.reduceat(xy[order, 1], starts), bbox[:, 3])
            # idempotence: a second run on the cached grid gives the same answer
            lab2, ncl2 = pts.block_dbscan(eps, 5)
            assert ncl2 == ncl and np.array_equal(lab2, lab)
    finally:
        pts.close()


@pytest.mark.parametrize("runs", [
    [(1000, 5), (500, 5), (200, 5)],             # chain of nested subsets (1000 -> 500 -> 200)
    [(1000, 4), (1000, 7), (500, 7), (500, 4), (200, 6)],
    [(1000, 5), (500, 5), (400, 5), (200, 5)],   # 500 -> 400 is not a safe pair: 400 falls back to the 1000 subset
    [(450, 5), (333, 5), (100, 5)],              # 450 -> 333 unsafe (every point), 450 -> 100 safe
    [(900, 6), (300, 3), (300, 8), (290, 6)],    # smaller minPts than the link: every point again
    [(200, 5), (1000, 5), (500, 5)],             # ascending eps after the first run: the chain is left alone
])
def test_sweep_hint_changes_nothing(gpu_ctx, runs):
    """cl2_points_sweep_hint: later runs work on the rows earlier runs left, labels / boxes stay those of the oracle
    (which knows nothing of the shortcut)."""
    xy = synth.cis_chrom(6_000_000, 90_000, 31, "hitrac")
    pts = _lib.Points(gpu_ctx, xy)
    try:
        pts.sweep_hint(True)
        for eps, minPts in runs:
            lab, ncl = pts.block_dbscan(eps, minPts)
            bbox, size = pts.cluster_bbox()
            want, wn = oracle.block_dbscan(xy, eps, minPts)
            wb, ws = oracle.cluster_bbox(xy, want, wn)
            assert ncl == wn and np.array_equal(lab, want), (eps, minPts)
            assert np.array_equal(bbox, wb) and np.array_equal(size, ws), (eps, minPts)
        keep = pts.block_noise_keep(runs[-1][0], runs[-1][1])          # always on every point
        assert np.array_equal(keep, oracle.block_noise_keep(xy, runs[-1][0], runs[-1][1]))
        pts.sweep_hint(False)
        lab, ncl = pts.block_dbscan(runs[-1][0], runs[-1][1])
        want, wn = oracle.block_dbscan(xy, runs[-1][0], runs[-1][1])
        assert ncl == wn and np.array_equal(lab, want)
    finally:
        pts.close()


def test_sweep_hint_random_small_cases(gpu_ctx):
    rng = np.random.default_rng(77)
    for it in range(60):
        xy, eps, minPts = random_case(rng, it)
        big = eps * int(rng.choice([2, 3, 5]))
        pts = _lib.Points(gpu_ctx, xy)
        try:
            if pts.n == 0:
                continue
            pts.sweep_hint(True)
            for e, m in ((big, minPts), (eps, minPts + int(rng.integers(0, 3)))):
                lab, ncl = pts.block_dbscan(e, m)
                want, wn = oracle.block_dbscan(xy, e, m)
                assert ncl == wn and np.array_equal(lab, want), (it, e, m)
        finally:
            pts.close()


def test_resident_points_through_another_context(gpu_ctx):
    """Points uploaded through one context, clustered and indexed through another (Points.on): same answers, and the
    owner can still use and free them afterwards."""
    xy = synth.cis_chrom(3_000_000, 40_000, 5, "hitrac")
    other = _lib.Context(gpu_ctx.device)
    pts = _lib.Points(gpu_ctx, xy)
    try:
        want, wn = oracle.block_dbscan(xy, 500, 5)
        view = pts.on(other)
        lab, ncl = view.block_dbscan(500, 5)
        assert ncl == wn and np.array_equal(lab, want)
        bbox, size = view.cluster_bbox()
        idx = _lib.Index(view)
        core, _, _, _, _ = idx.loop_counts(bbox[:50], win=5, mode=0, windows=False, anchors=False)
        idx.close()
        view.drop_cache()
        view.close()                                   # a view owns nothing
        lab2, ncl2 = pts.block_dbscan(500, 5)
        assert ncl2 == wn and np.array_equal(lab2, want)
        idx = _lib.Index(pts)
        core2, _, _, _, _ = idx.loop_counts(bbox[:50], win=5, mode=0, windows=False, anchors=False)
        idx.close()
        assert np.array_equal(core, core2)
    finally:
        pts.close()
        other.close()
