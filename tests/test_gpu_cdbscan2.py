"""cDBSCAN2 (-hic) on the GPU through the C ABI vs the oracle / the reference's golden labels.  Bit-exact."""
import numpy as np
import pytest

import oracle
from cloops2_b200 import _lib, synth
from cloops2_b200.callCisLoops import cis_chromosome, cols_to_loops
from conftest import load_golden, random_case, golden_cis_files, txt_rows

pytestmark = pytest.mark.gpu


def gpu_c2(ctx, xy, eps, minPts):
    pts = _lib.Points(ctx, xy)
    try:
        if pts.n == 0:
            return np.zeros(0, np.int32), 0, np.zeros((0, 4), np.int64), np.zeros(0, np.int64)
        lab, ncl = pts.cdbscan2(eps, minPts)
        bbox, size = pts.cluster_bbox()
        return lab, ncl, bbox, size
    finally:
        pts.close()


def test_golden_labels(gpu_ctx):
    g = load_golden("dbscan_small.npz")
    offs = g["offsets"]
    for i, (eps, minPts) in enumerate(g["params"].tolist()):
        xy = g["xy"][offs[i]:offs[i + 1]]
        lab, ncl, _, _ = gpu_c2(gpu_ctx, xy, eps, minPts)
        assert np.array_equal(lab, g["labels_c2"][offs[i]:offs[i + 1]]), (i, eps, minPts)


def test_random_small_cases(gpu_ctx):
    rng = np.random.default_rng(4242)
    for it in range(400):
        xy, eps, minPts = random_case(rng, it)
        want, wn = oracle.cdbscan2(xy, eps, minPts)
        lab, ncl, bbox, size = gpu_c2(gpu_ctx, xy, eps, minPts)
        assert ncl == wn and np.array_equal(lab, want), (it, len(xy), eps, minPts)
        wb, ws = oracle.cluster_bbox(xy, want, wn)
        assert np.array_equal(bbox, wb) and np.array_equal(size, ws), it


@pytest.mark.parametrize("profile,n,length,eps,minPts", [
    ("hic", 200_000, 3_000_000, 5000, 20), ("hic", 200_000, 3_000_000, 10000, 50),
    ("hitrac", 300_000, 46_709_983, 500, 5), ("hic", 150_000, 10_000_000, 2000, 10)])
def test_synthetic_chromosomes(gpu_ctx, profile, n, length, eps, minPts):
    xy = synth.cis_chrom(length, n, 99 + eps, profile)
    want, wn = oracle.cdbscan2(xy, eps, minPts)
    lab, ncl, bbox, size = gpu_c2(gpu_ctx, xy, eps, minPts)
    assert ncl == wn
    assert np.array_equal(lab, want)
    wb, ws = oracle.cluster_bbox(xy, want, wn)
    assert np.array_equal(bbox, wb) and np.array_equal(size, ws)


def test_scratch_arena_with_many_small_chunks(monkeypatch):
    """A context whose scratch arena grows in 1 MB chunks (CL2_ARENA_CHUNK_MB): dozens of chunks, many of them adjacent
    in the address space.  A block must stay inside one chunk (the runtime's memset / memcpy reject a range that spans
    two allocations), so both clusterings and the counting path must run and give the oracle's results."""
    monkeypatch.setenv("CL2_ARENA_CHUNK_MB", "1")
    ctx = _lib.Context(0)
    try:
        xy = synth.cis_chrom(6_000_000, 600_000, 5, "hic")
        for eps, minPts in ((5000, 20), (2000, 10)):
            want, wn = oracle.cdbscan2(xy, eps, minPts)
            lab, ncl, _, _ = gpu_c2(ctx, xy, eps, minPts)
            assert ncl == wn and np.array_equal(lab, want), (eps, minPts)
            pts = _lib.Points(ctx, xy)
            try:
                lab, ncl = pts.block_dbscan(eps, minPts)
            finally:
                pts.close()
            want, wn = oracle.block_dbscan(xy, eps, minPts)
            assert ncl == wn and np.array_equal(lab, want), (eps, minPts)
    finally:
        ctx.close()


def test_dense_5m_labels(gpu_ctx):
    """-hic on 5 M Hi-C-like PETs at the density of BASELINE configs[2] (324 k PETs/Mb), eps 5000, minPts 20: crowded
    cells (core counting with early exit, heavy cell pairs, border release rounds).  Labels identical to the oracle."""
    xy = synth.cis_chrom(15_432_098, 5_000_000, 33, "hic")
    want, wn = oracle.cdbscan2(xy, 5000, 20)
    lab, ncl, bbox, size = gpu_c2(gpu_ctx, xy, 5000, 20)
    assert ncl == wn and np.array_equal(lab, want)
    wb, ws = oracle.cluster_bbox(xy, want, wn)
    assert np.array_equal(bbox, wb) and np.array_equal(size, ws)


def test_hic_golden_rows(gpu_ctx):
    g, files = golden_cis_files()
    loops = []
    for k, xy in files.items():
        final, st = cis_chromosome(gpu_ctx, tuple(k.split("-")), xy, int(g["tot"]), [200, 500, 1000], [5], hic=True, exact=True)
        loops.extend(cols_to_loops(tuple(k.split("-")), final))
    assert oracle.loops_to_rows(loops) == txt_rows(g["hic_txt"])


def test_switching_grid_kinds_on_one_points_object(gpu_ctx):
    xy = synth.cis_chrom(5_000_000, 60_000, 31, "hitrac")
    pts = _lib.Points(gpu_ctx, xy)
    try:
        for _ in range(2):
            lab, _ = pts.block_dbscan(400, 5)
            assert np.array_equal(lab, oracle.block_dbscan(xy, 400, 5)[0])
            lab, _ = pts.cdbscan2(400, 5)
            assert np.array_equal(lab, oracle.cdbscan2(xy, 400, 5)[0])
    finally:
        pts.close()
