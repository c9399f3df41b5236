"""XY index + rectangle/window counting kernel vs the oracle's set-based counts (identical integers)."""
import numpy as np
import pytest

import oracle
from cloops2_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _cands(xy, rng, L):
    """Candidate-like boxes around real PETs plus adversarial ones (overlapping anchors, at the origin, huge)."""
    i = rng.integers(0, len(xy), size=L)
    wa = rng.integers(1, 3000, size=L)
    wb = rng.integers(1, 3000, size=L)
    xs = np.maximum(0, xy[i, 0] - wa // 2)
    ys = np.maximum(0, xy[i, 1] - wb // 2)
    c = np.stack([xs, xs + wa, ys, ys + wb], 1).astype(np.int64)
    c[0] = [0, 50, 10, 400]
    c[1] = [0, 1, 0, 1]
    c[2] = [100, 100000, 50000, 300000]
    return c


@pytest.mark.parametrize("mode", [0, 1])
def test_counts_match_oracle(gpu_ctx, mode):
    rng = np.random.default_rng(11 + mode)
    if mode == 0:
        xy = synth.cis_chrom(3_000_000, 80_000, 3, "hitrac")
    else:
        xy = synth.trans_pair(2_000_000, 2_500_000, 60_000, 4, per=1500)
    cands = _cands(xy, rng, 300)
    pts = _lib.Points(gpu_ctx, xy)
    idx = _lib.Index(pts)
    try:
        got = idx.loop_counts(cands, win=5, mode=mode)
    finally:
        idx.close()
        pts.close()
    want = oracle.XYIndex(xy).loop_counts(cands, win=5, mode=mode)
    for name, g, w in zip(("core", "na", "nb", "nab", "anchors"), got, want):
        assert np.array_equal(g, w), name


@pytest.mark.parametrize("case", ["cis", "trans", "hot_column", "dense"])
def test_counts_with_x_order_left_by_the_grid(gpu_ctx, case):
    """cl2_points_index_hint: the blockDBSCAN run before the index build leaves the x-order (columns of the grid sorted
    locally); the counts must be those of the oracle whether the local sort applies (cis, trans), gives up on a column
    longer than its walk limit (hot_column -> radix sort) or is not attempted at all (dense: > 64 rows per column)."""
    rng = np.random.default_rng(23)
    mode, eps = 0, 1000
    if case == "cis":
        xy = synth.cis_chrom(3_000_000, 80_000, 3, "hitrac")
    elif case == "trans":
        xy, mode, eps = synth.trans_pair(2_000_000, 2_500_000, 60_000, 4, per=1500), 1, 5000
    elif case == "hot_column":
        x = rng.integers(0, 10_000_000, size=200_000)
        hx = rng.integers(5_000_000, 5_000_500, size=6000)
        x = np.concatenate([x, hx])
        xy = np.stack([x, x + rng.integers(100, 2_000_000, size=len(x))], 1).astype(np.int64)
        xy = xy[rng.permutation(len(xy))]
    else:
        x = rng.integers(0, 200_000, size=60_000)
        xy = np.stack([x, x + rng.integers(100, 100_000, size=len(x))], 1).astype(np.int64)
    cands = _cands(xy, rng, 300)
    pts = _lib.Points(gpu_ctx, xy)
    try:
        pts.index_hint(True)
        pts.block_dbscan(eps, 5, want_labels=False)
        idx = _lib.Index(pts)
        try:
            got = idx.loop_counts(cands, win=5, mode=mode)
        finally:
            idx.close()
        # a second index on the same points (nothing left behind now) must agree with the first
        pts.index_hint(False)
        idx = _lib.Index(pts)
        try:
            again = idx.loop_counts(cands, win=5, mode=mode)
        finally:
            idx.close()
    finally:
        pts.close()
    want = oracle.XYIndex(xy).loop_counts(cands, win=5, mode=mode)
    for name, g, a, w in zip(("core", "na", "nb", "nab", "anchors"), got, again, want):
        assert np.array_equal(g, w), name
        assert np.array_equal(a, w), name


def test_counts_other_window_sizes_and_partial_outputs(gpu_ctx):
    rng = np.random.default_rng(5)
    xy = synth.cis_chrom(1_000_000, 30_000, 6, "hitrac")
    cands = _cands(xy, rng, 64)
    pts = _lib.Points(gpu_ctx, xy)
    idx = _lib.Index(pts)
    try:
        for win in (1, 3, 8):
            got = idx.loop_counts(cands, win=win)
            want = oracle.XYIndex(xy).loop_counts(cands, win=win)
            for g, w in zip(got, want):
                assert np.array_equal(g, w), win
        c, na, nb, nab, anc = idx.loop_counts(cands, windows=False, anchors=False)
        assert na is None and anc is None and np.array_equal(c, oracle.XYIndex(xy).loop_counts(cands)[0])
        assert idx.loop_counts(np.zeros((0, 4), np.int64))[0].shape == (0, 4)
        with pytest.raises(_lib.Cl2Error):
            idx.loop_counts(cands, win=9)
    finally:
        idx.close()
        pts.close()


def test_full_size_counts_properties(gpu_ctx):
    """8 M PETs (chr1 of configs[1]): counting identities that hold at any size, plus oracle equality on a sample."""
    n = int(round(100e6 * synth.HG38[0][1] / synth.HG38_TOTAL))
    xy = synth.cis_chrom(synth.HG38[0][1], n, 20261019 + 2000, "hitrac")
    rng = np.random.default_rng(1)
    cands = _cands(xy, rng, 2000)
    pts = _lib.Points(gpu_ctx, xy)
    idx = _lib.Index(pts)
    try:
        core, na, nb, nab, anc = idx.loop_counts(cands)
    finally:
        idx.close()
        pts.close()
    ra, rb, rab = core[:, 0], core[:, 1], core[:, 2]
    assert (rab <= ra).all() and (rab <= rb).all() and (core >= 0).all()
    assert (nab.reshape(-1, 10, 10) <= na[:, :, None]).all() and (nab.reshape(-1, 10, 10) <= nb[:, None, :]).all()
    assert np.array_equal(anc[:, 0], ra) and np.array_equal(anc[:, 2], rb)      # estAnchorSig counts the anchor itself
    assert (anc[:, 1] >= anc[:, 0]).all() and (anc[:, 3] >= anc[:, 2]).all()     # the +-5x window contains the anchor
    # |P(l,r)| by brute force on numpy for a few loops
    for i in range(0, 2000, 400):
        l, r = cands[i, 0], cands[i, 1]
        brute = int((((xy[:, 0] >= l) & (xy[:, 0] <= r)) | ((xy[:, 1] >= l) & (xy[:, 1] <= r))).sum())
        assert brute == ra[i]
    want = oracle.XYIndex(xy).loop_counts(cands[:200])
    for g, w in zip((core[:200], na[:200], nb[:200], nab[:200], anc[:200]), want):
        assert np.array_equal(g, w)
