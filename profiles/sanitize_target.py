#!/usr/bin/env python
"""Small single-process target for compute-sanitizer (memcheck / racecheck / synccheck, one tool per run): every kernel
family of the path on inputs small enough for a ~50x slowdown -- blockDBSCAN with a sweep chain (subsets), cDBSCAN2,
labels, the XY index, counting + statistics in the warp and the cooperative form (a dense piece), raw counts, the
de-duplication of `pre`, one whole per-file call in cis, -hic and trans mode.  Results are checked against the oracle.

    compute-sanitizer --tool racecheck python profiles/sanitize_target.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CL2_NO_POOL", "1")

import oracle  # noqa: E402
from cloops2_b200 import _lib, synth  # noqa: E402
from cloops2_b200.callCisLoops import cis_chromosome, cols_to_loops  # noqa: E402
from cloops2_b200.callTransLoops import trans_pair  # noqa: E402

ctx = _lib.Context(0)
rows = oracle.loops_to_rows

# sparse Hi-TrAC-like piece: sweep chain 1000 -> 500 -> 200, labels, counting in the warp form
xy = synth.cis_chrom(3_000_000, 90_000, 5, "hitrac")
pts = _lib.Points(ctx, xy)
pts.sweep_hint(1)
for eps in (1000, 500, 200):
    lab, ncl = pts.block_dbscan(eps, 5)
    want, wn = oracle.block_dbscan(xy, eps, 5)
    assert ncl == wn and np.array_equal(lab, want), ("block", eps)
pts.sweep_hint(0)
lab, ncl = pts.cdbscan2(500, 5)
want, wn = oracle.cdbscan2(xy, 500, 5)
assert ncl == wn and np.array_equal(lab, want), "cdbscan2"
u = pts.unique()
assert len(u) == len(np.unique(xy, axis=0))
final, st = cis_chromosome(ctx, ("chr1", "chr1"), None, len(xy), [200, 500, 1000], [5], pts=pts)
want, _ = oracle.call_cis_loops({"chr1-chr1": xy}, len(xy), [200, 500, 1000], [5])
assert [r[:18] for r in rows(cols_to_loops(("chr1", "chr1"), final))] == [r[:18] for r in rows(want)], "cis"
pts.close()

# dense Hi-C-like piece: crowded cells, heavy pairs, the cooperative counting kernels, -hic
xy = synth.cis_chrom(1_200_000, 400_000, 6, "hic")
for hic in (False, True):
    final, st = cis_chromosome(ctx, ("chr2", "chr2"), xy, len(xy), [5000, 10000], [50, 20], hic=hic)
    want, _ = oracle.call_cis_loops({"chr2-chr2": xy}, len(xy), [5000, 10000], [50, 20], hic=hic)
    assert [r[:18] for r in rows(cols_to_loops(("chr2", "chr2"), final))] == [r[:18] for r in rows(want)], ("dense", hic)

# raw counts incl. the wide-window form (win 8)
pts = _lib.Points(ctx, xy)
idx = _lib.Index(pts)
rng = np.random.default_rng(3)
a = rng.integers(0, 1_000_000, size=(40, 1))
loops = np.concatenate([a, a + rng.integers(200, 20000, size=(40, 1)), a + 60000, a + 60000 + rng.integers(200, 20000, size=(40, 1))], 1)
for win in (1, 5, 8):
    got = idx.loop_counts(loops, win=win)
    want = oracle.XYIndex(xy).loop_counts(loops, win=win)
    assert all(np.array_equal(g, w) for g, w in zip(got, want)), ("counts", win)
idx.close()
pts.close()

# one trans pair
txy = synth.trans_pair(2_000_000, 2_400_000, 60_000, 11, sigma=800, per=1500)
final, st = trans_pair(ctx, ("chr1", "chr2"), txy, [2000, 4000], [10, 5])
want = oracle.call_trans_loops({"chr1-chr2": txy}, [2000, 4000], [10, 5])
assert [r[:18] for r in rows(cols_to_loops(("chr1", "chr2"), final, cis=False))] == [r[:18] for r in rows(want)], "trans"
print("sanitize target ok: launches", ctx.launch_count())
