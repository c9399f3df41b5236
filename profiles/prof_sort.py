#!/usr/bin/env python
"""The radix sort alone: y-order + x-order of one chromosome of the bench workload (cl2_index_build = two 4-pass sorts of
(u32 key, u32 row) pairs) and the 3-pass grid sorts of a clustering sweep, timed by the library's CUDA-event timers.

    python profiles/prof_sort.py [pets] [reps]          (CL2_SORT_CFG=0|1|2 selects the tile shape)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import _lib, synth  # noqa: E402

pets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_060_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
xy = synth.cis_chrom(synth.HG38[0][1], pets, 20261019 + 2000, "hitrac")
ctx = _lib.Context(0)
pts = _lib.Points(ctx, xy)
for r in range(reps):
    ctx.timing_reset()
    ctx.timing_enable(True)
    idx = _lib.Index(pts)
    idx.close()
    for eps in (1000, 500, 200):
        pts.block_dbscan(eps, 5, want_labels=False)
    pts.drop_cache()
    tm = ctx.timing_read()
    ctx.timing_enable(False)
    s, h = tm["sort_scatter"], tm["sort_hist"]
    print("rep %d: sort_scatter %.3f ms in %d passes = %.1f us per pass, %.0f GB/s algorithmic; sort_hist %.3f ms"
          % (r, s["ms"], s["launches"], 1e3 * s["ms"] / s["launches"], s["bytes"] / 1e9 / (s["ms"] / 1e3), h["ms"]), flush=True)
