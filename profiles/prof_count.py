#!/usr/bin/env python
"""Counting stage alone on one chromosome of the bench workload: clusters once, then times cl2_est_loops (k_est_bg +
k_est_anchor + k_est_stats) over the candidates with the library's own CUDA-event timers.

    python profiles/prof_count.py [pets] [reps]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import _lib, synth, hostops  # noqa: E402
from cloops2_b200.callCisLoops import sweep_passes  # noqa: E402

pets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_060_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
xy = synth.cis_chrom(synth.HG38[0][1], pets, 20261019 + 2000, "hitrac")
ctx = _lib.Context(0)
pts = _lib.Points(ctx, xy)
cands = np.concatenate([c for c, _ in sweep_passes(pts, [(e, 5) for e in (200, 500, 1000)])])
ext = pts.extent()
rpb = pts.n / float(int(ext[3]) - int(ext[0]))
index = _lib.Index(pts)
for r in range(reps):
    ctx.timing_reset()
    ctx.timing_enable(True)
    t = time.perf_counter()
    sel, core, hyp, st = index.est_loops(cands, rpb, 5)
    dt = time.perf_counter() - t
    tm = ctx.timing_read()
    ctx.timing_enable(False)
    c = tm.get("count", {"ms": 0, "bytes": 0})
    print("rep %d: %d candidates -> %d survivors, wall %.2f ms, count family %.3f ms, %.0f GB/s algorithmic, checksum %d"
          % (r, len(cands), len(sel), 1e3 * dt, c["ms"], c["bytes"] / 1e9 / (c["ms"] / 1e3) if c["ms"] else 0,
             int(core.sum()) + int(sel.sum())), flush=True)
