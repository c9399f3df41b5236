#!/usr/bin/env python
"""Timing probe for the dense / skewed regime (BASELINE configs[2] style): Hi-C-like PETs packed into a short
chromosome, eps 5000/10000, minPts 20/50, blockDBSCAN and cDBSCAN2, plus one trans pair (configs[3] style).

    python profiles/prof_dense.py [pets] [length]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import _lib, synth  # noqa: E402
from cloops2_b200.callCisLoops import cis_chromosome  # noqa: E402
from cloops2_b200.callTransLoops import trans_pair  # noqa: E402

pets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
length = int(float(sys.argv[2])) if len(sys.argv) > 2 else 30_000_000
ctx = _lib.Context(0)
xy = synth.cis_chrom(length, pets, 5, "hic")
pts = _lib.Points(ctx, xy)
for name in ("block", "c2"):
    for eps in (5000, 10000):
        for mp in (50, 20):
            fn = pts.block_dbscan if name == "block" else pts.cdbscan2
            t = time.time()
            _, ncl = fn(eps, mp, want_labels=False)
            ctx.sync()
            print("%-5s eps %5d minPts %2d: %8.1f ms  %d clusters  (%d PETs, %.1f k PETs/Mb)"
                  % (name, eps, mp, 1e3 * (time.time() - t), ncl, pts.n, pts.n / (length / 1e6) / 1e3), flush=True)
for hic in (False, True):
    t = time.time()
    final, st = cis_chromosome(ctx, ("chr1", "chr1"), None, pts.n, [5000, 10000], [50, 20], hic=hic, pts=pts)
    print("callLoops hic=%s: %.1f ms %s" % (hic, 1e3 * (time.time() - t), {k: v for k, v in st.items() if k != "host_s"}), flush=True)
txy = synth.trans_pair(248_956_422, 242_193_529, 2_600_000, 9)
t = time.time()
final, st = trans_pair(ctx, ("chr1", "chr2"), txy, [5000], [20, 10])
print("trans chr1-chr2 %d PETs: %.1f ms %s" % (len(txy), 1e3 * (time.time() - t), st), flush=True)
