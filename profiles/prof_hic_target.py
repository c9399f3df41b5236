#!/usr/bin/env python
"""Small single-process ncu target for the dense regime (BASELINE configs[2] density, 324 k PETs/Mb): one blockDBSCAN
and one cDBSCAN2 run at eps 5000 / minPts 20 on a Hi-C-like chromosome (no fork, no torch; CL2_NO_POOL=1 selects plain
cudaMalloc, which ncu needs).

    CL2_NO_POOL=1 python profiles/prof_hic_target.py [pets]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import _lib, synth  # noqa: E402

pets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000
length = int(pets / 0.324)
xy = synth.cis_chrom(length, pets, 33, "hic")
ctx = _lib.Context(0)
pts = _lib.Points(ctx, xy)
for name, fn in (("blockDBSCAN", pts.block_dbscan), ("cDBSCAN2", pts.cdbscan2)):
    t = time.time()
    _, ncl = fn(5000, 20, want_labels=False)
    ctx.sync()
    print("%s eps 5000 minPts 20: %.1f ms, %d clusters, %d PETs" % (name, 1e3 * (time.time() - t), ncl, pts.n), flush=True)
print("launches", ctx.launch_count())
