#!/usr/bin/env python
"""Small single-process target for ncu: one chromosome of the bench workload through the whole per-chromosome path
(no fork, no torch, no nvidia-smi child; CL2_NO_POOL=1 selects plain cudaMalloc).

    CL2_NO_POOL=1 python profiles/prof_target.py [pets] [runs]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import _lib, synth  # noqa: E402
from cloops2_b200.callCisLoops import cis_chromosome  # noqa: E402

pets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_060_000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
xy = synth.cis_chrom(synth.HG38[0][1], pets, 20261019 + 2000, "hitrac")
ctx = _lib.Context(0)
pts = _lib.Points(ctx, xy)
for r in range(runs):
    t = time.time()
    final, st = cis_chromosome(ctx, ("chr1", "chr1"), None, 100_000_000, [200, 500, 1000], [5], pts=pts)
    ctx.sync()
    print("run", r, "%.1f ms" % (1e3 * (time.time() - t)), {k: v for k, v in st.items() if k != "host_s"},
          {k: round(1e3 * v, 1) for k, v in st["host_s"].items()}, flush=True)
print("launches", ctx.launch_count())
