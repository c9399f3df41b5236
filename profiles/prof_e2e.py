#!/usr/bin/env python
"""e2e probe: the per-chromosome call from a pinned host array, repeated (upload inside the loop)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cloops2_b200 import _lib, synth
from cloops2_b200.callCisLoops import cis_chromosome
pets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_060_000
xy = synth.cis_chrom(synth.HG38[0][1], pets, 20261019 + 2000, "hitrac")
ctx = _lib.Context(0)
pin = ctx.pinned_empty(xy.shape, np.int64); np.copyto(pin, xy)
for r in range(6):
    t = time.time()
    final, st = cis_chromosome(ctx, ("chr1", "chr1"), pin, 100_000_000, [200, 500, 1000], [5])
    ctx.sync()
    print("run", r, "%.1f ms" % (1e3 * (time.time() - t)), {k: round(1e3 * v, 1) for k, v in st["host_s"].items()}, flush=True)
