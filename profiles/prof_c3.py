#!/usr/bin/env python
"""One chromosome at the density of BASELINE configs[2] (1 B Hi-C-like PETs over hg38 = 324 k PETs/Mb) through the
whole per-chromosome path, blockDBSCAN and -hic (cDBSCAN2), with the library's per-kernel-family CUDA-event timers.

    python profiles/prof_c3.py [chrom index, default 20 = chr21] [block|hic|both]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import _lib, synth  # noqa: E402
from cloops2_b200.callCisLoops import cis_chromosome  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2] if len(sys.argv) > 2 else "both"
name, length = synth.HG38[ci]
n = int(round(1e9 * length / synth.HG38_TOTAL))
t = time.time()
xy = synth.cis_chrom(length, n, 20261019 + 3000 + ci, "hic")
print("%s: %d PETs generated in %.1f s" % (name, len(xy), time.time() - t), flush=True)
ctx = _lib.Context(0)
pts = _lib.Points(ctx, xy)
for hic in ([False, True] if which == "both" else [which == "hic"]):
    for rep in range(2):
        ctx.timing_reset()
        ctx.timing_enable(True)
        t = time.time()
        final, st = cis_chromosome(ctx, (name, name), None, 10**9, [5000, 10000], [50, 20], hic=hic, pts=pts)
        ctx.sync()
        dt = time.time() - t
        tm = ctx.timing_read()
        ctx.timing_enable(False)
        print("hic=%s rep %d: %.1f ms  %s  host_s %s" % (hic, rep, 1e3 * dt, {k: v for k, v in st.items() if k != "host_s"},
                                                       {k: round(v, 3) for k, v in st.get("host_s", {}).items()}), flush=True)
        if rep == 1:
            for k, v in sorted(tm.items(), key=lambda kv: -kv[1]["ms"]):
                print("    %-13s %9.3f ms %5d launches %8.0f GB/s" % (k, v["ms"], v["launches"], v["bytes"] / 1e9 / (v["ms"] / 1e3) if v["ms"] else 0))
