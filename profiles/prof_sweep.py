#!/usr/bin/env python
"""BASELINE configs[4]: eps x minPts sweep on ChIA-PET-like PETs (300 M over hg38 = 97 k PETs/Mb; d_min 1000, sigma 500),
clustering kernels only -- blockDBSCAN and cDBSCAN2 on one chromosome of that genome, every (eps, minPts) a fresh run
(grid cache dropped), per-run time and per-kernel-family algorithmic GB/s from the library's CUDA-event timers.

    python profiles/prof_sweep.py [chrom index, default 0 = chr1: 24.2 M PETs] > profiles/r2_sweep.md
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import _lib, synth  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 0
name, length = synth.HG38[ci]
n = int(round(300e6 * length / synth.HG38_TOTAL))
xy = synth.cis_chrom(length, n, 20261019 + 5000 + ci, "chiapet")
ctx = _lib.Context(0)
pts = _lib.Points(ctx, xy)
EPS = [100, 200, 500, 1000, 2000, 5000, 10000, 20000]
MINPTS = [3, 5, 10, 20, 50, 100]
FAMS = ["sort_scatter", "sort_hist", "gather", "cells", "nbr", "noise", "cellstat", "adj", "core", "union", "comp", "border", "bbox", "label"]
peak = 6553.0
print("# eps x minPts sweep, %s of a 300 M-PET ChIA-PET-like genome (%d PETs), clustering only\n" % (name, pts.n))
print("ms per run (eps grid built fresh for every run) and PETs/s; algorithmic bytes = 20 B per PET per run (SURVEY.md 8d)\n")
for algo in ("blockDBSCAN", "cDBSCAN2"):
    print("## %s\n" % algo)
    print("| eps \\ minPts | " + " | ".join(str(m) for m in MINPTS) + " |")
    print("|---|" + "---:|" * len(MINPTS))
    worst = {}
    for eps in EPS:
        cells = []
        for mp in MINPTS:
            pts.drop_cache()
            ctx.sync()
            ctx.timing_reset()
            ctx.timing_enable(True)
            t = time.perf_counter()
            if algo == "blockDBSCAN":
                _, ncl = pts.block_dbscan(eps, mp, want_labels=False)
            else:
                _, ncl = pts.cdbscan2(eps, mp, want_labels=False)
            ctx.sync()
            dt = time.perf_counter() - t
            tm = ctx.timing_read()
            ctx.timing_enable(False)
            dev = sum(v["ms"] for v in tm.values())
            cells.append("%.1f ms (dev %.1f), %d cl, %.0f GB/s" % (1e3 * dt, dev, ncl, 20.0 * pts.n / 1e9 / dt))
            for k, v in tm.items():
                if v["ms"] > 0 and v["bytes"] > 0:
                    g = v["bytes"] / 1e9 / (v["ms"] / 1e3)
                    w = worst.setdefault(k, [1e30, 0.0, 0.0])
                    w[0] = min(w[0], g)
                    w[1] = max(w[1], g)
                    w[2] = max(w[2], v["ms"])
        print("| %d | " % eps + " | ".join(cells) + " |")
    print("\nper-kernel-family algorithmic GB/s over the sweep (min .. max; fraction of the measured %.0f GB/s; largest time of one run):\n" % peak)
    print("| family | min GB/s | max GB/s | min frac | max frac | max ms |")
    print("|---|---:|---:|---:|---:|---:|")
    for k in FAMS:
        if k in worst:
            w = worst[k]
            print("| %s | %.0f | %.0f | %.3f | %.3f | %.2f |" % (k, w[0], w[1], w[0] / peak, w[1] / peak, w[2]))
    print()
