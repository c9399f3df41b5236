#!/usr/bin/env python
"""selSigLoops (callCisLoops.py:379-422) at Hi-C-scale loop counts, host only: the native sweep-based selection of
libcl2.so (cl2_sel_sig_loops, applied twice as the reference does, :586-595) against the reference's own function on
Loop objects where that is still bearable.  Runs without a GPU.

    python profiles/prof_selsig.py > profiles/r2_selsig_host.md
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import hostops  # noqa: E402
from oracle import ref_shim  # noqa: E402


def loops(L, rng, span):
    xs = rng.integers(0, span, size=L)
    ys = xs + rng.integers(20_000, 1_000_000, size=L)
    return (np.stack([xs, xs + rng.integers(500, 6000, size=L), ys, ys + rng.integers(500, 6000, size=L)], 1).astype(np.int64),
            rng.uniform(2, 50, size=L))


rng = np.random.default_rng(3)
print("# selSigLoops on the host: native sweep (cl2_sel_sig_loops, twice) vs the reference's list scans\n")
print("loops on one chromosome of 248 Mb, anchors 0.5-6 kb, distance 20 kb - 1 Mb, all marked significant\n")
print("| loops | kept | native, twice (s) | reference selSigLoops, twice (s) | same selection |")
print("|---:|---:|---:|---:|---|")
ref = ref_shim.load() if ref_shim.available() else None
for L in (2_000, 20_000, 200_000, 1_000_000):
    c, es = loops(L, rng, 248_000_000)
    cols = {"coords": c, "ES": es}
    t = time.time()
    cur = dict(cols)
    for _ in range(2):
        pick = hostops.sel_sig(cur["coords"], cur["ES"])
        cur = {k: v[pick] for k, v in cur.items()}
    dt = time.time() - t
    rt, same = "-", "-"
    if ref is not None and L <= 20_000:
        from cLoops2.ds import Loop
        from cLoops2.callCisLoops import selSigLoops
        ls = []
        for i in range(L):
            lp = Loop()
            lp.chromX = lp.chromY = "chr1"
            lp.x_start, lp.x_end, lp.y_start, lp.y_end = (int(v) for v in c[i])
            lp.ES = float(es[i])
            lp.significant = 1
            ls.append(lp)
        t = time.time()
        _, r1 = selSigLoops("chr1-chr1", ls)
        _, r2 = selSigLoops("chr1-chr1", r1)
        rt = "%.2f" % (time.time() - t)
        got = [tuple(int(v) for v in row) for row in cur["coords"]]
        want = [(lp.x_start, lp.x_end, lp.y_start, lp.y_end) for lp in r2]
        same = "yes" if got == want else "NO"
    print("| %d | %d | %.3f | %s | %s |" % (L, cur["coords"].shape[0], dt, rt, same))
