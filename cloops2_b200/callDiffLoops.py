"""
`callDiffLoops.quantDloops` on B200 -- drop-in for cLoops2/callDiffLoops.py:110-161, the per-chromosome work unit of
`cLoops2 callDiffLoops`: the loops of the merged set are counted in the target and in the control sample (queryLoop)
together with their permuted nearby windows (getLoopNearbyPETs, win=1 by default).  It is the counting kernel of the
callLoops path, called once per sample (first "next" row).  The MANorm fit, the Poisson test and
the plots that follow in the reference are host-side statistics on these numbers and are not part of this package.
"""
import numpy as np

from . import _lib
from .ds import DiffLoop
from .io import loadIxy


def _sample_counts(ctx, fixy, cands, cut, mcut, win):
    pts = _lib.Points(ctx, loadIxy(fixy), cut=cut, mcut=mcut)
    try:
        index = _lib.Index(pts)
        try:
            core, _, _, nab, _ = index.loop_counts(cands, win=win, mode=_lib.MODE_CIS, anchors=False)
        finally:
            index.close()
    finally:
        pts.close()
    return core, nab


def quantDloops(key, loops, fixya, fixyb, cut=0, mcut=-1, win=1, pseudo=1.0, ctx=None):
    """Drop-in for callDiffLoops.py:110-161.  Returns (key, ts, cs, dloops) -- ts / cs are the nearby-window counts of
    every loop in the target / control sample, in the reference's order -- or None without loops."""
    if len(loops) == 0:
        return None
    ctx = ctx or _lib.default_context()
    cands = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops], dtype=np.int64).reshape(-1, 4)
    acore, anab = _sample_counts(ctx, fixya, cands, cut, mcut, win)
    bcore, bnab = _sample_counts(ctx, fixyb, cands, cut, mcut, win)
    print("Quantify %s loops for %s." % (len(loops), key))
    ts, cs, dloops = [], [], []
    for i, loop in enumerate(loops):
        arabs, brabs = anab[i].astype(np.float64), bnab[i].astype(np.float64)
        d = DiffLoop()
        d.id = loop.id
        d.chromX, d.chromY = loop.chromX, loop.chromY
        d.x_start, d.x_end, d.x_center = loop.x_start, loop.x_end, loop.x_center
        d.y_start, d.y_end, d.y_center = loop.y_start, loop.y_end, loop.y_center
        d.distance = loop.y_center - loop.x_center
        d.raw_trt_ra, d.raw_trt_rb = int(acore[i, 0]), int(acore[i, 1])
        d.raw_con_ra, d.raw_con_rb = int(bcore[i, 0]), int(bcore[i, 1])
        d.raw_trt_rab, d.raw_con_rab = int(acore[i, 2]), int(bcore[i, 2])
        d.size = d.x_end - d.x_start + d.y_end - d.y_start
        d.raw_trt_mrab = np.mean(arabs)
        d.raw_con_mrab = np.mean(brabs)
        ts.extend(arabs)
        cs.extend(brabs)
        d.trt_es = d.raw_trt_rab / max(d.raw_trt_mrab, pseudo)
        d.con_es = d.raw_con_rab / max(d.raw_con_mrab, pseudo)
        dloops.append(d)
    return key, ts, cs, dloops
