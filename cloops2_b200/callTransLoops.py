"""
Inter-chromosomal loop calling on B200 -- drop-in for cLoops2/callTransLoops.py (runTransDBSCANLoops :33,
estLoopSig :106, markSigLoops :196, callTransLoops :209).  Same device-resident organisation as the cis path; the
differences the reference has are kept: blockDBSCAN only (:29), no cut / anchor-size / order filters (:58-79),
distance = -1 (:78), P2LL window lower-left in both axes (:142-145), mean instead of median background (:171-172),
binomial n = ra*rb (:181), density over the file's own PET count (:185-187), no early-exit filters, and the
significance rule binomial <= 1e-10, FDR <= 0.05, ES >= 2 (:200).
"""
import json

import numpy as np

from . import _lib, hostops, shard
from .ds import Loop
from .io import ixyKey, loadIxy, loops2txt, cols2txt, loops2juiceTxt, loops2washuTxt  # noqa: F401
from .callCisLoops import cluster_pass, sweep_passes, cols_to_loops, selSigLoops, _loops_to_cands, exact_default  # noqa: F401


def trans_pair(ctx, key, xy, eps, minPts, pts=None, exact=False):
    """Whole per-pair path (sweep -> combine -> count -> test -> mark -> select) around one native call
    (cl2_call_loops); returns (final cols, stats).  exact as in callCisLoops.cis_chromosome."""
    own = pts is None
    if own:
        pts = _lib.Points(ctx, xy, cut=0, mcut=-1)
    try:
        stats = {"pets": pts.n, "candidates": 0, "tested": 0, "loops": 0, "first_pass": -1}
        if pts.n == 0:
            return hostops._empty_cols(), stats
        passes = [(ep, mp) for ep in eps for mp in minPts]
        ctx.set_pvalue_margin(0.0)                            # the trans estLoopSig has no cut on a tail probability
        # combineLoops only (callTransLoops.py:230-234): coordinates already seen in an EARLIER run are dropped,
        # duplicates inside one run survive; no cut / anchor-size / order filters (:58-79); countDiffCut 10 (:107)
        res = pts.call_loops(passes, mode=_lib.MODE_TRANS, hic=False, cut=0, min_pts_test=max(minPts), countDiffCut=10)
        stats["candidates"] = int(res["totals"][1])
        nz = np.nonzero(res["pass_cand"])[0]
        stats["first_pass"] = int(nz[0]) if len(nz) else -1
        if stats["candidates"] == 0:
            return hostops._empty_cols(), stats
        ext = res["totals"][4:8]
        span = float(int(ext[3]) - int(ext[0]))
        rpb = float(pts.n) / span
        cols = hostops._cols_from_stats(res["coords"], res["core"], res["hyp"], res["stats"], 1, pts.n)
        if exact:
            cols = hostops.exact_cuts(cols, pts.n, rpb, cis=False)
        stats["tested"] = int(cols["coords"].shape[0])
        final = hostops.select_twice(cols, hostops.mark_sig_trans(cols))
        if exact and exact != "defer":
            final = hostops.scipy_pvalues(final, pts.n, rpb, cis=False)
        final["N"], final["rpb"] = pts.n, rpb
        if exact == "defer":
            hostops.submit_pvalues(final, cis=False)          # evaluated by the process-wide pool while other files run
        stats["loops"] = int(final["coords"].shape[0])
        return final, stats
    finally:
        if own:
            pts.close()


def _combine_across_passes(per_pass):
    seen = set()
    out = []
    for c in per_pass:
        rows = [tuple(r) for r in c.tolist()]
        out.extend(r for r in rows if r not in seen)
        seen.update(rows)
    return np.array(out, dtype=np.int64).reshape(-1, 4)


def runTransDBSCANLoops(fixy, eps, minPts):
    """Drop-in for callTransLoops.py:33-84."""
    key = ixyKey(fixy)
    ctx = _lib.default_context()
    pts = _lib.Points(ctx, loadIxy(fixy))
    try:
        c, s = cluster_pass(pts, eps, minPts, hic=False, cis=False) if pts.n else (np.zeros((0, 4), np.int64), np.zeros(0, np.int64))
    finally:
        pts.close()
    loops = []
    for i in range(c.shape[0]):
        lp = Loop()
        lp.rab = int(s[i])
        lp.chromX, lp.chromY = key
        lp.x_start, lp.x_end, lp.y_start, lp.y_end = (int(v) for v in c[i])
        lp.x_center = (lp.x_start + lp.x_end) / 2
        lp.y_center = (lp.y_start + lp.y_end) / 2
        lp.distance = -1
        lp.cis = False
        loops.append(lp)
    return "-".join(key), loops


def estLoopSig(key, loops, fixy, minPts=5, pseudo=1, peakPcut=1e-5, peakFccut=2, countDiffCut=10):
    """Drop-in for callTransLoops.py:106-193."""
    ctx = _lib.default_context()
    pts = _lib.Points(ctx, loadIxy(fixy))
    try:
        ext = pts.extent()
        index = _lib.Index(pts)
        try:
            cols = hostops.est_trans(index, _loops_to_cands(loops), pts.n, float(int(ext[3]) - int(ext[0])), minPts,
                                     pseudo=pseudo, countDiffCut=countDiffCut, exact=exact_default())
        finally:
            index.close()
    finally:
        pts.close()
    out = cols_to_loops(tuple(key.split("-")), cols, cis=False)
    for lp in out:
        del lp.significant
    return key, out


def markSigLoops(key, loops):
    """Drop-in for callTransLoops.py:196-206."""
    for lp in loops:
        lp.significant = 1 if (lp.binomial_p_value <= 1e-10 and lp.FDR <= 0.05 and lp.ES >= 2) else 0
    return key, loops


def callTransLoops(predir, fout, logger, eps=[2000, 5000], minPts=[5, 10], cpu=1, filter=False, washU=False,
                   juicebox=False):
    """Drop-in for callTransLoops.py:209-275; chromosome pairs sharded over ranks by PET count."""
    meta = json.loads(open(predir + "/petMeta.json").read())
    keys = list(meta["data"]["trans"].keys())
    files = {k: meta["data"]["trans"][k]["ixy"] for k in keys}
    rank, world = shard.rank_world()
    mine = shard.assign(keys, [shard.ixy_rows(files[k]) for k in keys], world)[rank]
    ex = "defer" if exact_default() else False
    if ex:
        hostops.pvalue_pool()              # the scipy workers start (and import scipy) while the files are loaded
    ctxs = shard.worker_contexts()

    def one(ctx, k):
        return trans_pair(ctx, tuple(k.split("-")), shard.load_pinned(ctx, files[k]), eps, minPts, exact=ex)

    results = {}
    first = {}
    stats = np.zeros(4, dtype=np.int64)
    prev_mode = shard.set_sync_mode(True) if ex else None      # sleeping waits leave the cores to the scipy pool
    try:
        done = shard.map_units(ctxs, mine, one)
    finally:
        if ex:
            shard.set_sync_mode(prev_mode)
    for k, (final, st) in zip(mine, done):
        stats += np.array([st["pets"], st["candidates"], st["tested"], st["loops"]], dtype=np.int64)
        if st["first_pass"] >= 0:
            results[k] = final
            first[k] = st["first_pass"]
    if ex:
        hostops.wait_pvalues([results[k] for k in results])
    stats = shard.allreduce_stats(stats)
    gathered = shard.gather_results({k: (first[k], results[k]) for k in results})
    if rank != 0:
        return
    logger.info("Estimating loop statstical significance.")
    logger.info("Selecting the most significant loops of overlapped ones. ")
    loops = []
    # dict order of the reference: a pair enters in the run that first finds a candidate for it (:100-103, geo.py:94-96)
    order = sorted((k for k in keys if k in gathered), key=lambda k: (gathered[k][0], keys.index(k)))
    logger.info("Output %s loops to %s_loops.txt" % (sum(int(gathered[k][1]["coords"].shape[0]) for k in order), fout))
    cols2txt([(tuple(k.split("-")), gathered[k][1]) for k in order], fout + "_trans_loops.txt", cis=False)
    if washU or juicebox:
        for k in order:
            loops.extend(cols_to_loops(tuple(k.split("-")), gathered[k][1], cis=False))
    if washU:
        loops2washuTxt(loops, fout + "_trans_loops_washU.txt")
    if juicebox:
        loops2juiceTxt(loops, fout + "_trans_loops_juicebox.txt")
    return stats
