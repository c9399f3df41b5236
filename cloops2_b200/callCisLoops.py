"""
Intra-chromosomal loop calling on B200 -- drop-in for cLoops2/callCisLoops.py on the `callLoops` path.

Same public functions and argument meaning as the reference (runCisDBSCANLoops :58, estLoopSig :250, markSigLoops
:357, selSigLoops :379, callCisLoops :471), but the work is organised around the device: each chromosome file is
uploaded ONCE (pinned staging -> int32 coordinates in HBM), stays resident across the whole eps x minPts sweep
(clustering + candidate boxes per pass), then the XY index is built on the device and the candidates are counted by
the rectangle-count kernel; only candidate boxes and per-loop counts cross PCIe.  The reference's joblib fan-out
over chromosomes (:134, :561) becomes chromosome sharding over ranks (one process per GPU, cloops2_b200/shard.py).
"""
import json
import os
from glob import glob

import numpy as np

from . import _lib, hostops, shard
from .ds import Loop
from .io import (ixyKey, loadIxy, parseIxy, updateJson, loops2txt, cols2txt, loops2ucscTxt, loops2juiceTxt,  # noqa: F401
                 loops2washuTxt, loops2NewWashuTxt)

logger = None


def _log(msg):
    if logger is not None:
        logger.info(msg)


def _sweep(eps, minPts, emPair):
    """Order of clustering runs (callCisLoops.py:529-551): eps-major, or zipped pairs with -emPair."""
    return list(zip(eps, minPts)) if emPair else [(e, m) for e in eps for m in minPts]


# ---------------------------------------------------------------------------------------------- device passes
def cluster_pass(pts, eps, minPts, hic=False, cis=True):
    """One DBSCAN run on resident points + candidate filters.  Returns (cands int64[C,4], sizes int64[C]) in
    ascending cluster id, which is the order the reference collects them in (callCisLoops.py:92-120)."""
    if hic:
        pts.cdbscan2(eps, minPts, want_labels=False)
    else:
        pts.block_dbscan(eps, minPts, want_labels=False)
    bbox, size = pts.cluster_bbox()
    ok = hostops.cis_candidate_mask(bbox) if cis else hostops.trans_candidate_mask(bbox)
    return bbox[ok], size[ok]


def _rows_carry_over(E, m, e, m2):
    """May a run (e, m2) skip the rows in noise cells removed by run (E, m)?  (include/cl2.h, cl2_points_sweep_hint)"""
    import math
    return e < E and m2 >= m and E >= 3 * e - math.gcd(int(e), int(E))


def sweep_passes(pts, passes, hic=False, cis=True, log=None):
    """All clustering runs of a sweep, results in sweep order.  The blockDBSCAN runs are executed from the largest
    eps down (smallest minPts first): every run then tells the later ones which rows sit in noise cells it removed,
    and those runs bin and sort only the remaining rows (cl2_points_sweep_hint).  The result of a run does not depend
    on the order."""
    order = list(range(len(passes)))
    hint = not hic and len(set(e for e, _ in passes)) > 1
    if hint:
        order.sort(key=lambda i: (-passes[i][0], passes[i][1], i))
    out = [None] * len(passes)
    for pos, i in enumerate(order):
        ep, mp = passes[i]
        if hint:
            # remembering rows costs a pass over the points: only when a later run can use them (the library checks
            # the same condition again before it does)
            useful = any(_rows_carry_over(ep, mp, *passes[j]) for j in order[pos + 1:])
            pts.sweep_hint(1 if useful else 2)
        out[i] = cluster_pass(pts, ep, mp, hic=hic, cis=cis)
        if log is not None:
            log(ep, mp, out[i])
    return out


def exact_default():
    """Whether final loops get their tail probabilities re-evaluated by scipy (the reference's printed digits) or keep
    the device fp64 values.  CL2_PVALUES=device|scipy overrides; the drop-in entry points default to scipy."""
    return os.environ.get("CL2_PVALUES", "scipy").lower() != "device"


def cis_chromosome(ctx, key, xy, tot, eps, minPts, cut=0, mcut=-1, hic=False, emPair=False, pts=None, exact=False):
    """Whole per-chromosome path: sweep -> combine -> count -> test -> mark -> select, the clustering runs, candidate
    filters, estLoopSig and the de-duplication in ONE native call (cl2_call_loops).
    xy: host int64[N,2] (ignored when a resident `pts` is passed).  Returns (final cols, stats dict).
    exact: True re-evaluates the tail probabilities of the FINAL loops with scipy (hostops.scipy_pvalues) so the
    printed digits are the reference's; "defer" leaves that to the caller (hostops.scipy_pvalues_many over all
    files of a rank at once) and only makes the cut decisions exact."""
    own = pts is None
    if own:
        pts = _lib.Points(ctx, xy, cut=cut, mcut=mcut)
    try:
        stats = {"pets": pts.n, "candidates": 0, "tested": 0, "loops": 0, "first_pass": -1}
        if pts.n == 0:
            return hostops._empty_cols(), stats
        passes = _sweep(eps, minPts, emPair)
        mm = min(minPts) if emPair else max(minPts)                                 # :557-560
        ctx.set_pvalue_margin(hostops.PV_MARGIN if exact else 0.0)
        res = pts.call_loops(passes, mode=_lib.MODE_CIS, hic=hic, cut=cut, min_pts_test=mm, want_unique=logger is not None)
        for (ep, mp), nc, npets in zip(passes, res["pass_cand"], res["pass_pets"]):
            _log("Clustering %s and %s using eps %s, minPts %s,pre-set distance cutoff > %s finished. Estimated %s "
                 "candidate loops with %s PETs." % (key[0], key[1], ep, mp, cut, int(nc), int(npets)))
        tested, unique = int(res["totals"][1]), int(res["totals"][2])
        stats["candidates"] = unique if unique >= 0 else tested
        nz = np.nonzero(res["pass_cand"])[0]
        stats["first_pass"] = int(nz[0]) if len(nz) else -1     # the reference's dicts list a chromosome from that run on
        stats["host_s"] = dict(zip(("cluster", "index", "test", "cluster_cpu", "index_cpu", "test_cpu"),
                                   (float(v) for v in res["seconds"])))
        if tested == 0:
            return hostops._empty_cols(), stats
        ext = res["totals"][4:8]
        span = float(int(ext[3]) - int(ext[0]))
        rpb = float(pts.n) / span
        _log("Estimate significance for %s candidate interactions in %s with %s PETs distance > =%s and <=%s,requiring minPts >=%s."
             % (stats["candidates"], "-".join(key), pts.n, cut, mcut, mm))
        import time
        t3, c3 = time.perf_counter(), time.thread_time()
        cols = hostops._cols_from_stats(res["coords"], res["core"], res["hyp"], res["stats"], 1, tot)
        if exact:
            cols = hostops.exact_cuts(cols, pts.n, rpb, cis=True, hic=hic)
        stats["tested"] = int(cols["coords"].shape[0])
        sig = hostops.mark_sig_cis(cols, hic=hic)
        final = hostops.select_twice(cols, sig)
        if exact and exact != "defer":
            final = hostops.scipy_pvalues(final, pts.n, rpb, cis=True)
        final["N"], final["rpb"] = pts.n, rpb
        if exact == "defer":
            hostops.submit_pvalues(final, cis=True)          # evaluated by the process-wide pool while other files run
        stats["host_s"]["select"] = time.perf_counter() - t3
        stats["host_s"]["select_cpu"] = time.thread_time() - c3
        stats["loops"] = int(final["coords"].shape[0])
        return final, stats
    finally:
        if own:
            pts.close()


def cols_to_loops(key, cols, cis=True):
    """Column arrays -> Loop objects (cLoops2/ds.py:245-281), Python scalars so str() prints as the reference's."""
    out = []
    c = cols["coords"]
    for i in range(c.shape[0]):
        lp = Loop()
        lp.chromX, lp.chromY = key[0], key[1]
        lp.x_start, lp.x_end, lp.y_start, lp.y_end = (int(v) for v in c[i])
        lp.x_center = (lp.x_start + lp.x_end) / 2
        lp.y_center = (lp.y_start + lp.y_end) / 2
        lp.cis = cis
        lp.distance = abs(lp.y_center - lp.x_center) if cis else -1
        lp.ra, lp.rb, lp.rab = int(cols["ra"][i]), int(cols["rb"][i]), int(cols["rab"][i])
        lp.P2LL, lp.FDR, lp.ES, lp.density = (float(cols[k][i]) for k in ("P2LL", "FDR", "ES", "density"))
        lp.hypergeometric_p_value = float(cols["hyp"][i])
        lp.poisson_p_value = float(cols["pop"][i])
        lp.binomial_p_value = float(cols["nbp"][i])
        lp.x_peak_poisson_p_value, lp.x_peak_es = float(cols["px"][i]), float(cols["esx"][i])
        lp.y_peak_poisson_p_value, lp.y_peak_es = float(cols["py"][i]), float(cols["esy"][i])
        lp.significant = 1
        out.append(lp)
    return out


# ---------------------------------------------------------------------------------------------- reference API
def runCisDBSCANLoops(fixy, eps, minPts, cut=0, mcut=-1, hic=False):
    """Drop-in for callCisLoops.py:58-124: ("chrA-chrA", [Loop]) or None.  (`hic` selects cDBSCAN2, which the
    reference does through the module-global DBSCAN, :498-504.)"""
    key = ixyKey(fixy)
    if key[0] != key[1]:
        return None
    ctx = _lib.default_context()
    pts = _lib.Points(ctx, loadIxy(fixy), cut=cut, mcut=mcut)
    try:
        if pts.n == 0:
            _log("No PETs found in %s, maybe due to cut > %s" % (fixy, cut))
            return None
        c, s = cluster_pass(pts, eps, minPts, hic=hic, cis=True)
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
        lp.cis = True
        lp.distance = abs(lp.y_center - lp.x_center)
        loops.append(lp)
    return "-".join(key), loops


def _loops_to_cands(loops):
    return np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops], dtype=np.int64).reshape(-1, 4)


def estLoopSig(key, loops, fixy, tot, minPts=5, pseudo=1, cut=0, mcut=-1, peakPcut=1e-5, win=5, countDiffCut=20,
               hic=False):
    """Drop-in for callCisLoops.py:250-354: (key, [Loop]) with the surviving loops annotated."""
    ctx = _lib.default_context()
    pts = _lib.Points(ctx, loadIxy(fixy), cut=cut, mcut=mcut)
    try:
        ext = pts.extent()
        index = _lib.Index(pts)
        try:
            cols = hostops.est_cis(index, _loops_to_cands(loops), pts.n, float(int(ext[3]) - int(ext[0])), tot, minPts,
                                   hic=hic, pseudo=pseudo, peakPcut=peakPcut, win=win, countDiffCut=countDiffCut,
                                   exact=exact_default())
        finally:
            index.close()
    finally:
        pts.close()
    out = cols_to_loops(tuple(key.split("-")), cols, cis=True)
    for lp in out:
        del lp.significant
    return key, out


def markSigLoops(key, loops, hic=False):
    """Drop-in for callCisLoops.py:357-376."""
    for lp in loops:
        sig = lp.FDR <= 0.05 and lp.ES >= 2 and lp.hypergeometric_p_value <= 1e-5 and lp.poisson_p_value <= 1e-5
        lp.significant = 1 if (sig and lp.binomial_p_value < (1e-3 if hic else 1e-1)) else 0
    return key, loops


def selSigLoops(key, loops):
    """Drop-in for callCisLoops.py:379-422 (native selection in libcl2.so)."""
    loops = [lp for lp in loops if lp.significant > 0]
    if not loops:
        return key, []
    pick = hostops.sel_sig(_loops_to_cands(loops), np.array([lp.ES for lp in loops], dtype=np.float64))
    return key, [loops[i] for i in pick]


def getAllAnchors(loops, margin=1):
    """Merged anchor intervals (callCisLoops.py:425-447).  The reference enumerates every covered base pair; here the
    intervals are merged directly.  Its scan closes the LAST interval one covered position early (`end = cov[j-1]`
    after the inner loop runs out, :441-444) -- reproduced."""
    iv = sorted([(lp.x_start, lp.x_end) for lp in loops] + [(lp.y_start, lp.y_end) for lp in loops])
    merged = []
    for s, e in iv:
        if merged and s - merged[-1][1] <= margin:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])
    if merged:
        merged[-1][1] -= 1           # anchors hold >= 2 positions (x_start != x_end), so this never empties one
    return merged


def filterPETs(key, predir, fixy, loops, margin=1):
    """-filter (callCisLoops.py:450-468): keep PETs with an end inside any merged anchor.  The reference collects the
    row ids in Python sets and writes `mat[list(set)]`; the same insertions are replayed in the same order (rows by
    ascending coordinate, ties by row, X hits then Y hits, anchor by anchor) so the row order of the new .ixy is the
    reference's."""
    _log("Filtering PETs of %s with %s loops." % (key, len(loops)))
    anchors = getAllAnchors(loops, margin=margin)
    key2, mat = parseIxy(fixy)
    ordx = np.argsort(mat[:, 0], kind="stable")
    ordy = np.argsort(mat[:, 1], kind="stable")
    xs, ys = mat[ordx, 0], mat[ordy, 1]
    rs = set()
    for l, r in anchors:
        xps, yps = set(), set()
        xps.update(ordx[np.searchsorted(xs, l, side="left"):np.searchsorted(xs, r, side="right")].tolist())
        yps.update(ordy[np.searchsorted(ys, l, side="left"):np.searchsorted(ys, r, side="right")].tolist())
        rs.update(xps.union(yps))
    rs = list(rs)
    if len(rs) == 0:
        return
    import joblib
    joblib.dump(mat[rs, ], predir + "/" + "-".join(key2) + ".ixy")


def callCisLoops(predir, fout, log, eps=[2000, 5000], minPts=[5, 10], cpu=1, cut=0, mcut=-1, plot=False,
                 max_cut=False, hic=False, filter=False, ucsc=False, juicebox=False, washU=False, emPair=False):
    """Drop-in for callCisLoops.py:471-628.  `cpu` is accepted for compatibility; parallelism is one process per GPU
    (launch under torchrun for more than one), chromosomes sharded by PET count.  Returns the job's counters
    (PETs, candidates, tested, loops) on rank 0, None on the other ranks and on the reference's early returns."""
    global logger
    logger = log
    if hic:
        _log("-hic option selected, cDBSCAN2 is used instead of blockDBSCAN.")
    if emPair and len(eps) != len(minPts):
        _log("-emPair option selected, number of eps not equal to minPts, return.")
        return
    meta = json.loads(open(predir + "/petMeta.json").read())
    tot = meta["Unique PETs"]
    rank, world = shard.rank_world()
    fdir = fout + "_filtered"
    if filter:
        # the reference's early return (:518-524) is taken by every rank or by none: rank 0 looks, all ranks hear
        ok = 1
        if rank == 0:
            _log("-filter option chosed, will filter raw PETs based on called loops, for any PET that any end overlaps loop anchors will be kept. ")
            if not os.path.exists(fdir):
                os.mkdir(fdir)
            elif len(os.listdir(fdir)) > 0:
                if logger is not None:
                    logger.error("working directory %s exists and not empty." % fdir)
                ok = 0
        if shard.broadcast_flag(ok) == 0:
            return
    keys = list(meta["data"]["cis"].keys())
    files = {k: meta["data"]["cis"][k]["ixy"] for k in keys}
    mine = shard.assign(keys, [shard.ixy_rows(files[k]) for k in keys], world)[rank]
    ex = "defer" if exact_default() else False
    if ex:
        hostops.pvalue_pool()              # the scipy workers start (and import scipy) while the files are loaded
    ctxs = shard.worker_contexts()

    def one(ctx, k):
        return cis_chromosome(ctx, tuple(k.split("-")), shard.load_pinned(ctx, files[k]), tot, eps, minPts, cut=cut, mcut=mcut,
                              hic=hic, emPair=emPair, exact=ex)

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
    _log("Estimating loop statstical significance.")
    _log("Selecting the most significant loops of overlapped ones. ")
    per_key = {}
    loops = []
    # the reference's dict order: chromosomes enter `loops` in the run that first finds a candidate for them, in meta
    # key order within a run (parallelRunCisDBSCANLoops :140-143 skips empty results, combineLoops appends new keys)
    order = sorted((k for k in keys if k in gathered), key=lambda k: (gathered[k][0], keys.index(k)))
    nloops = sum(int(gathered[k][1]["coords"].shape[0]) for k in order)
    _log("Output %s loops to %s_loops.txt" % (nloops, fout))
    # the loop table straight from the column arrays (same bytes as loops2txt on Loop objects, tests/test_text_format.py);
    # Loop objects are only built for the track writers and -filter
    cols2txt([(tuple(k.split("-")), gathered[k][1]) for k in order], fout + "_loops.txt", cis=True)
    if ucsc or juicebox or washU or filter:
        for k in order:
            per_key[k] = cols_to_loops(tuple(k.split("-")), gathered[k][1], cis=True)
            loops.extend(per_key[k])
    if ucsc:
        loops2ucscTxt(loops, fout + "_loops_ucsc.interact")
    if juicebox:
        loops2juiceTxt(loops, fout + "_loops_juicebox.txt")
    if washU:
        loops2washuTxt(loops, fout + "_loops_legacyWashU.txt")
        loops2NewWashuTxt(loops, fout + "_loops_newWashU.txt")
    if filter:
        for k in per_key:
            filterPETs(k, fdir, files[k], per_key[k], margin=max(eps))
        ixyfs = glob(fdir + "/*.ixy")
        n = 0
        for f in ixyfs:
            n += loadIxy(f).shape[0]
        with open(fdir + "/petMeta.json", "w") as fo:
            json.dump({"Unique PETs": n}, fo)
        updateJson(ixyfs, fdir + "/petMeta.json")
    return stats
