"""
TEST INFRASTRUCTURE ONLY -- plain-Python restatement of the host-side logic of the reference's callLoops path.
Each function cites the reference lines (path:line under /root/reference) it follows.  Small-case loops on
purpose: this is the checker, not the product (the product's host code is cloops2_b200/*.py, vectorised and
driven by the CUDA library).
"""
import numpy as np
from scipy.stats import hypergeom, binom, poisson


class OLoop(object):
    """Field-for-field stand-in for cLoops2/ds.py:245-281 (Loop)."""
    __slots__ = ["chromX", "chromY", "x_start", "x_end", "x_center", "ra", "y_start", "y_end", "y_center", "rb",
                 "rab", "cis", "distance", "density", "ES", "P2LL", "FDR", "hypergeometric_p_value",
                 "poisson_p_value", "binomial_p_value", "x_peak_poisson_p_value", "x_peak_es",
                 "y_peak_poisson_p_value", "y_peak_es", "significant", "id"]

    def coords(self):
        return (self.chromX, self.x_start, self.x_end, self.chromY, self.y_start, self.y_end)


def parse_ixy_filter(xy, cut=0, mcut=-1):
    """cLoops2/io.py:274-291 (parseIxy) on an in-memory int64[N,2] array."""
    xy = np.asarray(xy, dtype=np.int64)
    if cut > 0:
        xy = xy[(xy[:, 1] - xy[:, 0]) >= cut]
    if mcut > 0:
        xy = xy[(xy[:, 1] - xy[:, 0]) <= mcut]
    return xy


def _bbox_loops(key, xy, labels, ncl, cis):
    from . import cluster_bbox
    bbox, size = cluster_bbox(xy, labels, ncl)
    out = []
    for c in range(ncl):
        if size[c] == 0:
            continue
        x0, x1, y0, y1 = (int(v) for v in bbox[c])
        if x0 == x1 or y0 == y1:                      # callCisLoops.py:97-99
            continue
        lp = OLoop()
        lp.rab = int(size[c])
        lp.chromX, lp.chromY = key[0], key[1]
        lp.x_start, lp.x_end = x0, x1
        lp.x_center = (x0 + x1) / 2
        lp.y_start, lp.y_end = y0, y1
        lp.y_center = (y0 + y1) / 2
        lp.cis = cis
        out.append(lp)
    return out


def cis_candidates(key, xy, eps, minPts, cut=0, hic=False):
    """callCisLoops.py:58-124 (runCisDBSCANLoops) after parseIxy; `xy` is already cut/mcut filtered."""
    from . import block_dbscan, cdbscan2
    if key[0] != key[1]:
        return None
    if cut > 0:                                       # :75-79 (idempotent second filter)
        xy = xy[(xy[:, 1] - xy[:, 0]) >= cut]
    if len(xy) == 0:
        return None
    labels, ncl = (cdbscan2 if hic else block_dbscan)(xy, eps, minPts)
    loops = []
    for lp in _bbox_loops(key, xy, labels, ncl, True):
        lp.distance = abs(lp.y_center - lp.x_center)
        if lp.x_end - lp.x_start + lp.y_end - lp.y_start < 200:     # :114
            continue
        if lp.x_end < lp.y_start:                                   # :116
            loops.append(lp)
    return "-".join(key), loops


def trans_candidates(key, xy, eps, minPts):
    """callTransLoops.py:33-84 (runTransDBSCANLoops): blockDBSCAN only, no anchor-size / order filters."""
    from . import block_dbscan
    labels, ncl = block_dbscan(xy, eps, minPts)
    loops = _bbox_loops(key, xy, labels, ncl, False)
    for lp in loops:
        lp.distance = -1
    return "-".join(key), loops


def combine_loops(loops, loops_2):
    """geo.py:90-112."""
    for key in loops_2.keys():
        if key not in loops:
            loops[key] = loops_2[key]
        else:
            seen = set(lp.coords() for lp in loops[key])
            for lp in loops_2[key]:
                if lp.coords() not in seen:
                    loops[key].append(lp)
    return loops


def filter_by_dis(loops, cut):
    """callCisLoops.py:146-156."""
    return {k: [lp for lp in v if lp.distance > cut] for k, v in loops.items()}


def filter_dup(loops):
    """callCisLoops.py:159-169: first position, last object."""
    out = {}
    for k, v in loops.items():
        d = {}
        for lp in v:
            d[lp.coords()] = lp
        out[k] = list(d.values())
    return out


def est_loop_sig_cis(key, loops, xy, tot, minPts=5, pseudo=1, peakPcut=1e-5, win=5, countDiffCut=20, hic=False):
    """callCisLoops.py:250-354 (estLoopSig); `xy` is the cut/mcut-filtered array (ixy2pet, io.py:294-300)."""
    from . import XYIndex
    if len(loops) == 0:
        return key, []
    idx = XYIndex(xy)
    N = idx.n
    span = float(np.max(xy[:, 1]) - np.min(xy[:, 0]))
    rpb = float(N) / span
    arr = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops], dtype=np.int64)
    core, na, nb, nab, anc = idx.loop_counts(arr, win=win, mode=0)
    out = []
    for i, lp in enumerate(loops):
        la, lb = lp.x_end - lp.x_start, lp.y_end - lp.y_start
        if la / lb > countDiffCut or lb / la > countDiffCut:        # :282-286
            continue
        ra, rb, rab, lower = (int(v) for v in core[i])
        if rab < minPts:                                            # :290
            continue
        if ra == rab or rb == rab:                                  # :293
            continue
        if ra / float(rb) > countDiffCut or rb / float(ra) > countDiffCut:   # :296
            continue
        lp.ra, lp.rb, lp.rab = ra, rb, rab
        lp.P2LL = float(rab) / max(lower, pseudo)                   # :306
        if hic and lp.P2LL < 1:
            continue
        hyp = max([1e-300, hypergeom.sf(rab - 1.0, N, ra, rb)])     # :310
        if hyp > 1e-2:
            continue
        rabs, nbps = [], []                                          # :207-223
        for a in range(2 * win):
            nac = float(na[i, a])
            for b in range(2 * win):
                nbc = float(nb[i, b])
                nrab = float(nab[i, a * 2 * win + b])
                if nrab > 0:
                    rabs.append(nrab)
                    nbps.append(nrab / (nac * nbc))
                else:
                    rabs.append(0)
                    nbps.append(0.0)
        rabs, nbps = np.array(rabs), np.array(nbps)
        mrabs = float(np.median(rabs))
        mbps = np.median(nbps)
        fdr = len(rabs[rabs > rab]) / float(len(rabs)) if len(rabs) > 0 else 0.0
        if mrabs >= rab or fdr >= 0.1:                              # :322
            continue
        es = rab / max(mrabs, pseudo)
        if es < 1:
            continue
        pop = max([1e-300, poisson.sf(rab - 1.0, mrabs)])
        nbp = max([1e-300, binom.sf(rab - 1.0, ra * rb - rab, mbps)])
        lp.FDR, lp.ES = fdr, es
        lp.density = float(rab) / (la + lb) / tot * 10.0**9
        lp.hypergeometric_p_value, lp.poisson_p_value, lp.binomial_p_value = hyp, pop, nbp
        res = []
        for a, (left, right) in enumerate(((lp.x_start, lp.x_end), (lp.y_start, lp.y_end))):   # :226-247
            length = right - left
            count = int(anc[i, 2 * a])
            r = (int(anc[i, 2 * a + 1]) - count) / 5 / 2
            c = float(max([r, rpb * length, 1]))
            res.append((max([1e-300, poisson.sf(count - 1.0, c)]), count / c))
        (px, esx), (py, esy) = res
        if hic is False and not (px < peakPcut and py < peakPcut):  # :347
            continue
        lp.x_peak_poisson_p_value, lp.x_peak_es = px, esx
        lp.y_peak_poisson_p_value, lp.y_peak_es = py, esy
        out.append(lp)
    return key, out


def est_loop_sig_trans(key, loops, xy, minPts=5, pseudo=1, countDiffCut=10, win=5):
    """callTransLoops.py:106-193 (trans estLoopSig)."""
    from . import XYIndex
    if len(loops) == 0:
        return key, []
    idx = XYIndex(xy)
    N = idx.n
    span = float(np.max(xy[:, 1]) - np.min(xy[:, 0]))
    rpb = float(N) / span
    arr = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops], dtype=np.int64)
    core, na, nb, nab, anc = idx.loop_counts(arr, win=win, mode=1)
    out = []
    for i, lp in enumerate(loops):
        ra, rb, rab, lower = (int(v) for v in core[i])
        if rab < minPts:
            continue
        lp.ra, lp.rb, lp.rab = ra, rb, rab
        if ra / float(rb) > countDiffCut or rb / float(ra) > countDiffCut:
            continue
        la, lb = lp.x_end - lp.x_start, lp.y_end - lp.y_start
        if la / lb > countDiffCut or lb / la > countDiffCut:
            continue
        lp.P2LL = float(rab) / max(lower, pseudo)
        res = []
        for a, (left, right) in enumerate(((lp.x_start, lp.x_end), (lp.y_start, lp.y_end))):
            length = right - left
            count = int(anc[i, 2 * a])
            r = (int(anc[i, 2 * a + 1]) - count) / 5 / 2
            c = float(max([r, rpb * length, 1]))
            res.append((max([1e-300, poisson.sf(count - 1.0, c)]), count / c))
        (lp.x_peak_poisson_p_value, lp.x_peak_es), (lp.y_peak_poisson_p_value, lp.y_peak_es) = res
        hyp = max([1e-300, hypergeom.sf(rab - 1.0, N, ra, rb)])
        rabs, nbps = [], []
        for a in range(2 * win):
            nac = float(na[i, a])
            for b in range(2 * win):
                nbc = float(nb[i, b])
                nrab = float(nab[i, a * 2 * win + b])
                rabs.append(nrab)
                if nac > 0 and nbc > 0:
                    nbps.append(nrab / (nac * nbc))
                else:
                    nbps.append(0.0)
        rabs, nbps = np.array(rabs), np.array(nbps)
        mrabs = float(np.mean(rabs))
        mbps = np.mean(nbps)
        fdr = len(rabs[rabs > rab]) / float(len(rabs))
        es = rab / max(mrabs, pseudo)
        pop = max([1e-300, poisson.sf(rab - 1.0, mrabs)])
        nbp = max([1e-300, binom.sf(rab - 1.0, ra * rb, mbps)])
        lp.FDR, lp.ES = fdr, es
        lp.density = float(rab) / (la + lb) / N * 10.0**9
        lp.hypergeometric_p_value, lp.poisson_p_value, lp.binomial_p_value = hyp, pop, nbp
        out.append(lp)
    return key, out


def mark_sig_cis(loops, hic=False):
    """callCisLoops.py:357-376."""
    for lp in loops:
        sig = lp.FDR <= 0.05 and lp.ES >= 2 and lp.hypergeometric_p_value <= 1e-5 and lp.poisson_p_value <= 1e-5
        if sig:
            lp.significant = 1 if lp.binomial_p_value < (1e-3 if hic else 1e-1) else 0
        else:
            lp.significant = 0
    return loops


def mark_sig_trans(loops):
    """callTransLoops.py:196-206.  The reference's lambda reads `loop.ES` from the enclosing loop variable,
    which at call time is the loop being tested, so it is the plain three-way test."""
    for lp in loops:
        lp.significant = 1 if (lp.binomial_p_value <= 1e-10 and lp.FDR <= 0.05 and lp.ES >= 2) else 0
    return loops


def _anchor_ov(xa, xb, ya, yb):
    """geo.py:65-73."""
    if (ya <= xa <= yb) or (ya <= xb <= yb) or (ya <= xa <= xb <= yb):
        return True
    if (xa <= ya <= xb) or (xa <= yb <= xb) or (xa <= ya <= yb <= xb):
        return True
    return False


def _loop_ov(a, b):
    """geo.py:76-87."""
    if a.chromX != b.chromX or a.chromY != b.chromY:
        return False
    return _anchor_ov(a.x_start, a.x_end, b.x_start, b.x_end) and _anchor_ov(a.y_start, a.y_end, b.y_start, b.y_end)


def sel_sig_loops(loops):
    """callCisLoops.py:379-422."""
    loops = [lp for lp in loops if lp.significant > 0]
    groups, skips = [], set()
    for i in range(len(loops)):
        if i in skips:
            continue
        n = [loops[i]]
        for j in range(i + 1, len(loops)):
            for p in n:
                if _loop_ov(p, loops[j]):
                    n.append(loops[j])
                    skips.add(j)
                    break
        groups.append(n)
    reps = []
    for n in groups:
        if len(n) == 1:
            reps.append(n[0])
        else:
            for i in range(len(n) - 1):
                for j in range(i + 1, len(n)):
                    if n[i].ES < n[j].ES:
                        n[i], n[j] = n[j], n[i]
            reps.append(n[0])
    for la in loops:
        if not any(_loop_ov(la, lb) for lb in reps):
            reps.append(la)
    return reps


def call_cis_loops(files, tot, eps, minPts, cut=0, mcut=-1, hic=False, emPair=False):
    """callCisLoops.py:471-599 on in-memory inputs.  files: ordered dict "chrA-chrA" -> int64[N,2]."""
    loops = {}
    pairs = list(zip(eps, minPts)) if emPair else [(e, m) for e in eps for m in minPts]
    filt = {k: parse_ixy_filter(v, cut, mcut) for k, v in files.items()}
    for ep, mp in pairs:
        cur = {}
        for k in files:
            r = cis_candidates(tuple(k.split("-")), filt[k], ep, mp, cut=cut, hic=hic)
            if r is not None and len(r[1]) > 0:
                cur[r[0]] = r[1]
        loops = combine_loops(loops, cur)
    loops = filter_dup(filter_by_dis(loops, cut))
    mm = min(minPts) if emPair else max(minPts)
    final = []
    per_key = {}
    for k in loops:
        _, ls = est_loop_sig_cis(k, loops[k], filt[k], tot, minPts=mm, hic=hic)
        ls = mark_sig_cis(ls, hic=hic)
        ls = sel_sig_loops(sel_sig_loops(ls))
        per_key[k] = ls
        final.extend(ls)
    return final, per_key


def call_trans_loops(files, eps, minPts):
    """callTransLoops.py:209-275 on in-memory inputs.  files: ordered dict "chrA-chrB" -> int64[N,2]."""
    loops = {}
    for ep in eps:
        for mp in minPts:
            cur = {}
            for k in files:
                r = trans_candidates(tuple(k.split("-")), np.asarray(files[k], dtype=np.int64), ep, mp)
                if len(r[1]) > 0:
                    cur[r[0]] = r[1]
            loops = combine_loops(loops, cur)
    final = []
    for k in loops:
        _, ls = est_loop_sig_trans(k, loops[k], np.asarray(files[k], dtype=np.int64), minPts=max(minPts))
        ls = mark_sig_trans(ls)
        ls = sel_sig_loops(sel_sig_loops(ls))
        final.extend(ls)
    return final


def loops_to_rows(loops):
    """Rows of io.py:517-561 (loops2txt) as lists of str, header excluded."""
    rows = []
    for i, lp in enumerate(loops):
        line = ["loop_%s-%s-%s" % (lp.chromX, lp.chromY, i), lp.chromX, lp.x_start, lp.x_end, lp.chromY, lp.y_start,
                lp.y_end, lp.distance, lp.x_center, lp.y_center, lp.ra, lp.rb, lp.cis, lp.rab, lp.density, lp.ES,
                lp.P2LL, lp.FDR, lp.binomial_p_value, lp.hypergeometric_p_value, lp.poisson_p_value,
                lp.x_peak_poisson_p_value, lp.y_peak_poisson_p_value, lp.significant]
        rows.append(list(map(str, line)))
    return rows


def quant_loops(key, loops, xy, tot, pseudo=1, offp=False):
    """cLoops2/quant.py:129-237 (_quantLoops) on an in-memory, already cut/mcut-filtered array.  loops: OLoop list
    with coordinates set; returns the same objects annotated."""
    from . import XYIndex
    tot = float(tot)
    idx = XYIndex(xy)
    N = idx.n
    rpb = float(N) / (np.max(xy[:, 1]) - np.min(xy[:, 0]))
    arr = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops], dtype=np.int64)
    core, na, nb, nab, anc = idx.loop_counts(arr, win=5, mode=1)      # P2LL window lower-left in both axes (:146-149)
    for i, lp in enumerate(loops):
        ra, rb, rab, lower = (int(v) for v in core[i])
        lp.P2LL = float(rab) / max(lower, pseudo)
        lp.ra, lp.rb, lp.rab = ra, rb, rab
        size = lp.x_end - lp.x_start + lp.y_end - lp.y_start
        if offp:
            lp.FDR, lp.ES = 1, 0
            lp.density = float(rab) / size / tot * 10.0**9
            lp.hypergeometric_p_value = lp.poisson_p_value = lp.binomial_p_value = 1
            lp.x_peak_poisson_p_value, lp.x_peak_es, lp.y_peak_poisson_p_value, lp.y_peak_es = 1, 0, 1, 0
            continue
        res = []
        for a, (left, right) in enumerate(((lp.x_start, lp.x_end), (lp.y_start, lp.y_end))):
            length = right - left
            count = int(anc[i, 2 * a])
            r = (int(anc[i, 2 * a + 1]) - count) / 5 / 2
            c = float(max([r, rpb * length, 1]))
            res.append((max([1e-300, poisson.sf(count - 1.0, c)]), count / c))
        (lp.x_peak_poisson_p_value, lp.x_peak_es), (lp.y_peak_poisson_p_value, lp.y_peak_es) = res
        if rab > 0:
            rabs, nbps = [], []
            for a in range(10):
                for b in range(10):
                    nrab = float(nab[i, a * 10 + b])
                    if nrab > 0:
                        rabs.append(nrab)
                        nbps.append(nrab / (float(na[i, a]) * float(nb[i, b])))
                    else:
                        rabs.append(0)
                        nbps.append(0.0)
            rabs, nbps = np.array(rabs), np.array(nbps)
            mrabs = float(np.median(rabs)) if np.median(rabs) > 0 else pseudo
            mbps = np.median(nbps) if np.median(nbps) > 0 else 1e-10
            lp.FDR = len(rabs[rabs > rab]) / float(len(rabs))
            lp.ES = rab / mrabs
            lp.density = float(rab) / size / tot * 10.0**9
            lp.hypergeometric_p_value = max([1e-300, hypergeom.sf(rab - 1.0, N, ra, rb)])
            lp.poisson_p_value = max([1e-300, poisson.sf(rab - 1.0, mrabs)])
            lp.binomial_p_value = max([1e-300, binom.sf(rab - 1.0, ra * rb - rab, mbps)])
        else:
            lp.FDR, lp.ES, lp.density = 1, 0, 0
            lp.hypergeometric_p_value = lp.poisson_p_value = lp.binomial_p_value = 1
    return key, loops


def quant_rows(loops):
    """Rows of quant.py:240-280 (_loops2txt), header excluded."""
    rows = []
    for lp in loops:
        line = [lp.id, lp.chromX, lp.x_start, lp.x_end, lp.chromY, lp.y_start, lp.y_end, lp.distance, lp.x_center,
                lp.y_center, lp.ra, lp.rb, lp.cis, lp.rab, lp.density, lp.ES, lp.P2LL, lp.FDR, lp.binomial_p_value,
                lp.hypergeometric_p_value, lp.poisson_p_value, lp.x_peak_poisson_p_value, lp.y_peak_poisson_p_value]
        rows.append(list(map(str, line)))
    return rows


# ---------------------------------------------------------------------------------------------------------------
# "next" rows of SURVEY.md 8f that reuse the path's kernels: callDiffLoops.quantDloops and callPeaks' clustering
def quant_dloops(loops, xy_a, xy_b, win=1, pseudo=1.0):
    """cLoops2/callDiffLoops.py:110-161 (quantDloops) on in-memory, already cut/mcut-filtered arrays of the target
    (a) and control (b) samples.  loops: int64[L,4].  Returns (ts, cs, rows): the concatenated nearby counts of both
    samples (:154-155) and per loop (raw_trt_ra, raw_trt_rb, raw_con_ra, raw_con_rb, raw_trt_rab, raw_con_rab,
    raw_trt_mrab, raw_con_mrab, trt_es, con_es)."""
    from . import XYIndex
    loops = np.asarray(loops, dtype=np.int64).reshape(-1, 4)
    ia, ib = XYIndex(xy_a), XYIndex(xy_b)
    ca, _, _, naba, _ = ia.loop_counts(loops, win=win, mode=0)
    cb, _, _, nabb, _ = ib.loop_counts(loops, win=win, mode=0)
    ts, cs, rows = [], [], []
    for i in range(len(loops)):
        arabs = [float(v) for v in naba[i]]                    # getLoopNearbyPETs order: a-windows outer, b-windows inner
        brabs = [float(v) for v in nabb[i]]
        tm, cm = float(np.mean(np.array(arabs))), float(np.mean(np.array(brabs)))       # :152-153
        ts.extend(arabs)
        cs.extend(brabs)
        rows.append((int(ca[i, 0]), int(ca[i, 1]), int(cb[i, 0]), int(cb[i, 1]), int(ca[i, 2]), int(cb[i, 2]), tm, cm,
                     int(ca[i, 2]) / max(tm, pseudo), int(cb[i, 2]) / max(cm, pseudo)))         # :156-157
    return ts, cs, rows


def split_mat(xy, splitExt):
    """callPeaks.py:51-60 (getSplitMat): every PET becomes two single-end intervals."""
    xy = np.asarray(xy, dtype=np.int64)
    out = np.empty((2 * len(xy), 2), dtype=np.int64)
    out[0::2, 0] = xy[:, 0] - splitExt
    out[0::2, 1] = xy[:, 0] + splitExt
    out[1::2, 0] = xy[:, 1] - splitExt
    out[1::2, 1] = xy[:, 1] + splitExt
    return out


def stich_peaks(peaks, margin=1):
    """cLoops2/geo.py:35-61 (stichPeaks) on [(start, end)] closed intervals, replayed on the covered positions."""
    cov = set()
    for s, e in peaks:
        cov.update(range(s, e + 1))
    cov = sorted(cov)
    out = []
    i = 0
    while i < len(cov) - 1:
        j = i + 1
        for j in range(i + 1, len(cov)):
            if cov[j] - cov[j - 1] > margin:
                break
        out.append((cov[i], cov[j - 1]))
        i = j
    return out


def dbscan_peaks(xy, eps, minPts, split=False, splitExt=50):
    """callPeaks.py:63-138 (runDBSCANPeaks) on an in-memory, already cut/mcut-filtered cis array: blockDBSCAN,
    clusters whose anchors are within eps of each other become candidate peaks (:113), then stichPeaks.
    Returns (candidates [(start, end)], stitched [(start, end)], reads in candidates)."""
    from . import block_dbscan, cluster_bbox
    xy = np.asarray(xy, dtype=np.int64)
    if split:
        xy = split_mat(xy, splitExt)
    shift = min(0, int(xy.min())) if len(xy) else 0            # blockDBSCAN is translation invariant (cells from minX/minY)
    labels, ncl = block_dbscan(xy - shift, eps, minPts)
    bbox, size = cluster_bbox(xy - shift, labels, ncl)
    cands, reads = [], 0
    for c in range(ncl):                                       # set(labels) of 0..k-1 iterates ascending
        if size[c] == 0:
            continue
        xs, xe, ys, ye = (int(v) + shift for v in bbox[c])
        if xs == xe or ys == ye:                               # :104-106
            continue
        if xe + eps >= ys and size[c] >= minPts:               # :113
            cands.append((xs, ye))
            reads += int(size[c])
    return cands, (stich_peaks(cands, margin=eps) if cands else []), reads


# ---------------------------------------------------------------------------------------------------------------
# `cLoops2 pre` (third "next" row): BEDPE -> per-chromosome-pair unique (cA, cB) rows + counters
def pre_bedpe(lines, mapq=1, cs=(), cut=0, mcut=-1, cis=False):
    """cLoops2/io.py:71-157 (parseBedpe) + ds.py:42-88 (PET) on an iterable of text lines, in plain Python.  Returns
    ({"chrA-chrB": sorted unique [(cA, cB)]}, meta) where meta holds the reference's petMeta.json counters.  The
    reference writes the unique rows in Python-set order (random per process), so rows are compared as sorted sets."""
    rows = {}
    total = hiq = t = c = closeCis = farCis = 0
    for line in lines:
        d = line.split("\n")[0].split("\t")
        try:                                                   # PET.__init__, ds.py:42-88; any failure skips the line
            chromA, startA, endA, strandA = d[0], int(d[1]), int(d[2]), d[8]
            chromB, startB, endB, strandB = d[3], int(d[4]), int(d[5]), d[9]
            try:
                q = int(d[7])
            except Exception:
                q = 255
            is_cis = chromA == chromB
            if is_cis:
                if startA + endA > startB + endB:
                    startA, startB, endA, endB = startB, startA, endB, endA
            elif chromA > chromB:
                chromA, chromB = chromB, chromA
                startA, startB, endA, endB = startB, startA, endB, endA
            cA, cB = int((startA + endA) / 2), int((startB + endB) / 2)
            distance = int(abs(cB - cA)) if is_cis else None
        except Exception:
            continue
        if len(cs) > 0 and not (chromA in cs and chromB in cs):      # io.py:102-103
            continue
        total += 1
        if q < mapq:
            continue
        hiq += 1
        if is_cis:
            c += 1
        else:
            t += 1
        if cis and not is_cis:
            continue
        if cut > 0 and is_cis and distance < cut:
            closeCis += 1
            continue
        if mcut > 0 and is_cis and distance > mcut:
            farCis += 1
            continue
        rows.setdefault("%s-%s" % (chromA, chromB), set()).add((cA, cB))
    uniques = sum(len(v) for v in rows.values())
    meta = {"Total PETs": total, "High Mapq(>=%s) PETs" % mapq: hiq, "Total Trans PETs": t, "Total Cis PETs": c,
            "Filtered too close (<%s) PETs" % cut: closeCis, "Filtered too distant (>%s) PETs" % mcut: farCis,
            "Unique PETs": uniques, "Cis PETs Redundancy": (1 - uniques / 1.0 / (c - closeCis)) if c > 0 else 0}
    return {k: sorted(v) for k, v in rows.items()}, meta


def pre_pairs(lines, cs=(), cut=0, mcut=-1, cis=False):
    """cLoops2/io.py:166-245 (parsePairs) + ds.py:111-137 (Pair) on an iterable of text lines, in plain Python; same
    return convention as pre_bedpe (no MAPQ counter: .pairs has no MAPQ column)."""
    rows = {}
    total = t = c = closeCis = farCis = 0
    for line in lines:
        if line.startswith("#"):
            continue
        d = line.split("\n")[0].split("\t")
        if len(d) == 8:
            if d[7] not in ["UU", "MR", "MU", "RU", "UR"]:
                continue
        elif len(d) != 7:
            continue
        try:                                                   # Pair.__init__, ds.py:111-137
            chromA, cA, chromB, cB = d[1], int(d[2]), d[3], int(d[4])
        except Exception:
            continue
        is_cis = chromA == chromB
        if is_cis:
            if cA > cB:
                cA, cB = cB, cA
            distance = int(abs(cB - cA))
        elif chromA > chromB:
            chromA, chromB, cA, cB = chromB, chromA, cB, cA
        if len(cs) > 0 and not (chromA in cs and chromB in cs):
            continue
        total += 1
        if is_cis:
            c += 1
        else:
            t += 1
        if cis and not is_cis:
            continue
        if cut > 0 and is_cis and distance < cut:
            closeCis += 1
            continue
        if mcut > 0 and is_cis and distance > mcut:
            farCis += 1
            continue
        rows.setdefault("%s-%s" % (chromA, chromB), set()).add((cA, cB))
    uniques = sum(len(v) for v in rows.values())
    meta = {"Total PETs": total, "Total Trans PETs": t, "Total Cis PETs": c,
            "Filtered too close (<%s) PETs" % cut: closeCis, "Filtered too distant (>%s) PETs" % mcut: farCis,
            "Unique PETs": uniques, "Cis PETs Redundancy": (1 - uniques / 1.0 / (c - closeCis)) if c > 0 else 0}
    return {k: sorted(v) for k, v in rows.items()}, meta
