"""
Array-level host logic of the callLoops path.  The reference walks Python lists of Loop objects; at 10^5-10^6
candidates per run that is the bottleneck once clustering and counting are on the GPU, so everything between the
kernels is numpy on struct-of-arrays here, with the reference lines each step reproduces cited inline.
Float expressions are evaluated in the same order as the reference so results are bit-identical.
"""
import ctypes

import numpy as np
from scipy.stats import hypergeom, binom, poisson

from . import _lib


# ---------------------------------------------------------------------------------------------- candidates
def cis_candidate_mask(bbox):
    """callCisLoops.py:97-99 (degenerate box), :114 (anchor sizes sum < 200), :116 (x_end < y_start)."""
    x0, x1, y0, y1 = bbox[:, 0], bbox[:, 1], bbox[:, 2], bbox[:, 3]
    ok = (x0 != x1) & (y0 != y1)
    ok &= (x1 - x0 + y1 - y0) >= 200
    ok &= x1 < y0
    return ok


def trans_candidate_mask(bbox):
    """callTransLoops.py:62-64 (degenerate box only)."""
    return (bbox[:, 0] != bbox[:, 1]) & (bbox[:, 2] != bbox[:, 3])


def combine_unique(cands_per_pass):
    """combineLoops over the sweep (geo.py:90-112) followed by filterDupLoops (callCisLoops.py:159-169):
    coordinates in order of first appearance."""
    if not cands_per_pass:
        return np.zeros((0, 4), dtype=np.int64)
    allc = np.concatenate(cands_per_pass, axis=0) if len(cands_per_pass) > 1 else cands_per_pass[0]
    if allc.shape[0] == 0:
        return allc.reshape(0, 4)
    _, first = np.unique(allc, axis=0, return_index=True)
    return allc[np.sort(first)]


def cis_distance(c):
    """distance = abs(y_center - x_center) with centers (start+end)/2 as floats (callCisLoops.py:107-112)."""
    xc = (c[:, 0] + c[:, 1]) / 2
    yc = (c[:, 2] + c[:, 3]) / 2
    return np.abs(yc - xc)


# ---------------------------------------------------------------------------------------------- statistics
def _floor300(p):
    return np.maximum(1e-300, p)


def anchor_sig(count, wide, length, rpb):
    """estAnchorSig (callCisLoops.py:226-247) from the two counts: returns (p, es)."""
    r = (wide - count) / 5 / 2
    c = np.maximum(np.maximum(r, rpb * length), 1.0)
    p = _floor300(poisson.sf(count - 1.0, c))
    return p, count / c


def permuted_stats(na, nb, nab, rab, cis=True):
    """getLoopNearbyPETs (callCisLoops.py:201-223) / the trans variant (callTransLoops.py:155-172):
    returns (mrabs, mbps, fdr)."""
    L, W = na.shape
    rabs = nab.astype(np.float64)                                    # [L, W*W], a-major
    den = (na.astype(np.float64)[:, :, None] * nb.astype(np.float64)[:, None, :]).reshape(L, W * W)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = rabs / den
    if cis:
        nbps = np.where(rabs > 0, q, 0.0)                            # :213-221
        mrabs = np.median(rabs, axis=1)
        mbps = np.median(nbps, axis=1)
    else:
        nbps = np.where(den > 0, q, 0.0)                             # callTransLoops.py:165-169 (nac>0 and nbc>0)
        mrabs = np.mean(rabs, axis=1)
        mbps = np.mean(nbps, axis=1)
    fdr = (rabs > rab[:, None]).sum(axis=1) / float(W * W)
    return mrabs, mbps, fdr


def est_cis(index, cands, N, span, tot, minPts, hic=False, pseudo=1, peakPcut=1e-5, win=5, countDiffCut=20):
    """estLoopSig (callCisLoops.py:250-354) for one chromosome.  `index` is a device XY index; cands int64[L,4].
    Returns a dict of column arrays for the loops that survive the filter cascade, in candidate order."""
    L = cands.shape[0]
    cols = _empty_cols()
    if L == 0:
        return cols
    la = (cands[:, 1] - cands[:, 0])
    lb = (cands[:, 3] - cands[:, 2])
    keep = ~((la / lb > countDiffCut) | (lb / la > countDiffCut))                    # :282-286
    idx = np.nonzero(keep)[0]
    if idx.size == 0:
        return cols
    core, _, _, _, _ = index.loop_counts(cands[idx], win=win, mode=_lib.MODE_CIS, windows=False, anchors=False)
    ra, rb, rab, lower = core[:, 0], core[:, 1], core[:, 2], core[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        k = rab >= minPts                                                             # :290
        k &= ~((ra == rab) | (rb == rab))                                             # :293
        k &= ~((ra / rb.astype(np.float64) > countDiffCut) | (rb / ra.astype(np.float64) > countDiffCut))   # :296
    p2ll = rab.astype(np.float64) / np.maximum(lower, pseudo)                         # :306
    if hic:
        k &= ~(p2ll < 1)                                                              # :307
    idx, ra, rb, rab, p2ll = idx[k], ra[k], rb[k], rab[k], p2ll[k]
    if idx.size == 0:
        return cols
    hyp = _floor300(hypergeom.sf(rab - 1.0, N, ra, rb))                               # :310
    k = ~(hyp > 1e-2)
    idx, ra, rb, rab, p2ll, hyp = idx[k], ra[k], rb[k], rab[k], p2ll[k], hyp[k]
    if idx.size == 0:
        return cols
    _, na, nb, nab, anc = index.loop_counts(cands[idx], win=win, mode=_lib.MODE_CIS, core=False)
    mrabs, mbps, fdr = permuted_stats(na, nb, nab, rab, cis=True)
    k = ~((mrabs >= rab) | (fdr >= 0.1))                                              # :322
    es = rab / np.maximum(mrabs, pseudo)                                              # :325
    k &= ~(es < 1)
    idx, ra, rb, rab, p2ll, hyp = idx[k], ra[k], rb[k], rab[k], p2ll[k], hyp[k]
    mrabs, mbps, fdr, es, anc = mrabs[k], mbps[k], fdr[k], es[k], anc[k]
    if idx.size == 0:
        return cols
    pop = _floor300(poisson.sf(rab - 1.0, mrabs))                                     # :329
    nbp = _floor300(binom.sf(rab - 1.0, ra * rb - rab, mbps))                         # :331-333
    c = cands[idx]
    la, lb = c[:, 1] - c[:, 0], c[:, 3] - c[:, 2]
    density = rab.astype(np.float64) / (la + lb) / tot * 10.0**9                      # :337-339
    rpb = float(N) / span                                                             # :231
    px, esx = anchor_sig(anc[:, 0], anc[:, 1], la, rpb)
    py, esy = anchor_sig(anc[:, 2], anc[:, 3], lb, rpb)
    if not hic:
        k = (px < peakPcut) & (py < peakPcut)                                         # :347
    else:
        k = np.ones(idx.size, dtype=bool)
    cols = dict(coords=c[k], ra=ra[k], rb=rb[k], rab=rab[k], P2LL=p2ll[k], FDR=fdr[k], ES=es[k], density=density[k],
                hyp=hyp[k], pop=pop[k], nbp=nbp[k], px=px[k], py=py[k], esx=esx[k], esy=esy[k])
    return cols


def est_trans(index, cands, N, span, minPts, pseudo=1, countDiffCut=10, win=5):
    """trans estLoopSig (callTransLoops.py:106-193)."""
    L = cands.shape[0]
    cols = _empty_cols()
    if L == 0:
        return cols
    core, na, nb, nab, anc = index.loop_counts(cands, win=win, mode=_lib.MODE_TRANS)
    ra, rb, rab, lower = core[:, 0], core[:, 1], core[:, 2], core[:, 3]
    la, lb = cands[:, 1] - cands[:, 0], cands[:, 3] - cands[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        k = rab >= minPts                                                             # :127
        k &= ~((ra / rb.astype(np.float64) > countDiffCut) | (rb / ra.astype(np.float64) > countDiffCut))   # :133
        k &= ~((la / lb > countDiffCut) | (lb / la > countDiffCut))                   # :135-139
    c, ra, rb, rab, lower, la, lb = cands[k], ra[k], rb[k], rab[k], lower[k], la[k], lb[k]
    na, nb, nab, anc = na[k], nb[k], nab[k], anc[k]
    if c.shape[0] == 0:
        return cols
    p2ll = rab.astype(np.float64) / np.maximum(lower, pseudo)
    rpb = float(N) / span
    px, esx = anchor_sig(anc[:, 0], anc[:, 1], la, rpb)
    py, esy = anchor_sig(anc[:, 2], anc[:, 3], lb, rpb)
    hyp = _floor300(hypergeom.sf(rab - 1.0, N, ra, rb))
    mrabs, mbps, fdr = permuted_stats(na, nb, nab, rab, cis=False)
    es = rab / np.maximum(mrabs, pseudo)
    pop = _floor300(poisson.sf(rab - 1.0, mrabs))
    nbp = _floor300(binom.sf(rab - 1.0, ra * rb, mbps))                               # :181
    density = rab.astype(np.float64) / (la + lb) / N * 10.0**9                        # :185-187
    return dict(coords=c, ra=ra, rb=rb, rab=rab, P2LL=p2ll, FDR=fdr, ES=es, density=density, hyp=hyp, pop=pop,
                nbp=nbp, px=px, py=py, esx=esx, esy=esy)


COLS = ["coords", "ra", "rb", "rab", "P2LL", "FDR", "ES", "density", "hyp", "pop", "nbp", "px", "py", "esx", "esy"]


def _empty_cols():
    d = {k: np.zeros(0) for k in COLS}
    d["coords"] = np.zeros((0, 4), dtype=np.int64)
    for k in ("ra", "rb", "rab"):
        d[k] = np.zeros(0, dtype=np.int64)
    return d


def take(cols, sel):
    return {k: v[sel] for k, v in cols.items()}


def mark_sig_cis(cols, hic=False):
    """markSigLoops (callCisLoops.py:357-376) -> int array significant."""
    sig = (cols["FDR"] <= 0.05) & (cols["ES"] >= 2) & (cols["hyp"] <= 1e-5) & (cols["pop"] <= 1e-5)
    sig &= cols["nbp"] < (1e-3 if hic else 1e-1)
    return sig.astype(np.int64)


def mark_sig_trans(cols):
    """trans markSigLoops (callTransLoops.py:196-206)."""
    return ((cols["nbp"] <= 1e-10) & (cols["FDR"] <= 0.05) & (cols["ES"] >= 2)).astype(np.int64)


def sel_sig(coords, es):
    """selSigLoops (callCisLoops.py:379-422) on the significant loops of one chromosome (native, libcl2.so)."""
    L = coords.shape[0]
    if L == 0:
        return np.zeros(0, dtype=np.int64)
    lib = _lib.load()
    coords = np.ascontiguousarray(coords, dtype=np.int64)
    es = np.ascontiguousarray(es, dtype=np.float64)
    out = np.empty(L, dtype=np.int64)
    n = ctypes.c_int64(0)
    rc = lib.cl2_sel_sig_loops(coords.ctypes.data_as(_lib._i64p), es.ctypes.data_as(_lib._f64p), L,
                               out.ctypes.data_as(_lib._i64p), ctypes.byref(n))
    if rc != 0:
        raise _lib.Cl2Error("cl2_sel_sig_loops failed rc=%d" % rc)
    return out[:n.value]


def select_twice(cols, significant):
    """The reference applies selSigLoops twice (callCisLoops.py:586-595)."""
    s = np.nonzero(significant > 0)[0]
    cur = take(cols, s)
    for _ in range(2):
        pick = sel_sig(cur["coords"], cur["ES"])
        cur = take(cur, pick)
    return cur
