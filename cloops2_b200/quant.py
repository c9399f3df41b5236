"""
`cLoops2 quant -loops` on B200 -- drop-in for the loop part of cLoops2/quant.py (_quantLoops :129-237, quantLoops
:283-316, _loops2txt :240-280): re-quantify a given set of loops against a sample.  It is the counting kernel of the
callLoops path with different fall-backs and no filters (first "next" row): P2LL window
lower-left in both axes (:146-149), median background replaced by `pseudo` / 1e-10 when it is zero (:196-203),
every loop is kept, loops without PETs get the constant row of :221-228.
"""
import json

import numpy as np
from scipy.stats import hypergeom, binom, poisson

from . import _lib, hostops, shard
from .ds import Loop
from .io import loadIxy


def parseTxt2Loops(f, cut=0):
    """`_loops.txt` -> {"chrA-chrB": [Loop]} (cLoops2/io.py:710-740)."""
    loops = {}
    for i, line in enumerate(open(f)):
        if i == 0:
            continue
        line = line.split("\n")[0].split("\t")
        if len(line) < 8:
            continue
        lp = Loop()
        lp.id = line[0]
        lp.chromX, lp.chromY = line[1], line[4]
        lp.x_start, lp.x_end = int(float(line[2])), int(float(line[3]))
        lp.y_start, lp.y_end = int(float(line[5])), int(float(line[6]))
        lp.distance = int(float(line[7]))
        if lp.distance < cut:
            continue
        lp.x_center = (lp.x_start + lp.x_end) / 2
        lp.y_center = (lp.y_start + lp.y_end) / 2
        lp.cis = lp.chromX == lp.chromY
        loops.setdefault(lp.chromX + "-" + lp.chromY, []).append(lp)
    return loops


def _quantLoops(key, loops, fixy, tot, pcut=0, mcut=-1, pseudo=1, offp=False, ctx=None):
    """Drop-in for quant.py:129-237.  Counts come from the device (cl2_loop_counts); the per-loop arithmetic below is
    the reference's, expression for expression."""
    tot = float(tot)
    ctx = ctx or _lib.default_context()
    pts = _lib.Points(ctx, loadIxy(fixy), cut=pcut, mcut=mcut)
    try:
        N = pts.n
        ext = pts.extent()
        cands = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops], dtype=np.int64).reshape(-1, 4)
        index = _lib.Index(pts)
        try:
            core, na, nb, nab, anc = index.loop_counts(cands, win=5, mode=_lib.MODE_TRANS, windows=not offp,
                                                       anchors=not offp)
        finally:
            index.close()
    finally:
        pts.close()
    if not offp:
        rpb = float(N) / (int(ext[3]) - int(ext[0]))
        la, lb = cands[:, 1] - cands[:, 0], cands[:, 3] - cands[:, 2]
        px, esx = hostops.anchor_sig(anc[:, 0], anc[:, 1], la, rpb)
        py, esy = hostops.anchor_sig(anc[:, 2], anc[:, 3], lb, rpb)
        rabs = nab.astype(np.float64)
        den = (na.astype(np.float64)[:, :, None] * nb.astype(np.float64)[:, None, :]).reshape(rabs.shape)
        with np.errstate(divide="ignore", invalid="ignore"):
            nbps = np.where(rabs > 0, rabs / den, 0.0)
        med_r, med_b = np.median(rabs, axis=1), np.median(nbps, axis=1)
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
        lp.x_peak_poisson_p_value, lp.x_peak_es = px[i], esx[i]
        lp.y_peak_poisson_p_value, lp.y_peak_es = py[i], esy[i]
        if rab > 0:
            hyp = max([1e-300, hypergeom.sf(rab - 1.0, N, ra, rb)])
            mrabs = float(med_r[i]) if med_r[i] > 0 else pseudo                  # :196-199
            mbps = med_b[i] if med_b[i] > 0 else 1e-10                           # :200-203
            fdr = int((rabs[i] > rab).sum()) / float(rabs.shape[1])
            lp.FDR = fdr
            lp.ES = rab / mrabs
            lp.density = float(rab) / size / tot * 10.0**9
            lp.hypergeometric_p_value = hyp
            lp.poisson_p_value = max([1e-300, poisson.sf(rab - 1.0, mrabs)])
            lp.binomial_p_value = max([1e-300, binom.sf(rab - 1.0, ra * rb - rab, mbps)])
        else:
            lp.FDR, lp.ES, lp.density = 1, 0, 0
            lp.hypergeometric_p_value = lp.poisson_p_value = lp.binomial_p_value = 1
    return key, loops


def _loops2txt(loops, fout):
    """23-column table of quant.py:240-280 (no `significant` column, ids kept)."""
    header = ["loopId", "chrA", "startA", "endA", "chrB", "startB", "endB", "distance(bp)", "centerA", "centerB",
              "readsA", "readsB", "cis", "PETs", "density", "enrichmentScore", "P2LL", "FDR", "binomialPvalue",
              "hypergeometricPvalue", "poissonPvalue", "poissonPvaluePeakA", "poissonPvaluePeakB"]
    with open(fout, "w") as fo:
        fo.write("\t".join(header) + "\n")
        for lp in loops:
            line = [lp.id, lp.chromX, lp.x_start, lp.x_end, lp.chromY, lp.y_start, lp.y_end, lp.distance, lp.x_center,
                    lp.y_center, lp.ra, lp.rb, lp.cis, lp.rab, lp.density, lp.ES, lp.P2LL, lp.FDR, lp.binomial_p_value,
                    lp.hypergeometric_p_value, lp.poisson_p_value, lp.x_peak_poisson_p_value,
                    lp.y_peak_poisson_p_value]
            fo.write("\t".join(map(str, line)) + "\n")


def quantLoops(predir, loopf, output, logger, cut=0, mcut=-1, cpu=1, offp=False):
    """Drop-in for quant.py:283-316.  The reference orders chromosomes by iterating a Python set of keys (hash order);
    here they follow petMeta.json, rows inside a chromosome keep the input order as in the reference."""
    loops = parseTxt2Loops(loopf, cut=0)
    meta = json.loads(open(predir + "/petMeta.json").read())
    tot = meta["Unique PETs"]
    keys = [k for k in meta["data"]["cis"].keys() if k in loops]
    ctxs = shard.worker_contexts()
    res = shard.map_units(ctxs, keys, lambda c, k: _quantLoops(k, loops[k], meta["data"]["cis"][k]["ixy"], tot, pcut=cut,
                                                               mcut=mcut, offp=offp, ctx=c))
    out = []
    for _, ls in res:
        out.extend(ls)
    _loops2txt(out, output + "_loops.txt")
    return out
