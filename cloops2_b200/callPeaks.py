"""
The clustering step of `cLoops2 callPeaks` on B200 -- drop-in for cLoops2/callPeaks.py:51-138 (getSplitMat,
runDBSCANPeaks) and cLoops2/geo.py:35-61 (stichPeaks): blockDBSCAN over the (optionally split) PETs of one
chromosome, clusters whose two ends lie within eps of each other become candidate peaks, close candidates are
stitched.  The clustering and the cluster boxes come from the device (cl2_block_dbscan + cl2_cluster_bbox), unchanged
from the callLoops path (second "next" row).  Peak significance (callPeaks.py:141-) is a
different computation and not part of this package.
"""
import numpy as np

from . import _lib
from .ds import Peak
from .io import ixyKey, loadIxy

logger = None


def getSplitMat(mat, splitExt):
    """Paired-end PETs split into single-end intervals [end - ext, end + ext] (callPeaks.py:51-60)."""
    mat = np.asarray(mat, dtype=np.int64)
    out = np.empty((2 * mat.shape[0], 2), dtype=np.int64)
    out[0::2, 0] = mat[:, 0] - splitExt
    out[0::2, 1] = mat[:, 0] + splitExt
    out[1::2, 0] = mat[:, 1] - splitExt
    out[1::2, 1] = mat[:, 1] + splitExt
    return out


def _new_peak(chrom, start, end):
    p = Peak()
    p.chrom = chrom
    p.start = start
    p.end = end
    p.length = end - start + 1
    return p


def stichPeaks(peaks, margin=1):
    """geo.py:35-61 in interval arithmetic.  The reference walks the sorted set of covered positions and opens a new
    peak wherever two consecutive covered positions are more than `margin` apart; its last peak ends at the
    second-to-last covered position and a final single position yields no peak -- both kept here."""
    if len(peaks) == 0:
        return []
    chrom = peaks[0].chrom
    ivs = sorted((int(p.start), int(p.end)) for p in peaks if p.end >= p.start)
    if not ivs:
        return []
    if margin < 1:                                 # every covered position is its own run: replay the reference
        cov = sorted(set(v for s, e in ivs for v in range(s, e + 1)))
        out, i = [], 0
        while i < len(cov) - 1:
            j = i + 1
            for j in range(i + 1, len(cov)):
                if cov[j] - cov[j - 1] > margin:
                    break
            out.append(_new_peak(chrom, cov[i], cov[j - 1]))
            i = j
        return out
    runs, members = [], []
    cs, ce = ivs[0]
    cur = [ivs[0]]
    for s, e in ivs[1:]:
        if s - ce > margin:
            runs.append((cs, ce))
            members.append(cur)
            cs, ce, cur = s, e, [(s, e)]
        else:
            ce = max(ce, e)
            cur.append((s, e))
    runs.append((cs, ce))
    members.append(cur)
    out = [_new_peak(chrom, s, e) for s, e in runs[:-1]]
    s, e = runs[-1]
    below = [min(b, e - 1) for a, b in members[-1] if a <= e - 1]      # covered positions of the last run below its end
    if below:
        out.append(_new_peak(chrom, s, max(below)))
    return out


def runDBSCANPeaks(fixy, eps, minPts, cut=0, mcut=-1, split=False, splitExt=50, ctx=None):
    """Drop-in for callPeaks.py:63-138: candidate peaks of one cis `.ixy` file, stitched; None for a trans file or an
    empty one."""
    key = ixyKey(fixy)
    if key[0] != key[1]:
        return None
    ctx = ctx or _lib.default_context()
    xy = loadIxy(fixy)
    shift = 0
    if split:
        # the split intervals are made on the host from the cut/mcut-filtered PETs; blockDBSCAN works on coordinates
        # relative to the minimum, so lifting everything by a constant changes no cell and no distance
        d = xy[:, 1] - xy[:, 0]
        keep = np.ones(len(xy), dtype=bool)
        if cut > 0:
            keep &= d >= cut
        if mcut > 0:
            keep &= d <= mcut
        xy = getSplitMat(xy[keep], splitExt)
        shift = min(0, int(xy.min())) if len(xy) else 0
        if shift:
            xy = xy - shift
        pts = _lib.Points(ctx, xy)
    else:
        pts = _lib.Points(ctx, xy, cut=cut, mcut=mcut)
    try:
        if pts.n == 0:
            if logger is not None:
                logger.info("No PETs found in %s." % (fixy))
            return None
        if logger is not None:
            logger.info("Clustering %s to find candidate peaks using eps as %s, minPts as %s." % (key[0], eps, minPts))
        pts.block_dbscan(eps, minPts, want_labels=False)
        bbox, size = pts.cluster_bbox()
    finally:
        pts.close()
    bbox = bbox + shift
    ok = (size > 0) & (bbox[:, 0] != bbox[:, 1]) & (bbox[:, 2] != bbox[:, 3])          # :104-106
    ok &= (bbox[:, 1] + eps >= bbox[:, 2]) & (size >= minPts)                           # :113
    peaks = [_new_peak(key[0], int(b[0]), int(b[3])) for b in bbox[ok]]
    readS = int(size[ok].sum())
    if logger is not None:
        logger.info("Clustering %s finished. Estimated %s PETs in candidate %s peaks." % (key[0], readS, len(peaks)))
        logger.info("Stiching %s candidate close %s peaks." % (key[0], len(peaks)))
    peaks = stichPeaks(peaks, margin=eps)
    if logger is not None:
        logger.info("Merged %s %s peaks" % (key[0], len(peaks)))
    return peaks
