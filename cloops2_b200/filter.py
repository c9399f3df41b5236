"""
`cLoops2 filterPETs -knn` on B200 -- drop-in for cLoops2/filter.py:223-293 (_filterPETsByKNNs, filterPETsByKNNs): keep
the PETs that survive blockDBSCAN's noise-cell removal (the first three kernels of the clustering path; SURVEY.md
section 8f, second "next" row).  The surviving rows are written cell by cell in the reference's dict order of the
cells and in row order inside a cell (filter.py:273-278), so the output .ixy is identical row for row.
"""
import json
from glob import glob

import numpy as np

from . import _lib, shard
from .io import ixyKey, loadIxy, updateJson


def _filterPETsByKNNs(f, outdir, eps, minPts, ctx=None):
    key = ixyKey(f)
    mat = loadIxy(f)
    if mat.shape[0] == 0:
        return
    ctx = ctx or _lib.default_context()
    pts = _lib.Points(ctx, mat)
    try:
        keep, first = pts.block_noise_keep(eps, minPts, with_cell_order=True)
    finally:
        pts.close()
    rows = np.nonzero(keep)[0]
    if rows.size == 0:
        return
    rows = rows[np.argsort(first[rows], kind="stable")]       # cells in dict order, rows ascending inside a cell
    import joblib
    joblib.dump(mat[rows, ], outdir + "/" + "-".join(key) + ".ixy")


def writeNewJson(fdir):
    """petMeta.json for a directory of .ixy files (cLoops2/io.py:408-420)."""
    ixyfs = glob(fdir + "/*.ixy")
    tot = sum(loadIxy(f).shape[0] for f in ixyfs)
    with open(fdir + "/petMeta.json", "w") as fo:
        json.dump({"Unique PETs": tot}, fo)
    updateJson(ixyfs, fdir + "/petMeta.json")


def filterPETsByKNNs(predir, outdir, eps=1000, minPts=5, cpu=1):
    """Drop-in for filter.py:281-293."""
    meta = json.loads(open(predir + "/petMeta.json").read())
    keys = list(meta["data"]["cis"].keys())
    ctxs = shard.worker_contexts()
    shard.map_units(ctxs, keys, lambda c, k: _filterPETsByKNNs(meta["data"]["cis"][k]["ixy"], outdir, eps, minPts, ctx=c))
    writeNewJson(outdir)
