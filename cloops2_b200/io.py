"""
On-disk formats of the callLoops path: `.ixy` + petMeta.json in, `_loops.txt` and the optional track files out.
Byte-compatible with cLoops2/io.py (parseIxy :274-291, updateJson :390-405, loops2txt :517-561, loops2ucscTxt
:564-608, loops2juiceTxt :611-660, loops2washuTxt :663-684, loops2NewWashuTxt :687-707): every field goes through
Python str() exactly as the reference does, so the text files are identical when the numbers are.
"""
import json
import os

import numpy as np


def ixyKey(f):
    """("chrA","chrB") from the file name (io.py:280-281)."""
    return tuple(os.path.split(f)[1].replace(".ixy", "").split("-"))


def ixyMap(f):
    """The [N,2] matrix of a `.ixy` (joblib pickle, io.py:55-67) as a read-only memory map where joblib can map it."""
    import joblib
    try:
        mat = joblib.load(f, mmap_mode="r")
    except Exception:
        mat = joblib.load(f)
    mat = np.asarray(mat) if not isinstance(mat, np.ndarray) else mat
    if mat.ndim != 2 or mat.shape[1] != 2:
        raise ValueError("%s does not hold an [N,2] matrix" % f)
    return mat


def loadIxy(f, out=None):
    """Raw int64[N,2] of a `.ixy`.  With `out` (e.g. a pinned buffer) the rows are copied into it from the memory
    map, avoiding an intermediate pageable copy."""
    mat = ixyMap(f)
    if out is not None:
        np.copyto(out, mat, casting="unsafe")
        return out
    return np.ascontiguousarray(mat, dtype=np.int64)


def parseIxy(f, cut=0, mcut=-1):
    """Host-side parseIxy (io.py:274-291): key and the cut/mcut filtered matrix.  The GPU path applies the same
    filter on the device at upload (cl2_points_upload); this form exists for API compatibility."""
    key = ixyKey(f)
    mat = loadIxy(f)
    if cut > 0:
        mat = mat[(mat[:, 1] - mat[:, 0]) >= cut]
    if mcut > 0:
        mat = mat[(mat[:, 1] - mat[:, 0]) <= mcut]
    return key, mat


def updateJson(ixyfs, metaf):
    """Rewrite the `data` section of petMeta.json for the given .ixy files (io.py:390-405)."""
    meta = json.loads(open(metaf).read())
    meta["data"] = {"cis": {}, "trans": {}}
    for f in ixyfs:
        key = os.path.split(f)[1].replace(".ixy", "")
        a = key.split("-")
        meta["data"]["cis" if a[0] == a[1] else "trans"][key] = {"ixy": os.path.abspath(f)}
    with open(metaf, "w") as fo:
        json.dump(meta, fo)


LOOP_HEADER = ["loopId", "chrA", "startA", "endA", "chrB", "startB", "endB", "distance(bp)", "centerA", "centerB",
               "readsA", "readsB", "cis", "PETs", "density", "enrichmentScore", "P2LL", "FDR", "binomialPvalue",
               "hypergeometricPvalue", "poissonPvalue", "poissonPvaluePeakA", "poissonPvaluePeakB", "significant"]


def _lid(loop, i):
    return loop.id if hasattr(loop, "id") else "loop_%s-%s-%s" % (loop.chromX, loop.chromY, i)


def loops2txt(loops, fout):
    """24-column loop table (io.py:517-561)."""
    with open(fout, "w") as fo:
        fo.write("\t".join(LOOP_HEADER) + "\n")
        for i, lp in enumerate(loops):
            line = [_lid(lp, i), lp.chromX, lp.x_start, lp.x_end, lp.chromY, lp.y_start, lp.y_end, lp.distance,
                    lp.x_center, lp.y_center, lp.ra, lp.rb, lp.cis, lp.rab, lp.density, lp.ES, lp.P2LL, lp.FDR,
                    lp.binomial_p_value, lp.hypergeometric_p_value, lp.poisson_p_value, lp.x_peak_poisson_p_value,
                    lp.y_peak_poisson_p_value, lp.significant]
            fo.write("\t".join(map(str, line)) + "\n")


def cols2txt(items, fout, cis=True):
    """The same file as loops2txt(cols_to_loops(...)) straight from the column arrays of the per-file results, without
    a Loop object and 24 str() calls per row: items = [((chrA, chrB), cols), ...] in output order.  The rows are
    formatted by cl2_loops_text (native; floats laid out as Python's repr() does)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    with open(fout, "wb") as fo:
        fo.write(("\t".join(LOOP_HEADER) + "\n").encode())
        base = 0
        for key, cols in items:
            n = int(cols["coords"].shape[0])
            if n == 0:
                continue
            a = [np.ascontiguousarray(cols[k], dtype=np.int64) for k in ("coords", "ra", "rb", "rab")]
            f = [np.ascontiguousarray(cols[k], dtype=np.float64) for k in ("density", "ES", "P2LL", "FDR", "nbp", "hyp", "pop",
                                                                          "px", "py")]
            cap = n * (640 + 2 * (len(key[0]) + len(key[1])))
            buf = ctypes.create_string_buffer(cap)
            m = lib.cl2_loops_text(key[0].encode(), key[1].encode(), base, n, 1 if cis else 0,
                                   *[x.ctypes.data_as(_lib._i64p) for x in a], *[x.ctypes.data_as(_lib._f64p) for x in f],
                                   buf, cap)
            if m < 0:
                raise _lib.Cl2Error("cl2_loops_text failed rc=%d" % m)
            fo.write(buf.raw[:m])
            base += n


def loops2ucscTxt(loops, fout, significant=1):
    """UCSC interact track (io.py:564-608)."""
    print("Converting %s loops to UCSC interact track." % (len(loops)))
    cols = ["#chrom", "chromStart", "chromEnd", "name", "score", "value", "exp", "color", "sourceChrom",
            "sourceStart", "sourceEnd", "sourceName", "sourceStrand", "targetChrom", "targetStart", "targetENd",
            "targetName", "targetStrand"]
    with open(fout, "w") as f:
        f.write('track type=interact name="loops" description="loops" interactDirectional=false visibility=full\n')
        f.write("\t".join(cols) + "\n")
        for i, lp in enumerate(loops):
            if significant > 0 and lp.significant < 1:
                continue
            line = [lp.chromX, lp.x_start, lp.y_end, "loop_%s-%s-%s" % (lp.chromX, lp.chromY, i), int(lp.rab),
                    lp.density, ".", "#0000FF", lp.chromX, lp.x_start, lp.x_end, ".", ".", lp.chromY, lp.y_start,
                    lp.y_end, ".", "."]
            f.write("\t".join(map(str, line)) + "\n")
    print("Converting to UCSC interact track finished for %s." % fout)


def loops2juiceTxt(loops, fout, significant=1):
    """Juicebox 2D annotation (io.py:611-660); p-values as -log10."""
    print("Converting %s loops to Juicebox 2D annotation feature." % (len(loops)))
    cols = ["chromosome1", "x1", "x2", "chromosome2", "y1", "y2", "color", "observed", "loopId", "FDR",
            "EnrichmentScore", "P2LL", "distance", "counts_X", "counts_Y", "density", "-log10(binomal_p-value)",
            "-log10(poisson_p-value)", "-log10(hypergeometric_p-value)", "-log10(poisson_p-value_peak_x)",
            "-log10(poisson_p-value_peak_y)"]
    with open(fout, "w") as f:
        f.write("\t".join(cols) + "\n")
        for i, lp in enumerate(loops):
            if significant > 0 and lp.significant < 1:
                continue
            line = [lp.chromX, lp.x_start, lp.x_end, lp.chromY, lp.y_start, lp.y_end, '"0,255,255"', lp.rab,
                    "%s-%s-%s" % (lp.chromX, lp.chromY, i), lp.FDR, lp.ES, lp.P2LL, lp.distance, lp.ra, lp.rb,
                    lp.density, -np.log10(lp.binomial_p_value), -np.log10(lp.poisson_p_value),
                    -np.log10(lp.hypergeometric_p_value), -np.log10(lp.x_peak_poisson_p_value),
                    -np.log10(lp.y_peak_poisson_p_value)]
            f.write("\t".join(map(str, line)) + "\n")
    print("Converting to Juicebox 2D annotation feature finished for %s." % fout)


def loops2washuTxt(loops, fout, significant=1):
    """Legacy washU long-range track (io.py:663-684)."""
    print("Converting %s loops to washU long range interaction track." % len(loops))
    with open(fout, "w") as f:
        for lp in loops:
            if significant > 0 and lp.significant < 1:
                continue
            f.write("\t".join(["%s:%s-%s" % (lp.chromX, lp.x_start, lp.x_end),
                               "%s:%s-%s" % (lp.chromY, lp.y_start, lp.y_end), "1"]) + "\n")
    print("Converted to washU long range interaction track %s finished." % fout)


def loops2NewWashuTxt(loops, fout, significant=1):
    """New washU long-range track, two rows per loop (io.py:687-707)."""
    print("Converting %s loops to new washU long range interaction track." % len(loops))
    with open(fout, "w") as f:
        for lp in loops:
            if significant > 0 and lp.significant < 1:
                continue
            a = [lp.chromX, lp.x_start, lp.x_end, "%s:%s-%s,%s" % (lp.chromY, lp.y_start, lp.y_end, lp.rab)]
            b = [lp.chromY, lp.y_start, lp.y_end, "%s:%s-%s,%s" % (lp.chromX, lp.x_start, lp.x_end, lp.rab)]
            f.write("\t".join(map(str, a)) + "\n")
            f.write("\t".join(map(str, b)) + "\n")
    print("Converted to new washU long range interaction track %s finished." % fout)
