"""
`cLoops2 pre` on B200 -- drop-in for cLoops2/io.py:71-164 (parseBedpe): BEDPE(.gz) text -> one `.ixy` file of unique
(cA, cB) PET centres per chromosome pair + petMeta.json (third "next" row).

Where the reference loops over the lines in Python, keeps one text file per chromosome pair and de-duplicates each
through a Python set (getUniqueTxt, io.py:40-52), this module
  * tokenises the text with the library's native parser (cl2_bedpe_parse: the field handling of PET.__init__,
    ds.py:42-57, same accept / skip decisions),
  * applies the PET arithmetic (end swap, centres, distance, ds.py:58-88) and parseBedpe's filters and counters
    (io.py:101-125) to whole chunks with numpy,
  * de-duplicates every chromosome pair on the GPU (cl2_points_unique: radix sort by the packed (cA, cB) key, first
    row of every run).
The `.ixy` rows come out ascending by (cA, cB); the reference's order is that of a Python set of strings (different in
every process), so files agree as sets of rows.  Counters and petMeta.json keys are the reference's.
"""
import ctypes
import gzip
import json
import os

import numpy as np

from . import _lib
from .io import updateJson

CHUNK = 16 << 20


class BedpeParser:
    """cl2_bedpe: text chunks -> (chrom ids int32[n,2], positions int64[n,4], mapq int32[n])."""

    def __init__(self):
        self.lib = _lib.load()
        self.h = ctypes.c_void_p()
        if self.lib.cl2_bedpe_create(ctypes.byref(self.h)) != 0:
            raise _lib.Cl2Error("cl2_bedpe_create failed")

    def close(self):
        if self.h:
            self.lib.cl2_bedpe_destroy(self.h)
            self.h = None

    __del__ = close

    def parse(self, buf):
        cap = buf.count(b"\n") + 1
        chrom = np.empty((cap, 2), dtype=np.int32)
        pos = np.empty((cap, 4), dtype=np.int64)
        mapq = np.empty(cap, dtype=np.int32)
        n = self.lib.cl2_bedpe_parse(self.h, buf, len(buf), cap, chrom.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                     pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                     mapq.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None)
        if n < 0:
            raise _lib.Cl2Error("cl2_bedpe_parse failed rc=%d" % n)
        return chrom[:n], pos[:n], mapq[:n]

    def parse_pairs(self, buf):
        """.pairs text -> (chrom ids int32[n,2], (cA, cB) int64[n,2])."""
        cap = buf.count(b"\n") + 1
        chrom = np.empty((cap, 2), dtype=np.int32)
        pos = np.empty((cap, 2), dtype=np.int64)
        n = self.lib.cl2_pairs_parse(self.h, buf, len(buf), cap, chrom.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                     pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), None)
        if n < 0:
            raise _lib.Cl2Error("cl2_pairs_parse failed rc=%d" % n)
        return chrom[:n], pos[:n]

    def names(self):
        k = ctypes.c_int32(0)
        s = self.lib.cl2_bedpe_names(self.h, ctypes.byref(k))
        return s.decode().split("\n") if k.value else []


def _chunks(f):
    """Whole-line byte chunks of a plain or gzipped text file."""
    of = gzip.open(f, "rb") if f.endswith(".gz") else open(f, "rb")
    with of:
        rest = b""
        while True:
            b = of.read(CHUNK)
            if not b:
                break
            b = rest + b
            cutp = b.rfind(b"\n")
            if cutp < 0:
                rest = b
                continue
            rest = b[cutp + 1:]
            yield b[:cutp + 1]
        if rest:
            yield rest


def _chunk_pets(buf, mapq, cs, cut, mcut, cis, fmt):
    """One text chunk -> ({(chrA, chrB): int64[k,2]}, counters).  Self-contained (own name table), so chunks can be
    handled by several threads; the tokeniser and most of numpy run without the GIL."""
    parser = BedpeParser()
    try:
        if fmt == "pairs":
            chrom, pos = parser.parse_pairs(buf)
            mq = np.full(len(pos), 255, dtype=np.int32)
        else:
            chrom, pos, mq = parser.parse(buf)
        names = parser.names()
    finally:
        parser.close()
    cnt = {"total": 0, "hiq": 0, "trans": 0, "cis": 0, "closeCis": 0, "farCis": 0}
    if len(mq) == 0:
        return {}, cnt
    rank = np.empty(len(names), dtype=np.int64)                          # string order of the names (ds.py:82, :134)
    rank[np.array(sorted(range(len(names)), key=lambda i: names[i]), dtype=np.int64)] = np.arange(len(names))
    ida, idb = chrom[:, 0].astype(np.int64), chrom[:, 1].astype(np.int64)
    is_cis = ida == idb
    if fmt == "pairs":
        pA, pB = pos[:, 0], pos[:, 1]
        swap = np.where(is_cis, pA > pB, rank[ida] > rank[idb])           # ds.py:124-126, 134-137
        cA, cB = np.where(swap, pB, pA), np.where(swap, pA, pB)
    else:
        sA, eA, sB, eB = pos[:, 0], pos[:, 1], pos[:, 2], pos[:, 3]
        swap = np.where(is_cis, sA + eA > sB + eB, rank[ida] > rank[idb])  # ds.py:61-64, 82-86
        sumA, sumB = np.where(swap, sB + eB, sA + eA), np.where(swap, sA + eA, sB + eB)
        cA = np.trunc(sumA / 2).astype(np.int64)                          # int((s + e) / 2)
        cB = np.trunc(sumB / 2).astype(np.int64)
    ida, idb = np.where(swap, idb, ida), np.where(swap, ida, idb)
    keep = np.ones(len(mq), dtype=bool)
    if len(cs) > 0:                                                      # io.py:102-103
        incs = np.array([nm in cs for nm in names], dtype=bool)
        keep &= incs[ida] & incs[idb]
    cnt["total"] = int(keep.sum())
    if fmt != "pairs":
        keep &= mq >= mapq                                               # io.py:105-106
    cnt["hiq"] = int(keep.sum())
    cnt["cis"] = int((keep & is_cis).sum())
    cnt["trans"] = int((keep & ~is_cis).sum())
    if cis:
        keep &= is_cis                                                   # io.py:113-114
    dist = np.abs(cB - cA)
    if cut > 0:                                                          # io.py:116-118
        close = keep & is_cis & (dist < cut)
        cnt["closeCis"] = int(close.sum())
        keep &= ~close
    if mcut > 0:                                                         # io.py:119-121
        far = keep & is_cis & (dist > mcut)
        cnt["farCis"] = int(far.sum())
        keep &= ~far
    ida, idb, cA, cB = ida[keep], idb[keep], cA[keep], cB[keep]
    kid = ida * len(names) + idb
    parts = {}
    for k in np.unique(kid).tolist():
        m = kid == k
        parts[(names[k // len(names)], names[k % len(names)])] = np.stack([cA[m], cB[m]], axis=1)
    return parts, cnt


def collect_pets(fs, logger=None, mapq=1, cs=[], cut=0, mcut=-1, cis=False, fmt="bedpe", cpu=1):
    """The parsing loop of parseBedpe (io.py:84-133) or, with fmt="pairs", of parsePairs (io.py:178-234; no MAPQ)
    without the de-duplication: returns ({"chrA-chrB": int64[k,2] (cA, cB) rows in file order}, counters dict with
    total / hiq / trans / cis / closeCis / farCis).  cpu > 1: that many threads work on the text chunks."""
    from concurrent.futures import ThreadPoolExecutor
    cnt = {"total": 0, "hiq": 0, "trans": 0, "cis": 0, "closeCis": 0, "farCis": 0}
    parts = {}                                   # (chrA, chrB) -> [int64[k,2]], chunks in file order

    def absorb(res):
        p, c = res
        for k, v in c.items():
            cnt[k] += v
        for k, v in p.items():
            parts.setdefault(k, []).append(v)

    nthreads = max(1, int(cpu))
    with ThreadPoolExecutor(nthreads) as pool:
        for f in fs:
            if logger is not None:
                logger.info("Parsing PETs from %s, requiring initial distance cutoff >%s and <%s" % (f, cut, mcut))
            pending = []
            for buf in _chunks(f):
                pending.append(pool.submit(_chunk_pets, buf, mapq, cs, cut, mcut, cis, fmt))
                while len(pending) > 2 * nthreads:            # bounded look-ahead: chunks are large
                    absorb(pending.pop(0).result())
            for fut in pending:
                absorb(fut.result())
    rows = {"%s-%s" % k: np.ascontiguousarray(np.concatenate(ch, axis=0), dtype=np.int64) for k, ch in parts.items()}
    return rows, cnt


def _write_dir(pets, cnt, fs, fout, logger, cut, mcut, mapq, ctx):
    """getUniqueTxt + txt2ixy + petMeta.json (io.py:134-164 / 235-265) for the collected rows."""
    import joblib
    total, hiq, t, c, closeCis, farCis = (cnt[k] for k in ("total", "hiq", "trans", "cis", "closeCis", "farCis"))
    uniques = 0
    ixyfs = []
    for key, xy in pets.items():
        pts = _lib.Points(ctx, xy)
        try:
            rows = pts.unique()
        finally:
            pts.close()
        uniques += len(rows)
        fn = os.path.join(fout, key + ".ixy")
        joblib.dump(np.ascontiguousarray(rows), fn)
        ixyfs.append(fn)
    if logger is not None:
        logger.info("Totaly %s PETs in target chromosomes from %s, in which %s high quality unqiue PETs" % (
            total, ",".join(fs), uniques))
    nr = 1 - uniques / 1.0 / (c - closeCis) if c > 0 else 0
    ds = {"Total PETs": total}
    if mapq is not None:
        ds["High Mapq(>=%s) PETs" % mapq] = hiq
    ds.update({
        "Total Trans PETs": t,
        "Total Cis PETs": c,
        "Filtered too close (<%s) PETs" % cut: closeCis,
        "Filtered too distant (>%s) PETs" % mcut: farCis,
        "Unique PETs": uniques,
        "Cis PETs Redundancy": nr,
    })
    metaf = fout + "/petMeta.json"
    with open(metaf, "w") as fo:
        json.dump(ds, fo)
    updateJson(ixyfs, metaf)
    return metaf


def parseBedpe(fs, fout, logger, mapq=1, cs=[], cut=0, mcut=-1, cis=False, cpu=1, ctx=None):
    """Drop-in for io.py:71-164.  fs: BEDPE / BEDPE.gz files of the replicates; fout: output directory (must exist,
    as in the reference); cs: wanted chromosomes; cpu: threads that parse the text."""
    ctx = ctx or _lib.default_context()
    pets, cnt = collect_pets(fs, logger, mapq=mapq, cs=cs, cut=cut, mcut=mcut, cis=cis, cpu=cpu)
    return _write_dir(pets, cnt, fs, fout, logger, cut, mcut, mapq, ctx)


def parsePairs(fs, fout, logger, cs=[], cut=0, mcut=-1, cis=False, cpu=1, ctx=None):
    """Drop-in for io.py:166-265: the same for .pairs / .pairs.gz input (no MAPQ column, so no MAPQ counter)."""
    ctx = ctx or _lib.default_context()
    pets, cnt = collect_pets(fs, logger, cs=cs, cut=cut, mcut=mcut, cis=cis, fmt="pairs", cpu=cpu)
    return _write_dir(pets, cnt, fs, fout, logger, cut, mcut, None, ctx)
