"""
TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported
through oracle/ref_shim.py) in the build container.  The reference ships no golden vectors of its own
(SURVEY.md section 4), so these are the pinned answers; the GPU box has no /root/reference and only reads the
committed files.

    python -m oracle.gen_golden            # rewrites tests/golden/

Contents
  dbscan_small.npz     120 random point sets (sparse, clustered, duplicate-heavy, X>Y rows) with the reference's
                       blockDBSCAN and cDBSCAN `.labels` (as int32 arrays, -1 = absent from the dict)
  cis_pipeline.npz     two synthetic chromosomes + the reference's callCisLoops `_loops.txt` (default and -hic,
                       eps 200,500,1000 / minPts 5; plus a two-minPts, cut/mcut run, an -emPair run and the
                       OUT_filtered/*.ixy matrices of a -filter run)
  trans_pipeline.npz   two synthetic chromosome pairs + the reference's callTransLoops `_trans_loops.txt`
  estsig_cis.npz       per-candidate fields of the reference's estLoopSig on one chromosome (every attribute)
  quant_loops.npz      the reference's `quant -loops` output (quant.py quantLoops, with and without -offp)
  filter_knn.npz       the reference's `filterPETs -knn` output matrices (filter.py _filterPETsByKNNs)
  dloops.npz           the reference's callDiffLoops.quantDloops fields + background lists for two samples
  peaks.npz            the reference's callPeaks.runDBSCANPeaks peaks (paired and -split)
  pre_bedpe.npz        two synthetic BEDPE files + the reference's `pre` output (io.py parseBedpe): rows per key, counters
  pre_pairs.npz        the same for .pairs input (io.py parsePairs)
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from cloops2_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_labels(cls, xy, eps, minPts):
    n = len(xy)
    mat = np.zeros((n, 3))
    mat[:, 0] = np.arange(n)
    mat[:, 1] = xy[:, 0]
    mat[:, 2] = xy[:, 1]
    mat = mat.astype("int")
    db = cls(mat, eps, minPts)
    lab = -np.ones(n, dtype=np.int32)
    for k, v in db.labels.items():
        lab[int(k)] = v
    return lab


def random_case(rng, it):
    n = int(rng.integers(1, 600))
    span = int(rng.choice([20, 50, 200, 1000, 5000]))
    eps = int(rng.integers(1, 60))
    minPts = int(rng.integers(2, 14))
    kind = it % 4
    if kind == 0:      # clustered blobs
        k = int(rng.integers(1, 7))
        cx = rng.integers(0, span, size=k)
        cy = cx + rng.integers(0, span, size=k)
        r = rng.integers(0, k, size=n)
        s = eps * rng.uniform(0.3, 2.5)
        x = np.abs(np.rint(rng.normal(cx[r], s))).astype(np.int64)
        y = np.abs(np.rint(rng.normal(cy[r], s))).astype(np.int64)
        xy = np.stack([np.minimum(x, y), np.maximum(x, y)], 1)
    elif kind == 1:    # sparse uniform
        x = rng.integers(0, span, size=n)
        xy = np.stack([x, x + rng.integers(0, span, size=n)], 1)
    elif kind == 2:    # duplicate heavy
        x = rng.integers(0, max(2, span // 10), size=n) * 3
        xy = np.stack([x, x + rng.integers(0, max(2, span // 10), size=n) * 2], 1)
    else:              # unordered rows (X > Y allowed), as a trans file would have
        xy = np.stack([rng.integers(0, span, size=n), rng.integers(0, span, size=n)], 1)
    return xy.astype(np.int64), eps, minPts


def gen_dbscan(ref):
    rng = np.random.default_rng(20261019)
    xs, offs, params, lb, lc = [], [0], [], [], []
    for it in range(120):
        xy, eps, minPts = random_case(rng, it)
        xs.append(xy)
        offs.append(offs[-1] + len(xy))
        params.append((eps, minPts))
        lb.append(ref_labels(ref.blockDBSCAN, xy, eps, minPts))
        lc.append(ref_labels(ref.cDBSCAN, xy, eps, minPts))
    np.savez_compressed(os.path.join(GOLD, "dbscan_small.npz"), xy=np.concatenate(xs), offsets=np.array(offs),
                        params=np.array(params), labels_block=np.concatenate(lb), labels_c2=np.concatenate(lc))


def run_ref_cis(ref, files, want_filtered=False, **kw):
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir(files, tmp + "/d")
    ref.cis.callCisLoops(tmp + "/d", tmp + "/out", ref.logger, cpu=1, filter=want_filtered, **kw)
    meta = json.load(open(tmp + "/d/petMeta.json"))
    if want_filtered:
        import joblib
        filt = {k: joblib.load(tmp + "/out_filtered/%s.ixy" % k) for k in files
                if os.path.exists(tmp + "/out_filtered/%s.ixy" % k)}
        ftot = json.load(open(tmp + "/out_filtered/petMeta.json"))["Unique PETs"]
        return open(tmp + "/out_loops.txt").read(), filt, ftot
    return open(tmp + "/out_loops.txt").read(), list(meta["data"]["cis"].keys()), meta["Unique PETs"]


def gen_cis(ref):
    files = {"chr21-chr21": synth.cis_chrom(4_000_000, 36000, 7, "hitrac"),
             "chr22-chr22": synth.cis_chrom(3_000_000, 24000, 8, "hitrac")}
    out = {}
    txt, order, tot = run_ref_cis(ref, files, eps=[200, 500, 1000], minPts=[5])
    out["default_txt"] = txt
    out["hic_txt"] = run_ref_cis(ref, files, eps=[200, 500, 1000], minPts=[5], hic=True)[0]
    out["cut_txt"] = run_ref_cis(ref, files, eps=[300, 800], minPts=[8, 4], cut=2000, mcut=1500000)[0]
    out["empair_txt"] = run_ref_cis(ref, files, eps=[800, 300], minPts=[4, 8], emPair=True)[0]
    ftxt, filt, ftot = run_ref_cis(ref, files, want_filtered=True, eps=[200, 500, 1000], minPts=[5])
    assert ftxt == txt
    out["filter_tot"] = ftot
    for k, v in filt.items():                     # -filter output: rows AND row order of OUT_filtered/<key>.ixy
        out["filter_" + k.replace("-", "_")] = v
    np.savez_compressed(os.path.join(GOLD, "cis_pipeline.npz"), order=np.array(order), tot=tot,
                        **{k.replace("-", "_"): v for k, v in files.items()}, **out)
    # per-candidate estLoopSig fields on one chromosome
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir({"chr21-chr21": files["chr21-chr21"]}, tmp + "/d")
    fixy = tmp + "/d/chr21-chr21.ixy"
    ref.cis.DBSCAN = ref.blockDBSCAN
    _, cand = ref.cis.runCisDBSCANLoops(fixy, 500, 5)
    coords = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in cand], dtype=np.int64)
    _, est = ref.cis.estLoopSig("chr21-chr21", cand, fixy, tot, minPts=5)
    fields = ["ra", "rb", "rab", "P2LL", "FDR", "ES", "density", "hypergeometric_p_value", "poisson_p_value",
              "binomial_p_value", "x_peak_poisson_p_value", "x_peak_es", "y_peak_poisson_p_value", "y_peak_es"]
    kept = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in est], dtype=np.int64)
    vals = np.array([[float(getattr(lp, f)) for f in fields] for lp in est], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "estsig_cis.npz"), candidates=coords, kept=kept, fields=np.array(fields),
                        values=vals, tot=tot)


def gen_trans(ref):
    files = {"chr21-chr22": synth.trans_pair(2_000_000, 2_400_000, 30000, 11, sigma=800, per=1500),
             "chr20-chr22": synth.trans_pair(1_500_000, 2_400_000, 20000, 12, sigma=800, per=1500)}
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir(files, tmp + "/d")
    ref.trans.callTransLoops(tmp + "/d", tmp + "/out", ref.logger, eps=[2000, 4000], minPts=[10, 5], cpu=1)
    meta = json.load(open(tmp + "/d/petMeta.json"))
    np.savez_compressed(os.path.join(GOLD, "trans_pipeline.npz"), order=np.array(list(meta["data"]["trans"].keys())),
                        trans_txt=open(tmp + "/out_trans_loops.txt").read(),
                        **{k.replace("-", "_"): v for k, v in files.items()})


def gen_quant(ref):
    """cLoops2/quant.py quantLoops on the cis golden files: the default run's loops plus shifted copies (some of them
    without any PET), with and without -offp."""
    from cLoops2 import quant as rq
    g = np.load(os.path.join(GOLD, "cis_pipeline.npz"))
    files = {k: g[k.replace("-", "_")] for k in g["order"].tolist()}
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir(files, tmp + "/d")
    meta = json.load(open(tmp + "/d/petMeta.json"))
    meta["Unique PETs"] = int(g["tot"])
    json.dump(meta, open(tmp + "/d/petMeta.json", "w"))
    lines = str(g["default_txt"]).rstrip("\n").split("\n")
    extra = []
    for i, line in enumerate(lines[1:]):
        f = line.split("\t")
        f[0] = "shift_%d" % i
        for j in (2, 3):
            f[j] = str(int(f[j]) + 40000)
        extra.append("\t".join(f))
    open(tmp + "/in_loops.txt", "w").write("\n".join(lines + extra) + "\n")
    out = {}
    for offp in (False, True):
        rq.quantLoops(tmp + "/d", tmp + "/in_loops.txt", tmp + "/q%d" % offp, ref.logger, offp=offp)
        out["quant_offp%d" % offp] = open(tmp + "/q%d_loops.txt" % offp).read()
    np.savez_compressed(os.path.join(GOLD, "quant_loops.npz"), in_loops=open(tmp + "/in_loops.txt").read(), **out)


def gen_filter_knn(ref):
    """cLoops2/filter.py _filterPETsByKNNs (filterPETs -knn) on the cis golden files, eps 500, minPts 5."""
    import joblib
    from cLoops2 import filter as rf
    g = np.load(os.path.join(GOLD, "cis_pipeline.npz"))
    files = {k: g[k.replace("-", "_")] for k in g["order"].tolist()}
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir(files, tmp + "/d")
    os.makedirs(tmp + "/o")
    out = {}
    for k in files:
        rf._filterPETsByKNNs(tmp + "/d/%s.ixy" % k, tmp + "/o", 500, 5)
        out["knn_" + k.replace("-", "_")] = joblib.load(tmp + "/o/%s.ixy" % k)
    np.savez_compressed(os.path.join(GOLD, "filter_knn.npz"), eps=500, minPts=5, **out)


def gen_dloops(ref):
    """cLoops2/callDiffLoops.py quantDloops: the default run's loops quantified against two samples (chr21 of the cis
    golden files as target, a differently seeded chromosome of the same shape as control), win=1 and win=2."""
    from cLoops2 import callDiffLoops as rd
    from cLoops2 import io as rio
    g = np.load(os.path.join(GOLD, "cis_pipeline.npz"))
    a = g["chr21_chr21"]
    # control: every second PET of the target plus an independent background, so both samples see the loops
    b = np.concatenate([a[::2], synth.cis_chrom(4_000_000, 12000, 77, "hitrac")])
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir({"chr21-chr21": a}, tmp + "/a")
    synth.write_ixy_dir({"chr21-chr21": b}, tmp + "/b")
    open(tmp + "/loops.txt", "w").write(str(g["default_txt"]))
    lines = str(g["default_txt"]).rstrip("\n").split("\n")
    extra = []
    for sh in (-1500, 900, 6000, 50000):                      # shifted copies: partial overlap, flanks, empty regions
        for i, line in enumerate(lines[1:]):
            f = line.split("\t")
            if f[1] != "chr21":
                continue
            f[0] = "shift%d_%d" % (sh, i)
            for j in (2, 3, 5, 6):
                f[j] = str(max(0, int(f[j]) + sh))
            extra.append("\t".join(f))
    open(tmp + "/loops.txt", "w").write("\n".join(lines + extra) + "\n")
    loops = rio.parseTxt2Loops(tmp + "/loops.txt")["chr21-chr21"]
    coords = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in loops], dtype=np.int64)
    out = {}
    fields = ["raw_trt_ra", "raw_trt_rb", "raw_con_ra", "raw_con_rb", "raw_trt_rab", "raw_con_rab", "raw_trt_mrab",
              "raw_con_mrab", "trt_es", "con_es", "distance", "size"]
    for win, cut in ((1, 0), (2, 3000)):
        _, ts, cs, dl = rd.quantDloops("chr21-chr21", loops, tmp + "/a/chr21-chr21.ixy", tmp + "/b/chr21-chr21.ixy",
                                       cut=cut, win=win)
        out["ts_w%d" % win] = np.array(ts, dtype=np.float64)
        out["cs_w%d" % win] = np.array(cs, dtype=np.float64)
        out["rows_w%d" % win] = np.array([[float(getattr(d, f)) for f in fields] for d in dl], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "dloops.npz"), loops=coords, control=b, fields=np.array(fields),
                        loops_txt=open(tmp + "/loops.txt").read(), **out)


def gen_peaks(ref):
    """cLoops2/callPeaks.py runDBSCANPeaks (clustering + stichPeaks) on chr21 of the cis golden files, PETs closer than
    1 kb as `callPeaks` uses them, paired and -split."""
    from cLoops2 import callPeaks as rp
    rp.logger = ref.logger
    g = np.load(os.path.join(GOLD, "cis_pipeline.npz"))
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir({"chr21-chr21": g["chr21_chr21"]}, tmp + "/d")
    fixy = tmp + "/d/chr21-chr21.ixy"
    out = {}
    for name, kw in (("e200", dict(eps=200, minPts=3, mcut=1000)), ("e300", dict(eps=300, minPts=4, mcut=2000)),
                     ("split", dict(eps=150, minPts=5, mcut=1000, split=True, splitExt=50))):
        peaks = rp.runDBSCANPeaks(fixy, kw.pop("eps"), kw.pop("minPts"), **kw)
        out[name] = np.array([[p.start, p.end, p.length] for p in peaks], dtype=np.int64).reshape(-1, 3)
    np.savez_compressed(os.path.join(GOLD, "peaks.npz"), **out)


def synth_bedpe(seed, n):
    """BEDPE text with everything parseBedpe has a branch for: cis / trans pairs on three chromosomes, ends in either
    order, exact and centre-only duplicates, low and missing MAPQ, short and non-numeric lines."""
    rng = np.random.default_rng(seed)
    chroms = ["chr21", "chr22", "chrX"]
    out = []
    for i in range(n):
        u = rng.random()
        ca = chroms[int(rng.integers(0, 3))]
        cb = ca if u < 0.8 else chroms[int(rng.integers(0, 3))]
        x = int(rng.integers(100, 200_000))
        y = x + int(np.exp(rng.uniform(np.log(50), np.log(100_000)))) if ca == cb else int(rng.integers(100, 200_000))
        la, lb = int(rng.integers(30, 150)), int(rng.integers(30, 150))
        a = (ca, x - la // 2, x - la // 2 + la, "+-"[int(rng.integers(0, 2))])
        b = (cb, y - lb // 2, y - lb // 2 + lb, "+-"[int(rng.integers(0, 2))])
        if rng.random() < 0.5:
            a, b = b, a
        q = rng.random()
        mq = "0" if q < 0.1 else ("." if q < 0.15 else str(int(rng.integers(1, 61))))
        line = [a[0], str(a[1]), str(a[2]), b[0], str(b[1]), str(b[2]), "r%d" % i, mq, a[3], b[3]]
        out.append("\t".join(line))
        if rng.random() < 0.08:                                  # exact duplicate
            out.append("\t".join(line))
        if rng.random() < 0.05:                                  # same centres, different fragment
            line2 = list(line)
            line2[1], line2[2] = str(int(line[1]) - 2), str(int(line[2]) + 2)
            out.append("\t".join(line2))
        if rng.random() < 0.02:
            out.append("\t".join(line[:7]))                      # too few columns
        if rng.random() < 0.02:
            bad = list(line)
            bad[4] = "*"
            out.append("\t".join(bad))                           # non-numeric coordinate
    return "\n".join(out) + "\n"


def gen_pre(ref):
    """cLoops2/io.py parseBedpe (`cLoops2 pre`) on two synthetic BEDPE files (one gzipped): per-key rows (sorted; the
    reference writes them in set order) and the petMeta.json counters, for the default options and for
    -c chr21,chr22 -cut 1000 -mcut 50000 -trans off."""
    import gzip
    import joblib
    tmp = tempfile.mkdtemp()
    t1, t2 = synth_bedpe(101, 6000), synth_bedpe(202, 4000)
    open(tmp + "/a.bedpe", "w").write(t1)
    with gzip.open(tmp + "/b.bedpe.gz", "wt") as fo:
        fo.write(t2)
    out = {"bedpe_a": t1, "bedpe_b": t2}
    for name, kw in (("default", dict(mapq=1, cs=[], cut=0, mcut=-1, cis=False)),
                     ("filtered", dict(mapq=10, cs=["chr21", "chr22"], cut=1000, mcut=50000, cis=True))):
        d = tmp + "/" + name
        os.makedirs(d)
        ref.io.parseBedpe([tmp + "/a.bedpe", tmp + "/b.bedpe.gz"], d, ref.logger, cpu=1, **kw)
        meta = json.load(open(d + "/petMeta.json"))
        keys = sorted(list(meta["data"]["cis"].keys()) + list(meta["data"]["trans"].keys()))
        out[name + "_keys"] = np.array(keys)
        for k in keys:
            m = np.asarray(joblib.load(d + "/%s.ixy" % k), dtype=np.int64).reshape(-1, 2)
            out[name + "_" + k.replace("-", "_")] = m[np.lexsort((m[:, 1], m[:, 0]))]
        meta.pop("data")
        out[name + "_meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(GOLD, "pre_bedpe.npz"), **out)


def synth_pairs(seed, n):
    """.pairs text: header lines, 7- and 8-column records, all pair types, ends in either order, duplicates, bad
    lines."""
    rng = np.random.default_rng(seed)
    chroms = ["chr21", "chr22", "chrX"]
    types = ["UU", "UR", "RU", "MU", "MR", "NN", "DD", "WW", "MM"]
    out = ["## pairs format v1.0", "#columns: readID chr1 pos1 chr2 pos2 strand1 strand2 pair_type"]
    for i in range(n):
        ca = chroms[int(rng.integers(0, 3))]
        cb = ca if rng.random() < 0.8 else chroms[int(rng.integers(0, 3))]
        x = int(rng.integers(100, 200_000))
        y = x + int(np.exp(rng.uniform(np.log(50), np.log(100_000)))) if ca == cb else int(rng.integers(100, 200_000))
        a, b = (ca, x), (cb, y)
        if rng.random() < 0.5:
            a, b = b, a
        f = ["r%d" % i, a[0], str(a[1]), b[0], str(b[1]), "+-"[int(rng.integers(0, 2))], "+-"[int(rng.integers(0, 2))]]
        if rng.random() < 0.7:
            f.append(types[int(rng.integers(0, len(types)))])
        out.append("\t".join(f))
        if rng.random() < 0.08:
            out.append("\t".join(f))
        if rng.random() < 0.02:
            out.append("\t".join(f[:5]))                          # too few columns
        if rng.random() < 0.02:
            out.append("\t".join(f + ["extra", "cols"]))          # too many columns
        if rng.random() < 0.02:
            bad = list(f)
            bad[2] = "!"
            out.append("\t".join(bad))                            # non-numeric position
    return "\n".join(out) + "\n"


def gen_pre_pairs(ref):
    """cLoops2/io.py parsePairs on two synthetic .pairs files (one gzipped), default options and filtered."""
    import gzip
    import joblib
    tmp = tempfile.mkdtemp()
    t1, t2 = synth_pairs(303, 6000), synth_pairs(404, 4000)
    open(tmp + "/a.pairs", "w").write(t1)
    with gzip.open(tmp + "/b.pairs.gz", "wt") as fo:
        fo.write(t2)
    out = {"pairs_a": t1, "pairs_b": t2}
    import contextlib
    import io as _io
    for name, kw in (("default", dict(cs=[], cut=0, mcut=-1, cis=False)),
                     ("filtered", dict(cs=["chr21", "chr22"], cut=1000, mcut=50000, cis=True))):
        d = tmp + "/" + name
        os.makedirs(d)
        with contextlib.redirect_stdout(_io.StringIO()):          # "Missing information ..." for every odd line
            ref.io.parsePairs([tmp + "/a.pairs", tmp + "/b.pairs.gz"], d, ref.logger, cpu=1, **kw)
        meta = json.load(open(d + "/petMeta.json"))
        keys = sorted(list(meta["data"]["cis"].keys()) + list(meta["data"]["trans"].keys()))
        out[name + "_keys"] = np.array(keys)
        for k in keys:
            m = np.asarray(joblib.load(d + "/%s.ixy" % k), dtype=np.int64).reshape(-1, 2)
            out[name + "_" + k.replace("-", "_")] = m[np.lexsort((m[:, 1], m[:, 0]))]
        meta.pop("data")
        out[name + "_meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(GOLD, "pre_pairs.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.load()
    gen_dbscan(ref)
    gen_cis(ref)
    gen_trans(ref)
    gen_quant(ref)
    gen_filter_knn(ref)
    gen_dloops(ref)
    gen_peaks(ref)
    gen_pre(ref)
    gen_pre_pairs(ref)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
