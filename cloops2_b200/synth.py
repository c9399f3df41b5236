"""
Deterministic synthetic PET generator: hg38 main chromosomes, distance-decay background
(d log-uniform, i.e. density ~ 1/d) plus injected Gaussian loops, exact duplicates removed the way `cLoops2 pre`
would (cLoops2/io.py:40-52), rows shuffled, int64[N,2] with X <= Y for cis.

Used by bench.py and the tests; writes real `.ixy` (joblib pickle of int64[N,2], io.py:55-67) + petMeta.json
(io.py:390-405) so the reference and this implementation read the same files.
"""
import json
import os

import numpy as np

# hg38 main chromosomes (lengths as in the reference's data/hg38.chrom.sizes)
HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
        ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
        ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
        ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
        ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
        ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]
HG38_TOTAL = sum(l for _, l in HG38)

PROFILES = {
    # name: d_min, sigma range, PETs per loop range, loops per PET
    "hitrac": dict(d_min=100, sigma=(150, 300), npl=(10, 80), per=5000),
    "c1": dict(d_min=100, sigma=(150, 150), npl=(10, 80), per=5000),      # SURVEY.md 8d C1' (chr21 stand-in)
    "hic": dict(d_min=1000, sigma=(2000, 5000), npl=(40, 400), per=5000),
    "chiapet": dict(d_min=1000, sigma=(500, 500), npl=(10, 80), per=5000),
}


def _unique_shuffled(a, b, rng):
    """Exact duplicate rows removed (as `pre` does) and rows shuffled; packed 64-bit keys keep it O(n log n) fast."""
    key = np.unique((a.astype(np.int64) << 32) | b.astype(np.int64))
    key = key[rng.permutation(key.shape[0])]
    xy = np.empty((key.shape[0], 2), dtype=np.int64)
    xy[:, 0] = key >> 32
    xy[:, 1] = key & 0xFFFFFFFF
    return xy


def cis_chrom(length, n, seed, profile="hitrac", max_d=2_000_000):
    """One chromosome: int64[~n,2], X<=Y, unique rows, shuffled."""
    p = PROFILES[profile]
    rng = np.random.default_rng(seed)
    n = int(n)
    k = max(1, n // p["per"])
    npl = rng.integers(p["npl"][0], p["npl"][1] + 1, size=k)
    n_loop = int(npl.sum())
    n_bg = max(0, n - n_loop)
    x = rng.integers(0, max(1, length - p["d_min"] - 1), size=n_bg)
    hi = np.minimum(max_d, length - x).astype(np.float64)
    lo = float(p["d_min"])
    hi = np.maximum(hi, lo + 1)
    d = np.exp(rng.uniform(np.log(lo), np.log(hi))).astype(np.int64)
    y = np.minimum(x + d, length - 1)
    a = rng.integers(0, max(1, length - 1_100_000), size=k)
    b = a + rng.integers(20_000, 1_000_000, size=k)
    sig = rng.uniform(p["sigma"][0], p["sigma"][1] + 1e-9, size=k)
    rep = np.repeat(np.arange(k), npl)
    lx = np.rint(rng.normal(a[rep], sig[rep])).astype(np.int64)
    ly = np.rint(rng.normal(b[rep], sig[rep])).astype(np.int64)
    xs = np.clip(np.concatenate([x, lx]), 0, length - 1)
    ys = np.clip(np.concatenate([y, ly]), 0, length - 1)
    lo_, hi_ = np.minimum(xs, ys), np.maximum(xs, ys)
    return _unique_shuffled(lo_, hi_, rng)


def trans_pair(len_a, len_b, n, seed, sigma=1000, per=20000, npl=(20, 80)):
    """One chromosome pair: uniform background + injected 2-D Gaussian clusters."""
    rng = np.random.default_rng(seed)
    n = int(n)
    k = max(1, n // per)
    cnt = rng.integers(npl[0], npl[1] + 1, size=k)
    n_bg = max(0, n - int(cnt.sum()))
    x = rng.integers(0, len_a, size=n_bg)
    y = rng.integers(0, len_b, size=n_bg)
    a = rng.integers(0, len_a, size=k)
    b = rng.integers(0, len_b, size=k)
    rep = np.repeat(np.arange(k), cnt)
    lx = np.rint(rng.normal(a[rep], sigma)).astype(np.int64)
    ly = np.rint(rng.normal(b[rep], sigma)).astype(np.int64)
    xs = np.clip(np.concatenate([x, lx]), 0, len_a - 1)
    ys = np.clip(np.concatenate([y, ly]), 0, len_b - 1)
    return _unique_shuffled(xs, ys, rng)


def cis_genome(n_total, config=2, profile="hitrac", chroms=None):
    """Ordered dict key -> int64[N,2]; PETs per chromosome proportional to length.
    seed = 20261019 + 1000*config + chromIndex."""
    chroms = HG38 if chroms is None else chroms
    tot_len = sum(l for _, l in chroms)
    out = {}
    for ci, (name, length) in enumerate(HG38):
        if (name, length) not in chroms:
            continue
        n = int(round(n_total * length / tot_len))
        if n <= 0:
            continue
        out["%s-%s" % (name, name)] = cis_chrom(length, n, 20261019 + 1000 * config + ci, profile)
    return out


def trans_genome(n_total, config=4, chroms=None, max_pairs=None):
    chroms = HG38 if chroms is None else chroms
    pairs = [(i, j) for i in range(len(chroms)) for j in range(i + 1, len(chroms))]
    if max_pairs:
        pairs = pairs[:max_pairs]
    w = np.array([chroms[i][1] * float(chroms[j][1]) for i, j in pairs])
    w /= w.sum()
    out = {}
    for pi, (i, j) in enumerate(pairs):
        n = int(round(n_total * w[pi]))
        if n <= 0:
            continue
        out["%s-%s" % (chroms[i][0], chroms[j][0])] = trans_pair(chroms[i][1], chroms[j][1], n,
                                                                  20261019 + 1000 * config + pi)
    return out


def write_ixy_dir(files, outdir):
    """Write <key>.ixy (joblib pickle, io.py:55-67) and petMeta.json (io.py:390-405)."""
    import joblib
    os.makedirs(outdir, exist_ok=True)
    meta = {"Unique PETs": int(sum(len(v) for v in files.values())), "data": {"cis": {}, "trans": {}}}
    for key, xy in files.items():
        f = os.path.abspath(os.path.join(outdir, key + ".ixy"))
        joblib.dump(np.ascontiguousarray(xy, dtype=np.int64), f)
        a, b = key.split("-")
        meta["data"]["cis" if a == b else "trans"][key] = {"ixy": f}
    with open(os.path.join(outdir, "petMeta.json"), "w") as fo:
        json.dump(meta, fo)
    return os.path.join(outdir, "petMeta.json")
