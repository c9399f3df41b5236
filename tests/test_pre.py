"""`cLoops2 pre` (third "next" row): the oracle restatement of parseBedpe vs the reference's golden
output, the native tokeniser + vectorised PET logic vs the oracle (CPU), and the whole command with the GPU
de-duplication vs the golden output."""
import gzip
import json

import numpy as np
import pytest

import oracle
from conftest import load_golden

CASES = (("default", dict(mapq=1, cs=[], cut=0, mcut=-1, cis=False)),
         ("filtered", dict(mapq=10, cs=["chr21", "chr22"], cut=1000, mcut=50000, cis=True)))
META = {"total": "Total PETs", "trans": "Total Trans PETs", "cis": "Total Cis PETs"}


def _write_inputs(g, d):
    (d / "a.bedpe").write_text(str(g["bedpe_a"]))
    with gzip.open(str(d / "b.bedpe.gz"), "wt") as fo:
        fo.write(str(g["bedpe_b"]))
    return [str(d / "a.bedpe"), str(d / "b.bedpe.gz")]


@pytest.mark.parametrize("name,kw", CASES)
def test_oracle_pre_matches_reference(name, kw):
    g = load_golden("pre_bedpe.npz")
    lines = (str(g["bedpe_a"]) + str(g["bedpe_b"])).split("\n")[:-1]
    rows, meta = oracle.pre_bedpe(lines, **kw)
    assert meta == json.loads(str(g[name + "_meta"]))
    assert sorted(rows) == g[name + "_keys"].tolist()
    for k, v in rows.items():
        assert np.array_equal(np.array(v, dtype=np.int64), g[name + "_" + k.replace("-", "_")]), k


@pytest.mark.parametrize("name,kw", CASES)
def test_native_tokeniser_and_pet_logic(tmp_path, name, kw):
    """collect_pets (library tokeniser + numpy PET arithmetic, no GPU) vs the golden rows / counters."""
    from cloops2_b200.pre import collect_pets
    g = load_golden("pre_bedpe.npz")
    pets, cnt = collect_pets(_write_inputs(g, tmp_path), None, **kw)
    want = json.loads(str(g[name + "_meta"]))
    for k, mk in META.items():
        assert cnt[k] == want[mk], k
    assert cnt["hiq"] == want["High Mapq(>=%s) PETs" % kw["mapq"]]
    assert cnt["closeCis"] == want["Filtered too close (<%s) PETs" % kw["cut"]]
    assert cnt["farCis"] == want["Filtered too distant (>%s) PETs" % kw["mcut"]]
    assert sorted(pets) == g[name + "_keys"].tolist()
    for k, xy in pets.items():
        assert np.array_equal(np.unique(xy, axis=0), g[name + "_" + k.replace("-", "_")]), k


def test_tokeniser_accepts_and_skips_like_python(tmp_path):
    from cloops2_b200.pre import collect_pets
    good = "chr1\t100\t200\tchr1\t5000\t5100\tid\t30\t+\t-"
    lines = [good,
             "chr1\t100\t200\tchr1\t5000\t5100\tid\t30\t+",            # 9 fields
             "chr1\t1e3\t200\tchr1\t5000\t5100\tid\t30\t+\t-",         # float-looking coordinate
             "chr1\t 100 \t200\tchr1\t5000\t5100\tid\tNA\t+\t-\textra\tcols",   # blanks, mapq not an int, extra fields
             "",
             "#comment",
             "chr2\t7000\t7100\tchr1\t10\t20\tid\t60\t-\t+",           # trans, names out of order
             "chr1\t9000\t9100\tchr1\t300\t400\tid\t60\t-\t+"]         # cis, ends out of order
    f = tmp_path / "t.bedpe"
    f.write_text("\n".join(lines))                                    # no trailing newline
    pets, cnt = collect_pets([str(f)], None, mapq=1)
    want, meta = oracle.pre_bedpe(lines, mapq=1)
    assert cnt["total"] == meta["Total PETs"] == 4
    assert {k: np.unique(v, axis=0).tolist() for k, v in pets.items()} == {k: [list(r) for r in v] for k, v in want.items()}
    assert pets["chr1-chr2"].tolist() == [[15, 7050]]


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", CASES)
def test_gpu_pre_command(gpu_ctx, tmp_path, name, kw):
    import joblib
    from cloops2_b200.pre import parseBedpe
    g = load_golden("pre_bedpe.npz")
    out = tmp_path / "out"
    out.mkdir()
    parseBedpe(_write_inputs(g, tmp_path), str(out), None, ctx=gpu_ctx, **kw)
    meta = json.load(open(str(out / "petMeta.json")))
    data = meta.pop("data")
    assert meta == json.loads(str(g[name + "_meta"]))
    keys = sorted(list(data["cis"]) + list(data["trans"]))
    assert keys == g[name + "_keys"].tolist()
    assert all((k.split("-")[0] == k.split("-")[1]) == (k in data["cis"]) for k in keys)
    for k in keys:
        m = joblib.load(data["cis" if k in data["cis"] else "trans"][k]["ixy"])
        assert m.dtype == np.int64 and np.array_equal(m, g[name + "_" + k.replace("-", "_")]), k


@pytest.mark.gpu
def test_gpu_unique_rows_random(gpu_ctx):
    from cloops2_b200 import _lib
    rng = np.random.default_rng(9)
    for n, span in ((1, 10), (5000, 40), (200_000, 3000), (300_000, 2_000_000_000)):
        xy = rng.integers(0, span, size=(n, 2)).astype(np.int64)
        xy = np.concatenate([xy, xy[: n // 3]])
        pts = _lib.Points(gpu_ctx, xy)
        try:
            assert np.array_equal(pts.unique(), np.unique(xy, axis=0))
        finally:
            pts.close()


def test_tokeniser_fuzz_against_python_parser(tmp_path):
    """Random well-formed and malformed lines: the native tokeniser + numpy PET logic keeps exactly the rows and
    counters the line-by-line restatement keeps."""
    from cloops2_b200.pre import collect_pets
    rng = np.random.default_rng(123)
    chroms = ["chr1", "chr10", "chr2", "chrX", "chrM", "1", "scaffold_12"]
    # (int() would also take non-ASCII decimal digits and "1_000"; the tokeniser does not -- DESIGN.md section 1)
    junk = ["", "*", ".", "NA", "1.5", "-", "+7", " 12", "12 ", "0x10", "1e3", "12a", "-5", "--5", "+", "1 2"]

    def field_int():
        if rng.random() < 0.06:
            return junk[int(rng.integers(0, len(junk)))]
        return str(int(rng.integers(0, 5_000_000)))

    lines = []
    for _ in range(4000):
        a, b = chroms[int(rng.integers(0, len(chroms)))], chroms[int(rng.integers(0, len(chroms)))]
        if rng.random() < 0.6:
            b = a
        f = [a, field_int(), field_int(), b, field_int(), field_int(), "id", junk[int(rng.integers(0, len(junk)))]
             if rng.random() < 0.2 else str(int(rng.integers(0, 61))), "+", "-"]
        k = rng.random()
        if k < 0.04:
            f = f[:int(rng.integers(1, 10))]                   # short line
        elif k < 0.08:
            f = f + ["x", "y"]                                 # extra columns
        lines.append("\t".join(f))
    for kw in (dict(mapq=1), dict(mapq=20, cs=["chr1", "chr2", "chrX"], cut=1000, mcut=2_000_000, cis=True)):
        path = tmp_path / ("fuzz%d.bedpe" % kw["mapq"])
        path.write_text("\n".join(lines) + "\n", encoding="utf-8")
        pets, cnt = collect_pets([str(path)], None, **kw)
        want, meta = oracle.pre_bedpe(lines, **kw)
        assert cnt["total"] == meta["Total PETs"] and cnt["cis"] == meta["Total Cis PETs"]
        assert cnt["trans"] == meta["Total Trans PETs"] and cnt["hiq"] == meta["High Mapq(>=%s) PETs" % kw["mapq"]]
        assert sorted(pets) == sorted(want)
        for k, xy in pets.items():
            assert np.unique(xy, axis=0).tolist() == [list(r) for r in want[k]], k


# ---- .pairs input (parsePairs) ------------------------------------------------------------------------------------
PCASES = (("default", dict(cs=[], cut=0, mcut=-1, cis=False)),
          ("filtered", dict(cs=["chr21", "chr22"], cut=1000, mcut=50000, cis=True)))


def _write_pairs(g, d):
    (d / "a.pairs").write_text(str(g["pairs_a"]))
    with gzip.open(str(d / "b.pairs.gz"), "wt") as fo:
        fo.write(str(g["pairs_b"]))
    return [str(d / "a.pairs"), str(d / "b.pairs.gz")]


@pytest.mark.parametrize("name,kw", PCASES)
def test_oracle_pre_pairs_matches_reference(name, kw):
    g = load_golden("pre_pairs.npz")
    lines = (str(g["pairs_a"]) + str(g["pairs_b"])).split("\n")[:-1]
    rows, meta = oracle.pre_pairs(lines, **kw)
    assert meta == json.loads(str(g[name + "_meta"]))
    assert sorted(rows) == g[name + "_keys"].tolist()
    for k, v in rows.items():
        assert np.array_equal(np.array(v, dtype=np.int64), g[name + "_" + k.replace("-", "_")]), k


@pytest.mark.parametrize("name,kw", PCASES)
def test_native_pairs_tokeniser_and_logic(tmp_path, name, kw):
    from cloops2_b200.pre import collect_pets
    g = load_golden("pre_pairs.npz")
    pets, cnt = collect_pets(_write_pairs(g, tmp_path), None, fmt="pairs", **kw)
    want = json.loads(str(g[name + "_meta"]))
    for k, mk in META.items():
        assert cnt[k] == want[mk], k
    assert cnt["closeCis"] == want["Filtered too close (<%s) PETs" % kw["cut"]]
    assert cnt["farCis"] == want["Filtered too distant (>%s) PETs" % kw["mcut"]]
    assert sorted(pets) == g[name + "_keys"].tolist()
    for k, xy in pets.items():
        assert np.array_equal(np.unique(xy, axis=0), g[name + "_" + k.replace("-", "_")]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", PCASES)
def test_gpu_pre_pairs_command(gpu_ctx, tmp_path, name, kw):
    import joblib
    from cloops2_b200.pre import parsePairs
    g = load_golden("pre_pairs.npz")
    out = tmp_path / "out"
    out.mkdir()
    parsePairs(_write_pairs(g, tmp_path), str(out), None, ctx=gpu_ctx, **kw)
    meta = json.load(open(str(out / "petMeta.json")))
    data = meta.pop("data")
    assert meta == json.loads(str(g[name + "_meta"]))
    keys = sorted(list(data["cis"]) + list(data["trans"]))
    assert keys == g[name + "_keys"].tolist()
    for k in keys:
        m = joblib.load(data["cis" if k in data["cis"] else "trans"][k]["ixy"])
        assert np.array_equal(m, g[name + "_" + k.replace("-", "_")]), k
