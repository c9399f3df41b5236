"""End-to-end loop calling on the GPU vs the reference's golden `_loops.txt` and vs the oracle on larger inputs."""
import json
import logging
import os

import numpy as np
import pytest

import oracle
from cloops2_b200 import _lib, synth
from cloops2_b200.callCisLoops import cis_chromosome, cols_to_loops, callCisLoops
from cloops2_b200.callTransLoops import trans_pair, callTransLoops
from conftest import golden_cis_files, golden_trans_files, txt_rows

pytestmark = pytest.mark.gpu


def _rows(loops):
    return oracle.loops_to_rows(loops)


@pytest.mark.parametrize("name,kw", [
    ("default_txt", dict(eps=[200, 500, 1000], minPts=[5])),
    ("cut_txt", dict(eps=[300, 800], minPts=[8, 4], cut=2000, mcut=1500000)),
    ("empair_txt", dict(eps=[800, 300], minPts=[4, 8], emPair=True))])
def test_cis_golden_rows(gpu_ctx, name, kw):
    g, files = golden_cis_files()
    loops = []
    for k, xy in files.items():
        final, st = cis_chromosome(gpu_ctx, tuple(k.split("-")), xy, int(g["tot"]), exact=True, **kw)
        loops.extend(cols_to_loops(tuple(k.split("-")), final))
    assert _rows(loops) == txt_rows(g[name])


def test_trans_golden_rows(gpu_ctx):
    g, files = golden_trans_files()
    loops = []
    for k, xy in files.items():
        final, st = trans_pair(gpu_ctx, tuple(k.split("-")), xy, [2000, 4000], [10, 5], exact=True)
        loops.extend(cols_to_loops(tuple(k.split("-")), final, cis=False))
    assert _rows(loops) == txt_rows(g["trans_txt"])


def test_cli_files_match_golden(gpu_ctx, tmp_path):
    """callCisLoops / callTransLoops on a real .ixy directory: `_loops.txt` byte-identical to the reference's."""
    g, files = golden_cis_files()
    gt, tfiles = golden_trans_files()
    allf = dict(files)
    allf.update(tfiles)
    d = str(tmp_path / "d")
    synth.write_ixy_dir(allf, d)
    # the reference ran on cis-only / trans-only directories: restore its totals and key order
    meta = json.load(open(d + "/petMeta.json"))
    meta["Unique PETs"] = int(g["tot"])
    meta["data"]["cis"] = {k: meta["data"]["cis"][k] for k in g["order"].tolist()}
    meta["data"]["trans"] = {k: meta["data"]["trans"][k] for k in gt["order"].tolist()}
    json.dump(meta, open(d + "/petMeta.json", "w"))
    log = logging.getLogger("t")
    out = str(tmp_path / "o")
    callCisLoops(d, out, log, eps=[200, 500, 1000], minPts=[5], ucsc=True, juicebox=True, washU=True)
    assert open(out + "_loops.txt").read() == str(g["default_txt"])
    for suf in ("_loops_ucsc.interact", "_loops_juicebox.txt", "_loops_legacyWashU.txt", "_loops_newWashU.txt"):
        assert os.path.getsize(out + suf) > 0
    callTransLoops(d, out, log, eps=[2000, 4000], minPts=[10, 5])
    assert open(out + "_trans_loops.txt").read() == str(gt["trans_txt"])


def test_cis_larger_vs_oracle(gpu_ctx):
    files = {"chr21-chr21": synth.cis_chrom(46_709_983, 400_000, 21, "hitrac"),
             "chr22-chr22": synth.cis_chrom(50_818_468, 300_000, 22, "hitrac")}
    tot = sum(len(v) for v in files.values())
    want, _ = oracle.call_cis_loops(files, tot, [200, 500, 1000], [5])
    loops = []
    for k, xy in files.items():
        final, st = cis_chromosome(gpu_ctx, tuple(k.split("-")), xy, tot, [200, 500, 1000], [5], exact=True)
        loops.extend(cols_to_loops(tuple(k.split("-")), final))
    assert len(want) > 20
    assert _rows(loops) == _rows(want)


P_COLS = (18, 19, 20, 21, 22)      # binomial, hypergeometric, poisson, peakA, peakB columns of _loops.txt


def _assert_rows_close(got, want, rtol):
    """Identical text except the five tail probabilities, which must agree to rtol (relative)."""
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for j, (a, b) in enumerate(zip(g, w)):
            if j in P_COLS:
                fa, fb = float(a), float(b)
                assert abs(fa - fb) <= rtol * abs(fb), (j, a, b)
            else:
                assert a == b, (j, a, b)


def test_device_pvalues_within_tolerance_of_reference(gpu_ctx):
    """Default device mode (no scipy on the path): same loops, same counts / ES / FDR / density text, tail
    probabilities within 1e-9 relative of the reference's (north_star tolerance); tails reach 1e-300."""
    g, files = golden_cis_files()
    loops = []
    for k, xy in files.items():
        final, st = cis_chromosome(gpu_ctx, tuple(k.split("-")), xy, int(g["tot"]), [200, 500, 1000], [5])
        loops.extend(cols_to_loops(tuple(k.split("-")), final))
    _assert_rows_close(_rows(loops), txt_rows(g["default_txt"]), 1e-9)
    gt, tfiles = golden_trans_files()
    loops = []
    for k, xy in tfiles.items():
        final, st = trans_pair(gpu_ctx, tuple(k.split("-")), xy, [2000, 4000], [10, 5])
        loops.extend(cols_to_loops(tuple(k.split("-")), final, cis=False))
    _assert_rows_close(_rows(loops), txt_rows(gt["trans_txt"]), 1e-9)


def test_fused_stats_kernels_against_oracle_counts(gpu_ctx):
    """cl2_loop_core / cl2_loop_tests vs the same quantities derived from the oracle's set-based counts."""
    from cloops2_b200 import hostops
    from test_hostops import FakeIndex
    rng = np.random.default_rng(8)
    for mode, xy in ((0, synth.cis_chrom(3_000_000, 80_000, 3, "hitrac")),
                     (1, synth.trans_pair(2_000_000, 2_500_000, 60_000, 4, per=1500))):
        i = rng.integers(0, len(xy), size=200)
        wa, wb = rng.integers(50, 3000, size=200), rng.integers(50, 3000, size=200)
        xs, ys = np.maximum(0, xy[i, 0] - wa // 2), np.maximum(0, xy[i, 1] - wb // 2)
        cands = np.stack([xs, xs + wa, ys, ys + wb], 1).astype(np.int64)
        pts = _lib.Points(gpu_ctx, xy)
        idx = _lib.Index(pts)
        try:
            core, hyp = idx.loop_core(cands, mode=mode)
            rpb = len(xy) / float(xy[:, 1].max() - xy[:, 0].min())
            st = idx.loop_tests(cands, core, rpb, mode=mode)
        finally:
            idx.close()
            pts.close()
        fk = FakeIndex(xy)
        wcore, whyp = fk.loop_core(cands, mode=mode)
        wst = fk.loop_tests(cands, wcore, rpb, mode=mode)
        assert np.array_equal(core, wcore)
        assert np.array_equal(st[:, :3], wst[:, :3])            # mrabs, mbps, FDR: bit-identical
        assert np.array_equal(st[:, 9:], wst[:, 9:])            # anchor counts
        assert np.array_equal(st[:, [6, 8]], wst[:, [6, 8]])    # anchor enrichment
        for a, b in ((hyp, whyp), (st[:, 3], wst[:, 3]), (st[:, 4], wst[:, 4]), (st[:, 5], wst[:, 5]), (st[:, 7], wst[:, 7])):
            ok = np.abs(a - b) <= 1e-12 * np.abs(b) + 1e-305     # device libm vs glibc: last-ulp differences only
            assert ok.all(), (mode, a[~ok][:3], b[~ok][:3])


def test_fused_est_loops_equals_step_by_step(gpu_ctx):
    """cl2_est_loops (filters on the device, survivors only) vs cl2_loop_core + cl2_loop_tests + host filters."""
    from cloops2_b200 import hostops

    class TwoPhase:            # hides est_loops so hostops takes the step-by-step path
        def __init__(self, idx):
            self.loop_core, self.loop_tests = idx.loop_core, idx.loop_tests

    rng = np.random.default_rng(12)
    for mode, xy in ((0, synth.cis_chrom(3_000_000, 80_000, 3, "hitrac")),
                     (1, synth.trans_pair(2_000_000, 2_500_000, 60_000, 4, per=1500))):
        i = rng.integers(0, len(xy), size=600)
        wa, wb = rng.integers(20, 3000, size=600), rng.integers(20, 3000, size=600)
        xs, ys = np.maximum(0, xy[i, 0] - wa // 2), np.maximum(0, xy[i, 1] - wb // 2)
        cands = np.stack([xs, xs + wa, ys, ys + wb], 1).astype(np.int64)
        span = float(xy[:, 1].max() - xy[:, 0].min())
        pts = _lib.Points(gpu_ctx, xy)
        idx = _lib.Index(pts)
        try:
            for hic in ((False, True) if mode == 0 else (False,)):
                if mode == 0:
                    a = hostops.est_cis(idx, cands, len(xy), span, len(xy), 3, hic=hic)
                    b = hostops.est_cis(TwoPhase(idx), cands, len(xy), span, len(xy), 3, hic=hic)
                else:
                    a = hostops.est_trans(idx, cands, len(xy), span, 3)
                    b = hostops.est_trans(TwoPhase(idx), cands, len(xy), span, 3)
                assert a["coords"].shape[0] > 0
                for k in hostops.COLS:
                    assert np.array_equal(a[k], b[k]), (mode, hic, k)
        finally:
            idx.close()
            pts.close()


def test_trans_larger_vs_oracle(gpu_ctx):
    """configs[3]-style shard (one chromosome pair at the 200 M-PET density: ~0.7 M PETs) vs the oracle pipeline."""
    files = {"chr5-chr9": synth.trans_pair(181_538_259, 138_394_717, 700_000, 59, sigma=1000, per=20000)}
    want = oracle.call_trans_loops(files, eps=[5000], minPts=[20, 10])
    final, st = trans_pair(gpu_ctx, ("chr5", "chr9"), files["chr5-chr9"], [5000], [20, 10], exact=True)
    assert len(want) > 5
    assert _rows(cols_to_loops(("chr5", "chr9"), final, cis=False)) == _rows(want)


def test_hic_larger_vs_oracle(gpu_ctx):
    """-hic (cDBSCAN2) whole path on a Hi-C-like chromosome, two eps x two minPts (configs[2] parameters)."""
    files = {"chr20-chr20": synth.cis_chrom(20_000_000, 300_000, 77, "hic")}
    tot = len(files["chr20-chr20"])
    want, _ = oracle.call_cis_loops(files, tot, [5000, 10000], [50, 20], hic=True)
    final, st = cis_chromosome(gpu_ctx, ("chr20", "chr20"), files["chr20-chr20"], tot, [5000, 10000], [50, 20], hic=True,
                               exact=True)
    assert len(want) > 5
    assert _rows(cols_to_loops(("chr20", "chr20"), final)) == _rows(want)


def test_command_line_subprocess(gpu_ctx, tmp_path):
    """`python -m cloops2_b200.cli callLoops ...` end to end (flag parsing, eps/minPts ordering, -trans, outputs)."""
    import subprocess
    import sys
    g, files = golden_cis_files()
    gt, tfiles = golden_trans_files()
    allf = dict(files)
    allf.update(tfiles)
    d = str(tmp_path / "d")
    synth.write_ixy_dir(allf, d)
    meta = json.load(open(d + "/petMeta.json"))
    meta["Unique PETs"] = int(g["tot"])
    meta["data"]["cis"] = {k: meta["data"]["cis"][k] for k in g["order"].tolist()}
    meta["data"]["trans"] = {k: meta["data"]["trans"][k] for k in gt["order"].tolist()}
    json.dump(meta, open(d + "/petMeta.json", "w"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "o")
    # eps given unsorted: the command sorts ascending (cLoops2.py:69-70), so the result equals the golden run
    r = subprocess.run([sys.executable, "-m", "cloops2_b200.cli", "callLoops", "-d", d, "-o", out, "-eps", "1000,200,500",
                        "-minPts", "5", "-w", "-j"], cwd=str(tmp_path), env=dict(os.environ, PYTHONPATH=root),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out + "_loops.txt").read() == str(g["default_txt"])
    assert os.path.getsize(out + "_loops_juicebox.txt") > 0 and os.path.getsize(out + "_loops_newWashU.txt") > 0
    assert os.path.exists(str(tmp_path / "cLoops2.log"))
    r = subprocess.run([sys.executable, "-m", "cloops2_b200.cli", "callLoops", "-d", d, "-o", out + "2", "-eps", "0",
                        "-minPts", "5"], cwd=str(tmp_path), env=dict(os.environ, PYTHONPATH=root), capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 1 and not os.path.exists(out + "2_loops.txt")        # "Input eps is 0. Return."


def test_filter_option_files_match_golden(gpu_ctx, tmp_path):
    """callCisLoops(..., filter=True): OUT_filtered/<key>.ixy (rows and row order) and its petMeta.json."""
    import joblib
    g, files = golden_cis_files()
    d = str(tmp_path / "d")
    synth.write_ixy_dir(files, d)
    meta = json.load(open(d + "/petMeta.json"))
    meta["data"]["cis"] = {k: meta["data"]["cis"][k] for k in g["order"].tolist()}
    json.dump(meta, open(d + "/petMeta.json", "w"))
    out = str(tmp_path / "o")
    callCisLoops(d, out, logging.getLogger("t"), eps=[200, 500, 1000], minPts=[5], filter=True)
    assert open(out + "_loops.txt").read() == str(g["default_txt"])
    for k in files:
        assert np.array_equal(joblib.load(out + "_filtered/%s.ixy" % k), g["filter_" + k.replace("-", "_")]), k
    fmeta = json.load(open(out + "_filtered/petMeta.json"))
    assert fmeta["Unique PETs"] == int(g["filter_tot"]) and sorted(fmeta["data"]["cis"]) == sorted(files)


@pytest.mark.parametrize("hic", [False, True])
def test_dense_hic_like_vs_oracle(gpu_ctx, hic, monkeypatch):
    """BASELINE configs[2] density (324 k PETs/Mb, Hi-C-like clusters), eps 5000,10000 x minPts 50,20, blockDBSCAN and
    cDBSCAN2: candidates whose windows hold tens of thousands of rows, counted both ways the library has for them (the
    cooperative scan kernel, CL2_WM=0, and the wavelet-matrix range counts, CL2_WM=1), crowded cells (heavy cell
    pairs), one diagonal component holding most cells.  Same loops, same text as the oracle."""
    files = {"chr21-chr21": synth.cis_chrom(4_600_000, 1_500_000, 33, "hic")}
    tot = len(files["chr21-chr21"])
    want, _ = oracle.call_cis_loops(files, tot, [5000, 10000], [50, 20], hic=hic)
    assert len(want) > 10
    for wm in ("0", "1"):
        monkeypatch.setenv("CL2_WM", wm)
        final, st = cis_chromosome(gpu_ctx, ("chr21", "chr21"), files["chr21-chr21"], tot, [5000, 10000], [50, 20], hic=hic,
                                   exact=True)
        got = cols_to_loops(("chr21", "chr21"), final)
        assert st["candidates"] > 500
        assert _rows(got) == _rows(want), wm


def test_wavelet_counts_on_unordered_and_clipped_loops(gpu_ctx, monkeypatch):
    """The wavelet form of estLoopSig's counts on what the pipeline never produces but the drop-in estLoopSig may be
    handed: overlapping anchors, windows clipped at 0, trans-like rows with X > Y -- against the scan form."""
    from cloops2_b200 import hostops
    rng = np.random.default_rng(12)
    n = 400_000
    x = rng.integers(0, 600_000, n)
    xy = np.stack([x, x + np.exp(rng.uniform(0, np.log(200_000), n)).astype(np.int64)], 1)
    xy2 = np.stack([rng.integers(0, 500_000, n), rng.integers(0, 500_000, n)], 1)
    a = rng.integers(0, 400_000, size=(300, 1))
    la, lb, gap = rng.integers(2_000, 40_000, size=(300, 1)), rng.integers(2_000, 40_000, size=(300, 1)), rng.integers(-30_000, 90_000, size=(300, 1))
    loops = np.concatenate([a, a + la, np.maximum(0, a + la + gap), np.maximum(0, a + la + gap) + lb], 1)
    loops[:40, 0] = rng.integers(0, 3_000, 40)                      # windows hanging over position 0
    loops[:40, 1] = loops[:40, 0] + la[:40, 0]
    for mode, data in ((0, xy), (1, xy2)):
        core, na, nb, nab, anc = oracle.XYIndex(data).loop_counts(loops, win=5, mode=mode)
        mr, mb, fdr = hostops.permuted_stats(na, nb, nab, core[:, 2], cis=(mode == 0))
        pts = _lib.Points(gpu_ctx, data)
        idx = _lib.Index(pts)
        out = {}
        for wm in ("0", "1"):
            monkeypatch.setenv("CL2_WM", wm)
            out[wm] = sel, c, hyp, st = idx.est_loops(loops, 0.7, 3, mode=mode)
            assert len(sel) >= (3 if mode == 0 else 100)
            assert np.array_equal(c, core[sel])                              # ra, rb, rab, lower as the oracle counts them
            assert np.array_equal(st[:, 0], mr[sel]) and np.array_equal(st[:, 2], fdr[sel])
            assert np.array_equal(st[:, 9:13], anc[sel].astype(np.float64))
        idx.close()
        pts.close()
        for u, v in zip(out["0"], out["1"]):
            assert np.array_equal(u, v)
