"""
Host-side logic of the product (vectorised numpy + native selection) against the oracle's plain-Python restatement.
The device index is replaced by a stand-in that answers loop_counts from the oracle's C counting, so these run
without a GPU; the GPU tests exercise the same code with the real kernels.
"""
import os

import numpy as np

import oracle
from cloops2_b200 import hostops
from cloops2_b200.callCisLoops import cols_to_loops
from conftest import golden_cis_files, golden_trans_files, load_golden


class FakeIndex:
    def __init__(self, xy):
        self.ix = oracle.XYIndex(xy)

    def loop_counts(self, loops, win=5, mode=0, core=True, windows=True, anchors=True):
        c, na, nb, nab, anc = self.ix.loop_counts(loops, win=win, mode=mode)
        return (c if core else None, na if windows else None, nb if windows else None, nab if windows else None,
                anc if anchors else None)


    def loop_core(self, loops, mode=0):
        from cloops2_b200 import _lib
        L = _lib.load()
        c = self.ix.loop_counts(loops, win=5, mode=mode)[0]
        hyp = np.array([L.cl2_stat_hypergeom_sf(float(r[2]) - 1.0, float(self.ix.n), float(r[0]), float(r[1])) for r in c])
        return c, hyp

    def loop_tests(self, loops, core, rpb, win=5, mode=0):
        from cloops2_b200 import _lib
        L = _lib.load()
        _, na, nb, nab, anc = self.ix.loop_counts(loops, win=win, mode=mode)
        rab = core[:, 2]
        mrabs, mbps, fdr = hostops.permuted_stats(na, nb, nab, rab, cis=(mode == 0))
        st = np.zeros((len(loops), 13))
        st[:, 0], st[:, 1], st[:, 2] = mrabs, mbps, fdr
        for i in range(len(loops)):
            ra, rb, r = float(core[i, 0]), float(core[i, 1]), float(core[i, 2])
            st[i, 3] = L.cl2_stat_poisson_sf(r - 1.0, mrabs[i])
            st[i, 4] = L.cl2_stat_binom_sf(r - 1.0, ra * rb - r if mode == 0 else ra * rb, mbps[i])
            for a, (lo, hi) in enumerate(((loops[i][0], loops[i][1]), (loops[i][2], loops[i][3]))):
                count, wide = float(anc[i, 2 * a]), float(anc[i, 2 * a + 1])
                c = max((wide - count) / 5 / 2, rpb * float(hi - lo), 1.0)
                st[i, 5 + 2 * a] = L.cl2_stat_poisson_sf(count - 1.0, c)
                st[i, 6 + 2 * a] = count / c
        st[:, 9:13] = anc
        return st


def _rows(loops):
    return oracle.loops_to_rows(loops)


def test_est_cis_matches_oracle_and_reference_fields():
    e = load_golden("estsig_cis.npz")
    _, files = golden_cis_files()
    xy = files["chr21-chr21"]
    span = float(xy[:, 1].max() - xy[:, 0].min())
    cols = hostops.est_cis(FakeIndex(xy), e["candidates"], len(xy), span, int(e["tot"]), 5, exact=True)
    assert np.array_equal(cols["coords"], e["kept"])
    names = dict(ra="ra", rb="rb", rab="rab", P2LL="P2LL", FDR="FDR", ES="ES", density="density",
                 hypergeometric_p_value="hyp", poisson_p_value="pop", binomial_p_value="nbp",
                 x_peak_poisson_p_value="px", x_peak_es="esx", y_peak_poisson_p_value="py", y_peak_es="esy")
    for j, f in enumerate(e["fields"].tolist()):
        assert np.array_equal(np.asarray(cols[names[f]], dtype=np.float64), e["values"][:, j]), f


import pytest


@pytest.mark.parametrize("dedup_first", [True, False])
def test_full_cis_host_path_reproduces_reference_txt(dedup_first):
    """dedup_first=True is the reference's order (combine/dedup, then test); False is the product's (test every
    candidate, drop repeated coordinates among the survivors) -- both must give the reference's rows."""
    g, files = golden_cis_files()
    tot = int(g["tot"])
    for name, eps, minPts, hic in (("default_txt", [200, 500, 1000], [5], False), ("hic_txt", [200, 500, 1000], [5], True)):
        loops = []
        for k, xy in files.items():
            per_pass = []
            for ep in eps:
                for mp in minPts:
                    lab, ncl = (oracle.cdbscan2 if hic else oracle.block_dbscan)(xy, ep, mp)
                    bbox, _ = oracle.cluster_bbox(xy, lab, ncl)
                    per_pass.append(bbox[hostops.cis_candidate_mask(bbox)])
            cands = hostops.combine_unique(per_pass) if dedup_first else np.concatenate(per_pass)
            cands = cands[hostops.cis_distance(cands) > 0]
            span = float(xy[:, 1].max() - xy[:, 0].min())
            cols = hostops.est_cis(FakeIndex(xy), cands, len(xy), span, tot, max(minPts), hic=hic)
            if not dedup_first:
                cols = hostops.take(cols, hostops.first_occurrence(cols["coords"]))
            final = hostops.select_twice(cols, hostops.mark_sig_cis(cols, hic=hic))
            final = hostops.scipy_pvalues(final, len(xy), float(len(xy)) / span, cis=True)
            loops.extend(cols_to_loops(tuple(k.split("-")), final, cis=True))
        from conftest import txt_rows
        assert _rows(loops) == txt_rows(g[name]), name


def test_full_trans_host_path_reproduces_reference_txt():
    from cloops2_b200.callTransLoops import _combine_across_passes
    from conftest import txt_rows
    g, files = golden_trans_files()
    loops = []
    for k, xy in files.items():
        per_pass = []
        for ep in [2000, 4000]:
            for mp in [10, 5]:
                lab, ncl = oracle.block_dbscan(xy, ep, mp)
                bbox, _ = oracle.cluster_bbox(xy, lab, ncl)
                per_pass.append(bbox[hostops.trans_candidate_mask(bbox)])
        cands = _combine_across_passes(per_pass)
        span = float(xy[:, 1].max() - xy[:, 0].min())
        cols = hostops.est_trans(FakeIndex(xy), cands, len(xy), span, 10)
        # the product's order: test everything, then drop coordinates seen in an earlier pass
        allc = np.concatenate(per_pass)
        pid = np.concatenate([np.full(len(c), i) for i, c in enumerate(per_pass)])
        cols2 = hostops.est_trans(FakeIndex(xy), allc, len(xy), span, 10, pass_id=pid)
        cols2 = hostops.take(cols2, hostops.first_occurrence(cols2["coords"], cols2.pop("pass_id")))
        assert np.array_equal(cols2["coords"], cols["coords"]) and np.array_equal(cols2["nbp"], cols["nbp"])
        final = hostops.select_twice(cols, hostops.mark_sig_trans(cols))
        final = hostops.scipy_pvalues(final, len(xy), float(len(xy)) / span, cis=False)
        loops.extend(cols_to_loops(tuple(k.split("-")), final, cis=False))
    assert _rows(loops) == txt_rows(g["trans_txt"])


def test_native_selection_matches_plain_python():
    rng = np.random.default_rng(3)
    for it in range(200):
        L = int(rng.integers(0, 60))
        xs = rng.integers(0, 400, size=L)
        ys = xs + rng.integers(50, 600, size=L)
        coords = np.stack([xs, xs + rng.integers(1, 80, size=L), ys, ys + rng.integers(1, 80, size=L)], 1).astype(np.int64)
        es = rng.integers(1, 6, size=L).astype(np.float64)       # many ties
        loops = []
        for i in range(L):
            lp = oracle.OLoop()
            lp.chromX = lp.chromY = "c"
            lp.x_start, lp.x_end, lp.y_start, lp.y_end = (int(v) for v in coords[i])
            lp.ES = float(es[i])
            lp.significant = 1
            lp.ra = i                                            # identity tag
            loops.append(lp)
        want = [lp.ra for lp in oracle.sel_sig_loops(loops)]
        got = hostops.sel_sig(coords, es).tolist()
        assert got == want, it


def test_combine_unique_is_first_occurrence_order():
    a = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [1, 2, 3, 4]], dtype=np.int64)
    b = np.array([[9, 9, 9, 9], [5, 6, 7, 8]], dtype=np.int64)
    out = hostops.combine_unique([a, b])
    assert out.tolist() == [[1, 2, 3, 4], [5, 6, 7, 8], [9, 9, 9, 9]]
    assert hostops.combine_unique([]).shape == (0, 4)


def test_device_tail_functions_against_scipy_and_exact_values():
    """fp64 tails of cl2_stats.cuh (host build of the device code).  Tolerances: 1e-9 relative against scipy where
    scipy itself is accurate; scipy 1.18.1's hypergeom / binom drift up to ~1e-7 for populations >= 1e7 (checked
    against exact rational arithmetic when generating the constants below), there the exact values are the bar."""
    from scipy.stats import poisson, binom, hypergeom
    from cloops2_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(3000):
        mu = float(10 ** rng.uniform(-3, 3.5))
        k = float(rng.integers(0, int(mu * 3 + 60)))
        want = float(poisson.sf(k, mu))
        if want > 1e-290:
            assert abs(L.cl2_stat_poisson_sf(k, mu) - want) <= 1e-9 * want
    for _ in range(3000):
        n = float(int(10 ** rng.uniform(1, 6)))
        p = float(10 ** rng.uniform(-6, -0.5))
        k = float(rng.integers(0, int(min(n, n * p * 3 + 80)) + 1))
        want = float(binom.sf(k, n, p))
        if want > 1e-200:      # Boost's ibeta loses digits below ~1e-200 (1.9e-6 relative seen at 1e-280)
            assert abs(L.cl2_stat_binom_sf(k, n, p) - want) <= 1e-9 * want
    for _ in range(1500):
        M = int(10 ** rng.uniform(3, 5.5))
        n = int(rng.integers(1, M // 4 + 2))
        N = int(rng.integers(1, M // 4 + 2))
        k = float(rng.integers(0, min(n, N) + 1))
        want = float(hypergeom.sf(k, M, n, N))
        if want > 1e-290:
            assert abs(L.cl2_stat_hypergeom_sf(k, M, n, N) - want) <= 1e-9 * want, (k, M, n, N)
    # exact values (60-digit rational arithmetic) where scipy is off by 7e-8 / 2e-8
    assert abs(L.cl2_stat_hypergeom_sf(8, 235907837, 20298, 49454) - 0.029946224828182528793) < 1e-15
    assert abs(L.cl2_stat_binom_sf(2.0, 1106177179.0, 3.2149593661895347e-09) - 0.68944314367205768) < 1e-14
    # degenerate arguments the path produces (median background of 0)
    assert L.cl2_stat_poisson_sf(4.0, 0.0) == 0.0 == float(poisson.sf(4.0, 0.0))
    assert L.cl2_stat_binom_sf(4.0, 100.0, 0.0) == 0.0 == float(binom.sf(4.0, 100.0, 0.0))


def test_first_occurrence_rules():
    c = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [1, 2, 3, 4], [1, 2, 3, 4], [9, 9, 9, 9]], dtype=np.int64)
    assert hostops.first_occurrence(c).tolist() == [0, 1, 4]
    # trans: duplicates inside the first pass that holds them survive, later passes are dropped
    assert hostops.first_occurrence(c, np.array([0, 0, 0, 1, 1])).tolist() == [0, 1, 2, 4]


def test_filter_pets_matches_reference_rows_and_order(tmp_path):
    """-filter (callCisLoops.py:425-468): the rewritten .ixy matrices equal the reference's, row order included."""
    import joblib
    from cloops2_b200 import synth
    from cloops2_b200.callCisLoops import filterPETs
    g, files = golden_cis_files()
    d = str(tmp_path / "d")
    synth.write_ixy_dir(files, d)
    _, per_key = oracle.call_cis_loops(files, int(g["tot"]), [200, 500, 1000], [5])
    out = str(tmp_path / "f")
    os.makedirs(out)
    tot = 0
    for k, loops in per_key.items():
        filterPETs(k, out, d + "/%s.ixy" % k, loops, margin=1000)
        got = joblib.load(out + "/%s.ixy" % k)
        assert np.array_equal(got, g["filter_" + k.replace("-", "_")]), k
        tot += len(got)
    assert tot == int(g["filter_tot"])


def _fake_final(rng, n, N, length):
    """Columns shaped like the final loops of one file (only what the tail probabilities read)."""
    ra = rng.integers(200, 3000, size=n).astype(np.int64)
    rb = rng.integers(200, 3000, size=n).astype(np.int64)
    rab = rng.integers(10, 80, size=n).astype(np.int64)
    xs = rng.integers(0, length - 200000, size=n)
    la, lb = rng.integers(300, 2000, size=n), rng.integers(300, 2000, size=n)
    cols = hostops._empty_cols()
    cols.update(coords=np.stack([xs, xs + la, xs + 50000, xs + 50000 + lb], 1).astype(np.int64), ra=ra, rb=rb, rab=rab,
                mrabs=rng.uniform(0.5, 5, size=n), mbps=rng.uniform(1e-7, 1e-5, size=n),
                anc=np.stack([rng.integers(100, 3000, size=n), rng.integers(500, 9000, size=n),
                              rng.integers(100, 3000, size=n), rng.integers(500, 9000, size=n)], 1).astype(np.float64))
    for k in ("P2LL", "FDR", "ES", "density", "hyp", "pop", "nbp", "px", "py", "esx", "esy"):
        cols[k] = np.zeros(n)
    cols["N"], cols["rpb"] = N, N / float(length)
    return cols


def test_pvalue_worker_processes_give_scipys_values(monkeypatch):
    """The exact mode evaluates scipy's tails in worker processes (hostops.pvalue_pool; scipy's Boost ufuncs keep the
    interpreter lock): same bits as evaluating them inline, for cis and trans, through both entry points, and still
    when the workers have died (the caller then evaluates the chunks itself)."""
    from scipy.stats import binom, hypergeom, poisson
    monkeypatch.setenv("CL2_PV_WORKERS", "3")
    monkeypatch.setattr(hostops, "_pv_pool", None)
    rng = np.random.default_rng(5)
    files = [_fake_final(rng, n, N, 50_000_000) for n, N in ((1300, 8_060_000), (700, 4_400_000), (90, 1_500_000), (0, 1000))]
    try:
        for cis in (True, False):
            want = [hostops.scipy_pvalues(dict(f), f["N"], f["rpb"], cis=cis, threads=False) for f in files]
            f0 = files[0]                      # spot check of the inline path against scipy called as the reference does
            assert np.array_equal(want[0]["hyp"], np.maximum(1e-300, hypergeom.sf(f0["rab"] - 1.0, f0["N"], f0["ra"], f0["rb"])))
            assert np.array_equal(want[0]["pop"], np.maximum(1e-300, poisson.sf(f0["rab"] - 1.0, f0["mrabs"])))
            nn = f0["ra"] * f0["rb"] - f0["rab"] if cis else f0["ra"] * f0["rb"]
            assert np.array_equal(want[0]["nbp"], np.maximum(1e-300, binom.sf(f0["rab"] - 1.0, nn, f0["mbps"])))
            got = [dict(f) for f in files]
            for f in got:
                hostops.submit_pvalues(f, cis=cis)
            hostops.wait_pvalues(got)
            pooled = hostops.scipy_pvalues(dict(files[0]), files[0]["N"], files[0]["rpb"], cis=cis)
            for g, w in zip(got + [pooled], want + [want[0]]):
                for k in ("hyp", "pop", "nbp", "px", "py", "esx", "esy"):
                    assert np.array_equal(g[k], w[k]), (cis, k)
        pool = hostops.pvalue_pool()
        assert pool is not None and pool.alive == 3
        for p in pool.procs:                   # the workers die: results must not change, nothing may hang
            p.kill()
        for p in pool.procs:
            p.wait()
        got = [dict(f) for f in files]
        for f in got:
            hostops.submit_pvalues(f, cis=True)
        hostops.wait_pvalues(got)
        want = [hostops.scipy_pvalues(dict(f), f["N"], f["rpb"], cis=True, threads=False) for f in files]
        for g, w in zip(got, want):
            for k in ("hyp", "pop", "nbp", "px", "py", "esx", "esy"):
                assert np.array_equal(g[k], w[k]), k
    finally:
        pool = hostops._pv_pool
        if pool:
            pool.close()
        hostops._pv_pool = None
