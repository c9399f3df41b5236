"""
Host-side logic of the product (vectorised numpy + native selection) against the oracle's plain-Python restatement.
The device index is replaced by a stand-in that answers loop_counts from the oracle's C counting, so these run
without a GPU; the GPU tests exercise the same code with the real kernels.
"""
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


def _rows(loops):
    return oracle.loops_to_rows(loops)


def test_est_cis_matches_oracle_and_reference_fields():
    e = load_golden("estsig_cis.npz")
    _, files = golden_cis_files()
    xy = files["chr21-chr21"]
    span = float(xy[:, 1].max() - xy[:, 0].min())
    cols = hostops.est_cis(FakeIndex(xy), e["candidates"], len(xy), span, int(e["tot"]), 5)
    assert np.array_equal(cols["coords"], e["kept"])
    names = dict(ra="ra", rb="rb", rab="rab", P2LL="P2LL", FDR="FDR", ES="ES", density="density",
                 hypergeometric_p_value="hyp", poisson_p_value="pop", binomial_p_value="nbp",
                 x_peak_poisson_p_value="px", x_peak_es="esx", y_peak_poisson_p_value="py", y_peak_es="esy")
    for j, f in enumerate(e["fields"].tolist()):
        assert np.array_equal(np.asarray(cols[names[f]], dtype=np.float64), e["values"][:, j]), f


def test_full_cis_host_path_reproduces_reference_txt():
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
            cands = hostops.combine_unique(per_pass)
            cands = cands[hostops.cis_distance(cands) > 0]
            span = float(xy[:, 1].max() - xy[:, 0].min())
            cols = hostops.est_cis(FakeIndex(xy), cands, len(xy), span, tot, max(minPts), hic=hic)
            final = hostops.select_twice(cols, hostops.mark_sig_cis(cols, hic=hic))
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
        final = hostops.select_twice(cols, hostops.mark_sig_trans(cols))
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
