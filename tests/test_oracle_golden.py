"""The CPU oracle against the golden vectors produced by the unmodified reference (oracle/gen_golden.py)."""
import numpy as np

import oracle
from conftest import load_golden, golden_cis_files, golden_trans_files, txt_rows


def test_dbscan_labels_match_reference_vectors():
    g = load_golden("dbscan_small.npz")
    offs = g["offsets"]
    for i, (eps, minPts) in enumerate(g["params"].tolist()):
        xy = g["xy"][offs[i]:offs[i + 1]]
        lb, _ = oracle.block_dbscan(xy, eps, minPts)
        lc, _ = oracle.cdbscan2(xy, eps, minPts)
        assert np.array_equal(lb, g["labels_block"][offs[i]:offs[i + 1]]), ("blockDBSCAN", i, eps, minPts)
        assert np.array_equal(lc, g["labels_c2"][offs[i]:offs[i + 1]]), ("cDBSCAN2", i, eps, minPts)


def test_cis_pipeline_rows_match_reference_txt():
    g, files = golden_cis_files()
    tot = int(g["tot"])
    for name, kw in (("default_txt", dict(eps=[200, 500, 1000], minPts=[5])),
                     ("hic_txt", dict(eps=[200, 500, 1000], minPts=[5], hic=True)),
                     ("cut_txt", dict(eps=[300, 800], minPts=[8, 4], cut=2000, mcut=1500000)),
                     ("empair_txt", dict(eps=[800, 300], minPts=[4, 8], emPair=True))):
        final, _ = oracle.call_cis_loops(files, tot, **kw)
        assert oracle.loops_to_rows(final) == txt_rows(g[name]), name


def test_trans_pipeline_rows_match_reference_txt():
    g, files = golden_trans_files()
    final = oracle.call_trans_loops(files, eps=[2000, 4000], minPts=[10, 5])
    assert oracle.loops_to_rows(final) == txt_rows(g["trans_txt"])


def test_estsig_fields_match_reference():
    e = load_golden("estsig_cis.npz")
    g, files = golden_cis_files()
    xy = files["chr21-chr21"]
    loops = []
    for c in e["candidates"]:
        lp = oracle.OLoop()
        lp.chromX = lp.chromY = "chr21"
        lp.x_start, lp.x_end, lp.y_start, lp.y_end = (int(v) for v in c)
        loops.append(lp)
    _, out = oracle.est_loop_sig_cis("chr21-chr21", loops, xy, int(e["tot"]), minPts=5)
    kept = np.array([[lp.x_start, lp.x_end, lp.y_start, lp.y_end] for lp in out], dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(kept, e["kept"])
    vals = np.array([[float(getattr(lp, f)) for f in e["fields"].tolist()] for lp in out])
    assert np.array_equal(vals, e["values"])          # bit-exact: same scipy, same expressions


def test_empty_and_tiny_inputs():
    z = np.zeros((0, 2), dtype=np.int64)
    for fn in (oracle.block_dbscan, oracle.cdbscan2):
        lab, n = fn(z, 10, 3)
        assert lab.shape == (0,) and n == 0
        lab, n = fn(np.array([[5, 9]], dtype=np.int64), 10, 1)
        assert n == 1 and lab.tolist() == [0]
        lab, n = fn(np.array([[5, 9]], dtype=np.int64), 10, 2)
        assert n == 0 and lab.tolist() == [-1]
