"""
TEST INFRASTRUCTURE -- C1' golden: the UNMODIFIED reference's `callLoops` on the chr21 stand-in for
example/data (1 148 724 synthetic Hi-TrAC-like PETs, seed 20261019 + 1000*1 + 20), eps 200,500,1000, minPts 5 and
minPts 10 (example/test_run/run.sh:40), with every track writer switched on (-w -j and the ucsc file), plus the trans
writers (callTransLoops.py:271-275) on the small trans golden input.

    python -m oracle.gen_golden_c1          # ~10-15 min of the pure-Python reference; rewrites tests/golden/c1_*.npz

The input is regenerated from the seed by the tests (cloops2_b200.synth.cis_chrom), only the reference's output text
is stored.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from cloops2_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
C1_N = 1148724
C1_SEED = 20261019 + 1000 * 1 + 20
CHR21 = 46709983


def c1_input():
    return {"chr21-chr21": synth.cis_chrom(CHR21, C1_N, C1_SEED, "c1")}


def main():
    ref = ref_shim.load()
    files = c1_input()
    out = {"n": int(len(files["chr21-chr21"]))}
    for mp in (5, 10):
        tmp = tempfile.mkdtemp()
        synth.write_ixy_dir(files, tmp + "/d")
        t = time.time()
        ref.cis.callCisLoops(tmp + "/d", tmp + "/out", ref.logger, eps=[200, 500, 1000], minPts=[mp], cpu=1,
                             ucsc=True, juicebox=True, washU=True)
        out["seconds_minPts%d" % mp] = time.time() - t
        for tag, suffix in (("loops", "_loops.txt"), ("ucsc", "_loops_ucsc.interact"), ("juice", "_loops_juicebox.txt"),
                            ("washu", "_loops_legacyWashU.txt"), ("newwashu", "_loops_newWashU.txt")):
            out["%s_minPts%d" % (tag, mp)] = open(tmp + "/out" + suffix).read()
        print("minPts", mp, "done in %.0f s" % out["seconds_minPts%d" % mp], flush=True)
    np.savez_compressed(os.path.join(GOLD, "c1_cis.npz"), **out)
    # trans writers on the committed trans golden input
    g = np.load(os.path.join(GOLD, "trans_pipeline.npz"))
    tfiles = {k: g[k.replace("-", "_")] for k in g["order"].tolist()}
    tmp = tempfile.mkdtemp()
    synth.write_ixy_dir(tfiles, tmp + "/d")
    ref.trans.callTransLoops(tmp + "/d", tmp + "/out", ref.logger, eps=[2000, 4000], minPts=[10, 5], cpu=1, washU=True,
                             juicebox=True)
    assert open(tmp + "/out_trans_loops.txt").read() == str(g["trans_txt"])
    np.savez_compressed(os.path.join(GOLD, "trans_writers.npz"),
                        washu=open(tmp + "/out_trans_loops_washU.txt").read(),
                        juice=open(tmp + "/out_trans_loops_juicebox.txt").read())
    print("written")


if __name__ == "__main__":
    main()
