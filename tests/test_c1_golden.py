"""C1' (SURVEY.md 8d): the chr21 stand-in for the reference's example data -- 1 148 724 synthetic Hi-TrAC-like PETs,
callLoops -eps 200,500,1000 with -minPts 5 and -minPts 10 (example/test_run/run.sh:40) and every track writer on.
Golden text from the UNMODIFIED reference (oracle/gen_golden_c1.py; ~8 + 4 minutes of its pure-Python path)."""
import logging
import os

import numpy as np
import pytest

import oracle
from cloops2_b200 import synth
from conftest import GOLD, txt_rows

C1_N, C1_SEED, CHR21 = 1148724, 20261019 + 1000 * 1 + 20, 46709983


def c1_input():
    return {"chr21-chr21": synth.cis_chrom(CHR21, C1_N, C1_SEED, "c1")}


@pytest.mark.parametrize("mp", [5, 10])
def test_oracle_reproduces_c1_loops(mp):
    g = np.load(os.path.join(GOLD, "c1_cis.npz"))
    files = c1_input()
    assert len(files["chr21-chr21"]) == int(g["n"])
    final, _ = oracle.call_cis_loops(files, int(g["n"]), [200, 500, 1000], [mp])
    assert oracle.loops_to_rows(final) == txt_rows(g["loops_minPts%d" % mp])


@pytest.mark.gpu
@pytest.mark.parametrize("mp", [5, 10])
def test_gpu_c1_files_byte_identical(gpu_ctx, tmp_path, mp):
    """callCisLoops on the .ixy directory: _loops.txt and the four track files equal the reference's byte for byte."""
    from cloops2_b200.callCisLoops import callCisLoops
    g = np.load(os.path.join(GOLD, "c1_cis.npz"))
    d = str(tmp_path / "d")
    synth.write_ixy_dir(c1_input(), d)
    out = str(tmp_path / "o")
    callCisLoops(d, out, logging.getLogger("t"), eps=[200, 500, 1000], minPts=[mp], ucsc=True, juicebox=True, washU=True)
    for tag, suffix in (("loops", "_loops.txt"), ("ucsc", "_loops_ucsc.interact"), ("juice", "_loops_juicebox.txt"),
                        ("washu", "_loops_legacyWashU.txt"), ("newwashu", "_loops_newWashU.txt")):
        assert open(out + suffix).read() == str(g["%s_minPts%d" % (tag, mp)]), suffix


@pytest.mark.gpu
def test_gpu_trans_writer_files_byte_identical(gpu_ctx, tmp_path):
    """callTransLoops with -w -j: the washU and juicebox texts (callTransLoops.py:271-275) equal the reference's."""
    import json
    from cloops2_b200.callTransLoops import callTransLoops
    from conftest import golden_trans_files
    gt, tfiles = golden_trans_files()
    w = np.load(os.path.join(GOLD, "trans_writers.npz"))
    d = str(tmp_path / "d")
    synth.write_ixy_dir(tfiles, d)
    meta = json.load(open(d + "/petMeta.json"))
    meta["data"]["trans"] = {k: meta["data"]["trans"][k] for k in gt["order"].tolist()}
    json.dump(meta, open(d + "/petMeta.json", "w"))
    out = str(tmp_path / "o")
    callTransLoops(d, out, logging.getLogger("t"), eps=[2000, 4000], minPts=[10, 5], washU=True, juicebox=True)
    assert open(out + "_trans_loops.txt").read() == str(gt["trans_txt"])
    assert open(out + "_trans_loops_washU.txt").read() == str(w["washu"])
    assert open(out + "_trans_loops_juicebox.txt").read() == str(w["juice"])
