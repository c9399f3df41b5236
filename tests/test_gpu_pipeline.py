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
        final, st = cis_chromosome(gpu_ctx, tuple(k.split("-")), xy, int(g["tot"]), **kw)
        loops.extend(cols_to_loops(tuple(k.split("-")), final))
    assert _rows(loops) == txt_rows(g[name])


def test_trans_golden_rows(gpu_ctx):
    g, files = golden_trans_files()
    loops = []
    for k, xy in files.items():
        final, st = trans_pair(gpu_ctx, tuple(k.split("-")), xy, [2000, 4000], [10, 5])
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
        final, st = cis_chromosome(gpu_ctx, tuple(k.split("-")), xy, tot, [200, 500, 1000], [5])
        loops.extend(cols_to_loops(tuple(k.split("-")), final))
    assert len(want) > 20
    assert _rows(loops) == _rows(want)
