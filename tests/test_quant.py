"""`quant -loops` (first "next" row): oracle restatement vs the reference's golden output (CPU), and
the GPU implementation vs both."""
import json
import logging

import numpy as np
import pytest

import oracle
from conftest import golden_cis_files, load_golden, txt_rows


def _parse(txt):
    """{key: [OLoop]} from a _loops.txt (cLoops2/io.py:710-740), file order kept per chromosome."""
    loops = {}
    for line in str(txt).rstrip("\n").split("\n")[1:]:
        f = line.split("\t")
        lp = oracle.OLoop()
        lp.id, lp.chromX, lp.chromY = f[0], f[1], f[4]
        lp.x_start, lp.x_end, lp.y_start, lp.y_end = (int(float(v)) for v in (f[2], f[3], f[5], f[6]))
        lp.distance = int(float(f[7]))
        lp.x_center, lp.y_center = (lp.x_start + lp.x_end) / 2, (lp.y_start + lp.y_end) / 2
        lp.cis = lp.chromX == lp.chromY
        loops.setdefault(lp.chromX + "-" + lp.chromY, []).append(lp)
    return loops


def _by_key(rows):
    out = {}
    for r in rows:
        out.setdefault(r[1] + "-" + r[4], []).append(r)
    return out


@pytest.mark.parametrize("offp", [False, True])
def test_oracle_quant_matches_reference(offp):
    q = load_golden("quant_loops.npz")
    g, files = golden_cis_files()
    want = _by_key(txt_rows(q["quant_offp%d" % offp]))
    for key, loops in _parse(q["in_loops"]).items():
        _, out = oracle.quant_loops(key, loops, files[key], int(g["tot"]), offp=offp)
        assert oracle.quant_rows(out) == want[key], key      # the reference orders chromosomes by set iteration


@pytest.mark.gpu
@pytest.mark.parametrize("offp", [False, True])
def test_gpu_quant_matches_reference(gpu_ctx, tmp_path, offp):
    from cloops2_b200 import synth
    from cloops2_b200.quant import quantLoops
    q = load_golden("quant_loops.npz")
    g, files = golden_cis_files()
    d = str(tmp_path / "d")
    synth.write_ixy_dir(files, d)
    meta = json.load(open(d + "/petMeta.json"))
    meta["Unique PETs"] = int(g["tot"])
    json.dump(meta, open(d + "/petMeta.json", "w"))
    fin = str(tmp_path / "in_loops.txt")
    open(fin, "w").write(str(q["in_loops"]))
    out = str(tmp_path / "q")
    quantLoops(d, fin, out, logging.getLogger("t"), offp=offp)
    got = _by_key(txt_rows(open(out + "_loops.txt").read()))
    want = _by_key(txt_rows(q["quant_offp%d" % offp]))
    assert got == want
