"""Command-line surface of `callLoops` (cLoops2.py:60-85, 1431-1548) -- parsing only, no GPU."""
from cloops2_b200 import cli


def test_eps_minpts_ordering_rules():
    assert cli.parseEps("1000,200,500") == [200, 500, 1000]           # ascending
    assert cli.parseMinpts("5,20,10") == [20, 10, 5]                  # descending
    assert cli.parseEps("1000,200", emPair=True) == [1000, 200]       # -emPair keeps the given order
    assert cli.parseMinpts("5,20", emPair=True) == [5, 20]
    assert cli.parseEps(500) == [500] and cli.parseMinpts("7") == [7]


def test_flag_surface_matches_reference():
    op = cli.buildParser().parse_args(["callLoops", "-d", "D", "-o", "O", "-eps", "200,500", "-minPts", "5", "-p", "4",
                                       "-cut", "1000", "-mcut", "2000000", "-hic", "-trans", "-emPair", "-filter", "-i",
                                       "-j", "-w", "-plot", "-max_cut"])
    assert (op.predir, op.fnOut, op.cpu, op.cut, op.mcut) == ("D", "O", 4, 1000, 2000000)
    assert all([op.hic, op.trans, op.emPair, op.filterPETs, op.ucsc, op.juicebox, op.washU, op.plot, op.max_cut])
    dflt = cli.buildParser().parse_args(["callLoops"])
    assert (dflt.eps, dflt.minPts, dflt.cut, dflt.mcut, dflt.cpu, dflt.fnOut) == (0, 0, 0, -1, 1, "cLoops2_output")


def test_pre_and_quant_flag_surface():
    """`cLoops2 pre` (cLoops2.py:300-360, 2914-2975) and `quant` options and defaults."""
    op = cli.buildParser().parse_args(["pre", "-f", "a.bedpe.gz,b.bedpe", "-o", "trac", "-c", "chr21,chr22", "-mapq", "0",
                                       "-cut", "1000", "-mcut", "500000", "-trans", "-format", "pairs", "-p", "8"])
    assert (op.cmd, op.fnIn, op.fnOut, op.chroms, op.mapq, op.cut, op.mcut, op.trans, op.format, op.cpu) == (
        "pre", "a.bedpe.gz,b.bedpe", "trac", "chr21,chr22", 0, 1000, 500000, True, "pairs", 8)
    dflt = cli.buildParser().parse_args(["pre"])
    assert (dflt.mapq, dflt.cut, dflt.mcut, dflt.trans, dflt.format, dflt.chroms) == (10, 0, -1, False, "bedpe", "")
    q = cli.buildParser().parse_args(["quant", "-d", "D", "-loops", "L.txt", "-o", "Q", "-offp"])
    assert (q.cmd, q.predir, q.loops, q.fnOut, q.offp) == ("quant", "D", "L.txt", "Q", True)


def test_pre_refuses_missing_input_and_non_empty_output(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                # the CLI appends to ./cLoops2.log like the reference: keep it out of the checkout
    out = tmp_path / "out"
    out.mkdir()
    (out / "something").write_text("x")
    assert cli.main(["pre", "-f", str(tmp_path / "missing.bedpe"), "-o", str(out)]) == 1      # non-empty directory
    assert cli.main(["pre", "-f", str(tmp_path / "missing.bedpe"), "-o", str(tmp_path / "new")]) == 1   # no input file
