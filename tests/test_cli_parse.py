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
