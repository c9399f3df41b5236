"""
`cLoops2 callLoops` command line (cLoops2/cLoops2.py:1431-1548 parser, :3639-3734 dispatch, :60-85 eps / minPts
parsing) for the B200 path.  Flags, defaults, ordering rules and log messages follow the reference; the other 21
sub-commands of cLoops2 are out of scope.  Run as `python -m cloops2_b200.cli callLoops -d DIR -o OUT -eps ... -minPts ...`
(under torchrun for several GPUs).
"""
import argparse
import logging
import os
import sys
from datetime import datetime

from . import shard


def parseEps(eps, emPair=False):
    """cLoops2.py:60-71: comma list -> ints, ascending unless -emPair."""
    eps = list(map(int, eps.split(","))) if "," in str(eps) else [int(eps)]
    if not emPair:
        eps.sort()
    return eps


def parseMinpts(minPts, emPair=False):
    """cLoops2.py:74-85: comma list -> ints, descending unless -emPair."""
    minPts = list(map(int, minPts.split(","))) if "," in str(minPts) else [int(minPts)]
    if not emPair:
        minPts.sort(reverse=True)
    return minPts


def getLogger(fn="cLoops2.log"):
    """File + stdout logger like cLoops2/utils.py:71-99."""
    log = logging.getLogger("cLoops2")
    if not log.handlers:
        log.setLevel(logging.INFO)
        fmt = logging.Formatter("%(asctime)s %(name)-6s %(levelname)-8s %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
        for h in (logging.FileHandler(fn), logging.StreamHandler(sys.stdout)):
            h.setFormatter(fmt)
            log.addHandler(h)
    return log


def buildParser():
    p = argparse.ArgumentParser(prog="cLoops2", description="cLoops2 callLoops on B200 (cl2-b200)")
    sub = p.add_subparsers(dest="cmd")
    c = sub.add_parser("callLoops", help="Call intra- (and with -trans inter-) chromosomal loops.")
    # shared flags, cLoops2.py:165-235
    c.add_argument("-d", dest="predir", type=str, default="")
    c.add_argument("-o", dest="fnOut", type=str, default="cLoops2_output")
    c.add_argument("-p", dest="cpu", type=int, default=1)
    c.add_argument("-cut", dest="cut", type=int, default=0)
    c.add_argument("-mcut", dest="mcut", type=int, default=-1)
    # callLoops flags, cLoops2.py:1439-1548
    c.add_argument("-eps", dest="eps", default=0)
    c.add_argument("-minPts", dest="minPts", default=0)
    c.add_argument("-plot", dest="plot", action="store_true", default=False)
    c.add_argument("-i", dest="ucsc", action="store_true", default=False)
    c.add_argument("-j", dest="juicebox", action="store_true", default=False)
    c.add_argument("-w", dest="washU", action="store_true", default=False)
    c.add_argument("-max_cut", dest="max_cut", action="store_true", default=False)
    c.add_argument("-hic", dest="hic", action="store_true", default=False)
    c.add_argument("-filter", dest="filterPETs", action="store_true", default=False)
    c.add_argument("-trans", dest="trans", action="store_true", default=False)
    c.add_argument("-emPair", dest="emPair", action="store_true", default=False)
    # quant -loops (cLoops2.py quant sub-command; only the loop part is implemented here)
    r = sub.add_parser("pre", help="BEDPE(.gz) files -> per chromosome pair .ixy files + petMeta.json.")
    r.add_argument("-f", dest="fnIn", type=str, default="")
    r.add_argument("-o", dest="fnOut", type=str, default="cLoops2_output")
    r.add_argument("-p", dest="cpu", type=int, default=1)
    r.add_argument("-c", dest="chroms", type=str, default="")
    r.add_argument("-cut", dest="cut", type=int, default=0)
    r.add_argument("-mcut", dest="mcut", type=int, default=-1)
    r.add_argument("-mapq", dest="mapq", type=int, default=10)
    r.add_argument("-trans", dest="trans", action="store_true", default=False)
    r.add_argument("-format", dest="format", type=str, default="bedpe", choices=["bedpe", "pairs"])

    q = sub.add_parser("quant", help="Quantify given loops against a sample (quant -loops).")
    q.add_argument("-d", dest="predir", type=str, default="")
    q.add_argument("-o", dest="fnOut", type=str, default="cLoops2_output")
    q.add_argument("-p", dest="cpu", type=int, default=1)
    q.add_argument("-cut", dest="cut", type=int, default=0)
    q.add_argument("-mcut", dest="mcut", type=int, default=-1)
    q.add_argument("-loops", dest="loops", type=str, default="")
    q.add_argument("-offp", dest="offp", action="store_true", default=False)
    return p


def main(argv=None):
    from .callCisLoops import callCisLoops
    from .callTransLoops import callTransLoops
    op = buildParser().parse_args(argv)
    # keep the collector away from the objects numpy / scipy / torch created at import (a full pass over them costs
    # tens of milliseconds in the middle of the run)
    import gc
    gc.freeze()
    if op.cmd == "pre":                                        # cLoops2.py:2914-2975
        from .pre import parseBedpe, parsePairs
        logger = getLogger()
        if not os.path.isdir(op.fnOut):
            os.mkdir(op.fnOut)
        elif len(os.listdir(op.fnOut)) > 0:
            logger.error("working directory %s exists and not empty." % op.fnOut)
            return 1
        fs = op.fnIn.split(",")
        for f in fs:
            if not os.path.isfile(f):
                logger.error("Input file %s not exitst!" % f)
                return 1
        cs = set(op.chroms.split(",")) if op.chroms else []
        if op.format == "pairs":
            parsePairs(fs, op.fnOut, logger, cs=cs, cut=op.cut, mcut=op.mcut, cis=not op.trans, cpu=op.cpu)
        else:
            parseBedpe(fs, op.fnOut, logger, mapq=op.mapq, cs=cs, cut=op.cut, mcut=op.mcut, cis=not op.trans, cpu=op.cpu)
        return 0
    if op.cmd == "quant":
        from .quant import quantLoops
        logger = getLogger()
        if op.loops == "" or not os.path.isfile(op.loops) or not os.path.exists(op.predir):
            logger.error("quant needs -d DIR and -loops FILE. Return.")
            return 1
        quantLoops(op.predir, op.loops, op.fnOut, logger, cut=op.cut, mcut=op.mcut, cpu=op.cpu, offp=op.offp)
        return 0
    if op.cmd != "callLoops":
        buildParser().print_help()
        return 1
    rank, world = shard.init_from_env()
    logger = getLogger() if rank == 0 else logging.getLogger("cLoops2.rank%d" % rank)
    start = datetime.now()
    logger.info("Command: cLoops2 {} -d {} -eps {} -minPts {} -p {} -o {} -cut {} -mcut {} -filter {} -i {} -j {} -w {} -hic {} -max_cut {} -trans {} -emPair {}".format(
        op.cmd, op.predir, op.eps, op.minPts, op.cpu, op.fnOut, op.cut, op.mcut, op.filterPETs, op.ucsc, op.juicebox,
        op.washU, op.hic, op.max_cut, op.trans, op.emPair))
    if op.predir == "":
        logger.error("No -d assigned! Return.")
        return 1
    if not os.path.exists(op.predir):
        logger.error("%s not exists! Return." % op.predir)
        return 1
    eps = parseEps(op.eps, op.emPair)
    if len(eps) == 1 and eps[0] == 0:
        logger.error("Input eps is 0. Return.")
        return 1
    minPts = parseMinpts(op.minPts, op.emPair)
    if len(minPts) == 1 and minPts[0] == 0:
        logger.error("Input minPts is 0. Return.")
        return 1
    if os.path.isfile(op.fnOut + "_loop.txt"):          # the reference checks this name (cLoops2.py:3683)
        logger.error("Output file %s exists, return." % (op.fnOut + "_loop.txt"))
        return 1
    callCisLoops(op.predir, op.fnOut, logger, eps=eps, minPts=minPts, cpu=op.cpu, cut=op.cut, mcut=op.mcut,
                 plot=op.plot, max_cut=op.max_cut, hic=op.hic, filter=op.filterPETs, ucsc=op.ucsc,
                 juicebox=op.juicebox, washU=op.washU, emPair=op.emPair)
    if op.trans:
        if os.path.isfile(op.fnOut + "_trans_loop.txt"):
            logger.error("Output file %s exists, return." % (op.fnOut + "_trans_loop.txt"))
            return 1
        callTransLoops(op.predir, op.fnOut, logger, eps=eps, minPts=minPts, cpu=op.cpu, filter=op.filterPETs,
                       washU=op.washU, juicebox=op.juicebox)
    logger.info("cLoops2 %s finished. Used time: %s." % (op.cmd, datetime.now() - start) + "\n" * 3)
    return 0


if __name__ == "__main__":
    sys.exit(main())
