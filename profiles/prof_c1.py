#!/usr/bin/env python
"""BASELINE configs[0] stand-in (C1'): chr21, 1 148 724 synthetic Hi-TrAC-like PETs, `callLoops -eps
200,500,1000 -minPts 5` with every track writer on -- the one configuration the UNMODIFIED Python reference was run on
in the build container (476 s, tests/golden/c1_cis.npz holds its outputs).  This program runs the same job on the GPU
through the drop-in callCisLoops (.ixy directory in, five text files out, byte-identical mode = scipy tail
probabilities), checks the files against the reference's byte for byte and prints the wall times.

    python profiles/prof_c1.py
"""
import logging
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import synth  # noqa: E402
from cloops2_b200.callCisLoops import callCisLoops  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "c1_cis.npz"))
tmp = tempfile.mkdtemp(prefix="cl2c1_")
try:
    xy = synth.cis_chrom(46709983, 1148724, 20261019 + 1000 * 1 + 20, "c1")
    synth.write_ixy_dir({"chr21-chr21": xy}, tmp + "/d")
    log = logging.getLogger("c1")
    log.addHandler(logging.NullHandler())
    log.propagate = False
    ts = []
    for it in range(4):
        out = tmp + "/o%d" % it
        t = time.perf_counter()
        callCisLoops(tmp + "/d", out, log, eps=[200, 500, 1000], minPts=[5], ucsc=True, juicebox=True, washU=True)
        ts.append(time.perf_counter() - t)
    same = all(open(out + sfx).read() == str(g["%s_minPts5" % tag]) for tag, sfx in
               (("loops", "_loops.txt"), ("ucsc", "_loops_ucsc.interact"), ("juice", "_loops_juicebox.txt"),
                ("washu", "_loops_legacyWashU.txt"), ("newwashu", "_loops_newWashU.txt")))
    nl = open(out + "_loops.txt").read().count("\n") - 1
    print("C1': %d PETs, %d loops; callCisLoops wall seconds per call: %s (first call creates the contexts and starts the "
          "scipy workers)" % (len(xy), nl, ", ".join("%.3f" % v for v in ts)))
    print("five output files byte-identical to the reference's: %s" % same)
    print("reference (unmodified, one process, build container): %.0f s -> %.0f x with the best call"
          % (float(g["seconds_minPts5"]), float(g["seconds_minPts5"]) / min(ts)))
    sys.exit(0 if same else 1)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
