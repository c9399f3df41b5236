#!/usr/bin/env python
"""Host side of `cLoops2 pre` on this machine's CPU (no GPU needed): the native BEDPE tokeniser + the numpy PET
arithmetic and filters of cloops2_b200.pre (collect_pets) in lines per second, and -- when the reference checkout is
present (build container only) -- the unmodified reference's parseBedpe on the same file (text parse, per-pair text
files, set de-duplication, .ixy writing: its whole `pre`).  The GPU de-duplication (cl2_points_unique) and the .ixy
writer of this package are not part of the timed region here.

    python profiles/prof_pre_host.py [lines, default 1000000] > profiles/r2_pre_host.md
"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloops2_b200 import pre, synth  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(7)
chroms = [c for c, _ in synth.HG38[:6]]
lens = dict(synth.HG38)
tmp = tempfile.mkdtemp(prefix="cl2pre_")
try:
    f = os.path.join(tmp, "pets.bedpe")
    ca = rng.integers(0, len(chroms), size=n)
    cb = np.where(rng.random(n) < 0.9, ca, rng.integers(0, len(chroms), size=n))
    with open(f, "w") as fh:
        for i in range(n):
            a, b = chroms[ca[i]], chroms[cb[i]]
            x = int(rng.integers(0, lens[a] - 2000))
            y = int(rng.integers(0, lens[b] - 2000))
            fh.write("%s\t%d\t%d\t%s\t%d\t%d\tr%d\t%d\t+\t-\n" % (a, x, x + 50, b, y, y + 50, i, int(rng.integers(0, 61))))
    size = os.path.getsize(f)
    print("# host side of `pre`: %d BEDPE lines (%.1f MB), %d host cores\n" % (n, size / 1e6, len(os.sched_getaffinity(0))))
    print("| what | seconds | lines/s |")
    print("|---|---:|---:|")
    for cpu in (1, min(8, len(os.sched_getaffinity(0)))):
        t = time.time()
        pets, cnt = pre.collect_pets([f], mapq=10, cut=0, mcut=-1, cis=False, fmt="bedpe", cpu=cpu)[:2]
        dt = time.time() - t
        kept = int(sum(len(v) for v in pets.values())) if hasattr(pets, "values") else -1
        print("| cloops2_b200.pre.collect_pets (native tokeniser + numpy filters), -p %d | %.2f | %.2f M |" % (cpu, dt, n / dt / 1e6))
    sys.path.insert(0, os.path.join(ROOT))
    from oracle import ref_shim
    if ref_shim.available():
        ns = ref_shim.load()
        import logging
        log = logging.getLogger("pre_ref")
        log.addHandler(logging.NullHandler())
        log.propagate = False
        from cLoops2 import io as rio
        out = os.path.join(tmp, "ref_out")
        os.mkdir(out)                         # the reference's command line creates it (cLoops2.py pre)
        import contextlib
        t = time.time()
        with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):      # it prints a progress line per 100 k PETs
            rio.parseBedpe([f], out, log, mapq=10, cs=[], cut=0, mcut=-1, cis=False, cpu=1)
        dt = time.time() - t
        print("| reference cLoops2.io.parseBedpe (whole `pre`, 1 process) | %.2f | %.3f M |" % (dt, n / dt / 1e6))
    else:
        print("| reference | not present on this machine | |")
finally:
    shutil.rmtree(tmp, ignore_errors=True)
