#!/usr/bin/env python
"""Timeline of one resident step of the bench workload: when each chromosome file starts / ends on which worker.

    python profiles/prof_timeline.py [pets] [threads]
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from cloops2_b200 import _lib, shard, synth  # noqa: E402
from cloops2_b200.callCisLoops import cis_chromosome  # noqa: E402

pets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
files = {}
for i, (name, length) in enumerate(synth.HG38):
    files[name] = synth.cis_chrom(length, pets * length / synth.HG38_TOTAL, 20261019 + 1000 * i, "hitrac")
ctxs = shard.worker_contexts(0, nthreads)
keys = sorted(files, key=lambda k: -len(files[k]))
owner = {k: ctxs[i % len(ctxs)] for i, k in enumerate(keys)}
res = {k: _lib.Points(owner[k], files[k]) for k in keys}
tot = sum(len(v) for v in files.values())


def step(log):
    def work(c, wi):
        for k in keys:
            if owner[k] is c:
                t0 = time.perf_counter()
                cis_chromosome(c, (k, k), None, tot, [200, 500, 1000], [5], pts=res[k])
                log.append((wi, k, t0, time.perf_counter()))
    ths = [threading.Thread(target=work, args=(c, i)) for i, c in enumerate(ctxs)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()


for r in range(4):
    log = []
    t0 = time.perf_counter()
    step(log)
    for c in ctxs:
        c.sync()
    t1 = time.perf_counter()
    print("step %d: %.1f ms" % (r, 1e3 * (t1 - t0)))
log.sort(key=lambda e: (e[0], e[2]))
for wi, k, a, b in log:
    print("worker %d %-6s %7.1f -> %7.1f ms  (%5.1f ms, %.2f M PETs)" % (wi, k, 1e3 * (a - t0), 1e3 * (b - t0), 1e3 * (b - a),
                                                                       len(files[k]) / 1e6))
print("sum of unit times %.1f ms over %d workers" % (1e3 * sum(b - a for _, _, a, b in log), len(ctxs)))
