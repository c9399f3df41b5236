#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python profiles/summarize.py gpurun_out/launches_r1.csv > profiles/r1_launches_summary.md
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        rows.append(dict(zip(header, r)))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values()) or 1.0
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("| %s | %d | %.3f | %.1f%% | %.1f |" % (name, cnt[name], ms, 100 * ms / total, 1e3 * ms / cnt[name]))
    print("\ntotal: %d launches, %.3f ms of kernel time (cold-cache, serialised under ncu: compare shares, not absolutes)"
          % (sum(cnt.values()), total))


if __name__ == "__main__":
    main(sys.argv[1])
