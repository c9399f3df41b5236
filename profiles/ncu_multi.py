#!/usr/bin/env python
"""Markdown summary of an `ncu --set full --import-source on` report that holds several kernels: one headline table
(a column per captured launch, named) and, per kernel name given, the stall reasons and most-stalled SASS lines.

    python profiles/ncu_multi.py gpurun_out/a/prof_hic_full.ncu-rep "title" k_c2_core k_c2_edges ... > profiles/r2_k_dense.md
"""
import collections
import csv
import io
import re
import sys

from ncu_summary import KEYS, ncu


def main(rep, title, kernels):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(hdr)}
    names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]) for r in rows[2:]]
    print("# ncu --set full: %s\n" % title)
    print("| metric | " + " | ".join(names) + " | unit |")
    print("|---|" + "---:|" * len(names) + "---|")
    for k in KEYS:
        if k in idx:
            print("| %s | " % k + " | ".join(r[idx[k]] for r in rows[2:]) + " | %s |" % units[idx[k]])
    for kernel in kernels:
        src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--kernel-name", "regex:^%s$" % kernel))))
        blocks, cur = [], None
        for r in src:
            if len(r) >= 2 and r[0] == "Kernel Name":
                cur = {"rows": []}
                blocks.append(cur)
            elif cur is not None and r and r[0] == "Address":
                cur["hdr"] = r
            elif cur is not None and r:
                cur["rows"].append(r)
        if not blocks or "hdr" not in blocks[-1]:
            continue
        b = blocks[-1]
        h = {n: i for i, n in enumerate(b["hdr"])}
        stalls = [n for n in b["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        tot = collections.Counter()
        for r in b["rows"]:
            for s in stalls:
                try:
                    tot[s] += float(r[h[s]] or 0)
                except (ValueError, IndexError):
                    pass
        T = sum(tot.values()) or 1.0
        print("\n## %s: warp stall samples (last captured launch, %d samples, %d SASS instructions)\n" % (kernel, T, len(b["rows"])))
        print(", ".join("%s %.1f%%" % (s, 100 * v / T) for s, v in tot.most_common(6)))
        print("\n| samples | executed | SASS | top stall |")
        print("|---:|---:|---|---|")
        for r in sorted(b["rows"], key=lambda r: -float(r[h["# Samples"]] or 0))[:8]:
            st = max(((float(r[h[s]] or 0), s) for s in stalls))
            print("| %s | %s | `%s` | %s |" % (r[h["# Samples"]], r[h["Instructions Executed"]], r[h["Source"]].strip()[:70], st[1]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
