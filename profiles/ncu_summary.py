#!/usr/bin/env python
"""Markdown summary of an `ncu --set full --import-source on` report: headline metrics per captured launch, stall
reasons and the most-stalled SASS lines of the last launch.

    python profiles/ncu_summary.py gpurun_out/prof_est_r1.ncu-rep k_est_loops > profiles/r1_k_est_loops.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main(rep, kernel):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(hdr)}
    print("# ncu --set full: %s\n" % kernel)
    print("| metric | " + " | ".join("launch %d" % i for i in range(len(rows) - 2)) + " | unit |")
    print("|---|" + "---:|" * (len(rows) - 2) + "---|")
    for k in KEYS:
        if k in idx:
            print("| %s | " % k + " | ".join(r[idx[k]] for r in rows[2:]) + " | %s |" % units[idx[k]])
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--kernel-name", kernel))))
    blocks, cur = [], None
    for r in src:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    if not blocks:
        return
    b = blocks[-1]
    h = {n: i for i, n in enumerate(b["hdr"])}
    stalls = [n for n in b["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    for r in b["rows"]:
        for s in stalls:
            try:
                tot[s] += float(r[h[s]] or 0)
            except ValueError:
                pass
    T = sum(tot.values()) or 1.0
    print("\n## warp stall samples (last captured launch, %d samples, %d SASS instructions)\n" % (T, len(b["rows"])))
    print(", ".join("%s %.1f%%" % (s, 100 * v / T) for s, v in tot.most_common(8)))
    print("\n## most-stalled SASS lines\n")
    print("| samples | executed | SASS | top stall |")
    print("|---:|---:|---|---|")
    for r in sorted(b["rows"], key=lambda r: -float(r[h["# Samples"]] or 0))[:15]:
        st = max(((float(r[h[s]] or 0), s) for s in stalls))
        print("| %s | %s | `%s` | %s |" % (r[h["# Samples"]], r[h["Instructions Executed"]], r[h["Source"]].strip()[:70], st[1]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
