#!/usr/bin/env python
"""Markdown tables of DESIGN.md section 5 from the bench lines kept under profiles/ (one JSON line per file).

    python profiles/measure_table.py > /tmp/tables.md
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return None
    for line in reversed(open(p).read().strip().splitlines()):
        try:
            return json.loads(line)
        except ValueError:
            continue
    return None


def g(v, nd=2):
    return "-" if v is None else ("%.*f" % (nd, v / 1e9))


def main(files):
    print("| file | config | GPUs | scaling | PETs in the job | resident G PETs/s (ms/step) | e2e G PETs/s (ms/step) | host CPU ms/step | roofline family: GB/s (frac) |")
    print("|---|---|---:|---|---:|---:|---:|---:|---|")
    for f in files:
        d = load(f)
        if not d:
            continue
        e = d.get("e2e") or {}
        r = d.get("roofline") or {}
        roof = "%s: %.0f (%.2f)" % (r.get("kernel"), r.get("achieved") or 0, r.get("frac") or 0) if r.get("kernel") else "-"
        print("| `%s` | %s | %d | %s | %.1f M | %s (%.1f) | %s (%.1f) | %s | %s |" % (
            f, d["config"].get("config"), d["n_gpus"], d.get("scaling"), d["job"]["pets"] / 1e6, g(d["value"]), d["ms_per_step"],
            g(e.get("value")), e.get("ms_per_step") or 0, d.get("host_cpu_ms_per_step"), roof))


def families(f):
    d = load(f)
    if not d or not d.get("kernels"):
        return
    print("\n| family | ms/step | share | algorithmic GB/s | of %.0f GB/s |" % d["roofline"]["peak"])
    print("|---|---:|---:|---:|---:|")
    for nm, v in d["kernels"].items():
        gb = v.get("GBps")
        print("| %s | %.2f | %.1f %% | %s | %s |" % (nm, v["ms_per_step"], 100 * v["share"], "-" if not gb else "%.0f" % gb,
                                                  "-" if not gb else "%.2f" % (gb / d["roofline"]["peak"])))
    print("\nsum of the families: %.1f ms; step (8 streams): %.1f ms" % (sum(v["ms_per_step"] for v in d["kernels"].values()), d["ms_per_step"]))


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--families":
        families(args[1])
    else:
        main(args or sorted(f for f in os.listdir(HERE) if f.startswith("r2") and f.endswith(".json") and "bench" in f))
