#!/usr/bin/env python
"""
bench.py -- PETs clustered+tested per second for `callLoops` (BASELINE.json metric) on N B200s of one node.

Default workload (BASELINE.json configs[1], `--config c2`, named in config.workload): synthetic Hi-TrAC-like
intra-chromosomal hg38 PETs, 100 M per GPU, callLoops -eps 200,500,1000 -minPts 5 (blockDBSCAN), chromosome files
sharded over the ranks by PET count (weak scaling: at N GPUs the job is N synthetic genomes = 24*N files;
`--scaling strong` keeps the job at one genome).  One "step" = the whole hot path over the rank's files: the
clustering runs + candidate boxes, XY index, per-candidate counting, hypergeometric / Poisson / binomial tests,
significance marking and overlap selection -- everything `callCisLoops` does between reading the .ixy arrays and
writing the text file.  Other BASELINE configs: `--config c3` / `c3hic` (1 B Hi-C-like PETs, eps 5000,10000, minPts
20,50, blockDBSCAN / cDBSCAN2; the genome is cut into 8 PET-balanced shards and rank r processes shard r, so
`--gpus 8` is the whole job) and `--config c4` (200 M inter-chromosomal PETs over 276 pair files, callTransLoops).

  value  whole-job PETs/s with the PET arrays already resident in HBM (cl2_points) when the timed region starts
  e2e    the same metric through the public per-file call with HOST (pinned) int64 arrays: the upload of every
         file and the read-back of the surviving loops are inside the timed region
  e2e_files  (c2, N=1) file to file through callCisLoops(): .ixy directory in, _loops.txt out
  roofline  dominant kernel family by device time (CUDA events on the library's stream, one instrumented pass):
         algorithmic bytes per launch / average launch time against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (oracle/, single thread) on a bounded sample of the same workload (rank 0, N=1); its
         loops are compared with the GPU arm's for the same files (parity_check)

`--impl reference` times the CPU implementation of the path (the oracle port of the pure-Python reference, which
cannot travel to the GPU box) on all host cores, one process per file like the reference's joblib fan-out, on a
sample sized by a calibration pass so that warm-up + steps finish within ~3 minutes whatever the core count.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c2": dict(kind="cis", profile="hitrac", eps=[200, 500, 1000], minPts=[5], hic=False, pets=100e6, cid=2,
               workload="BASELINE configs[1]: synthetic Hi-TrAC-like intra-chromosomal hg38 PETs, callLoops "
                        "-eps 200,500,1000 -minPts 5 (blockDBSCAN), chromosome files sharded by PET count"),
    "c3": dict(kind="cis", profile="hic", eps=[5000, 10000], minPts=[50, 20], hic=False, pets=1e9, cid=3, shards=8,
               workload="BASELINE configs[2]: synthetic Hi-C-like 1 B intra-chromosomal hg38 PETs, callLoops "
                        "-eps 5000,10000 -minPts 20,50 (blockDBSCAN), 8 PET-balanced shards, rank r = shard r"),
    "c3hic": dict(kind="cis", profile="hic", eps=[5000, 10000], minPts=[50, 20], hic=True, pets=1e9, cid=3, shards=8,
                  workload="BASELINE configs[2] with -hic: synthetic Hi-C-like 1 B intra-chromosomal hg38 PETs, callLoops "
                           "-eps 5000,10000 -minPts 20,50 -hic (cDBSCAN2), 8 PET-balanced shards, rank r = shard r"),
    "c4": dict(kind="trans", eps=[5000], minPts=[20, 10], hic=False, pets=200e6, cid=4,
               workload="BASELINE configs[3]: synthetic inter-chromosomal 200 M PETs over the 276 hg38 chromosome pairs, "
                        "callTransLoops -eps 5000 -minPts 10,20, pair files sharded by PET count"),
}
CFG = CONFIGS["c2"]
METRIC = "PETs clustered+tested/sec (callLoops)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--pets", type=float, default=None, help="PETs per GPU (c2, c4) or of the whole genome (c3); "
                                                             "default: the figure of the BASELINE config")
    ap.add_argument("--cpu-sample-pets", type=float, default=5.5e6)
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="time the reference arm aims to spend in its passes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true", help="skip the instrumented single-stream pass")
    ap.add_argument("--no-exact", action="store_true", help="skip the e2e_exact_scipy leg")
    ap.add_argument("--no-files", action="store_true", help="skip the e2e_files leg")
    ap.add_argument("--extra-sync-modes", default="",
                    help="comma list of wait modes (0 spin, 1 sleep, 2 poll + yield) to time the resident arm with once "
                         "more each, after the default one (reported as sync_modes_ms_per_step)")
    ap.add_argument("--host-cores", type=int, default=0,
                    help="confine the process to this many host cores once the workload is generated (emulates the "
                         "per-rank share of a multi-GPU node, e.g. 4 = 8 ranks on 32 cores); 0 = leave the affinity alone")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- workload
def _gen_one(job):
    from cloops2_b200 import synth
    kind = job[0]
    if kind == "cis":
        _, key, length, n, seed, profile = job
        return key, synth.cis_chrom(length, n, seed, profile)
    _, key, la, lb, n, seed = job
    return key, synth.trans_pair(la, lb, n, seed)


def job_pets(job):
    return job[3] if job[0] == "cis" else job[4]


def workload_jobs(cfg, pets, world, scaling):
    """All files of the job.  c2 / c4: `world` synthetic genomes under weak scaling (replicate r has its own seeds),
    one genome under strong scaling.  c3: one genome; the caller keeps the shards it has ranks for."""
    from cloops2_b200 import synth
    reps = world if (scaling == "weak" and "shards" not in cfg) else 1
    jobs = []
    for r in range(reps):
        tag = (lambda nm: nm) if reps == 1 else (lambda nm, r=r: "g%d.%s" % (r, nm))
        if cfg["kind"] == "cis":
            for ci, (name, length) in enumerate(synth.HG38):
                n = int(round(pets * length / synth.HG38_TOTAL))
                jobs.append(("cis", "%s-%s" % (tag(name), tag(name)), length, n, 20261019 + 1000 * cfg["cid"] + ci + 100000 * r,
                             cfg["profile"]))
        else:
            pairs = [(i, j) for i in range(len(synth.HG38)) for j in range(i + 1, len(synth.HG38))]
            w = np.array([synth.HG38[i][1] * float(synth.HG38[j][1]) for i, j in pairs])
            w /= w.sum()
            for pi, (i, j) in enumerate(pairs):
                n = int(round(pets * w[pi]))
                if n > 0:
                    jobs.append(("trans", "%s-%s" % (tag(synth.HG38[i][0]), tag(synth.HG38[j][0])), synth.HG38[i][1],
                                 synth.HG38[j][1], n, 20261019 + 1000 * cfg["cid"] + pi + 100000 * r))
    return jobs


def generate(jobs, nproc):
    """Synthetic files in parallel host processes (forked before any CUDA context exists)."""
    if nproc <= 1 or len(jobs) <= 1:
        return dict(_gen_one(j) for j in jobs)
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(nproc, len(jobs))) as pool:
        return dict(pool.map(_gen_one, jobs, chunksize=1))


def workload_config(cfg, args, world, pets, nfiles):
    return {"workload": cfg["workload"], "config": args.config,
            ("pets_genome" if "shards" in cfg else "pets_per_gpu"): int(pets), "files": nfiles, "eps": cfg["eps"],
            "minPts": cfg["minPts"], "hic": cfg["hic"], "scaling": args.scaling,
            "l2": "inputs larger than L2 (>= 0.8 GB of resident PETs per GPU vs 126 MB)",
            "pvalues": "device fp64 series (cl2_stats.cuh), accurate to ~1e-14 against exact rational arithmetic; scipy "
                       "1.18 itself is up to ~1e-7 off at these populations (tests/test_pvalues_exact.py, DESIGN.md "
                       "section 2); e2e_exact_scipy is the byte-identical mode"}


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device, enabled=True):
        self.rows = []
        self.proc = None
        if not enabled:          # only rank 0 reports clocks; the other ranks keep their host cores for the work
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arms
def _cpu_one(args):
    """The oracle port on one file: returns (key, PETs, final loops as text rows, seconds)."""
    import oracle
    cfgname, key, xy, tot = args
    cfg = CONFIGS[cfgname]
    t = time.time()
    if cfg["kind"] == "cis":
        final, _ = oracle.call_cis_loops({key: xy}, tot, sorted(cfg["eps"]), sorted(cfg["minPts"], reverse=True), hic=cfg["hic"])
    else:
        final = oracle.call_trans_loops({key: xy}, sorted(cfg["eps"]), sorted(cfg["minPts"], reverse=True))
    return key, len(xy), oracle.loops_to_rows(final), time.time() - t


def cpu_sample_files(cfg, target_pets, jobs):
    """Bounded sample of the SAME workload for the single-thread cpu_baseline: the smallest files at full density."""
    out, tot = [], 0
    for j in sorted(jobs, key=job_pets):
        if out and tot + job_pets(j) > target_pets:
            break
        out.append(j)
        tot += job_pets(j)
    return out


def _piece(cfg, pets, m, seed):
    """A synthetic file of m PETs at the density of the workload (a shortened chromosome / a shrunken pair)."""
    from cloops2_b200 import synth
    if cfg["kind"] == "cis":
        length = int(m * synth.HG38_TOTAL / pets) + 1
        return ("cis", "chrP%d-chrP%d" % (seed, seed), length, int(m), 20261019 + 7000 + seed, cfg["profile"])
    la, lb = synth.HG38[0][1], synth.HG38[1][1]
    npair = pets * la * float(lb) / float(sum(a[1] * float(b[1]) for i, a in enumerate(synth.HG38) for b in synth.HG38[i + 1:]))
    f = (m / npair) ** 0.5
    return ("trans", "chrP%d-chrQ%d" % (seed, seed), max(10000, int(la * f)), max(10000, int(lb * f)), int(m), 20261019 + 7000 + seed)


def run_reference_arm(args, cfg, pets):
    """CPU implementation of the path on all host cores (one process per file, like joblib), time-bounded: a
    calibration pass on small pieces sizes the sample so that warm-up + steps fit the budget on this host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    nproc = max(1, cores)
    dens_pets = pets if "shards" not in cfg else pets          # density of the workload's genome either way
    passes = max(1, args.warmup + args.steps)
    t_begin = time.time()
    with mp.get_context("fork").Pool(nproc) as pool:
        m_c = 150_000
        cal = generate([_piece(cfg, dens_pets, m_c, s) for s in range(nproc)], nproc)
        work = [(args.config, k, v, m_c * nproc) for k, v in cal.items()]
        pool.map(_cpu_one, work, chunksize=1)     # first use builds / loads the oracle library in every process
        t = time.time()
        pool.map(_cpu_one, work, chunksize=1)
        t_c = max(1e-3, time.time() - t)
        per_pass = args.ref_budget_s / passes
        m = int(min(8e6, max(5e4, 0.7 * m_c * per_pass / t_c)))
        files = generate([_piece(cfg, dens_pets, m, 100 + s) for s in range(nproc)], nproc)
        tot = int(sum(len(v) for v in files.values()))
        work = [(args.config, k, v, tot) for k, v in files.items()]
        times = []
        for it in range(passes):
            t = time.time()
            pool.map(_cpu_one, work, chunksize=1)
            dt = time.time() - t
            if it >= args.warmup:
                times.append(dt)
            if time.time() - t_begin + dt > 2.2 * args.ref_budget_s and it + 1 < passes:
                # this host is slower than calibrated: stop with the passes that fit (steps reports how many were timed)
                if not times:
                    times.append(dt)
                break
    ms = 1000.0 * float(np.mean(times))
    val = tot / (ms / 1000.0)
    sample = "%d synthetic files of %d PETs each at the workload's density (%d PETs per pass), %d processes; sized by a " \
             "%.1f s calibration pass" % (len(work), m, tot, nproc, t_c)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "PETs/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(cfg, args, 1, pets, len(work)),
        "cpu_baseline": {"value": val, "unit": "PETs/s", "cores": nproc, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "PETs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t_begin,
        "note": "the reference is pure Python and cannot travel to the GPU box; this arm is the C/numpy oracle port of it "
                "(oracle/), pinned to the reference's outputs in the build container; the reference itself measures "
                "2-4 k PETs/s/core (BASELINE.md)"}))


# ---------------------------------------------------------------------------------------------- parity check
P_COLS = (18, 19, 20, 21, 22)      # binomial, hypergeometric, poisson, peakA, peakB columns of _loops.txt


def compare_rows(got, want, rtol=1e-6):
    """Rows of _loops.txt (lists of str): everything identical as text except the five tail probabilities, which the
    benched device mode computes itself (scipy's own error at these populations reaches ~1e-7).
    Returns (ok, worst relative difference of a tail probability)."""
    if len(got) != len(want):
        return False, None
    worst = 0.0
    for g, w in zip(got, want):
        for j, (a, b) in enumerate(zip(g, w)):
            if j in P_COLS:
                fa, fb = float(a), float(b)
                d = abs(fa - fb) / abs(fb) if fb != 0 else abs(fa)
                worst = max(worst, d)
                if d > rtol:
                    return False, worst
            elif a != b:
                return False, worst
    return True, worst


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    global CFG
    args = parse_args()
    cfg = CFG = CONFIGS[args.config]
    pets = float(args.pets) if args.pets else cfg["pets"]
    if args.impl == "reference":
        run_reference_arm(args, cfg, pets)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from cloops2_b200 import shard
    jobs = workload_jobs(cfg, pets, world, args.scaling)
    nshards = cfg.get("shards", world)
    if "shards" in cfg and world > nshards:
        raise SystemExit("config %s has %d shards" % (args.config, nshards))
    parts = shard.assign([j[1] for j in jobs], [job_pets(j) for j in jobs], nshards)
    mine_keys = set(parts[rank])
    my_jobs = [j for j in jobs if j[1] in mine_keys]
    nfiles_job = sum(len(parts[r]) for r in range(world))
    try:
        host_cores = len(os.sched_getaffinity(0))
    except AttributeError:
        host_cores = os.cpu_count() or 1
    cores = max(1, host_cores // world)
    t0 = time.time()
    files = generate(my_jobs, cores)              # fork happens before CUDA / NCCL initialisation
    gen_s = time.time() - t0
    if args.host_cores > 0 and world == 1:
        avail = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, avail[:max(1, min(args.host_cores, len(avail)))])
        host_cores = len(os.sched_getaffinity(0))
    kind = cfg["kind"]
    eps, minPts, hic = sorted(cfg["eps"]), sorted(cfg["minPts"], reverse=True), cfg["hic"]     # cLoops2.py:69-70, 83-84

    cpu_base, cpu_rows = None, {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cj = cpu_sample_files(cfg, args.cpu_sample_pets, my_jobs)
        ctot = int(sum(len(files[j[1]]) for j in cj))
        t = time.time()
        for j in cj:
            k, _, rows, _ = _cpu_one((args.config, j[1], files[j[1]], ctot))
            cpu_rows[k] = rows
        dt = time.time() - t
        cpu_base = {"value": ctot / dt, "unit": "PETs/s", "cores": 1, "kind": "port",
                    "sample": "%d smallest files of the workload (%d PETs), oracle port, %.1f s" % (len(cj), ctot, dt),
                    "tot": ctot}

    import torch
    shard.init_from_env()
    from cloops2_b200 import _lib
    from cloops2_b200.callCisLoops import cis_chromosome, cols_to_loops
    from cloops2_b200.callTransLoops import trans_pair
    _lib.build()
    ctx = _lib.default_context(local)
    keys = [j[1] for j in my_jobs]
    tot = int(sum(len(files[k]) for k in keys))
    job_tot = int(shard.allreduce_stats(np.array([tot], dtype=np.int64))[0])

    # host (pinned) copies for the e2e arm, resident points for the device arm.  A rank runs several files at once on
    # separate streams (cloops2_b200.shard.worker_contexts), largest first.
    ctxs = shard.worker_contexts(local)
    keys = sorted(keys, key=lambda k: -len(files[k]))
    pinned = {}
    for k in keys:
        p = ctx.pinned_empty(files[k].shape, np.int64)
        np.copyto(p, files[k])
        pinned[k] = p
    small_files = {k: files[k] for k in cpu_rows}           # kept for the parity check
    files = None
    owner = {k: ctxs[i % len(ctxs)] for i, k in enumerate(keys)}
    resident = {k: _lib.Points(owner[k], pinned[k]) for k in keys}
    counters = {}
    last = {}

    def unit(c, k, xy, pts, exact=False, tot_=None):
        kk = tuple(k.split("-"))
        if kind == "cis":
            return cis_chromosome(c, kk, xy, job_tot if tot_ is None else tot_, eps, minPts, hic=hic, pts=pts, exact=exact)
        return trans_pair(c, kk, xy, eps, minPts, pts=pts, exact=exact)

    def _sum(results, name):
        st = np.zeros(4, dtype=np.int64)
        host = {}
        for _, s in results:
            st += [s["pets"], s["candidates"], s["tested"], s["loops"]]
            for nm, v in s.get("host_s", {}).items():
                host[nm] = host.get(nm, 0.0) + 1e3 * v
        counters[name] = st
        counters[name + "_host_ms"] = host

    def step_resident():
        # the same dynamic hand-out as the e2e arm; a resident file is driven by whichever worker takes it
        _sum(shard.map_units(ctxs, keys, lambda c, k: unit(c, k, None, resident[k].on(c))), "stats")

    def step_e2e():
        res = shard.map_units(ctxs, keys, lambda c, k: unit(c, k, pinned[k], None))
        _sum(res, "stats_e2e")
        last["e2e"] = res

    def step_e2e_exact():
        from cloops2_b200 import hostops
        prev = shard.set_sync_mode(True)          # as callCisLoops does in this mode: waits sleep, the scipy pool gets the cores
        try:
            res = shard.map_units(ctxs, keys, lambda c, k: unit(c, k, pinned[k], None, exact="defer"))
        finally:
            shard.set_sync_mode(prev)
        hostops.wait_pvalues([f for f, _ in res])

    def timed(fn, warmup, steps):
        for _ in range(warmup):
            fn()
        for c in ctxs:
            c.sync()
        shard.barrier()
        torch.cuda.synchronize()
        for c in ctxs:
            c.launch_count_reset()
        sampler = ClockSampler(local, enabled=(rank == 0))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        ev0.record()
        step_marks = [t]
        cpu0 = time.process_time()
        for _ in range(steps):
            fn()
            step_marks.append(time.perf_counter())       # every unit's results are on the host when fn returns
        timed.last_step_ms = [round(1e3 * (b - a), 2) for a, b in zip(step_marks, step_marks[1:])]
        timed.last_cpu_ms = round(1e3 * (time.process_time() - cpu0) / max(1, steps), 1)   # host CPU (all threads) per step
        for c in ctxs:
            c.sync()
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t
        shard.barrier()
        clocks = sampler.stop()
        # the step mixes device work with host-side bookkeeping, so the step time is the host clock around the
        # synchronised region (the CUDA-event span on torch's stream is reported beside it)
        dev_ms = ev0.elapsed_time(ev1)
        xf = np.sum([c.transfer_bytes() for c in ctxs], axis=0)
        return shard.max_over_ranks(wall), dev_ms, clocks, sum(c.launch_count() for c in ctxs), (int(xf[0]), int(xf[1]))

    # A full (generation-2) collection walks every object torch / numpy / scipy created at import: ~45 ms, once every
    # 8-10 steps.  Park those objects in the permanent generation so the collector only sees what the steps allocate.
    import gc
    gc.collect()
    gc.freeze()
    wall, dev_ms, clocks, launches, _ = timed(step_resident, args.warmup, args.steps)
    steps_ms, cpu_ms = timed.last_step_ms, timed.last_cpu_ms
    wall_e, _, clocks_e, _, xfer = timed(step_e2e, args.warmup, args.steps)
    steps_ms_e, cpu_ms_e = timed.last_step_ms, timed.last_cpu_ms
    sync_ab = {}
    for m in [int(v) for v in args.extra_sync_modes.split(",") if v.strip() != ""]:
        prev = shard.set_sync_mode(m)
        try:
            w = timed(step_resident, 1, args.steps)[0]
            sync_ab[str(m)] = {"ms_per_step": round(1e3 * w / args.steps, 2), "host_cpu_ms_per_step": timed.last_cpu_ms}
        finally:
            shard.set_sync_mode(prev)

    # parity of the benched configuration itself: the oracle's loops for the cpu_baseline sample vs the GPU arm's loops
    # for the same files (the oracle ran them as a job of their own, so the GPU arm repeats them with that job's total)
    parity = None
    if cpu_rows:
        import oracle
        worst, ok = 0.0, True
        for k in cpu_rows:
            final, _ = unit(ctx, k, small_files[k], None, tot_=cpu_base["tot"])
            got = oracle.loops_to_rows(cols_to_loops(tuple(k.split("-")), final, cis=(kind == "cis")))
            o, w = compare_rows(got, cpu_rows[k])
            ok = ok and o
            worst = max(worst, w or 0.0)
        parity = {"status": "equal" if ok else "MISMATCH", "files": sorted(cpu_rows), "loops": int(sum(len(v) for v in cpu_rows.values())),
                  "max_rel_diff_tail_probability": worst,
                  "what": "rows of _loops.txt from the oracle port vs the GPU arm for the same files: identical text, tail "
                          "probabilities within 1e-6 relative (device fp64 series vs scipy)"}

    # Per-kernel roofline: the timed steps run several files concurrently on separate streams, where a CUDA-event pair
    # around one launch also spans other streams' kernels.  The per-kernel numbers therefore come from the same resident
    # step run once more on ONE stream at a time (same launches, same inputs), events enabled.
    fam = None
    if not args.no_kernel_timing:
        for c in ctxs:
            c.sync()
            c.timing_reset()
            c.timing_enable(True)
        _sum([unit(owner[k], k, None, resident[k]) for k in keys], "stats")
        for c in ctxs:
            c.sync()
        fam = {}
        for c in ctxs:
            for nm, v in c.timing_read().items():
                a = fam.setdefault(nm, {"ms": 0.0, "launches": 0, "bytes": 0.0})
                a["ms"] += v["ms"]
                a["launches"] += v["launches"]
                a["bytes"] += v["bytes"]
            c.timing_enable(False)
    xfer = shard.allreduce_stats(np.array(xfer, dtype=np.int64))
    wall_x, ex_steps = None, max(1, min(2, args.steps))
    if not args.no_exact:
        wall_x = timed(step_e2e_exact, 1, ex_steps)[0]

    # file to file: .ixy directory in, _loops.txt out, through the drop-in callCisLoops (page cache warm)
    e2e_files = None
    if kind == "cis" and world == 1 and not args.no_files and "shards" not in cfg:
        import logging
        import shutil
        import tempfile
        from cloops2_b200 import synth
        from cloops2_b200.callCisLoops import callCisLoops
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        tmp = tempfile.mkdtemp(prefix="cl2bench_", dir=base)
        try:
            synth.write_ixy_dir({k: np.asarray(pinned[k]) for k in sorted(keys, key=lambda k: [j[1] for j in my_jobs].index(k))}, tmp + "/d")
            os.environ["CL2_PVALUES"] = "device"
            log = logging.getLogger("cl2bench")
            log.addHandler(logging.NullHandler())
            log.propagate = False
            ts = []
            for it in range(3):
                t = time.perf_counter()
                callCisLoops(tmp + "/d", tmp + "/out%d" % it, None, eps=eps, minPts=minPts, hic=hic)
                ts.append(time.perf_counter() - t)
            nl = sum(1 for _ in open(tmp + "/out2_loops.txt")) - 1
            e2e_files = {"value": job_tot / min(ts[1:]), "unit": "PETs/s", "ms": [round(1e3 * v, 1) for v in ts], "loops_written": nl,
                         "note": "callCisLoops(predir, fout): joblib memory map -> pinned staging -> GPU -> _loops.txt; "
                                 "first call includes context creation; page cache warm; best of the last two"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    stats = shard.allreduce_stats(counters["stats"])
    if rank != 0:
        return
    ms_step = 1000.0 * wall / args.steps
    value = job_tot / (wall / args.steps)
    e2e_val = job_tot / (wall_e / args.steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    roof = None
    kernels = {}
    if fam:
        tot_ms = sum(v["ms"] for v in fam.values()) or 1.0
        for nm, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            kernels[nm] = {"ms_per_step": v["ms"], "launches_per_step": v["launches"],
                           "share": v["ms"] / tot_ms,
                           "GBps": (v["bytes"] / 1e9) / (v["ms"] / 1e3) if v["ms"] > 0 and v["bytes"] > 0 else None}
        top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        nm, v = top
        ach = (v["bytes"] / 1e9) / (v["ms"] / 1e3) if v["ms"] > 0 else 0.0
        traffic, traffic_note, traffic_alg = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(nm)
            traffic, traffic_note = tj["dram_bytes_per_launch"], tj["note"]
            traffic_alg = tj.get("capture_algorithmic_bytes")
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": nm, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_note": traffic_note, "traffic_capture_algorithmic_bytes": traffic_alg,
                "peak_source": peak_src,
                "measured_in": "single-stream instrumented pass of the resident step (kernels_note)",
                "avg_launch_ms": v["ms"] / max(1, v["launches"]),
                "algorithmic_bytes_per_launch": v["bytes"] / max(1, v["launches"])}
    h2d, d2h = int(xfer[0]) // args.steps, int(xfer[1]) // args.steps      # counted by the library per copy
    out = {
        "metric": METRIC, "value": value, "unit": "PETs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(cfg, args, world, pets, nfiles_job),
        "e2e": {"value": e2e_val, "unit": "PETs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1000.0 * wall_e / args.steps, "step_ms": steps_ms_e,
                "host_cpu_ms_per_step": cpu_ms_e},
        "step_ms": steps_ms, "host_cpu_ms_per_step": cpu_ms,
        "e2e_exact_scipy": None if wall_x is None else {
            "value": job_tot / (wall_x / ex_steps), "unit": "PETs/s", "steps": ex_steps,
            "note": "same call in the byte-identical mode: cuts decided with scipy's value wherever the device value is "
                    "within 1e-5 of a threshold, the final loops' tail probabilities re-evaluated by scipy in the rank's worker "
                    "processes (hostops.pvalue_pool) while other files are on the GPU"},
        "e2e_files": e2e_files,
        "parity_check": parity,
        "gpu_launches": int(launches),
        "clocks": clocks, "clocks_e2e": clocks_e,
        "roofline": roof, "cpu_baseline": {k: v for k, v in cpu_base.items() if k != "tot"} if cpu_base else None,
        "kernels": kernels,
        "kernels_note": "per-kernel-family CUDA-event times of one resident step run on a single stream (see bench.py)",
        "stage_wall_ms_per_step": counters.get("stats_host_ms"),
        "host_threads": len(ctxs), "host_cores": host_cores, "host_sync_mode": shard._sync_mode[0],
        "sync_modes_ms_per_step": sync_ab or None,
        "job": {"pets": job_tot, "candidates": int(stats[1]), "tested": int(stats[2]), "loops": int(stats[3]),
                "device_event_ms_per_step": dev_ms / args.steps, "generate_s": gen_s},
    }
    print(json.dumps(out))
    if parity and parity["status"] != "equal":
        raise SystemExit("parity check failed: GPU loops differ from the oracle's on %s" % ", ".join(parity["files"]))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    # exactly ONE line on stdout (the JSON): libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    import io
    _buf = io.StringIO()
    _py_stdout = sys.stdout
    sys.stdout = _buf
    _rc = 0
    try:
        main()
    except SystemExit as e:
        _rc = e.code if isinstance(e.code, int) else 1
        if not isinstance(e.code, int) and e.code:
            print(e.code, file=sys.stderr)
    finally:
        _shutdown()
        sys.stdout = _py_stdout
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        os.close(_real_stdout)
        out = _buf.getvalue().strip().splitlines()
        if out:
            print(out[-1], flush=True)
    sys.exit(_rc)
