#!/usr/bin/env python
"""
bench.py -- PETs clustered+tested per second for `callLoops` (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[1], named in config.workload): synthetic Hi-TrAC-like intra-chromosomal hg38 PETs,
100 M per GPU, callLoops -eps 200,500,1000 -minPts 5 (blockDBSCAN), chromosome files sharded over the ranks by PET
count (weak scaling: at N GPUs the job is N synthetic genomes = 24*N files).  One "step" = the whole hot path over
the rank's files: 3 clustering passes + candidate boxes, XY index, per-candidate counting, hypergeometric / Poisson /
binomial tests, significance marking and overlap selection -- everything `callCisLoops` does between reading the
.ixy arrays and writing the text file.

  value  whole-job PETs/s with the PET arrays already resident in HBM (cl2_points) when the timed region starts
  e2e    the same metric through the public per-chromosome call with HOST (pinned) int64 arrays: the upload of every
         chromosome and the read-back of candidate boxes / counts are inside the timed region
  roofline  dominant kernel family by device time (CUDA events on the library's stream, measured during the timed
         steps): algorithmic bytes per launch / average launch time against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (oracle/, single thread) on a bounded sample of the same workload (rank 0, N=1)

`--impl reference` times the CPU implementation of the path (the oracle port of the pure-Python reference, which
cannot travel to the GPU box) on all host cores, one process per chromosome file like the reference's joblib fan-out.
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

EPS = [200, 500, 1000]
MINPTS = [5]
PROFILE = "hitrac"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pets", type=float, default=100e6, help="PETs per GPU (default: the 100 M of configs[1])")
    ap.add_argument("--cpu-sample-pets", type=float, default=5.5e6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true", help="do not record per-kernel CUDA events in the timed steps")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- workload
def _gen_one(job):
    from cloops2_b200 import synth
    key, length, n, seed = job
    return key, synth.cis_chrom(length, n, seed, PROFILE)


def workload_jobs(pets_per_gpu, world):
    """All chromosome files of the job: `world` synthetic genomes (replicate r has its own seeds)."""
    from cloops2_b200 import synth
    jobs = []
    for r in range(world):
        for ci, (name, length) in enumerate(synth.HG38):
            n = int(round(pets_per_gpu * length / synth.HG38_TOTAL))
            key = "%s-%s" % (name, name) if world == 1 else "g%d.%s-g%d.%s" % (r, name, r, name)
            jobs.append((key, length, n, 20261019 + 1000 * 2 + ci + 100000 * r))
    return jobs


def generate(jobs, nproc):
    """Synthetic files in parallel host processes (forked before any CUDA context exists)."""
    if nproc <= 1 or len(jobs) <= 1:
        return dict(_gen_one(j) for j in jobs)
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(nproc, len(jobs))) as pool:
        return dict(pool.map(_gen_one, jobs, chunksize=1))


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
    import oracle
    key, xy, tot = args
    t = time.time()
    final, _ = oracle.call_cis_loops({key: xy}, tot, EPS, MINPTS)
    return key, len(xy), len(final), time.time() - t


def cpu_sample_files(target_pets, pets_per_gpu):
    """Bounded sample of the SAME workload: the smallest chromosomes of the synthetic genome at full density."""
    from cloops2_b200 import synth
    order = sorted(range(len(synth.HG38)), key=lambda i: synth.HG38[i][1])
    jobs, tot = [], 0
    for ci in order:
        name, length = synth.HG38[ci]
        n = int(round(pets_per_gpu * length / synth.HG38_TOTAL))
        if jobs and tot + n > target_pets:
            break
        jobs.append(("%s-%s" % (name, name), length, n, 20261019 + 1000 * 2 + ci))
        tot += n
    return jobs


def run_reference_arm(args):
    """CPU implementation of the path on all host cores (one process per chromosome file, like joblib)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    # sample sized so that W+K passes end within a few minutes
    jobs = cpu_sample_files(args.cpu_sample_pets * max(1, min(cores, 24)) / 2, args.pets)
    files = generate(jobs, cores)
    tot = int(sum(len(v) for v in files.values()))
    work = [(k, v, tot) for k, v in files.items()]
    nproc = min(cores, len(work))
    times = []
    with mp.get_context("fork").Pool(nproc) as pool:
        for it in range(args.warmup + args.steps):
            t = time.time()
            pool.map(_cpu_one, work, chunksize=1)
            dt = time.time() - t
            if it >= args.warmup:
                times.append(dt)
            if it == 0 and dt > 60:        # keep the whole arm within minutes on slow hosts
                args.warmup, args.steps = min(args.warmup, 1), min(args.steps, 2)
            if len(times) >= args.steps:
                break
    ms = 1000.0 * float(np.mean(times))
    val = tot / (ms / 1000.0)
    sample = "%d smallest chromosomes of the 100M-PET synthetic genome (%d PETs), %d processes" % (len(work), tot, nproc)
    print(json.dumps({
        "impl": "reference", "metric": "PETs clustered+tested/sec (callLoops)", "value": val, "unit": "PETs/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": "PETs/s", "cores": nproc, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "PETs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is pure Python and cannot travel to the GPU box; this arm is the C/numpy oracle port of it "
                "(oracle/), pinned to the reference's outputs in the build container; the reference itself measures "
                "2-4 k PETs/s/core (BASELINE.md)"}))


def workload_config(args, world):
    return {"workload": "BASELINE configs[1]: synthetic Hi-TrAC-like intra-chromosomal hg38 PETs, callLoops "
                        "-eps 200,500,1000 -minPts 5 (blockDBSCAN), chromosome files sharded by PET count",
            "pets_per_gpu": int(args.pets), "files": 24 * world, "eps": EPS, "minPts": MINPTS,
            "l2": "inputs larger than L2 (0.8 GB of resident PETs per GPU vs 126 MB)",
            "pvalues": "device fp64 (cl2_stats.cuh; <=1e-9 relative to scipy where scipy is that accurate, DESIGN.md section 2)"}


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from cloops2_b200 import shard
    jobs = workload_jobs(args.pets, world)
    mine_keys = set(shard.assign([j[0] for j in jobs], [j[2] for j in jobs], world)[rank])
    my_jobs = [j for j in jobs if j[0] in mine_keys]
    cores = max(1, (os.cpu_count() or 1) // world)
    t0 = time.time()
    files = generate(my_jobs, cores)              # fork happens before CUDA / NCCL initialisation
    gen_s = time.time() - t0

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cj = cpu_sample_files(args.cpu_sample_pets, args.pets)
        cf = {k: files[k] if k in files else _gen_one((k, l, n, s))[1] for k, l, n, s in cj}
        ctot = int(sum(len(v) for v in cf.values()))
        t = time.time()
        for k, v in cf.items():
            _cpu_one((k, v, ctot))
        dt = time.time() - t
        cpu_base = {"value": ctot / dt, "unit": "PETs/s", "cores": 1, "kind": "port",
                    "sample": "%d smallest chromosomes of the workload (%d PETs), oracle port, %.1f s" % (len(cf), ctot, dt)}

    import torch
    shard.init_from_env()
    from cloops2_b200 import _lib
    from cloops2_b200.callCisLoops import cis_chromosome
    _lib.build()
    ctx = _lib.default_context(local)
    keys = [j[0] for j in my_jobs]
    tot = int(sum(len(files[k]) for k in keys))
    job_tot = int(shard.allreduce_stats(np.array([tot], dtype=np.int64))[0])

    # host (pinned) copies for the e2e arm, resident points for the device arm.  A rank runs several chromosome files
    # at once on separate streams (cloops2_b200.shard.worker_contexts), largest first.
    ctxs = shard.worker_contexts(local)
    keys = sorted(keys, key=lambda k: -len(files[k]))
    pinned = {}
    for k in keys:
        p = ctx.pinned_empty(files[k].shape, np.int64)
        np.copyto(p, files[k])
        pinned[k] = p
    files = None
    owner = {k: ctxs[i % len(ctxs)] for i, k in enumerate(keys)}
    resident = {k: _lib.Points(owner[k], pinned[k]) for k in keys}
    counters = {}

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
        _sum(shard.map_units(ctxs, keys, lambda c, k: cis_chromosome(c, tuple(k.split("-")), None, job_tot, EPS, MINPTS,
                                                                      pts=resident[k].on(c))), "stats")

    def step_e2e():
        _sum(shard.map_units(ctxs, keys, lambda c, k: cis_chromosome(c, tuple(k.split("-")), pinned[k], job_tot, EPS, MINPTS)),
             "stats_e2e")

    def step_e2e_exact():
        shard.map_units(ctxs, keys, lambda c, k: cis_chromosome(c, tuple(k.split("-")), pinned[k], job_tot, EPS, MINPTS,
                                                                 exact=True))

    def timed(fn, warmup, steps, with_kernel_timing):
        for _ in range(warmup):
            fn()
        for c in ctxs:
            c.sync()
        shard.barrier()
        torch.cuda.synchronize()
        for c in ctxs:
            if with_kernel_timing:
                c.timing_reset()
                c.timing_enable(True)
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
        fam = None
        if with_kernel_timing:
            fam = {}
            for c in ctxs:
                for nm, v in c.timing_read().items():
                    a = fam.setdefault(nm, {"ms": 0.0, "launches": 0, "bytes": 0.0})
                    a["ms"] += v["ms"]
                    a["launches"] += v["launches"]
                    a["bytes"] += v["bytes"]
                c.timing_enable(False)
        # the step mixes device work with host-side statistics, so the step time is the host clock around the
        # synchronised region (the CUDA-event span on torch's stream is reported beside it)
        dev_ms = ev0.elapsed_time(ev1)
        xf = np.sum([c.transfer_bytes() for c in ctxs], axis=0)
        return shard.max_over_ranks(wall), dev_ms, clocks, fam, sum(c.launch_count() for c in ctxs), (int(xf[0]), int(xf[1]))

    # A full (generation-2) collection walks every object torch / numpy / scipy created at import: ~45 ms, once every
    # 8-10 steps.  Park those objects in the permanent generation so the collector only sees what the steps allocate.
    import gc
    gc.collect()
    gc.freeze()
    wall, dev_ms, clocks, _, launches, _ = timed(step_resident, args.warmup, args.steps, False)
    steps_ms, cpu_ms = timed.last_step_ms, timed.last_cpu_ms
    wall_e, _, clocks_e, _, _, xfer = timed(step_e2e, args.warmup, args.steps, False)
    steps_ms_e, cpu_ms_e = timed.last_step_ms, timed.last_cpu_ms
    # Per-kernel roofline: the timed steps run several chromosome files concurrently on separate streams, where a
    # CUDA-event pair around one launch also spans other streams' kernels.  The per-kernel numbers therefore come from
    # the same resident step run once more on ONE stream at a time (same launches, same inputs), events enabled.
    fam = None
    if not args.no_kernel_timing:
        def step_serial():
            _sum([cis_chromosome(owner[k], tuple(k.split("-")), None, job_tot, EPS, MINPTS, pts=resident[k]) for k in keys],
                 "stats")
        for c in ctxs:
            c.sync()
            c.timing_reset()
            c.timing_enable(True)
        t_serial = time.perf_counter()
        step_serial()
        for c in ctxs:
            c.sync()
        t_serial = time.perf_counter() - t_serial
        fam = {}
        for c in ctxs:
            for nm, v in c.timing_read().items():
                a = fam.setdefault(nm, {"ms": 0.0, "launches": 0, "bytes": 0.0})
                a["ms"] += v["ms"]
                a["launches"] += v["launches"]
                a["bytes"] += v["bytes"]
            c.timing_enable(False)
    xfer = shard.allreduce_stats(np.array(xfer, dtype=np.int64))
    ex_steps = max(1, min(2, args.steps))
    wall_x = timed(step_e2e_exact, 1, ex_steps, False)[0]

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
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(nm)
            traffic, traffic_note = tj["dram_bytes_per_launch"], tj["note"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": nm, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "measured_in": "single-stream instrumented pass of the resident step (kernels_note)",
                "avg_launch_ms": v["ms"] / max(1, v["launches"]),
                "algorithmic_bytes_per_launch": v["bytes"] / max(1, v["launches"])}
    h2d, d2h = int(xfer[0]) // args.steps, int(xfer[1]) // args.steps      # counted by the library per copy
    out = {
        "metric": "PETs clustered+tested/sec (callLoops)", "value": value, "unit": "PETs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_val, "unit": "PETs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1000.0 * wall_e / args.steps, "step_ms": steps_ms_e,
                "host_cpu_ms_per_step": cpu_ms_e},
        "step_ms": steps_ms, "host_cpu_ms_per_step": cpu_ms,
        "e2e_exact_scipy": {"value": job_tot / (wall_x / ex_steps), "unit": "PETs/s", "steps": ex_steps,
                            "note": "same call with the final loops' five tail probabilities re-evaluated by scipy on "
                                    "the host (byte-identical _loops.txt; scipy.hypergeom.sf ~150 us per loop)"},
        "gpu_launches": int(launches),
        "clocks": clocks, "clocks_e2e": clocks_e,
        "roofline": roof, "cpu_baseline": cpu_base,
        "kernels": kernels,
        "kernels_note": "per-kernel-family CUDA-event times of one resident step run on a single stream (see bench.py)",
        "stage_wall_ms_per_step": counters.get("stats_host_ms"),
        "host_threads": len(ctxs),
        "job": {"pets": job_tot, "candidates": int(stats[1]), "tested": int(stats[2]), "loops": int(stats[3]),
                "device_event_ms_per_step": dev_ms / args.steps, "generate_s": gen_s},
    }
    print(json.dumps(out))


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
    try:
        main()
    finally:
        _shutdown()
        sys.stdout = _py_stdout
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        os.close(_real_stdout)
        out = _buf.getvalue().strip().splitlines()
        if out:
            print(out[-1], flush=True)
