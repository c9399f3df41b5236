"""
Array-level host logic of the callLoops path.  The reference walks Python lists of Loop objects; at 10^5-10^6
candidates per run that is the bottleneck once clustering and counting are on the GPU, so everything between the
kernels is numpy on struct-of-arrays here, with the reference lines each step reproduces cited inline.
Float expressions are evaluated in the same order as the reference so results are bit-identical.
"""
import ctypes
import os

import numpy as np
from . import _lib, pvalues


# ---------------------------------------------------------------------------------------------- candidates
def cis_candidate_mask(bbox):
    """callCisLoops.py:97-99 (degenerate box), :114 (anchor sizes sum < 200), :116 (x_end < y_start)."""
    x0, x1, y0, y1 = bbox[:, 0], bbox[:, 1], bbox[:, 2], bbox[:, 3]
    ok = (x0 != x1) & (y0 != y1)
    ok &= (x1 - x0 + y1 - y0) >= 200
    ok &= x1 < y0
    return ok


def trans_candidate_mask(bbox):
    """callTransLoops.py:62-64 (degenerate box only)."""
    return (bbox[:, 0] != bbox[:, 1]) & (bbox[:, 2] != bbox[:, 3])


def combine_unique(cands_per_pass):
    """combineLoops over the sweep (geo.py:90-112) followed by filterDupLoops (callCisLoops.py:159-169):
    coordinates in order of first appearance.  (np.unique(axis=0) costs ~0.3 s per 150 k rows; a 64-bit mix of the
    four coordinates is sorted instead and the grouping is verified exactly, with an exact fallback.)"""
    if not cands_per_pass:
        return np.zeros((0, 4), dtype=np.int64)
    allc = np.concatenate(cands_per_pass, axis=0) if len(cands_per_pass) > 1 else cands_per_pass[0]
    n = allc.shape[0]
    if n == 0:
        return allc.reshape(0, 4)
    u = allc.astype(np.uint64)
    h = (u[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ (u[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F)) \
        ^ (u[:, 2] * np.uint64(0x165667B19E3779F9)) ^ (u[:, 3] * np.uint64(0x27D4EB2F165667C5))
    order = np.argsort(h, kind="stable")
    hs, cs = h[order], allc[order]
    same_h = hs[1:] == hs[:-1]
    same_c = (cs[1:] == cs[:-1]).all(axis=1)
    if (same_h & ~same_c).any():                      # a 64-bit collision between different boxes: exact path
        _, first = np.unique(allc, axis=0, return_index=True)
        return allc[np.sort(first)]
    head = np.ones(n, dtype=bool)
    head[1:] = ~same_h
    return allc[np.sort(order[head])]                 # stable sort => smallest original index of each group


def first_occurrence(coords, pass_id=None):
    """Indices of the rows to keep so that coordinates appear once, at their first position -- what combineLoops +
    filterDupLoops leave (geo.py:90-112, callCisLoops.py:159-169).  With `pass_id` only rows whose coordinates were
    seen in an EARLIER pass are dropped (callTransLoops has no filterDupLoops, so duplicates inside one clustering
    run survive, callTransLoops.py:230-234).  Meant for the few rows that survive the tests."""
    seen = {}
    keep = []
    rows = coords.tolist()
    for i, r in enumerate(rows):
        t = tuple(r)
        if t not in seen:
            seen[t] = pass_id[i] if pass_id is not None else 0
            keep.append(i)
        elif pass_id is not None and seen[t] == pass_id[i]:
            keep.append(i)
    return np.array(keep, dtype=np.int64)


def cis_distance(c):
    """distance = abs(y_center - x_center) with centers (start+end)/2 as floats (callCisLoops.py:107-112)."""
    xc = (c[:, 0] + c[:, 1]) / 2
    yc = (c[:, 2] + c[:, 3]) / 2
    return np.abs(yc - xc)


# ---------------------------------------------------------------------------------------------- statistics
_floor300 = pvalues.floor300
anchor_sig = pvalues.anchor_sig          # estAnchorSig (callCisLoops.py:226-247)


def permuted_stats(na, nb, nab, rab, cis=True):
    """getLoopNearbyPETs (callCisLoops.py:201-223) / the trans variant (callTransLoops.py:155-172):
    returns (mrabs, mbps, fdr)."""
    L, W = na.shape
    rabs = nab.astype(np.float64)                                    # [L, W*W], a-major
    den = (na.astype(np.float64)[:, :, None] * nb.astype(np.float64)[:, None, :]).reshape(L, W * W)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = rabs / den
    if cis:
        nbps = np.where(rabs > 0, q, 0.0)                            # :213-221
        mrabs = np.median(rabs, axis=1)
        mbps = np.median(nbps, axis=1)
    else:
        nbps = np.where(den > 0, q, 0.0)                             # callTransLoops.py:165-169 (nac>0 and nbc>0)
        mrabs = np.mean(rabs, axis=1)
        mbps = np.mean(nbps, axis=1)
    fdr = (rabs > rab[:, None]).sum(axis=1) / float(W * W)
    return mrabs, mbps, fdr


def _clamp(p):
    """max([1e-300, p]) as the reference wraps every tail probability (callCisLoops.py:246,310,329,331)."""
    return np.maximum(1e-300, p)


def est_cis(index, cands, N, span, tot, minPts, hic=False, pseudo=1, peakPcut=1e-5, win=5, countDiffCut=20,
            exact=False, stepwise=None):
    """estLoopSig (callCisLoops.py:250-354) for one chromosome.  `index` is a device XY index; cands int64[L,4].
    Counting, the permuted-window statistics and the tail probabilities are computed on the device (two launches:
    every candidate, then the survivors of the first filters); the filter cascade itself is applied here in the
    reference's order.  Returns a dict of column arrays for the surviving loops, in candidate order.
    exact=True re-evaluates the five tail probabilities of the surviving loops with scipy, which reproduces the
    reference's printed digits (the device values are accurate to ~1e-14 but scipy itself is not)."""
    L = cands.shape[0]
    cols = _empty_cols()
    if L == 0:
        return cols
    if stepwise is None:
        stepwise = not hasattr(index, "est_loops")          # an index that only offers the phase-by-phase calls
    if not stepwise:
        # fused device path: counting, statistics and the filter cascade below all run on the device; only the
        # survivors come back.  stepwise=True runs the same arithmetic phase by phase with the cascade applied here
        # (tests/test_hostops.py drives it with an oracle-backed index, tests/test_gpu_pipeline.py with the device)
        rpb = float(N) / span
        if exact:
            index.ctx.set_pvalue_margin(PV_MARGIN)
        try:
            sel, core, hyp, st = index.est_loops(cands, rpb, minPts, mode=_lib.MODE_CIS, win=win, hic=hic,
                                                 countDiffCut=countDiffCut, pseudo=pseudo, peakPcut=peakPcut)
        finally:
            if exact:
                index.ctx.set_pvalue_margin(0.0)
        cols = _cols_from_stats(cands[sel], core, hyp, st, pseudo, tot)
        if exact:
            cols = scipy_pvalues(exact_cuts(cols, N, rpb, cis=True, hic=hic, peakPcut=peakPcut, mark=False), N, rpb, cis=True)
        return cols
    la = (cands[:, 1] - cands[:, 0])
    lb = (cands[:, 3] - cands[:, 2])
    keep = ~((la / lb > countDiffCut) | (lb / la > countDiffCut))                    # :282-286
    c = cands[keep]
    if c.shape[0] == 0:
        return cols
    core, hyp = index.loop_core(c, mode=_lib.MODE_CIS)
    ra, rb, rab, lower = core[:, 0], core[:, 1], core[:, 2], core[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        k = rab >= minPts                                                             # :290
        k &= ~((ra == rab) | (rb == rab))                                             # :293
        k &= ~((ra / rb.astype(np.float64) > countDiffCut) | (rb / ra.astype(np.float64) > countDiffCut))   # :296
    p2ll = rab.astype(np.float64) / np.maximum(lower, pseudo)                         # :306
    if hic:
        k &= ~(p2ll < 1)                                                              # :307
    hyp = _clamp(hyp)                                                                 # :310
    k &= ~(hyp > 1e-2)                                                                # :311
    c, core, ra, rb, rab, p2ll, hyp = c[k], core[k], ra[k], rb[k], rab[k], p2ll[k], hyp[k]
    if c.shape[0] == 0:
        return cols
    rpb = float(N) / span                                                             # :231
    st = index.loop_tests(c, core, rpb, win=win, mode=_lib.MODE_CIS)
    mrabs, mbps, fdr = st[:, 0], st[:, 1], st[:, 2]
    k = ~((mrabs >= rab) | (fdr >= 0.1))                                              # :322
    es = rab / np.maximum(mrabs, pseudo)                                              # :325
    k &= ~(es < 1)
    pop, nbp = _clamp(st[:, 3]), _clamp(st[:, 4])                                     # :329, :331-333
    px, esx, py, esy = _clamp(st[:, 5]), st[:, 6], _clamp(st[:, 7]), st[:, 8]
    if not hic:
        k &= (px < peakPcut) & (py < peakPcut)                                        # :347
    la, lb = c[:, 1] - c[:, 0], c[:, 3] - c[:, 2]
    density = rab.astype(np.float64) / (la + lb) / tot * 10.0**9                      # :337-339
    cols = dict(coords=c[k], ra=ra[k], rb=rb[k], rab=rab[k], P2LL=p2ll[k], FDR=fdr[k], ES=es[k], density=density[k],
                hyp=hyp[k], pop=pop[k], nbp=nbp[k], px=px[k], py=py[k], esx=esx[k], esy=esy[k],
                mrabs=mrabs[k], mbps=mbps[k], anc=st[k, 9:13])
    if exact:
        cols = scipy_pvalues(cols, N, rpb, cis=True)
    return cols


def est_trans(index, cands, N, span, minPts, pseudo=1, countDiffCut=10, win=5, exact=False, pass_id=None,
              stepwise=None):
    """trans estLoopSig (callTransLoops.py:106-193).  `pass_id` (optional, per candidate) is carried through the
    filters so the caller can apply combineLoops' cross-pass de-duplication afterwards."""
    L = cands.shape[0]
    cols = _empty_cols()
    if L == 0:
        if pass_id is not None:
            cols["pass_id"] = np.zeros(0, dtype=np.int64)
        return cols
    if stepwise is None:
        stepwise = not hasattr(index, "est_loops")
    if not stepwise:
        rpb = float(N) / span
        sel, core, hyp, st = index.est_loops(cands, rpb, minPts, mode=_lib.MODE_TRANS, win=win,
                                             countDiffCut=countDiffCut, pseudo=pseudo)
        cols = _cols_from_stats(cands[sel], core, hyp, st, pseudo, N)
        if pass_id is not None:
            cols["pass_id"] = np.asarray(pass_id)[sel]
        return scipy_pvalues(cols, N, rpb, cis=False) if exact else cols
    core, hyp = index.loop_core(cands, mode=_lib.MODE_TRANS)
    ra, rb, rab, lower = core[:, 0], core[:, 1], core[:, 2], core[:, 3]
    la, lb = cands[:, 1] - cands[:, 0], cands[:, 3] - cands[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        k = rab >= minPts                                                             # :127
        k &= ~((ra / rb.astype(np.float64) > countDiffCut) | (rb / ra.astype(np.float64) > countDiffCut))   # :133
        k &= ~((la / lb > countDiffCut) | (lb / la > countDiffCut))                   # :135-139
    c, core, ra, rb, rab, lower, la, lb, hyp = cands[k], core[k], ra[k], rb[k], rab[k], lower[k], la[k], lb[k], hyp[k]
    if c.shape[0] == 0:
        if pass_id is not None:
            cols["pass_id"] = np.zeros(0, dtype=np.int64)
        return cols
    p2ll = rab.astype(np.float64) / np.maximum(lower, pseudo)
    rpb = float(N) / span
    st = index.loop_tests(c, core, rpb, win=win, mode=_lib.MODE_TRANS)
    mrabs, mbps, fdr = st[:, 0], st[:, 1], st[:, 2]
    es = rab / np.maximum(mrabs, pseudo)
    density = rab.astype(np.float64) / (la + lb) / N * 10.0**9                        # :185-187
    cols = dict(coords=c, ra=ra, rb=rb, rab=rab, P2LL=p2ll, FDR=fdr, ES=es, density=density, hyp=_clamp(hyp),
                pop=_clamp(st[:, 3]), nbp=_clamp(st[:, 4]), px=_clamp(st[:, 5]), py=_clamp(st[:, 7]), esx=st[:, 6],
                esy=st[:, 8], mrabs=mrabs, mbps=mbps, anc=st[:, 9:13])
    if pass_id is not None:
        cols["pass_id"] = np.asarray(pass_id)[k]
    if exact:
        cols = scipy_pvalues(cols, N, rpb, cis=False)
    return cols


PV_CHUNK = 256        # rows per task of the p-value pool (hypergeom.sf: 5-40 ms of work each)


def _tails_chunks(args, cis, pool=None):
    """pvalues.tails over row chunks of `args` (a tuple of equally long arrays), as futures of the p-value pool (or,
    without a pool, as already evaluated results behind the same interface)."""
    n = args[0].shape[0]
    out = []
    for o in range(0, n, PV_CHUNK):
        part = tuple(np.ascontiguousarray(a[o:o + PV_CHUNK]) for a in args)
        out.append(pool.submit(pvalues.tails, *part, cis=cis) if pool is not None else _Done(pvalues.tails(*part, cis=cis)))
    return out, args


class _Done:
    def __init__(self, v):
        self.v = v

    def result(self):
        return self.v


def _gather_tails(futs, args, cis):
    """Results of _tails_chunks in row order; a task whose worker process died is evaluated here instead."""
    parts = []
    for i, f in enumerate(futs):
        try:
            parts.append(f.result())
        except Exception:
            o = i * PV_CHUNK
            parts.append(pvalues.tails(*[a[o:o + PV_CHUNK] for a in args], cis=cis))
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, 7))


def _cols_from_stats(c, core, hyp, st, pseudo, tot):
    """Column dict from the device outputs of the surviving loops (same expressions as the step-by-step path)."""
    if c.shape[0] == 0:
        return _empty_cols()
    ra, rb, rab, lower = core[:, 0], core[:, 1], core[:, 2], core[:, 3]
    la, lb = c[:, 1] - c[:, 0], c[:, 3] - c[:, 2]
    mrabs = st[:, 0]
    return dict(coords=c, ra=ra, rb=rb, rab=rab, P2LL=rab.astype(np.float64) / np.maximum(lower, pseudo),
                FDR=st[:, 2], ES=rab / np.maximum(mrabs, pseudo),
                density=rab.astype(np.float64) / (la + lb) / tot * 10.0**9, hyp=_clamp(hyp), pop=_clamp(st[:, 3]),
                nbp=_clamp(st[:, 4]), px=_clamp(st[:, 5]), py=_clamp(st[:, 7]), esx=st[:, 6], esy=st[:, 8],
                mrabs=mrabs, mbps=st[:, 1], anc=st[:, 9:13])


def scipy_pvalues(cols, N, rpb, cis=True, threads=True):
    """Re-evaluate the tail probabilities of (few) surviving loops with scipy, argument for argument as the reference
    calls it (callCisLoops.py:246,310,329,332; callTransLoops.py:153,178,181), so the printed digits are the
    reference's.  Everything else in `cols` is already bit-identical.  N and rpb may be per-row arrays (rows of
    several files at once)."""
    n = cols["coords"].shape[0]
    if n == 0:
        return cols
    out = dict(cols)
    Nv = np.broadcast_to(np.asarray(N, dtype=np.float64), (n,))
    rv = np.broadcast_to(np.asarray(rpb, dtype=np.float64), (n,))
    args = _tail_args(cols, Nv, rv)
    pool = pvalue_pool() if (threads and n >= 2 * PV_CHUNK) else None
    futs, args = _tails_chunks(args, cis, pool)
    t = _gather_tails(futs, args, cis)
    out["hyp"], out["pop"], out["nbp"] = _clamp(t[:, 0]), _clamp(t[:, 1]), _clamp(t[:, 2])
    out["px"], out["esx"], out["py"], out["esy"] = t[:, 3], t[:, 4], t[:, 5], t[:, 6]
    return out


def _tail_args(cols, Nv, rv):
    c, anc = cols["coords"], cols["anc"]
    return (cols["rab"], cols["ra"], cols["rb"], Nv, rv, cols["mrabs"], cols["mbps"], c[:, 1] - c[:, 0], c[:, 3] - c[:, 2],
            anc[:, 0].astype(np.int64), anc[:, 1].astype(np.int64), anc[:, 2].astype(np.int64), anc[:, 3].astype(np.int64))


_pv_pool = None


class _PvPool:
    """Worker processes running `python -m cloops2_b200.pvalues --worker`, one feeder thread per process (blocked on
    the pipes, so the interpreter lock is free while a worker computes).  A task whose worker fails gets an exception
    (its caller then evaluates the chunk itself, _gather_tails); when no worker is left every task does."""

    def __init__(self, workers):
        import queue
        import subprocess
        import sys
        import threading
        self.q = queue.Queue()
        self.procs = []
        self.lock = threading.Lock()
        self.alive = 0
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            env[k] = "1"
        for _ in range(workers):
            # the worker's own pair of pipes (tasks down, results up); its stdin / stdout stay out of the protocol
            task_r, task_w = os.pipe()
            res_r, res_w = os.pipe()
            try:
                p = subprocess.Popen([sys.executable, "-m", "cloops2_b200.pvalues", "--worker", str(task_r), str(res_w)],
                                     stdin=subprocess.DEVNULL, pass_fds=(task_r, res_w), env=env, cwd=root)
            finally:
                os.close(task_r)
                os.close(res_w)
            p.cl2_tasks, p.cl2_results = os.fdopen(task_w, "wb"), os.fdopen(res_r, "rb")
            self.procs.append(p)
            self.alive += 1
            threading.Thread(target=self._feed, args=(p,), daemon=True).start()
        import atexit
        atexit.register(self.close)

    def _feed(self, p):
        try:
            while True:
                item = self.q.get()
                if item is None:
                    return
                fut, args, cis = item
                if not fut.set_running_or_notify_cancel():
                    continue
                try:
                    pvalues.send_msg(p.cl2_tasks, (args, cis))
                    ans = pvalues.recv_msg(p.cl2_results)
                    if ans is None or ans[0] != "ok":
                        raise RuntimeError("p-value worker failed: %r" % (ans,))
                    fut.set_result(ans[1])
                except Exception as e:
                    fut.set_exception(e)
                    if p.poll() is not None or isinstance(e, (OSError, ValueError)):
                        return                   # the process is gone: its share goes to the other feeders
        finally:
            with self.lock:
                self.alive -= 1
                last = self.alive <= 0
            if last:                             # nobody left to serve the queue
                self._fail_pending()

    def _fail_pending(self):
        import queue
        while True:
            try:
                item = self.q.get_nowait()
            except queue.Empty:
                return
            if item is not None and item[0].set_running_or_notify_cancel():
                item[0].set_exception(RuntimeError("no p-value worker left"))

    def submit(self, fn, *args, cis=True):
        from concurrent.futures import Future
        assert fn is pvalues.tails
        fut = Future()
        with self.lock:
            dead = self.alive <= 0
            if not dead:
                self.q.put((fut, args, cis))
        if dead:
            fut.set_exception(RuntimeError("no p-value worker left"))
        return fut

    def close(self):
        with self.lock:
            n = self.alive
        for _ in range(n):
            self.q.put(None)
        for p in self.procs:
            try:
                p.cl2_tasks.close()          # the worker leaves its loop at end of file
            except Exception:
                pass
        self.procs = []


def pvalue_pool():
    """One pool of worker PROCESSES per rank for the scipy evaluations of the exact mode (pvalues.py: the Boost-backed
    ufuncs keep the interpreter lock, so threads gain nothing).  The workers are plain child processes -- fresh
    interpreters with numpy / scipy only, which inherit the rank's core affinity and nothing else.  Work is handed
    over file by file as soon as a file's final loops are known, so it overlaps the GPU work on the remaining files.
    CL2_PV_WORKERS overrides the number of workers (0: evaluate in the calling thread)."""
    global _pv_pool
    if _pv_pool is None:
        from . import shard
        want = os.environ.get("CL2_PV_WORKERS")
        workers = int(want) if want not in (None, "") else max(2, shard.rank_cores() - 1)
        try:
            _pv_pool = _PvPool(workers) if workers > 0 else False
        except Exception:                          # no child processes to be had: the calling thread does the work
            _pv_pool = False
    return _pv_pool or None


def submit_pvalues(final, cis=True):
    """Queue the scipy tails of the rows of `final` (which carries its file's "N" and "rpb"); the futures are stored in
    final["_pending"] and the columns are replaced by wait_pvalues()."""
    n = final["coords"].shape[0]
    if n == 0:
        return final
    Nv = np.full(n, float(final["N"]))
    rv = np.full(n, float(final["rpb"]))
    futs, args = _tails_chunks(_tail_args(final, Nv, rv), cis, pvalue_pool())
    final["_pending"] = (futs, args, cis)
    return final


def wait_pvalues(finals):
    for f in finals:
        pend = f.pop("_pending", None)
        if not pend:
            continue
        futs, args, cis = pend
        t = _gather_tails(futs, args, cis)
        f["hyp"], f["pop"], f["nbp"] = _clamp(t[:, 0]), _clamp(t[:, 1]), _clamp(t[:, 2])
        f["px"], f["esx"], f["py"], f["esy"] = t[:, 3], t[:, 4], t[:, 5], t[:, 6]


def scipy_pvalues_many(finals, cis=True):
    """scipy_pvalues over the final loops of several files (every entry carries its own "N" and "rpb"), in place."""
    for f in finals:
        if "_pending" not in f:
            submit_pvalues(f, cis=cis)
    wait_pvalues(finals)


# Relative distance from a cut below which a device tail probability is not trusted to decide like scipy's would: the
# fp64 series of cl2_stats.cuh are accurate to ~1e-14, scipy 1.18 drifts up to ~1e-7 at bench-scale populations
# (DESIGN.md section 2); rows that close to a threshold are decided with scipy's own value in the exact mode.
PV_MARGIN = 1e-5


def _near(v, thresholds):
    near = np.zeros(v.shape[0], dtype=bool)
    for t in thresholds:
        near |= np.abs(v - t) <= 2.0 * PV_MARGIN * t
    return near


def exact_cuts(cols, N, rpb, cis=True, hic=False, peakPcut=1e-5, mark=True):
    """Exact-mode decisions.  The device applied its p-value cuts widened by PV_MARGIN (cl2_set_pvalue_margin); rows
    whose tails lie within the margin of any threshold the reference compares them with (estLoopSig :311, :347;
    markSigLoops :364-372 when mark=True; trans :200) get scipy's values, then the cuts are applied as written."""
    n = cols["coords"].shape[0]
    if n == 0:
        return cols
    if cis:
        near = _near(cols["hyp"], [1e-2] + ([1e-5] if mark else []))
        if not hic:
            near |= _near(cols["px"], [peakPcut]) | _near(cols["py"], [peakPcut])
        if mark:
            near |= _near(cols["pop"], [1e-5]) | _near(cols["nbp"], [1e-3 if hic else 1e-1])
    else:
        near = _near(cols["nbp"], [1e-10]) if mark else np.zeros(n, dtype=bool)
    if near.any():
        sub = scipy_pvalues(take(cols, near), N, rpb, cis=cis)
        cols = dict(cols)
        for k in ("hyp", "pop", "nbp", "px", "py", "esx", "esy"):
            v = cols[k].copy()
            v[near] = sub[k]
            cols[k] = v
    if cis:
        keep = ~(cols["hyp"] > 1e-2)                                                  # :311
        if not hic:
            keep &= (cols["px"] < peakPcut) & (cols["py"] < peakPcut)                 # :347
        if not keep.all():
            cols = take(cols, keep)
    return cols


COLS = ["coords", "ra", "rb", "rab", "P2LL", "FDR", "ES", "density", "hyp", "pop", "nbp", "px", "py", "esx", "esy",
        "mrabs", "mbps", "anc"]


def _empty_cols():
    d = {k: np.zeros(0) for k in COLS}
    d["coords"] = np.zeros((0, 4), dtype=np.int64)
    d["anc"] = np.zeros((0, 4))
    for k in ("ra", "rb", "rab"):
        d[k] = np.zeros(0, dtype=np.int64)
    return d


def take(cols, sel):
    return {k: v[sel] for k, v in cols.items()}


def mark_sig_cis(cols, hic=False):
    """markSigLoops (callCisLoops.py:357-376) -> int array significant."""
    sig = (cols["FDR"] <= 0.05) & (cols["ES"] >= 2) & (cols["hyp"] <= 1e-5) & (cols["pop"] <= 1e-5)
    sig &= cols["nbp"] < (1e-3 if hic else 1e-1)
    return sig.astype(np.int64)


def mark_sig_trans(cols):
    """trans markSigLoops (callTransLoops.py:196-206)."""
    return ((cols["nbp"] <= 1e-10) & (cols["FDR"] <= 0.05) & (cols["ES"] >= 2)).astype(np.int64)


def sel_sig(coords, es):
    """selSigLoops (callCisLoops.py:379-422) on the significant loops of one chromosome (native, libcl2.so)."""
    L = coords.shape[0]
    if L == 0:
        return np.zeros(0, dtype=np.int64)
    lib = _lib.load()
    coords = np.ascontiguousarray(coords, dtype=np.int64)
    es = np.ascontiguousarray(es, dtype=np.float64)
    out = np.empty(L, dtype=np.int64)
    n = ctypes.c_int64(0)
    rc = lib.cl2_sel_sig_loops(coords.ctypes.data_as(_lib._i64p), es.ctypes.data_as(_lib._f64p), L,
                               out.ctypes.data_as(_lib._i64p), ctypes.byref(n))
    if rc != 0:
        raise _lib.Cl2Error("cl2_sel_sig_loops failed rc=%d" % rc)
    return out[:n.value]


def select_twice(cols, significant):
    """The reference applies selSigLoops twice (callCisLoops.py:586-595)."""
    s = np.nonzero(significant > 0)[0]
    cur = take(cols, s)
    for _ in range(2):
        pick = sel_sig(cur["coords"], cur["ES"])
        cur = take(cur, pick)
    return cur
