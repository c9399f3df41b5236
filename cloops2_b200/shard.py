"""
Multi-GPU plumbing for the callLoops path: one process per GPU (torchrun), chromosome(-pair) files as the unit of
work.  The reference fans the same units out over joblib workers (callCisLoops.py:134-137, 561-572;
callTransLoops.py:95-97, 237-243); no data flows between units, so there is no data-path collective -- only a small
all-reduce (NCCL over NVLink when ranks own GPUs, gloo on CPU in the tests) of the PET / candidate / loop counters,
and a host gather of the per-chromosome loop tables to rank 0.
"""
import os

import numpy as np


def rank_world():
    """(rank, world) from torch.distributed if it is initialised, else from the torchrun environment, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def init_from_env(backend=None):
    """Initialise torch.distributed when launched by torchrun (WORLD_SIZE > 1).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
        pin_rank_cores()
    return dist.get_rank(), dist.get_world_size()


_pinned = [False]


def rank_cores():
    """Host cores this rank can count on: its own share once pinned, else an equal split of what the process may use."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 8
    if _pinned[0]:
        return max(1, cores)
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return max(1, cores // max(1, world))


def pin_rank_cores():
    """Give every rank of a node its own share of the host cores (CL2_PIN=0 leaves the affinity alone).  The ranks'
    worker threads wait actively; left to roam over all cores they keep displacing each other's launching threads,
    which a 32-core host with 8 ranks shows as a third of the throughput lost."""
    if os.environ.get("CL2_PIN", "1") in ("0", ""):
        return
    try:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        nlocal = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(1, nlocal)
        if nlocal > 1 and per >= 2:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
            _pinned[0] = True
    except (AttributeError, OSError, ValueError):
        pass


def ixy_rows(f):
    """PET count of a .ixy without loading it: joblib stores the raw int64[N,2] buffer after a small pickle header
    (~240 bytes), so rows ~= file size / 16.  Only used to balance shards."""
    try:
        return max(0, (os.path.getsize(f) - 240) // 16)
    except OSError:
        return 0


def assign(keys, weights, world):
    """Greedy longest-processing-time partition of `keys` by `weights` into `world` bins.  Every rank computes the
    same assignment (deterministic tie-breaks); within a bin the original key order is kept."""
    order = sorted(range(len(keys)), key=lambda i: (-int(weights[i]), i))
    load = [0] * world
    owner = [0] * len(keys)
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        owner[i] = r
        load[r] += int(weights[i])
    return [[keys[i] for i in range(len(keys)) if owner[i] == r] for r in range(world)]


_comm = {}


def _library_comm():
    """The library's own NCCL communicator (cl2_comm_init), created on first use.  Rank 0's unique id travels over the
    launcher's process group -- any host channel would do; a C caller would use a file or MPI."""
    if "h" not in _comm:
        import ctypes
        import torch.distributed as dist
        from . import _lib
        rank, world = rank_world()
        ctx = _lib.default_context()
        idb = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            ctx.check(ctx.lib.cl2_comm_unique_id(idb.ctypes.data_as(_lib._u8p)), "cl2_comm_unique_id")
        box = [idb.tobytes()]
        dist.broadcast_object_list(box, src=0)
        idb = np.frombuffer(box[0], dtype=np.uint8).copy()
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.cl2_comm_init(ctx.h, idb.ctypes.data_as(_lib._u8p), rank, world, ctypes.byref(h)), "cl2_comm_init")
        _comm["h"], _comm["ctx"] = h, ctx
    return _comm["ctx"], _comm["h"]


def allreduce_stats(vals):
    """Sum an int64 vector over ranks (the path's only collective): cl2_stats_allreduce (NCCL inside libcl2.so) when
    the ranks own GPUs, the launcher's gloo group in the CPU tests."""
    rank, world = rank_world()
    vals = np.ascontiguousarray(vals, dtype=np.int64).copy()
    if world == 1:
        return vals
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        from . import _lib
        ctx, h = _library_comm()
        ctx.check(ctx.lib.cl2_stats_allreduce(ctx.h, h, vals.ctypes.data_as(_lib._i64p), int(vals.shape[0])),
                  "cl2_stats_allreduce")
        return vals
    import torch
    t = torch.from_numpy(vals)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def broadcast_flag(v):
    """Rank 0's integer, on every rank (decisions that every rank has to take the same way)."""
    rank, world = rank_world()
    if world == 1:
        return int(v)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(v)], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0)
    return int(t.item())


def load_pinned(ctx, f):
    """Rows of a `.ixy` file in the worker's page-locked staging buffer, copied straight from the memory map joblib
    opens (no pageable intermediate).  The buffer belongs to `ctx` and is reused by its next file: the upload
    (cl2_points_upload) has finished with it when it returns."""
    from .io import ixyMap
    mat = ixyMap(f)
    out = ctx.staging(mat.shape, np.int64)
    np.copyto(out, mat, casting="unsafe")
    return out


def gather_results(results):
    """Per-chromosome result dicts of every rank merged on rank 0 (other ranks get {})."""
    rank, world = rank_world()
    if world == 1:
        return results
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(results, bucket, dst=0)
    if rank != 0:
        return {}
    merged = {}
    for part in bucket:
        merged.update(part)
    return merged


def max_over_ranks(x):
    """max of a float over ranks (timing)."""
    rank, world = rank_world()
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    rank, world = rank_world()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def wait_policy(share):
    """(worker threads, wait mode) for a rank that can count on `share` host cores.  8 workers that spin while they
    wait (mode 0) from 4 cores up; below that they sleep (mode 1).  Measured per 100 M-PET step (bench.py
    --extra-sync-modes; --host-cores 4 confines one rank like one of 8 on a 32-core node):
        1 GPU, 16 cores           spin 57-59 ms
        1 GPU, 4 cores            spin 59.1   poll + yield 63.3   sleep 66.2 (102 ms of host CPU per step instead of 210)
        8 GPUs, 4 cores per rank  spin 58.4   poll + yield 61.4   sleep 67.7
    (Before scope-bound scratch moved into the context's arena, eight ranks took 109 ms per step whichever way they
    waited: the ranks' cudaMallocAsync / cudaFreeAsync calls, not their waits, were what the node ran out of.)"""
    share = max(1, int(share))
    nthreads = 8 if share < 8 else max(2, min(8, share - 1))
    return nthreads, (0 if share >= 4 else 1)


def worker_contexts(device=None, nthreads=None):
    """One cl2 context (= CUDA stream + scratch) per host worker thread of this rank.  Chromosome files are
    independent units, so a rank processes several at once: while one thread's kernels run, another thread does its
    host-side bookkeeping (the per-file loop runs on the native side of the ABI, cl2_call_loops, so the workers hold no
    interpreter lock while they wait).  CL2_THREADS overrides the number of workers, CL2_BLOCKING_SYNC=0/1/2 how they
    wait (wait_policy)."""
    from . import _lib
    dflt_threads, mode = wait_policy(rank_cores())
    if nthreads is None:
        nthreads = int(os.environ.get("CL2_THREADS", str(dflt_threads)))
    if os.environ.get("CL2_BLOCKING_SYNC") is not None:
        mode = int(os.environ["CL2_BLOCKING_SYNC"] or 0)
    set_sync_mode(mode)
    nthreads = max(1, nthreads)
    first = _lib.default_context(device)
    # the workers of a device are kept for the life of the process: a second callCisLoops / callTransLoops call reuses
    # their streams, scratch reservations and page-locked staging buffers
    pool = _worker_pool.setdefault(first.device, [first])
    while len(pool) < nthreads:
        pool.append(_lib.Context(first.device))
    return pool[:nthreads]


_worker_pool = {}


_sync_mode = [0]


def set_sync_mode(mode):
    """How worker threads wait for their streams (cl2_set_sync_mode): 0 spinning (lowest latency), 1 / True sleeping
    (leaves the cores to other host work, e.g. the scipy pool of the exact mode), 2 polling with a yield between polls
    (more workers than cores).  Returns the previous mode."""
    from . import _lib
    prev = _sync_mode[0]
    _sync_mode[0] = 1 if mode is True else (int(mode) if int(mode) in (1, 2) else 0)
    _lib.load().cl2_set_sync_mode(_sync_mode[0])
    return prev


def map_units(ctxs, units, fn):
    """fn(ctx, unit) for every unit, results in unit order; units are handed to the worker threads dynamically."""
    if len(ctxs) <= 1 or len(units) <= 1:
        return [fn(ctxs[0], u) for u in units]
    import queue
    import threading
    q = queue.Queue()
    for i, u in enumerate(units):
        q.put((i, u))
    out = [None] * len(units)
    errs = []

    def work(ctx):
        while not errs:
            try:
                i, u = q.get_nowait()
            except queue.Empty:
                return
            try:
                out[i] = fn(ctx, u)
            except BaseException as e:      # surface the first failure in the caller
                errs.append(e)
                return

    ths = [threading.Thread(target=work, args=(c,)) for c in ctxs]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
    return out
