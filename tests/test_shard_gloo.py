"""world_size-2 gloo test of the multi-rank plumbing (assignment, stats all-reduce, gather to rank 0)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from cloops2_b200 import shard


def test_lpt_assignment_is_balanced_and_deterministic():
    from cloops2_b200.synth import HG38
    keys = [n for n, _ in HG38]
    w = [l for _, l in HG38]
    for world in (1, 2, 4, 8):
        parts = shard.assign(keys, w, world)
        assert sorted(sum(parts, [])) == sorted(keys)
        loads = [sum(w[keys.index(k)] for k in p) for p in parts]
        assert max(loads) <= 1.13 * (sum(w) / world) or world == 1
        assert parts == shard.assign(keys, w, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w = shard.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    keys = ["a", "b", "c", "d", "e"]
    mine = shard.assign(keys, [50, 40, 30, 20, 10], world)[rank]
    stats = shard.allreduce_stats(np.array([len(mine), rank + 1], dtype=np.int64))
    res = shard.gather_results({k: np.full(2, rank) for k in mine})
    t = shard.max_over_ranks(float(rank))
    shard.barrier()
    q.put((rank, mine, stats.tolist(), sorted(res.keys()), t))
    import torch.distributed as dist
    dist.destroy_process_group()


def _filter_worker(rank, world, port, q, predir, fout):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), CL2_PIN="0")
    shard.init_from_env(backend="gloo")
    from cloops2_b200.callCisLoops import callCisLoops
    res = callCisLoops(predir, fout, None, eps=[200], minPts=[5], filter=True)     # must return before any GPU work
    stats = shard.allreduce_stats(np.array([1], dtype=np.int64))                     # the ranks are still in step
    q.put((rank, res, stats.tolist()))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_filter_early_return_is_taken_by_every_rank(tmp_path):
    """`-filter` with a non-empty OUT_filtered directory (callCisLoops.py:518-524): only rank 0 looks at the directory,
    but every rank must return -- a rank that carried on into the collectives of the run would hang the job (and, with
    -trans, pair them with another phase's).  No GPU is touched on that path, so this runs on the CPU ranks of gloo."""
    import json
    predir, fout = str(tmp_path / "pre"), str(tmp_path / "out")
    os.mkdir(predir)
    json.dump({"Unique PETs": 10, "data": {"cis": {}, "trans": {}}}, open(predir + "/petMeta.json", "w"))
    os.mkdir(fout + "_filtered")
    open(fout + "_filtered/leftover.ixy", "w").write("x")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_filter_worker, args=(r, 2, port, q, predir, fout)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(0, None, [2]), (1, None, [2])]


def test_wait_policy_by_core_share():
    """Workers and how they wait, by the host cores a rank can count on (shard.wait_policy): 8 spinning workers from 4
    cores up (16-core single-GPU box, 4 cores per rank on an 8-GPU node), one worker fewer than cores above 8, sleeping
    waits only for a rank that is truly short of cores."""
    assert shard.wait_policy(16) == (8, 0)
    assert shard.wait_policy(8) == (7, 0)
    assert shard.wait_policy(4) == (8, 0)
    assert shard.wait_policy(2) == (8, 1)
    assert shard.wait_policy(0) == (8, 1)


def test_two_rank_gloo_roundtrip():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, s0, k0, t0), (r1, m1, s1, k1, t1) = out
    assert sorted(m0 + m1) == ["a", "b", "c", "d", "e"] and not set(m0) & set(m1)
    assert s0 == s1 == [5, 3]
    assert k0 == ["a", "b", "c", "d", "e"] and k1 == []
    assert t0 == t1 == 1.0


def test_map_units_order_and_error_propagation():
    """Worker-thread fan-out used inside a rank: results come back in unit order; the first failure is re-raised."""
    import pytest
    ctxs = ["c0", "c1", "c2"]
    out = shard.map_units(ctxs, list(range(20)), lambda c, u: (u * u, c))
    assert [o[0] for o in out] == [u * u for u in range(20)]
    assert {o[1] for o in out} <= set(ctxs)
    assert shard.map_units(ctxs[:1], [1, 2], lambda c, u: u + 1) == [2, 3]

    def boom(c, u):
        if u == 7:
            raise ValueError("unit 7")
        return u
    with pytest.raises(ValueError):
        shard.map_units(ctxs, list(range(20)), boom)
