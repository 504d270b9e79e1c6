"""N>1 host logic on CPU: world_size-2 gloo process group, replicas block-partitioned over ranks, the swap
decisions (the library's pure-host octo_pt_decide) identical on every rank and equal to a single-process run."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import octofitter_jl_b200 as octo


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _lls(R, rnd):
    rng = np.random.default_rng(1000 + rnd)
    return rng.normal(-50, 5, R), rng.normal(-80, 20, R)


def _worker(rank, world, port, R, rounds, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pt = octo.ParallelTempering(R, rank=rank, world=world, seed=7, backend="gloo")
    hist = []
    for rnd in range(rounds):
        ref, tgt = _lls(R, rnd)
        # each replica's likelihood depends on the state it carries (its replica index), not on its rung
        acc = pt.swap_round(ref[pt.local_slice], tgt[pt.local_slice])
        hist.append((acc.copy(), pt.chain_of_replica.copy()))
    q.put((rank, hist))
    dist.destroy_process_group()


def test_two_rank_swaps_match_single_process():
    R, rounds, world = 16, 12, 2
    single = octo.ParallelTempering(R, seed=7, backend="local")
    ref_hist = []
    for rnd in range(rounds):
        ref, tgt = _lls(R, rnd)
        acc = single.swap_round(ref, tgt)
        ref_hist.append((acc.copy(), single.chain_of_replica.copy()))
    assert any(a.sum() > 0 for a, _ in ref_hist), "test should exercise accepted swaps"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, rounds, q)) for r in range(world)]
    [p.start() for p in procs]
    out = dict(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    for r in range(world):
        for (a, c), (a0, c0) in zip(out[r], ref_hist):
            assert np.array_equal(a, a0) and np.array_equal(c, c0)


def test_swap_rules():
    R = 6
    pt = octo.ParallelTempering(R, seed=1, backend="local")
    # overwhelmingly favourable swap on pair (0,1): replica on rung 0 has the much better target
    ref = np.zeros(R); tgt = np.array([0.0, -1e6, -1e6, -1e6, -1e6, -1e6])
    acc = pt.swap_round(ref, tgt)            # round 0 pairs (0,1),(2,3),(4,5)
    assert acc[0] == 1 and acc[1] == 0 and acc[3] == 0
    assert sorted(pt.chain_of_replica) == list(range(R))
    assert pt.chain_of_replica[0] == 1 and pt.chain_of_replica[1] == 0
    acc = pt.swap_round(ref, tgt)            # round 1 pairs (1,2),(3,4): replica 0 (now rung 1) climbs again
    assert acc[1] == 1 and pt.chain_of_replica[0] == 2
    with pytest.raises(ValueError):
        octo.ParallelTempering(7, world=2)
    bad = octo.ParallelTempering(4, backend="local")
    bad.chain_of_replica[:] = 0
    with pytest.raises(RuntimeError, match="permutation"):
        bad.swap_round(np.zeros(4), np.zeros(4))
