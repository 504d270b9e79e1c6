"""Multi-GPU (needs >= 2 devices; skipped otherwise): chains sharded over ranks give the same numbers as one
GPU, and the NCCL swap round (octo_pt_swap_round: ncclAllGather inside libocto_b200) agrees with the pure-host
decision path."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import octofitter_jl_b200 as octo
    import workloads
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    spec, x = workloads.config("C4")                 # 64 tempered replicas
    R = x.shape[0]
    model = octo.LogDensityModel(spec, device=rank)
    pt = octo.ParallelTempering(R, rank=rank, world=world, seed=11, backend="nccl", model=model)
    ll_local = model.ln_like(x[pt.local_slice])      # shard: independent chains, no collective
    hist = []
    rng = np.random.default_rng(5)
    ref_all = rng.normal(-30, 3, R)
    for rnd in range(6):
        acc = pt.swap_round(ref_all[pt.local_slice], ll_local)
        hist.append((acc.copy(), pt.chain_of_replica.copy()))
    q.put((rank, ll_local, hist))
    pt.close(); model.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_sharded_chains_and_nccl_swap_round():
    import torch.multiprocessing as mp
    import octofitter_jl_b200 as octo
    import workloads
    world = 8 if _ngpu() >= 8 else (4 if _ngpu() >= 4 else 2)     # C4 proper on an 8-GPU box: 64 replicas, 8 per GPU
    spec, x = workloads.config("C4")
    R = x.shape[0]
    model = octo.LogDensityModel(spec, device=0)
    ll_all = model.ln_like(x)
    model.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    # sharding changes the launch geometry (hence the summation order), not the numbers
    ll_cat = np.concatenate([r[1] for r in res])
    assert np.max(np.abs(ll_cat - ll_all) / np.abs(ll_all)) < 1e-12
    # the NCCL all-gather path takes the same decisions as the pure-host path fed with the same values
    single = octo.ParallelTempering(R, seed=11, backend="local")
    rng = np.random.default_rng(5)
    ref_all = rng.normal(-30, 3, R)
    ref_hist = []
    for rnd in range(6):
        acc = single.swap_round(ref_all, ll_cat)
        ref_hist.append((acc.copy(), single.chain_of_replica.copy()))
    assert any(a.sum() > 0 for a, _ in ref_hist)
    for rank, ll_local, hist in res:
        for (a, c), (a0, c0) in zip(hist, ref_hist):
            assert np.array_equal(a, a0) and np.array_equal(c, c0)
