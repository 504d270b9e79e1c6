"""Multi-GPU (needs >= 2 devices; skipped otherwise): chains sharded over ranks give the same numbers as one
GPU, the NCCL swap round (octo_pt_swap_round: ncclAllGather inside libocto_b200) agrees with the pure-host
decision path, and the device-resident parallel tempering with its ladder sharded over the ranks
(octo_pt_hmc_run_dist: resident explorer + one ncclAllGather + one decision kernel per round, in stream order)
reproduces the single-GPU run of all replicas bit for bit."""
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


def _worker_dist(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import octofitter_jl_b200 as octo
    import workloads
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    R = 64
    spec_p, th_p = workloads.one_planet_with_priors(100, 100, R, seed=2)
    model = octo.LogDensityModel(spec_p, device=rank)
    lad = np.linspace(0.0, 1.0, R) ** 3
    pt = octo.ParallelTempering(R, rank=rank, world=world, seed=11, beta=lad, backend="nccl", model=model)
    kw = dict(n_iter=2, n_leapfrog=4, step_size=1e-3, inv_mass=np.full(spec_p.D, 1e-4), seed=11)
    res = octo.device_parallel_tempering_dist(model, pt, th_p[pt.local_slice], lad, 12, **kw)
    # device-ordered swap rounds on their own: pairs on the device, nothing synchronised until the end
    tens, addr = pt.device_swap_state(torch, f"cuda:{rank}")
    rng = np.random.default_rng(3)
    st = torch.cuda.current_stream()
    keep = []
    for rnd in range(20):
        ref, tgt = rng.normal(-30, 3, R), rng.normal(-80, 25, R)
        pair = torch.tensor(np.stack([ref, tgt], axis=1)[pt.local_slice], dtype=torch.float64, device=f"cuda:{rank}")
        keep.append(pair)
        pt.swap_round_device(pair.data_ptr(), addr, st.cuda_stream)
    torch.cuda.synchronize()
    q.put((rank, {k: v for k, v in res.items() if isinstance(v, np.ndarray)}, tens["rung_of_chain"].cpu().numpy(),
           tens["beta_local"].cpu().numpy()))
    pt.close(); model.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_sharded_ladder_reproduces_the_single_gpu_run():
    import torch.multiprocessing as mp
    import octofitter_jl_b200 as octo
    import workloads
    world = 8 if _ngpu() >= 8 else (4 if _ngpu() >= 4 else 2)
    R = 64
    spec_p, th_p = workloads.one_planet_with_priors(100, 100, R, seed=2)
    model = octo.LogDensityModel(spec_p, device=0)
    lad = np.linspace(0.0, 1.0, R) ** 3
    kw = dict(n_iter=2, n_leapfrog=4, step_size=1e-3, inv_mass=np.full(spec_p.D, 1e-4), seed=11)
    ref = octo.device_parallel_tempering(model, th_p, lad, 12, **kw)
    model.close()
    assert ref["swap_counts"].sum() > 0
    host = octo.ParallelTempering(R, seed=11, beta=lad, backend="local")
    rng = np.random.default_rng(3)
    for rnd in range(20):
        host.swap_round(rng.normal(-30, 3, R), rng.normal(-80, 25, R))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_dist, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    out = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    nl = R // world
    for rank, res, rung_dev, beta_dev in out:
        sl = slice(rank * nl, (rank + 1) * nl)
        for k in ("theta_final", "logpost_tempered", "loglike", "beta", "rung", "accept"):
            assert np.array_equal(res[k], ref[k][sl]), (rank, k)
        for k in ("swap_counts", "cold_trace"):
            assert np.array_equal(res[k], ref[k]), (rank, k)
        assert np.array_equal(rung_dev, host.chain_of_replica)
        assert np.array_equal(beta_dev, lad[host.chain_of_replica][sl])
