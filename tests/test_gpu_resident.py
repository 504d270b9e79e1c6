"""The trajectory-resident HMC kernel (k_hmc_resident: a chain group's whole run — transitions x leapfrogs — inside one
launch, state in shared memory) against the launch-per-leapfrog explorer it replaces.  With both on the same geometry
(same sub-lane count, same CTA width, no epoch splits across CTAs) the two must produce the same TRAJECTORIES bit for bit —
positions, gradients, accept decisions: same evaluation body, same per-coordinate arithmetic (octo_hmc_dev.cuh), same
counter-based random streams.  The log-posterior VALUES they report may differ in the last bit (relative 2e-16 seen):
since round 2 the evaluation is inlined into each kernel and the compiler contracts a*b+c on the value-only path
per instantiation (building with -fmad=false makes them identical again, at +4 % per leapfrog)."""
import os

import numpy as np
import pytest

import octofitter_jl_b200 as octo
import workloads
from helpers import reference_test_system

pytestmark = pytest.mark.gpu


def _same(res, ref, keys):
    for k in keys:
        if k.startswith("logpost"):
            np.testing.assert_allclose(res[k], ref[k], rtol=4e-15, atol=0, err_msg=k)
        else:
            assert np.array_equal(res[k], ref[k]), k


def _both(model, th0, n_iter, **kw):
    res = octo.device_hmc(model, th0, n_iter, **kw)
    n0 = model.kernel_launches
    octo.device_hmc(model, th0, 1, **kw)
    assert model.kernel_launches - n0 == 1                       # one launch for the whole run
    os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"] = "1"
    try:
        ref = octo.device_hmc(model, th0, n_iter, **kw)
    finally:
        del os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"]
    return res, ref


@pytest.mark.parametrize("force", ["1,1,1", "4,1,1", "32,1,1"])
def test_resident_equals_launch_per_leapfrog_bit_for_bit(monkeypatch, force):
    monkeypatch.setenv("OCTO_B200_FORCE", force)
    # the reference's 11-D test model (8 epochs) and the C2 tables with priors (200 epochs, D = 14, tempering off)
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(4)
    params, _ = model.guess_starting_position(rng, N=40_000, batch=20_000)
    start = model.link(params)
    inv_mass = octo.diagonal_metric(model, start)
    th0 = np.asfortranarray(start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((77, spec.D)))
    res, ref = _both(model, th0, 9, step_size=0.15, n_leapfrog=5, inv_mass=inv_mass, seed=99)
    _same(res, ref, ("theta", "logpost", "theta_final", "logpost_final", "accept"))
    assert 0.3 < res["accept_rate"] <= 1.0
    model.close()
    spec_p, th_p = workloads.one_planet_with_priors(100, 100, 200, seed=2)
    model = octo.LogDensityModel(spec_p)
    res, ref = _both(model, th_p, 4, step_size=1e-3, n_leapfrog=6, inv_mass=np.full(spec_p.D, 1e-4), seed=5)
    _same(res, ref, ("theta", "logpost", "theta_final", "logpost_final", "accept"))
    model.close()


def test_resident_parallel_tempering_equals_launch_per_leapfrog(monkeypatch):
    monkeypatch.setenv("OCTO_B200_FORCE", "4,1,1")
    spec_p, th_p = workloads.one_planet_with_priors(100, 100, 64, seed=2)
    model = octo.LogDensityModel(spec_p)
    lad = np.linspace(0.0, 1.0, 64) ** 3
    kw = dict(n_iter=2, n_leapfrog=4, step_size=1e-3, inv_mass=np.full(spec_p.D, 1e-4), seed=11)
    res = octo.device_parallel_tempering(model, th_p, lad, 12, **kw)
    os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"] = "1"
    try:
        ref = octo.device_parallel_tempering(model, th_p, lad, 12, **kw)
    finally:
        del os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"]
    for k in res:
        if isinstance(res[k], np.ndarray):
            assert np.array_equal(res[k], ref[k]), k
    assert res["swap_accept"].sum() > 0
    model.close()


def test_resident_automatic_geometry_two_planets_and_ragged_batch():
    """Automatic sub-lane choice; a 2-planet model; batch sizes that leave the last chain group ragged."""
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(3)
    params, _ = model.guess_starting_position(rng, N=40_000, batch=20_000)
    start = model.link(params)
    inv_mass = octo.diagonal_metric(model, start)
    for n in (1, 5, 33, 300):
        th0 = np.asfortranarray(start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((n, spec.D)))
        res, ref = _both(model, th0, 6, step_size=0.15, n_leapfrog=4, inv_mass=inv_mass, seed=7)
        assert np.allclose(res["theta_final"], ref["theta_final"], rtol=1e-8, atol=1e-10)
        assert np.array_equal(res["accept"], ref["accept"])
    model.close()


def test_sharded_ladder_entry_point_single_rank_equals_plain_run():
    """octo_pt_hmc_run_dist on a single rank (world = 1: the all-gather degenerates to a copy) takes the same decisions
    and moves every chain exactly like octo_pt_hmc_run — the pair packing in the resident kernel, the decision kernel on
    the gathered buffer and the replicated rung assignment are the code the multi-GPU run uses (tests/test_gpu_multi.py)."""
    spec_p, th_p = workloads.one_planet_with_priors(100, 100, 64, seed=2)
    model = octo.LogDensityModel(spec_p)
    lad = np.linspace(0.0, 1.0, 64) ** 3
    kw = dict(n_iter=2, n_leapfrog=4, step_size=1e-3, inv_mass=np.full(spec_p.D, 1e-4), seed=11)
    ref = octo.device_parallel_tempering(model, th_p, lad, 12, **kw)
    pt = octo.ParallelTempering(64, seed=11, beta=lad, backend="local", model=model)
    res = octo.device_parallel_tempering_dist(model, pt, th_p, lad, 12, **kw)
    for k in ref:
        if isinstance(ref[k], np.ndarray):
            assert np.array_equal(res[k], ref[k]), k
    assert res["swap_counts"].sum() > 0
    pt.close(); model.close()


def test_device_ordered_swap_round_matches_host_decisions():
    """octo_pt_swap_round_device (pairs on the device, decisions in stream order, no host sync) against the pure-host
    octo_pt_decide fed with the same values: same acceptances, same rung assignment, round after round."""
    import torch
    spec, x = workloads.config("C4")
    R = x.shape[0]
    model = octo.LogDensityModel(spec)
    lad = np.linspace(0.0, 1.0, R) ** 2
    dev = octo.ParallelTempering(R, seed=5, beta=lad, backend="local", model=model)
    host = octo.ParallelTempering(R, seed=5, beta=lad, backend="local")
    tens, addr = dev.device_swap_state(torch, "cuda")
    rng = np.random.default_rng(3)
    st = torch.cuda.current_stream()
    d_pairs = []
    expect = []
    for rnd in range(40):
        ref, tgt = rng.normal(-30, 3, R), rng.normal(-80, 25, R)
        pair = torch.tensor(np.stack([ref, tgt], axis=1), dtype=torch.float64, device="cuda")
        d_pairs.append(pair)
        dev.swap_round_device(pair.data_ptr(), addr, st.cuda_stream)          # enqueued only
        host.swap_round(ref, tgt)
        expect.append(host.chain_of_replica.copy())
    torch.cuda.synchronize()
    rung_of_chain = tens["rung_of_chain"].cpu().numpy()
    assert np.array_equal(rung_of_chain, expect[-1])
    assert np.array_equal(tens["chain_of_rung"].cpu().numpy()[rung_of_chain], np.arange(R))
    assert np.array_equal(tens["beta_local"].cpu().numpy(), lad[rung_of_chain])
    assert tens["swap_count"].sum().item() > 0
    dev.close(); model.close()
