"""The trajectory-resident HMC kernel (k_hmc_resident: a chain group's whole run — transitions x leapfrogs — inside one
launch, state in shared memory) against the launch-per-leapfrog explorer it replaces.  With both on the same geometry
(same sub-lane count, same CTA width, no epoch splits across CTAs) the two must agree BIT FOR BIT: same evaluation
body, same per-coordinate arithmetic (octo_hmc_dev.cuh), same counter-based random streams."""
import os

import numpy as np
import pytest

import octofitter_jl_b200 as octo
import workloads
from helpers import reference_test_system

pytestmark = pytest.mark.gpu


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
    for k in ("theta", "logpost", "theta_final", "logpost_final", "accept"):
        assert np.array_equal(res[k], ref[k]), k
    assert 0.3 < res["accept_rate"] <= 1.0
    model.close()
    spec_p, th_p = workloads.one_planet_with_priors(100, 100, 200, seed=2)
    model = octo.LogDensityModel(spec_p)
    res, ref = _both(model, th_p, 4, step_size=1e-3, n_leapfrog=6, inv_mass=np.full(spec_p.D, 1e-4), seed=5)
    for k in ("theta", "logpost", "theta_final", "logpost_final", "accept"):
        assert np.array_equal(res[k], ref[k]), k
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
