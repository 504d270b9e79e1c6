"""GPU tests of the device-side standard parameterisation (K0 forward/backward around K1): libocto_b200's
octo_logpost_grad against the oracle and the mpmath goldens."""
import os

import numpy as np
import pytest

import octofitter_jl_b200 as octo
from helpers import grad_err, load_post, post_cases, reference_test_system, rel_err

pytestmark = pytest.mark.gpu
LOGP_RTOL, GRAD_RTOL = 1e-10, 1e-8


def _model_from_golden(d, spec, consts):
    import ctypes as C
    lib = octo.load_library()
    h = C.c_void_p()
    assert lib.octo_create(C.byref(consts), C.byref(spec.packed.layout), spec.packed.blocks, spec.packed.n_blocks, 0,
                           C.byref(h)) == 0, lib.octo_last_error()
    assert lib.octo_set_parameterization(h, spec.priors, spec.D, spec.defs) == 0, lib.octo_last_error()
    return lib, h


def _logpost(lib, h, D, th, grad=True):
    th = np.asfortranarray(np.atleast_2d(np.asarray(th, dtype=np.float64)))
    n = th.shape[0]
    lp = np.empty(n); g = np.empty((n, D), order="F") if grad else None
    rc = lib.octo_logpost_grad(h, th.ctypes.data, n, n, lp.ctypes.data, g.ctypes.data if grad else None)
    assert rc == 0, lib.octo_last_error()
    return lp, g


@pytest.mark.parametrize("name", post_cases())
def test_cuda_logpost_matches_golden_and_oracle(oracle_lib, name):
    d, spec, consts = load_post(name)
    lib, h = _model_from_golden(d, spec, consts)
    lp, g = _logpost(lib, h, spec.D, d["theta_t"])
    assert rel_err(lp[0], d["lp"]) < LOGP_RTOL and grad_err(g, d["grad"]).max() < GRAD_RTOL
    rng = np.random.default_rng(8)
    th = np.array(d["theta_t"])[None, :] + 0.05 * rng.standard_normal((67, spec.D))
    lp, g = _logpost(lib, h, spec.D, th)
    lpv, _ = _logpost(lib, h, spec.D, th, grad=False)
    lp_o, g_o = oracle_lib.logpost(spec, consts, th, threads=4)
    assert rel_err(lp, lp_o).max() < LOGP_RTOL and rel_err(lpv, lp_o).max() < LOGP_RTOL
    assert grad_err(g, g_o).max() < GRAD_RTOL
    lib.octo_destroy(h)


def test_reference_test_model_on_device(oracle_lib):
    """The reference's 11-D test model through the host mirror: ℓπcallback / ∇ℓπcallback / invlink."""
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    assert model.D == 11 and model.dimension() == 11
    rng = np.random.default_rng(2)
    th = rng.normal(0, 0.8, (129, 11))
    th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(129)       # keep plx near its tight prior
    lp, g = model.ℓπcallback_grad(th)
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    assert rel_err(lp, lp_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    assert rel_err(model.ℓπcallback(th), lp_o).max() < LOGP_RTOL
    assert np.allclose(model.invlink(th), oracle_lib.invlink(spec, th), rtol=1e-14, atol=0)
    lp1, g1 = model.logdensity_and_gradient(th[0])
    assert rel_err(lp1, lp_o[0]) < LOGP_RTOL and model(th[0]) == pytest.approx(lp1, rel=1e-12)
    # `@test model.ℓπcallback(start) > -1000` style sanity at a sensible point (test/integration/sampling.jl:76)
    good = np.array([np.log(1.1), np.log(49.9), -1.99, -2.0, -0.5, 0.8, 0.6, 0.95, 0.28, -0.99, -0.13])
    assert np.isfinite(model.ℓπcallback(good))
    model.close()


def test_device_logpost_invalid_and_healed(oracle_lib):
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    th = np.zeros((4, spec.D)); th[:, 5:] = 0.7; th[:, 4] = -0.5      # (i = π/2 exactly makes tp ill-conditioned)
    th[1, 3] = np.nan
    th[2, 3] = 800.0                       # e clamps to its bound: healed prior, finite but hugely negative
    th[3, 9] = np.inf
    lp, g = model.ℓπcallback_grad(th)
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th)
    assert np.isfinite(lp[0]) and lp[1] == -np.inf and lp[3] == -np.inf and np.all(g[1] == 0) and np.all(g[3] == 0)
    assert lp[2] == lp_o[2] and lp[2] < -1e300
    assert rel_err(lp[0], lp_o[0]) < LOGP_RTOL and grad_err(g[:1], g_o[:1]).max() < GRAD_RTOL
    model.close()


def test_parameterisation_required():
    import workloads
    spec, x = workloads.config("C1")
    model = octo.LogDensityModel(spec)
    with pytest.raises(octo.OctoError, match="bare variable names"):
        model.ℓπcallback(np.zeros(3))
    lp = np.empty(1)
    assert model._lib.octo_logpost_grad(model._h, x.ctypes.data, 1, 1, lp.ctypes.data, None) == 4   # OCTO_ERR_STATE
    model.close()


def test_guess_starting_position_batched(oracle_lib):
    """initialization.jl:14-66 as a batched value-only consumer: best of N prior draws, evaluated on the device in
    large batches; the winner's log-posterior agrees with the oracle and link/invlink round-trip."""
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(1)
    params, lp = model.guess_starting_position(rng, N=40_000, batch=16384)
    th = model.link(params)
    assert np.allclose(model.invlink(th), params, rtol=1e-9, atol=1e-12)
    lp_o = oracle_lib.logpost(spec, octo.default_constants(), th, grad=False)[0]
    assert abs(lp - lp_o) <= 1e-10 * abs(lp_o)
    # a uniform draw from the prior is (much) worse than the best of 40k
    assert lp > np.median(model.ℓπcallback(model.link(model.sample_priors(rng, 2000))))
    model.close()


def test_batched_hmc_end_to_end():
    """The whole stack as a sampler sees it (reference: "Test fitting a chain", test/integration-tests.jl:52-58):
    256 chains of static HMC on the reference's 11-D test model, every leapfrog one device call.  Chains must move,
    accept, stay at plausible log-posterior, and the posterior must cover the orbit that generated the fixture."""
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(4)
    params, lp0 = model.guess_starting_position(rng, N=60_000, batch=20_000)
    start = model.link(params)
    inv_mass = octo.diagonal_metric(model, start)            # diagonal preconditioner from the local curvature
    th0 = start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((256, spec.D))
    res = octo.batched_hmc(model, th0, 150, step_size=0.15, n_leapfrog=12, rng=rng, inv_mass=inv_mass)
    assert 0.4 < res["accept_rate"] <= 1.0
    lp_end = res["logpost"][-1]
    assert np.all(np.isfinite(lp_end)) and np.median(lp_end) > -1000            # `@test all(chain[:logpost] .> -1000)`
    assert np.median(lp_end) >= lp0 - 30
    nat = model.invlink(res["theta"][-1])
    names = list(spec.theta_names)
    a, e = nat[:, names.index("b.a")], nat[:, names.index("b.e")]
    assert 5.0 < np.median(a) < 40.0 and 0.0 <= np.median(e) < 0.9            # generating orbit: a = 12, e = 0.11
    assert res["n_gradient_calls"] == 1 + 150 * 12
    model.close()


def test_batched_parallel_tempering_single_rank():
    """N2: tempered replicas advanced in lockstep on the device + deterministic even-odd swaps.  The tempering
    reference is the prior-only model, built as the reference does (observations blanked, same parameters:
    src/cross-validation.jl:60-99, ext/OctofitterPigeonsExt:61-67)."""
    system = reference_test_system()
    spec = octo.ModelSpec(system)
    model = octo.LogDensityModel(spec)
    # prior-only twin: same variables, no tables
    b0 = system.planets[0]
    b_ref = octo.Planet(name=b0.name, variables=dict(b0.var_specs), observations=[])
    ref_model = octo.LogDensityModel(octo.ModelSpec(octo.System(name="ref", variables=dict(system.var_specs), companions=[b_ref])))
    assert ref_model.D == model.D and ref_model.spec.total_epochs == 0
    rng = np.random.default_rng(6)
    params, _ = model.guess_starting_position(rng, N=30_000, batch=30_000)
    start = model.link(params)
    inv_mass = octo.diagonal_metric(model, start)
    R = 16
    pt = octo.ParallelTempering(R, seed=3, backend="local")
    th0 = start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((R, spec.D))
    res = octo.batched_parallel_tempering(model, ref_model.ℓπcallback_grad, pt, th0, 40, step_size=0.15, n_leapfrog=6, rng=rng,
                                          inv_mass=inv_mass)
    assert sorted(pt.chain_of_replica) == list(range(R)) and pt.round == 40
    assert res["swap_accept"].shape == (40, R - 1) and res["swap_accept"].sum() > 0
    assert np.all(np.isfinite(model.ℓπcallback(np.asfortranarray(res["theta"]))))
    model.close(); ref_model.close()


@pytest.mark.parametrize("name", post_cases())
def test_fused_parameterisation_equals_standalone_kernels(name, monkeypatch):
    """The parameterisation fused into K1 (one launch) and the stand-alone K0 forward / backward kernels share their
    device functions and summation orders: they agree to rounding (1e-12) for valid chains — the fused stage reuses
    intermediates of its forward pass (reciprocals, roots, log r² of a UniformCircular pair) where the stand-alone
    reverse kernel recomputes them, so the last bits may differ — and exactly on invalid and "healed" chains."""
    d, spec, consts = load_post(name)
    rng = np.random.default_rng(21)
    th = np.array(d["theta_t"])[None, :] + 0.1 * rng.standard_normal((131, spec.D))
    th[5, 0] = np.nan; th[9, 1] = np.inf; th[11, :] = 40.0; th[12, :] = -745.0
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("OCTO_B200_FUSE_PARAM", mode)
        lib, h = _model_from_golden(d, spec, consts)
        n0 = lib.octo_kernel_launches(h)
        lp, g = _logpost(lib, h, spec.D, th)
        launches = lib.octo_kernel_launches(h) - n0
        lpv, _ = _logpost(lib, h, spec.D, th, grad=False)
        out[mode] = (lp, g, lpv, launches)
        lib.octo_destroy(h)
    assert out["1"][3] == 1 and out["0"][3] == 3
    for a, b in zip(out["1"][:3], out["0"][:3]):
        fin = np.isfinite(b)
        assert np.array_equal(fin, np.isfinite(a)) and np.array_equal(a[~fin], b[~fin], equal_nan=True)
        scale = np.where(fin, np.abs(b), 0.0).max(axis=-1, keepdims=True) if b.ndim == 2 else np.abs(b)
        with np.errstate(invalid="ignore"):
            assert np.all(np.abs(a - b)[fin] <= 1e-12 * np.broadcast_to(np.maximum(scale, 1e-300), b.shape)[fin])
    assert np.array_equal(out["1"][0], out["1"][2], equal_nan=True)
    assert np.isneginf(out["1"][0][[5, 9]]).all() and (out["1"][1][[5, 9]] == 0).all()


def test_logpost_pinned_buffers_and_leading_dimension():
    """octo_logpost_grad with octo_alloc_pinned buffers (direct copies, results written straight to host) and a
    leading dimension larger than the batch gives the same bits as pageable arrays."""
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(4)
    n, ld, D = 77, 96, spec.D
    th = rng.normal(0, 0.8, (n, D))
    th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(n)
    lp0, g0 = model.ℓπcallback_grad(th)
    thp = model.pinned_empty((ld, D)); thp[...] = 0.0; thp[:n] = th
    lpp = model.pinned_empty(ld); gp = model.pinned_empty((ld, D)); gp[...] = 7.0
    rc = model._lib.octo_logpost_grad(model._h, thp.ctypes.data, n, ld, lpp.ctypes.data, gp.ctypes.data)
    assert rc == 0, model._lib.octo_last_error()
    assert np.array_equal(lpp[:n], lp0) and np.array_equal(gp[:n], g0) and (gp[n:] == 7.0).all()


def test_two_models_alive_with_different_shared_memory_sizes(oracle_lib):
    """The dynamic shared-memory opt-in is a per-function attribute: a small model created after a large one must
    not shrink it for the large one."""
    import workloads
    big, xb = workloads.many_planets(4, 64, seed=3)
    small, xs = workloads.one_planet(20, 0, 64, seed=4)
    mb = octo.LogDensityModel(big)
    ms = octo.LogDensityModel(small)
    for model, spec, x in ((mb, big, xb), (ms, small, xs), (mb, big, xb)):
        ll, g = model.ln_like_and_gradient(x)
        ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x)
        assert rel_err(ll, ll_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL


@pytest.mark.parametrize("fuse", ["1", "0"])
def test_logpost_device_entry_point(oracle_lib, monkeypatch, fuse):
    """octo_logpost_grad_device on device-resident buffers and the caller's stream (torch tensors), both as the
    fused single launch (no workspace) and as the three-launch fallback with a caller-provided workspace."""
    import torch
    monkeypatch.setenv("OCTO_B200_FUSE_PARAM", fuse)
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(12)
    n, ld, D = 200, 256, spec.D
    th = rng.normal(0, 0.8, (n, D))
    th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(n)
    d_th = torch.zeros((D, ld), dtype=torch.float64, device="cuda"); d_th[:, :n] = torch.from_numpy(th.T.copy()).cuda()
    d_lp = torch.zeros(ld, dtype=torch.float64, device="cuda")
    d_g = torch.full((D, ld), 7.0, dtype=torch.float64, device="cuda")
    nbytes = model.logpost_workspace_bytes(n)
    assert (nbytes == 0) == (fuse == "1")
    d_work = torch.empty(max(nbytes, 8), dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2):        # twice: tickets / partial buffers of the stream's workspace are reused
            model.enqueue_logpost_device(d_th.data_ptr(), n, ld, d_lp.data_ptr(), d_g.data_ptr(),
                                         d_work.data_ptr() if nbytes else 0, st.cuda_stream)
    st.synchronize()
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    assert rel_err(d_lp[:n].cpu().numpy(), lp_o).max() < LOGP_RTOL
    assert grad_err(d_g[:, :n].cpu().numpy().T, g_o).max() < GRAD_RTOL
    assert (d_g[:, n:] == 7.0).all()


def test_loglike_of_theta_and_rejection_sampler(oracle_lib, monkeypatch):
    """octo_loglike_theta (the likelihood part of the fused launch) against the oracle, fused and stand-alone, and
    the batched rejection sampler built on it (octofit_rejection, src/sampling.jl:168-258)."""
    spec = octo.ModelSpec(reference_test_system())
    rng = np.random.default_rng(31)
    th = rng.normal(0, 0.8, (150, spec.D))
    th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(150)
    th[3, 2] = np.nan
    ll_o = oracle_lib.loglike_theta(spec, octo.default_constants(), th, threads=4)
    res = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("OCTO_B200_FUSE_PARAM", fuse)
        model = octo.LogDensityModel(spec)
        res[fuse] = model.ln_like_of_theta(th)
        assert np.isneginf(res[fuse][3]) and np.isneginf(ll_o[3])
        fin = np.isfinite(ll_o)
        assert rel_err(res[fuse][fin], ll_o[fin]).max() < LOGP_RTOL
        # lp = prior part + likelihood part
        lp = model.ℓπcallback(th)
        assert np.all(lp[fin] <= res[fuse][fin] + 200.0)
    # fused and stand-alone evaluate the tables under different launch geometries (summation orders): rounding apart
    fin = np.isfinite(res["0"])
    assert np.array_equal(fin, np.isfinite(res["1"])) and rel_err(res["1"][fin], res["0"][fin]).max() < 1e-13
    monkeypatch.setenv("OCTO_B200_FUSE_PARAM", "1")
    model = octo.LogDensityModel(spec)
    out = octo.octofit_rejection(model, np.random.default_rng(5), draws=200_000, batch=50_000)
    info = out["info"]
    assert info["draws"] == 200_000 and info["n_accepted"] == len(out["loglike"]) >= 1
    assert out["theta"].shape == (info["n_accepted"], spec.D)
    # accepted draws: the oracle agrees on their likelihood, and logpost is what ℓπcallback returns for them
    ll_acc = oracle_lib.loglike_theta(spec, octo.default_constants(), out["theta_t"], threads=4)
    assert rel_err(out["loglike"], ll_acc).max() < 1e-9
    assert np.array_equal(out["logpost"], model.ℓπcallback(out["theta_t"]))


def test_hgca_model_with_priors_on_device(oracle_lib):
    """A proper-motion-anomaly model as the docs build it (docs/src/pma.md: HGCAInstantaneousObs, pmra/pmdec priors,
    mass prior, θ_at_epoch_to_tperi): full ℓπ and ∇ℓπ through the host mirror against the oracle."""
    row = dict(pmra_hip=10.1, pmdec_hip=-5.2, pmra_hip_error=0.9, pmdec_hip_error=0.8, pmra_pmdec_hip=0.2,
               pmra_hg=10.5, pmdec_hg=-5.0, pmra_hg_error=0.05, pmdec_hg_error=0.04, pmra_pmdec_hg=-0.1,
               pmra_gaia=11.2, pmdec_gaia=-4.6, pmra_gaia_error=0.12, pmdec_gaia_error=0.1, pmra_pmdec_gaia=0.35,
               epoch_ra_hip=1991.1, epoch_dec_hip=1991.3, epoch_ra_gaia=2016.0, epoch_dec_gaia=2016.2)
    hg = octo.HGCAInstantaneousObs(row, N_ave=5)
    b = octo.Planet(name="b", variables={
        "a": octo.LogUniform(1, 30), "e": octo.Uniform(0, 0.9), "i": octo.Sine(), "ω": octo.UniformCircular(),
        "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 57000.0),
        "mass": octo.LogUniform(1, 100)})
    system = octo.System(name="pma", companions=[b], observations=[hg], variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1),
        "pmra": octo.Normal(10.5, 5.0), "pmdec": octo.Normal(-5.0, 5.0)})
    spec = octo.ModelSpec(system)
    model = octo.LogDensityModel(spec)
    assert model.total_epochs == 0
    rng = np.random.default_rng(41)
    th = rng.normal(0, 0.7, (97, spec.D))
    th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(97)
    lp, g = model.ℓπcallback_grad(th)
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    assert np.isfinite(lp_o).all()
    assert rel_err(lp, lp_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    assert rel_err(model.ℓπcallback(th), lp_o).max() < LOGP_RTOL


def test_radial_velocity_orbit_basis(oracle_lib):
    """Planet(basis="RadialVelocityOrbit"): the RV-only orbit (no i, Ω, plx) is the same kernel with i = π/2, Ω = 0
    injected as constants; K carries no sin i factor (sin(π/2) == 1.0 exactly)."""
    rng = np.random.default_rng(52)
    ep = np.sort(rng.uniform(58000, 59500, 40))
    rv = octo.StarAbsoluteRVObs(octo.Table(epoch=ep, rv=30 * np.sin(ep / 60.0) + rng.normal(0, 5, 40), σ_rv=np.full(40, 5.0)),
                                name="HARPS", variables={"offset": octo.Normal(0, 100), "jitter": octo.LogUniform(0.1, 100.0)})
    b = octo.Planet(name="b", basis="RadialVelocityOrbit", variables={
        "a": octo.LogUniform(0.1, 10), "e": octo.Uniform(0, 0.9), "ω": octo.UniformCircular(),
        "tp": octo.Uniform(57500, 59000), "mass": octo.LogUniform(0.1, 100)})
    system = octo.System(name="rvonly", companions=[b], observations=[rv],
                         variables={"M": octo.truncated(octo.Normal(1.0, 0.05), lower=0.1)})
    spec = octo.ModelSpec(system)
    assert float(np.sin(np.pi / 2)) == 1.0
    model = octo.LogDensityModel(spec)
    th = rng.normal(0, 0.8, (64, spec.D))
    lp, g = model.ℓπcallback_grad(th)
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    assert np.isfinite(lp_o).all() and rel_err(lp, lp_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    with pytest.raises(ValueError):
        octo.Planet(name="c", basis="RadialVelocityOrbit", variables={"a": 1.0},
                    observations=[octo.PlanetRelAstromObs(octo.Table(epoch=[5e4], ra=[1.], dec=[1.], σ_ra=[1.], σ_dec=[1.]), name="x")])


def test_device_resident_hmc():
    """octo_hmc_run: the whole HMC run enqueued on one stream.  (1) One transition is reproduced step by step on the
    host from the same counter-based random numbers and the same device log-posterior; (2) runs are reproducible;
    (3) a longer run behaves like the host-driven batched_hmc on the reference's 11-D test model."""
    import time
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(4)
    params, lp0 = model.guess_starting_position(rng, N=60_000, batch=20_000)
    start = model.link(params)
    inv_mass = octo.diagonal_metric(model, start)
    n, D, eps, L, seed = 64, spec.D, 0.15, 5, 1234
    th0 = np.asfortranarray(start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((n, D)))
    res = octo.device_hmc(model, th0, 1, step_size=eps, n_leapfrog=L, inv_mass=inv_mass, seed=seed)
    # host replica of transition 0
    lp, g = model.ℓπcallback_grad(th0)
    zu = [octo.hmc_random(model, seed, 0, c, D) for c in range(n)]
    z = np.array([t[0] for t in zu]); u = np.array([t[1] for t in zu])
    p = z / np.sqrt(inv_mass)
    h0 = -lp + 0.5 * np.sum(p * p * inv_mass, axis=1)
    p = p + 0.5 * eps * g
    q = th0 + eps * p * inv_mass
    for k in range(L):
        lq, gq = model.ℓπcallback_grad(np.asfortranarray(q))
        gq = np.where(np.isfinite(lq)[:, None], gq, 0.0)
        p = p + (eps if k < L - 1 else 0.5 * eps) * gq
        if k < L - 1:
            q = q + eps * p * inv_mass
    h1 = -lq + 0.5 * np.sum(p * p * inv_mass, axis=1)
    accept = np.isfinite(lq) & (np.log(u) < h0 - h1)
    expect = np.where(accept[:, None], q, th0)
    assert 0 < accept.sum() <= n and np.array_equal(res["accept"] == 1.0, accept)
    assert np.allclose(res["theta_final"], expect, rtol=1e-11, atol=1e-13)
    assert np.allclose(res["logpost_final"], np.where(accept, lq, lp), rtol=1e-10)
    assert np.array_equal(res["theta"][0], res["theta_final"]) and np.array_equal(res["logpost"][0], res["logpost_final"])
    # reproducible
    r1 = octo.device_hmc(model, th0, 7, step_size=eps, n_leapfrog=L, inv_mass=inv_mass, seed=99)
    r2 = octo.device_hmc(model, th0, 7, step_size=eps, n_leapfrog=L, inv_mass=inv_mass, seed=99)
    assert np.array_equal(r1["theta"], r2["theta"]) and np.array_equal(r1["accept"], r2["accept"])
    r3 = octo.device_hmc(model, th0, 7, step_size=eps, n_leapfrog=L, inv_mass=inv_mass, seed=100)
    assert not np.array_equal(r1["theta"], r3["theta"])
    # the whole run is ONE launch of the trajectory-resident kernel
    n0 = model.kernel_launches
    octo.device_hmc(model, th0, 7, step_size=eps, n_leapfrog=L, inv_mass=inv_mass, seed=99)
    assert model.kernel_launches - n0 == 1
    # launch-per-leapfrog explorer: the leapfrog update folded into the log-posterior launch == the separate update
    # kernel, bit for bit
    os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"] = "1"
    try:
        n0 = model.kernel_launches
        r5 = octo.device_hmc(model, th0, 7, step_size=eps, n_leapfrog=L, inv_mass=inv_mass, seed=99)
        assert model.kernel_launches - n0 == 7 * (L + 1) + 2         # L posterior launches + 1 turn per transition, + start, + final turn
        os.environ["OCTO_B200_HMC_SEPARATE_LEAP"] = "1"
        r4 = octo.device_hmc(model, th0, 7, step_size=eps, n_leapfrog=L, inv_mass=inv_mass, seed=99)
    finally:
        del os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"]
        os.environ.pop("OCTO_B200_HMC_SEPARATE_LEAP", None)
    assert np.array_equal(r5["theta"], r4["theta"]) and np.array_equal(r5["logpost"], r4["logpost"])
    # and it agrees with the resident kernel to rounding (identical bits when both use the same geometry: test_gpu_resident.py)
    assert np.allclose(r1["theta"], r5["theta"], rtol=1e-9, atol=1e-11) and np.array_equal(r1["accept"], r5["accept"])
    # a real run: 256 chains x 150 transitions x 12 leapfrogs, against the host-driven explorer
    th256 = start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((256, D))
    t0 = time.perf_counter()
    dev = octo.device_hmc(model, th256, 150, step_size=0.15, n_leapfrog=12, inv_mass=inv_mass, seed=7)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    host = octo.batched_hmc(model, th256, 150, step_size=0.15, n_leapfrog=12, rng=rng, inv_mass=inv_mass)
    t_host = time.perf_counter() - t0
    print(f"device-resident HMC {t_dev*1e3:.1f} ms vs host-driven {t_host*1e3:.1f} ms for 256 chains x 150 x 12 leapfrogs")
    assert abs(dev["accept_rate"] - host["accept_rate"]) < 0.08 and dev["accept_rate"] > 0.4
    assert abs(np.median(dev["logpost"][-50:]) - np.median(host["logpost"][-50:])) < 2.0
    names = list(spec.theta_names)
    for nm in ("b.a", "b.e"):
        a_d = model.invlink(dev["theta"][-1])[:, names.index(nm)]; a_h = model.invlink(host["theta"][-1])[:, names.index(nm)]
        assert abs(np.median(a_d) - np.median(a_h)) < 0.35 * (np.std(a_h) + np.std(a_d))
    model.close()


def test_three_planet_parameterised_model(oracle_lib, monkeypatch):
    """A larger standard model — three planets, each with UniformCircular ω, Ω, θ and its own θ_at_epoch_to_tperi,
    astrometry with sampled jitter, star RV and marginalised RV (D = 38, three tperi definitions in one launch) —
    fused and stand-alone against the oracle."""
    rng = np.random.default_rng(61)
    planets = []
    for k, nm in enumerate("bcd"):
        a0 = 3.0 * 2.1 ** k
        ep = np.sort(rng.uniform(50000, 52500, 9 + k))
        tab = octo.Table(epoch=ep, ra=50 * a0 * np.cos(ep / (200.0 * (k + 1))) + rng.normal(0, 2, len(ep)),
                         dec=50 * a0 * np.sin(ep / (200.0 * (k + 1))) + rng.normal(0, 2, len(ep)),
                         σ_ra=np.full(len(ep), 2.0), σ_dec=np.full(len(ep), 2.5), cor=rng.uniform(-0.5, 0.5, len(ep)))
        obs = octo.PlanetRelAstromObs(tab, name=f"cam{k}", variables={"jitter": octo.LogUniform(0.01, 10.0)})
        planets.append(octo.Planet(name=nm, observations=[obs], variables={
            "a": octo.LogUniform(0.5 * a0, 2.0 * a0), "e": octo.Uniform(0, 0.8), "i": octo.Sine(), "ω": octo.UniformCircular(),
            "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 51000.0),
            "mass": octo.LogUniform(0.5, 50)}))
    eps = np.sort(rng.uniform(50000, 52500, 30))
    rv1 = octo.StarAbsoluteRVObs(octo.Table(epoch=eps[:15], rv=rng.normal(0, 30, 15), σ_rv=np.full(15, 5.0)), name="harps",
                                 variables={"offset": octo.Normal(0, 100), "jitter": octo.LogUniform(0.1, 100.0)})
    rv2 = octo.MarginalizedStarAbsoluteRVObs(octo.Table(epoch=eps[15:], rv=rng.normal(0, 30, 15), σ_rv=np.full(15, 5.0)), name="hires",
                                             variables={"jitter": octo.LogUniform(0.1, 100.0)})
    system = octo.System(name="three", companions=planets, observations=[rv1, rv2], variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(40.0, 0.5), lower=0.1)})
    spec = octo.ModelSpec(system)
    assert spec.D == 38 and sum(1 for d in spec.defs if d.op == 3) == 3
    th = rng.normal(0, 0.6, (90, spec.D))
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    fin = np.isfinite(lp_o)
    assert fin.sum() > 60
    out = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("OCTO_B200_FUSE_PARAM", fuse)
        model = octo.LogDensityModel(spec)
        n0 = model.kernel_launches
        lp, g = model.ℓπcallback_grad(th)
        out[fuse] = (lp, g, model.kernel_launches - n0)
        assert np.array_equal(np.isfinite(lp), fin)
        assert rel_err(lp[fin], lp_o[fin]).max() < LOGP_RTOL and grad_err(g[fin], g_o[fin]).max() < GRAD_RTOL
        model.close()
    assert out["1"][2] == 1 and out["0"][2] == 3
    # (not bit-equal here: the fused launch needs a smaller CTA for this model, which changes the summation tree)
    assert rel_err(out["1"][0][fin], out["0"][0][fin]).max() < 1e-12


def test_octofit_recovers_the_fixture_orbit():
    """The reference's "fit a chain" smoke test (test/integration-tests.jl:52-58, test/integration/sampling.jl:78-134)
    through the batched driver: adaptation + sampling on the 11-D test model; the posterior must be plausible
    (`logpost > -1000`), the chains must agree with each other, and the well-constrained quantity of this data set —
    the position angle at the reference epoch — must match the data."""
    import time
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    t0 = time.perf_counter()
    chain = octo.octofit(model, np.random.default_rng(3), n_chains=256, adaptation=240, iterations=200, n_init=60_000, seed=11)
    dt = time.perf_counter() - t0
    info = chain["info"]
    print(f"octofit: {info['n_gradient_calls']} batched gradient calls for 256 chains in {dt:.2f} s, accept {info['accept_rate']:.2f}, step {info['step_size']:.3g}")
    assert 0.5 < info["accept_rate"] <= 1.0
    lp = chain["logpost"]
    assert np.all(np.isfinite(lp[-1])) and np.all(lp[-1] > -1000)
    names = list(chain["names"])
    nat = chain["theta"][-100:].reshape(-1, len(names))
    th_x, th_y = nat[:, names.index("b.θx")], nat[:, names.index("b.θy")]
    theta = np.arctan2(th_y, th_x)
    # the fixture's first epoch is the reference epoch of θ (50000): PA of the data there
    d = load_post("post_fixture8")[0]
    ra0, dec0 = d["blocks"][0]["y1"][0], d["blocks"][0]["y2"][0]
    pa = np.arctan2(ra0, dec0)
    dpa = np.angle(np.exp(1j * (theta - pa)))
    assert abs(np.median(dpa)) < 0.05 and np.std(dpa) < 0.2
    # between-chain agreement (a crude R-hat on the log posterior): chains have mixed
    lp_tail = lp[-100:]
    between = lp_tail.mean(axis=0).var(); within = lp_tail.var(axis=0).mean()
    assert between < 0.5 * within
    model.close()


def test_device_parallel_tempering():
    """octo_pt_hmc_run.  (1) With every weight equal to 1 it is the plain explorer, bit for bit (tempering multiplies
    by 1.0; equal-weight swaps always happen and change nothing).  (2) Frozen states with widely separated likelihoods:
    the swap round takes the decisions of the even-odd rule.  (3) A real ladder on the reference's test model: rungs stay
    a permutation, weights follow rungs, neighbours do swap, and the chain on the last rung ends at a plausible log
    posterior."""
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(4)
    params, lp0 = model.guess_starting_position(rng, N=60_000, batch=20_000)
    start = model.link(params)
    inv_mass = octo.diagonal_metric(model, start)
    n, D = 32, spec.D
    th0 = np.asfortranarray(start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((n, D)))
    # (1)
    plain = octo.device_hmc(model, th0, 6, step_size=0.1, n_leapfrog=5, inv_mass=inv_mass, seed=77)
    ones = octo.device_parallel_tempering(model, th0, np.ones(n), 1, n_iter=6, n_leapfrog=5, step_size=0.1, inv_mass=inv_mass, seed=77)
    assert np.array_equal(ones["theta_final"], plain["theta_final"]) and np.array_equal(ones["logpost_tempered"], plain["logpost_final"])
    assert np.all(ones["swap_counts"][0::2] == 1) and np.all(ones["swap_counts"][1::2] == 0)          # round 0 pairs (0,1), (2,3), ...
    assert sorted(ones["rung"]) == list(range(n)) and np.all(ones["beta"] == 1.0)
    # (2) frozen states: chain c sits further from the start the larger c is => ln_like falls steeply with c
    spread = np.linspace(0.0, 6.0, n)[:, None]
    th_far = np.asfortranarray(start[None, :] + spread * np.sqrt(inv_mass)[None, :] * np.sign(rng.standard_normal((1, D))))
    ladder = np.linspace(0.0, 1.0, n)
    res = octo.device_parallel_tempering(model, th_far, ladder, 1, n_iter=1, n_leapfrog=1, step_size=1e-12, inv_mass=inv_mass, seed=5)
    ll = res["loglike"]                  # the data likelihood alone: the UnitLengthPrior terms belong to the reference too
    assert np.isfinite(ll).all() and np.all(np.diff(ll) < 0)
    # tempered density = full posterior - (1 - beta) ln_like, at the weights the chains hold after the swap round
    assert np.allclose(model.ℓπcallback(th_far) - res["logpost_tempered"], (1.0 - res["beta"]) * ll, rtol=1e-9, atol=1e-9)
    for i in range(0, n - 1, 2):
        log_ratio = (ladder[i] - ladder[i + 1]) * (ll[i + 1] - ll[i])
        if abs(log_ratio) > 40:                                       # far from any uniform draw's log
            assert res["swap_counts"][i] == (1.0 if log_ratio > 0 else 0.0), (i, log_ratio)
    assert np.all(res["swap_counts"][1::2] == 0)
    assert sorted(res["rung"]) == list(range(n)) and np.array_equal(res["beta"], ladder[res["rung"]])
    # (3) a real run from scattered starts
    th_pr = model.link(model.sample_priors(rng, n))
    th_pr[-1] = start
    ladder = np.linspace(0.0, 1.0, n) ** 3
    out = octo.device_parallel_tempering(model, th_pr, ladder, 60, n_iter=2, n_leapfrog=10, step_size=0.1, inv_mass=inv_mass, seed=9)
    assert sorted(out["rung"]) == list(range(n)) and np.array_equal(out["beta"], ladder[out["rung"]])
    assert np.all((out["swap_accept"] >= 0) & (out["swap_accept"] <= 1)) and out["swap_accept"].mean() > 0.05
    cold = int(np.argmax(out["rung"]))
    assert out["beta"][cold] == 1.0 and np.isfinite(out["logpost_tempered"][cold]) and out["logpost_tempered"][cold] > lp0 - 40
    assert np.array_equal(out["cold_trace"][-1], out["theta_final"][cold])
    model.close()


def test_host_terms_next_to_the_device_posterior(oracle_lib):
    """non-epoch terms that are arbitrary user code in the reference — a `UserLikelihood`
    (`mass_ratio ~ Normal`, src/variables.jl:332-380) and a planet-order prior (src/likelihoods/prior-planet-order.jl) —
    as host callbacks next to the device posterior: value and gradient of the full posterior against oracle + numpy."""
    import workloads
    spec, th = workloads.one_planet_with_priors(60, 40, 90, seed=5)
    model = octo.LogDensityModel(spec)
    names = list(spec.theta_names)
    ja, jm, jM = names.index("b.a"), names.index("b.mass"), names.index("M")

    def user_likelihood(nat):                 # (mass * mjup2msol / M) ~ Normal(0.008, 0.002)
        q = nat[:, jm] * 0.0009545942339693249 / nat[:, jM]
        z = (q - 0.008) / 0.002
        g = np.zeros_like(nat)
        g[:, jm] = -z / 0.002 * 0.0009545942339693249 / nat[:, jM]
        g[:, jM] = z / 0.002 * q / nat[:, jM]
        return -0.5 * z * z - np.log(0.002) - 0.5 * np.log(2 * np.pi), g

    def order_prior(nat):                     # smooth stand-in for `a_b < 30 AU` ordering constraints: log-sigmoid wall
        u = (30.0 - nat[:, ja]) / 0.5
        g = np.zeros_like(nat)
        g[:, ja] = -(1.0 / (1.0 + np.exp(u))) / 0.5
        return -np.log1p(np.exp(-u)), g
    lp0, g0 = model.ℓπcallback_grad(th)
    model.add_host_term(user_likelihood); model.add_host_term(order_prior)
    lp, g = model.ℓπcallback_grad(th)
    lpv = model.ℓπcallback(th)
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    nat = oracle_lib.invlink(spec, th)
    v1, gn1 = user_likelihood(nat); v2, gn2 = order_prior(nat)
    # d natural / d θ_t of the elementwise bijectors, by differences of the oracle's invlink
    h = 1e-6 * np.maximum(1.0, np.abs(th))
    dxdy = (oracle_lib.invlink(spec, th + h) - oracle_lib.invlink(spec, th - h)) / (2 * h)
    assert rel_err(lp, lp_o + v1 + v2).max() < LOGP_RTOL and rel_err(lpv, lp_o + v1 + v2).max() < LOGP_RTOL
    assert grad_err(g, g_o + (gn1 + gn2) * dxdy).max() < 1e-7
    assert np.array_equal(lp0, model.ℓπcallback_grad(th)[0] - (v1 + v2)) or rel_err(lp0 + v1 + v2, lp).max() < 1e-13
    model.close()


def test_batched_slice_sampler_and_tempering_with_it():
    """The value-only explorer the reference gives Pigeons (SliceSampler), replica-batched over K1v: (1) at tempering
    weight 0 the target is the prior-only reference model, whose marginals are known exactly: started from prior draws
    the sampler must keep reproducing them (invariance) — Uniform(0, 100), Uniform(0, 0.99), Sine, truncated Normal;
    (2) at weight 1, started on the posterior, the log density stays where HMC's is; (3) tempered with
    `ParallelTempering` swaps the cold rung stays on the posterior and swaps are accepted."""
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(12)
    names = list(spec.theta_names)
    n = 512
    th_prior = model.link(model.sample_priors(rng, n))
    n0 = model.kernel_launches
    sl = octo.batched_slice_sampler(model, th_prior, 6, rng=rng, w=2.0, beta=np.zeros(n))
    assert sl["n_evaluations"] > 6 * 3 * spec.D * 2 and model.kernel_launches - n0 >= sl["n_evaluations"]
    nat = model.invlink(sl["theta"][-1])
    a, e, inc, M = (nat[:, names.index(k)] for k in ("b.a", "b.e", "b.i", "M"))
    se = lambda sd: 4.0 * sd / np.sqrt(n)                                # 4 sigma of the sample mean
    assert abs(a.mean() - 50.0) < se(28.9) and abs(a.std() - 28.87) < 3.0
    assert abs(e.mean() - 0.495) < se(0.286) and abs(inc.mean() - np.pi / 2) < se(0.68)
    assert abs(M.mean() - 1.2) < se(0.1) and abs(M.std() - 0.1) < 0.02
    # (2) posterior: start from HMC draws, the density level is kept
    params, _ = model.guess_starting_position(rng, N=60_000, batch=20_000)
    start = model.link(params)
    inv_mass = octo.diagonal_metric(model, start)
    th0 = start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((128, spec.D))
    hmc = octo.device_hmc(model, th0, 200, step_size=0.15, n_leapfrog=12, inv_mass=inv_mass, seed=3)
    post = hmc["theta"][-1]
    sp = octo.batched_slice_sampler(model, post, 8, rng=rng, w=1.0)
    assert np.all(np.isfinite(sp["logdensity"]))
    assert abs(np.median(sp["logdensity"]) - np.median(hmc["logpost"][-1])) < 3.0
    # (3) tempering: 16 replicas; the cold replica stays on the posterior, swaps happen
    R = 16
    lad = np.linspace(0.0, 1.0, R) ** 3
    pt = octo.ParallelTempering(R, seed=4, beta=lad, backend="local")
    res = octo.batched_slice_parallel_tempering(model, pt, post[:R], 12, rng=rng, w=1.0)
    assert res["swap_accept"].sum() > 0
    cold = int(np.argmax(pt.chain_of_replica == R - 1))
    lp_cold = model.ℓπcallback(res["theta"][cold])
    assert lp_cold > np.median(hmc["logpost"][-1]) - 25.0
    model.close()
