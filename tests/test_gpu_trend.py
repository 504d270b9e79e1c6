"""`trend_function` of the radial-velocity observations (rv-absolute.jl:143, rv-absolute-margin.jl:111,
rv-relative.jl:131) offloaded when it is linear in the observation variables: every RV kind, one to three coefficient
variables plus a constant part, several launch geometries, the fused log posterior — all against the oracle
(1e-10 on logp, 1e-8 on the gradient); non-linear closures are refused."""
import numpy as np
import pytest

import octofitter_jl_b200 as octo
import workloads
from helpers import grad_err, rel_err

pytestmark = pytest.mark.gpu
LOGP_RTOL, GRAD_RTOL = 1e-10, 1e-8
U = lambda t: (t - 52000.0) / 1000.0


def _system(kind, n_trend, n_ep=70, seed=3):
    rng = np.random.default_rng(seed)
    ep = np.sort(rng.uniform(50000.0, 54000.0, n_ep))
    _, _, rvp, _ = workloads._state(workloads.TRUTH_B, ep)
    mu = workloads.TRUTH_B["mass"] * workloads.MJUP2MSOL / workloads.TRUTH_B["M"]
    trend_true = 2.0 + 4.0 * U(ep) - 1.5 * U(ep) ** 2 + 3.0 * np.sin(2 * np.pi * ep / 365.25)
    tvars = ["trend_slope", "trend_quad", "trend_amp"][:n_trend]
    f = {1: lambda th, t: 2.0 + th.trend_slope * U(t),
         2: lambda th, t: 2.0 + th.trend_slope * U(t) + th.trend_quad * U(t) ** 2,
         3: lambda th, t: th.trend_slope * U(t) + th.trend_quad * U(t) ** 2 + th.trend_amp * np.sin(2 * np.pi * t / 365.25)}[n_trend]
    pv = ["a", "e", "i", "ω", "Ω", "tp", "mass"]
    if kind == "planet":
        tab = octo.Table(epoch=ep, rv=rvp + trend_true + 30 * rng.standard_normal(n_ep), σ_rv=np.full(n_ep, 30.0))
        obs = octo.PlanetRelativeRVObs(tab, name="crires", variables=["jitter"] + tvars, trend_function=f)
        b = octo.Planet(name="b", variables=pv, observations=[obs])
        return octo.System(name="s", variables=["M", "plx"], companions=[b]), "b.crires."
    tab = octo.Table(epoch=ep, rv=150.0 - mu * rvp + trend_true + 5 * rng.standard_normal(n_ep), σ_rv=np.full(n_ep, 5.0))
    cls = octo.MarginalizedStarAbsoluteRVObs if kind == "margin" else octo.StarAbsoluteRVObs
    base = ["jitter"] if kind == "margin" else (["offset"] if kind == "star_nojit" else ["offset", "jitter"])
    obs = cls(tab, name="rv", variables=base + tvars, trend_function=f)
    b = octo.Planet(name="b", variables=pv, observations=[])
    return octo.System(name="s", variables=["M", "plx"], companions=[b], observations=[obs]), "rv."


def _inputs(spec, prefix, n, seed):
    rng = np.random.default_rng(seed)
    truth = {"M": 1.2, "plx": 50.0, "b.a": 10.0, "b.e": 0.3, "b.i": 1.0, "b.ω": 0.5, "b.Ω": 2.0, "b.tp": 50000.0, "b.mass": 10.0,
             prefix + "offset": 150.0, prefix + "jitter": 3.0, prefix + "trend_slope": 4.0, prefix + "trend_quad": -1.5, prefix + "trend_amp": 3.0}
    x0 = np.array([truth[nm] for nm in spec.input_names])
    x = x0[None, :] * (1.0 + 0.03 * rng.standard_normal((n, len(x0))))
    x[:, spec.column("b.e")] = np.clip(x[:, spec.column("b.e")], 0, 0.9)
    return np.asfortranarray(x)


@pytest.mark.parametrize("kind,n_trend", [("star", 1), ("star", 3), ("star_nojit", 2), ("margin", 2), ("margin", 3), ("planet", 1), ("planet", 3)])
@pytest.mark.parametrize("force", [None, "1,0,3", "4,1,1"])
def test_linear_trend_matches_oracle(oracle_lib, monkeypatch, kind, n_trend, force):
    if force:
        monkeypatch.setenv("OCTO_B200_FORCE", force)
    system, prefix = _system(kind, n_trend)
    spec = octo.ModelSpec(system)
    x = _inputs(spec, prefix, 77, seed=11)
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    llv = model.ln_like(x)
    model.close()
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=4)
    assert np.all(np.isfinite(ll_o))
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and rel_err(llv, ll_o).max() < LOGP_RTOL
    assert grad_err(g, g_o).max() < GRAD_RTOL
    for nm in ("trend_slope", "trend_quad", "trend_amp")[:n_trend]:          # the trend gradients themselves, relative to their own size
        k = spec.column(prefix + nm)
        assert np.max(np.abs(g[:, k] - g_o[:, k]) / np.maximum(np.abs(g_o[:, k]), 1e-12 * np.abs(g_o).max())) < 1e-7, nm


def test_trend_next_to_astrometry_with_priors_on_device(oracle_lib):
    """The docs' linear trend with a prior on its slope, next to an astrometry table: full log posterior on the device."""
    spec0, _ = workloads.one_planet(40, 0, 1, seed=2)
    astrom = octo.PlanetRelAstromObs(spec0.system.planets[0].observations[0].table, name="astrom")
    sysrv, _ = _system("star", 1, n_ep=50)
    tab = sysrv.observations[0].table
    rv = octo.StarAbsoluteRVObs({"epoch": tab["epoch"], "rv": tab["rv"], "σ_rv": tab["σ_rv"]}, name="rv",
                                variables={"offset": octo.Normal(150, 100), "jitter": octo.LogUniform(0.1, 100.0),
                                           "trend_slope": octo.Normal(0, 10)},
                                trend_function=lambda th, t: 2.0 + th.trend_slope * U(t))
    b = octo.Planet(name="b", observations=[astrom], variables={
        "a": octo.LogUniform(1, 100), "e": octo.Uniform(0, 0.99), "i": octo.Sine(), "ω": octo.UniformCircular(),
        "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 50000.0),
        "mass": octo.LogUniform(0.1, 100)})
    system = octo.System(name="s", companions=[b], observations=[rv], variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1)})
    spec = octo.ModelSpec(system)
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(9)
    th = rng.normal(0, 0.5, (45, spec.D)); th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(45)
    lp, g = model.ℓπcallback_grad(th)
    model.close()
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    fin = np.isfinite(lp_o)
    assert fin.sum() > 30 and rel_err(lp[fin], lp_o[fin]).max() < LOGP_RTOL and grad_err(g[fin], g_o[fin]).max() < GRAD_RTOL


def test_nonlinear_or_oversized_trends_are_refused():
    sysrv, _ = _system("star", 1, n_ep=10)
    tab = sysrv.observations[0].table
    t = {"epoch": tab["epoch"], "rv": tab["rv"], "σ_rv": tab["σ_rv"]}
    with pytest.raises(ValueError, match="not linear"):
        octo.StarAbsoluteRVObs(t, name="rv", variables=["offset", "jitter", "period"], trend_function=lambda th, e: np.sin(e / (1.0 + th.period ** 2)))
    with pytest.raises(ValueError, match="at most three"):
        octo.StarAbsoluteRVObs(t, name="rv", variables=["offset", "jitter", "c1", "c2", "c3", "c4"],
                               trend_function=lambda th, e: th.c1 + th.c2 * U(e) + th.c3 * U(e) ** 2 + th.c4 * U(e) ** 3)
    with pytest.raises(ValueError, match="offset / jitter"):
        octo.StarAbsoluteRVObs(t, name="rv", variables=["offset", "jitter"], trend_function=lambda th, e: th.offset * U(e))
    # a zero closure is the default trend
    ok = octo.StarAbsoluteRVObs(t, name="rv", variables=["offset", "jitter"], trend_function=lambda th, e: 0.0)
    assert ok.trend[0] == [] and ok.trend[2] is None
