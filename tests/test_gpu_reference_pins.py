"""Behavioural pins the reference itself ships, replayed through the CUDA path.

* test/unit/likelihoods.jl:32-95 — `northangle` sign convention (issue #141): data rotated by +ε in position angle
  must be undone by northangle = -ε in BOTH table formats; the reference scans a 2001-point grid one value at a time,
  here the grid is one 2001-chain batch per format.  Also `ll(0.0) == ll(-0.0)` exactly.
* test/integration/sampling.jl:136-192 — finite-difference vs AD gradient of ∇ℓπcallback at a prior draw, with the
  reference's own tolerances (`atol=1e-3 rtol=1e-4` of `isapprox`, i.e. on the 2-norm), on the device log posterior.
"""
import numpy as np
import pytest

import octofitter_jl_b200 as octo
import workloads

pytestmark = pytest.mark.gpu


def _northangle_tables():
    epochs = np.array([50000.0, 50300.0, 50600.0, 50900.0, 51200.0])
    el = dict(plx=50.0, M=1.2, a=15.0, e=0.2, i=0.6, w=0.3, W=1.1, tp=50000.0)
    ra_m, dec_m, _, _ = workloads._state(el, epochs)
    pa_m, sep_m = np.arctan2(ra_m, dec_m), np.hypot(ra_m, dec_m)
    eps = 0.05
    pa_d = pa_m + eps
    ra_d, dec_d = sep_m * np.sin(pa_d), sep_m * np.cos(pa_d)
    n = len(epochs)
    seppa = octo.Table(epoch=epochs, sep=sep_m, pa=pa_d, σ_sep=np.full(n, 1.0), σ_pa=np.full(n, 0.001))
    radec = octo.Table(epoch=epochs, ra=ra_d, dec=dec_d, σ_ra=np.full(n, 1.0), σ_dec=np.full(n, 1.0))
    return el, eps, seppa, radec


def _northangle_ll(tab, el, deltas):
    obs = octo.PlanetRelAstromObs(tab, name="inst", variables=["northangle"])
    pl = octo.Planet(name="b", variables=["M", "a", "e", "i", "ω", "Ω", "tp"], observations=[obs])
    spec = octo.ModelSpec(octo.System(name="northangle_test", variables=["plx"], companions=[pl]))
    truth = {"plx": el["plx"], "b.M": el["M"], "b.a": el["a"], "b.e": el["e"], "b.i": el["i"], "b.ω": el["w"],
             "b.Ω": el["W"], "b.tp": el["tp"], "b.inst.northangle": 0.0}
    x = np.tile(np.array([truth[n] for n in spec.input_names]), (len(deltas), 1))
    x[:, spec.column("b.inst.northangle")] = deltas
    model = octo.LogDensityModel(spec)
    ll = model.ln_like(np.asfortranarray(x))
    llg, g = model.ln_like_and_gradient(np.asfortranarray(x))
    model.close()
    assert np.array_equal(ll, llg)
    return ll, g[:, spec.column("b.inst.northangle")]


def test_northangle_sign_convention_both_table_formats():
    el, eps, seppa, radec = _northangle_tables()
    grid = np.linspace(-0.1, 0.1, 2001)
    ll_s, g_s = _northangle_ll(seppa, el, grid)
    ll_r, g_r = _northangle_ll(radec, el, grid)
    best_s, best_r = grid[np.argmax(ll_s)], grid[np.argmax(ll_r)]
    assert abs(best_s - (-eps)) < 1e-3, best_s
    assert abs(best_r - (-eps)) < 1e-3, best_r
    assert np.sign(best_s) == np.sign(best_r)
    # the analytic ∂ll/∂northangle changes sign at the same place in both formats
    for g in (g_s, g_r):
        k = np.argmax(g < 0)
        assert g[0] > 0 and g[-1] < 0 and abs(grid[k] - (-eps)) < 1e-3
    # a zero northangle is a no-op whatever its sign bit
    for tab in (seppa, radec):
        ll0, _ = _northangle_ll(tab, el, np.array([0.0, -0.0]))
        assert ll0[0] == ll0[1]


def _gradient_test_system():
    astrom = octo.PlanetRelAstromLikelihood(octo.Table(
        epoch=[50000, 50120, 50240, 50360], ra=[-505.76, -502.57, -498.21, -492.68], dec=[-66.93, -37.47, -7.93, 21.64],
        σ_ra=[10.0] * 4, σ_dec=[10.0] * 4, cor=[0.0] * 4), name="gradient_test")
    b = octo.Planet(name="b", observations=[astrom], variables={
        "a": octo.Uniform(0, 100), "e": octo.Uniform(0.0, 0.99), "i": octo.Sine(), "ω": octo.UniformCircular(),
        "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 50000)})
    return octo.System(name="GradTestSys", companions=[b], variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1)})


def test_finite_difference_vs_device_gradient_reference_tolerances():
    model = octo.LogDensityModel(_gradient_test_system())
    assert model.D == 11
    rng = np.random.default_rng(42)
    checked = 0
    for _ in range(12):
        theta = model.link(model.sample_priors(rng, 1)[0])
        lp, grad = model.ℓπcallback_grad(theta)
        if not np.isfinite(lp):
            continue
        # central differences, all 2 D perturbed points in one value-only batch
        h = 1e-6 * np.maximum(1.0, np.abs(theta))
        pts = np.tile(theta, (2 * model.D, 1))
        for j in range(model.D):
            pts[2 * j, j] += h[j]; pts[2 * j + 1, j] -= h[j]
        v = model.ℓπcallback(np.asfortranarray(pts))
        fd = (v[0::2] - v[1::2]) / (2 * h)
        # Julia isapprox(x, y; atol, rtol): norm(x - y) <= max(atol, rtol * max(norm(x), norm(y)))
        assert np.linalg.norm(fd - grad) <= max(1e-3, 1e-4 * max(np.linalg.norm(fd), np.linalg.norm(grad))), (fd, grad)
        checked += 1
    assert checked >= 8
    model.close()
