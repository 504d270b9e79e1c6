"""CPU tests of the standard-parameterisation layer (N1): host mirror of @variables, oracle vs mpmath goldens,
oracle gradient vs finite differences (the reference's own gradient test, test/integration/sampling.jl:136-192)."""
import numpy as np
import pytest

import octofitter_jl_b200 as octo
from helpers import grad_err, load_post, post_cases, reference_test_system, rel_err


def test_reference_test_model_has_11_parameters():
    spec = octo.ModelSpec(reference_test_system())
    assert spec.D == 11                       # `@test model.D == 11`, test/integration/sampling.jl:70
    assert spec.theta_names == ("M", "plx", "b.a", "b.e", "b.i", "b.ωx", "b.ωy", "b.Ωx", "b.Ωy", "b.θx", "b.θy")
    assert spec.input_names == ("M", "plx", "b.a", "b.e", "b.i", "b.ω", "b.Ω", "b.θ", "b.tp")
    ops = [d.op for d in spec.defs]
    assert ops == [0, 0, 0, 0, 0, 2, 2, 2, 3]
    assert list(spec.defs[8].a)[:7] == [7, 0, 3, 2, 4, 5, 6]      # θ, M, e, a, i, ω, Ω as kernel-input columns


def test_variable_vocabulary_validation():
    with pytest.raises(ValueError):
        octo.Uniform(1, 0)
    with pytest.raises(ValueError):
        octo.LogUniform(0, 1)
    with pytest.raises(ValueError):
        octo.truncated(octo.Uniform(0, 1), lower=0)
    t = octo.Table(epoch=[50000.], ra=[1.], dec=[1.], σ_ra=[1.], σ_dec=[1.])
    obs = octo.PlanetRelAstromObs(t, name="x")
    bad = octo.Planet(name="b", observations=[obs], variables={
        "tp": octo.θ_at_epoch_to_tperi("θ", 50000), "a": octo.Uniform(0, 1), "e": octo.Uniform(0, 0.9), "i": octo.Sine(),
        "ω": octo.UniformCircular(), "Ω": octo.UniformCircular(), "θ": octo.UniformCircular()})
    with pytest.raises(octo.OctoError, match="after the variables"):
        octo.ModelSpec(octo.System(name="s", companions=[bad], variables={"M": 1.0, "plx": 50.0}))
    # mixing bare names and priors => raw-input mode (no parameterisation)
    raw = octo.Planet(name="b", observations=[obs], variables=["a", "e", "i", "ω", "Ω", "tp"])
    assert octo.ModelSpec(octo.System(name="s", companions=[raw], variables={"M": 1.0, "plx": 50.0})).priors is None


@pytest.mark.parametrize("name", post_cases())
def test_oracle_logpost_matches_mpmath(oracle_lib, name):
    d, spec, consts = load_post(name)
    lp, g = oracle_lib.logpost(spec, consts, d["theta_t"])
    lpv = oracle_lib.logpost(spec, consts, d["theta_t"], grad=False)
    assert rel_err(lp[0], d["lp"]) < 1e-10 and lpv[0] == lp[0]
    assert grad_err(g, d["grad"]).max() < 1e-8


def test_oracle_gradient_vs_finite_differences(oracle_lib):
    spec = octo.ModelSpec(reference_test_system())
    c = octo.default_constants()
    rng = np.random.default_rng(3)
    th = rng.normal(0, 0.7, (4, spec.D))
    lp, g = oracle_lib.logpost(spec, c, th)
    h = 1e-6
    for j in range(spec.D):
        tp, tm = th.copy(), th.copy()
        tp[:, j] += h; tm[:, j] -= h
        fd = (oracle_lib.logpost(spec, c, tp, grad=False) - oracle_lib.logpost(spec, c, tm, grad=False)) / (2 * h)
        assert np.allclose(g[:, j], fd, rtol=1e-4, atol=1e-3)     # the reference's tolerances


def test_tperi_places_planet_at_position_angle(oracle_lib):
    """θ_at_epoch_to_tperi: with the returned tp the planet's position angle at t_ref equals θ."""
    import ctypes as C
    L = oracle_lib.lib()
    L.octo_oracle_tperi.restype = C.c_double
    L.octo_oracle_tperi.argtypes = [C.c_void_p] + [C.c_double] * 8
    c = octo.default_constants()
    rng = np.random.default_rng(5)
    for _ in range(50):
        theta, M, e, a = rng.uniform(-np.pi, np.pi), rng.uniform(0.5, 2), rng.uniform(0, 0.9), rng.uniform(1, 30)
        i, w, W = rng.uniform(0.1, 3.0), rng.uniform(0, 6.28), rng.uniform(0, 6.28)
        tp = L.octo_oracle_tperi(C.addressof(c), theta, 50000.0, M, e, a, i, w, W)
        ra, dec, _ = oracle_lib.orbit_radecrv(c, a, e, i, w, W, tp, M, 50.0, [50000.0])
        assert abs(np.angle(np.exp(1j * (np.arctan2(ra[0], dec[0]) - theta)))) < 1e-9


def test_invalid_and_healed(oracle_lib):
    spec = octo.ModelSpec(reference_test_system())
    c = octo.default_constants()
    th = np.zeros((3, spec.D)); th[:, 5:] = 0.7; th[:, 4] = -0.5      # (i = π/2 exactly makes tp ill-conditioned)
    th[1, 3] = np.nan                 # non-finite θ_t => -Inf (logdensitymodel.jl:120-124)
    th[2, 3] = 800.0                  # logistic saturates => e clamps to its upper bound => healed prior
    lp, g = oracle_lib.logpost(spec, c, th)
    assert np.isfinite(lp[0]) and lp[1] == -np.inf and np.all(g[1] == 0)
    assert lp[2] < -1e300 and np.isfinite(lp[2])
