"""The sub-lane mapping (S epoch sub-lanes per chain inside a warp, 32/S chains per CTA; octo_kernels.cu "SUB-LANES"):
every sub-lane count against the oracle on every kind of model, the automatic choice for small batches, and the
invariants the mapping promises (results independent of what else is in the batch, run-to-run bit-reproducible).
Tolerances: 1e-10 relative on logp, 1e-8 on the gradient (the north star's)."""
import numpy as np
import pytest

import octofitter_jl_b200 as octo
import workloads
from helpers import grad_err, load_post, post_cases, reference_test_system, rel_err

pytestmark = pytest.mark.gpu
LOGP_RTOL, GRAD_RTOL = 1e-10, 1e-8
SUBS = ["2", "4", "8", "16", "32"]


def _eval(spec, x, monkeypatch, sub):
    monkeypatch.setenv("OCTO_B200_SUBLANES", sub)
    model = octo.LogDensityModel(spec)
    geom = model.launch_geometry_full(x.shape[0])
    ll, g = model.ln_like_and_gradient(x)
    llv = model.ln_like(x)
    ll2, g2 = model.ln_like_and_gradient(x)
    model.close()
    assert np.array_equal(ll, ll2) and np.array_equal(g, g2)
    return geom, ll, llv, g


@pytest.mark.parametrize("sub", SUBS)
def test_every_sublane_count_baseline_configs(oracle_lib, monkeypatch, sub):
    for cfg, n in (("C1", 1), ("C2", 203), ("C3", 77), ("C4", 64)):
        spec, x = workloads.config(cfg)
        x = np.asfortranarray(x[:n])
        geom, ll, llv, g = _eval(spec, x, monkeypatch, sub)
        assert geom[4] == int(sub) and geom[0] == -(-n // (32 // int(sub)))
        ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=8)
        assert rel_err(ll, ll_o).max() < LOGP_RTOL and rel_err(llv, ll_o).max() < LOGP_RTOL, cfg
        assert grad_err(g, g_o).max() < GRAD_RTOL, cfg


@pytest.mark.parametrize("sub", ["4", "32"])
def test_sublanes_every_observation_kind(oracle_lib, monkeypatch, sub):
    """4 planets, every table kind incl. reflex from interior companions, marginalised RV, observable prior, HGCA."""
    spec, x = workloads.many_planets(4, 45, seed=21, extras=True)
    geom, ll, llv, g = _eval(spec, x, monkeypatch, sub)
    assert geom[4] == int(sub)
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=8)
    assert np.all(np.isfinite(ll_o))
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and rel_err(llv, ll_o).max() < LOGP_RTOL
    assert grad_err(g, g_o).max() < GRAD_RTOL


@pytest.mark.parametrize("sub", ["2", "8"])
def test_sublanes_with_epoch_splits_across_ctas(oracle_lib, monkeypatch, sub):
    """Few chains, many epochs: sub-lanes inside the warp AND splits across CTAs (the L2 combine) together."""
    spec, x = workloads.one_planet(3000, 500, 40, seed=9)
    geom, ll, llv, g = _eval(spec, x, monkeypatch, sub)
    assert geom[4] == int(sub) and geom[1] > 1
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=8)
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and rel_err(llv, ll_o).max() < LOGP_RTOL
    assert grad_err(g, g_o).max() < GRAD_RTOL


def test_automatic_choice_small_batches(oracle_lib, monkeypatch):
    """What the library picks by itself: C2 runs 8 chains x 4 sub-lanes per warp on 128 CTAs without a cross-CTA
    combine; a single chain spreads its epochs over the lanes; a large batch keeps lane = chain."""
    monkeypatch.delenv("OCTO_B200_SUBLANES", raising=False)
    spec, x = workloads.config("C2")
    model = octo.LogDensityModel(spec)
    gx, gy, block, slice_, sub, lat = model.launch_geometry_full(1024)
    assert sub > 1 and gy == 1 and gx * gy <= 148, (gx, gy, sub)
    assert model.launch_geometry_full(1)[4] >= 4
    assert model.launch_geometry_full(32 * 148 * 2)[4] == 1
    ll, g = model.ln_like_and_gradient(x)
    n0 = model.kernel_launches
    ll1, g1 = model.ln_like_and_gradient(x[517])
    assert model.kernel_launches == n0 + 1
    model.close()
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=8)
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    assert rel_err(ll1, ll_o[517]) < LOGP_RTOL and grad_err(g1, g_o[517]).max() < GRAD_RTOL


def test_result_of_a_chain_does_not_depend_on_its_neighbours(monkeypatch):
    """Same geometry (forced sub-lane count), a chain evaluated alone / inside different batches: identical bits.
    This is what lets replicas sharded over GPUs reproduce a single-GPU run (tests/test_gpu_multi.py)."""
    monkeypatch.setenv("OCTO_B200_SUBLANES", "4")
    spec, x = workloads.config("C4")
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    for lo, hi in ((0, 8), (8, 16), (56, 64), (3, 4), (5, 41)):
        ll_s, g_s = model.ln_like_and_gradient(np.asfortranarray(x[lo:hi]))
        assert np.array_equal(ll_s, ll[lo:hi]) and np.array_equal(g_s, g[lo:hi])
    model.close()


@pytest.mark.parametrize("sub", ["1", "4", "16"])
def test_sublanes_fused_parameterisation(oracle_lib, monkeypatch, sub):
    """θ_t -> log posterior + gradient with the parameterisation fused into the kernel, under sub-lanes."""
    monkeypatch.setenv("OCTO_B200_SUBLANES", sub)
    for name in post_cases():
        d, spec, consts = load_post(name)
        import ctypes as C
        lib = octo.load_library()
        h = C.c_void_p()
        assert lib.octo_create(C.byref(consts), C.byref(spec.packed.layout), spec.packed.blocks, spec.packed.n_blocks, 0, C.byref(h)) == 0
        assert lib.octo_set_parameterization(h, spec.priors, spec.D, spec.defs) == 0
        rng = np.random.default_rng(8)
        th = np.asfortranarray(np.array(d["theta_t"])[None, :] + 0.05 * rng.standard_normal((45, spec.D)))
        n = th.shape[0]
        lp, g, lpv = np.empty(n), np.empty((n, spec.D), order="F"), np.empty(n)
        assert lib.octo_logpost_grad(h, th.ctypes.data, n, n, lp.ctypes.data, g.ctypes.data) == 0
        assert lib.octo_logpost_grad(h, th.ctypes.data, n, n, lpv.ctypes.data, None) == 0
        lib.octo_destroy(h)
        lp_o, g_o = oracle_lib.logpost(spec, consts, th, threads=4)
        assert rel_err(lp, lp_o).max() < LOGP_RTOL and rel_err(lpv, lp_o).max() < LOGP_RTOL, name
        assert grad_err(g, g_o).max() < GRAD_RTOL, name
    # the reference's 11-D test model, single chain (what its samplers pass) and a ragged batch
    spec = octo.ModelSpec(reference_test_system())
    model = octo.LogDensityModel(spec)
    rng = np.random.default_rng(2)
    th = rng.normal(0, 0.8, (37, 11)); th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(37)
    lp, g = model.ℓπcallback_grad(th)
    lp1, g1 = model.ℓπcallback_grad(th[5])
    ll1 = model.ln_like_of_theta(th[5])
    model.close()
    lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), th, threads=4)
    assert rel_err(lp, lp_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    assert rel_err(lp1, lp_o[5]) < LOGP_RTOL and grad_err(g1, g_o[5]).max() < GRAD_RTOL
    assert np.isfinite(ll1) and ll1 != lp1          # the likelihood part alone, also for an inlined single-chain call
