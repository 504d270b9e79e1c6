"""Parity in the launch regimes the throughput numbers come from (VERDICT r1 "what's weak" 2).

The C5 sweep / `prof_kinds.py` figures run the THROUGHPUT instantiation on a multi-wave grid with tens to hundreds of
epochs per warp (many 32-record staging rounds per warp); C1-C4 and the older tests are all single-wave.  Here the CUDA
path is compared with the oracle in exactly those regimes, on a chain subset spread over the first, middle and last
chain groups (the oracle's dual-number gradient is the expensive side), plus bit-reproducibility of the full batch.
Tolerances: the north star's (1e-10 relative on logp, 1e-8 on the gradient)."""
import numpy as np
import pytest

import octofitter_jl_b200 as octo
import workloads
from helpers import grad_err, rel_err

pytestmark = pytest.mark.gpu
LOGP_RTOL, GRAD_RTOL = 1e-10, 1e-8
RESIDENT_SLOTS = 296          # 148 SMs x 2 CTAs of the throughput instantiation


def _subset(n, groups=(0, None, -1), per=21):
    """chains from the first, a middle and the last 32-chain group (ragged picks inside each)"""
    ng = (n + 31) // 32
    idx = []
    for g in groups:
        g = ng // 2 if g is None else (g % ng)
        lo, hi = 32 * g, min(32 * g + 32, n)
        idx += list(range(lo, hi))[:per] + [hi - 1]
    return np.unique(np.array(idx))


def _check(spec, x, oracle_lib, pick):
    model = octo.LogDensityModel(spec)
    n = x.shape[0]
    geom = model.launch_geometry(n)
    ll, g = model.ln_like_and_gradient(x)
    llv = model.ln_like(x)
    ll2, g2 = model.ln_like_and_gradient(x)
    model.close()
    assert np.array_equal(ll, ll2) and np.array_equal(g, g2)          # fixed reduction order
    orc = oracle_lib.Oracle(spec.packed, octo.default_constants())
    ll_o, g_o = orc.logp_grad(np.asfortranarray(x[pick]), threads=8)
    assert np.all(np.isfinite(ll_o))
    assert rel_err(ll[pick], ll_o).max() < LOGP_RTOL
    assert rel_err(llv[pick], ll_o).max() < LOGP_RTOL
    assert grad_err(g[pick], g_o).max() < GRAD_RTOL
    return geom


# the multi-wave geometry of the THROUGHPUT instantiation (two CTAs per SM, lane = chain, epoch splits combined through
# L2) — what the right end of the C5 sweep runs on — and whatever the library picks by default for the same batch (for
# 4096 chains x 4000 epochs: one wave of the latency instantiation, 500 dependent pairs per lane)
THROUGHPUT = {"OCTO_B200_LATENCY": "0", "OCTO_B200_SUBLANES": "1"}


@pytest.mark.parametrize("env", [THROUGHPUT, {}], ids=["throughput-multiwave", "default"])
def test_multiwave_throughput_regime_astrometry(oracle_lib, monkeypatch, env):
    """4096 chains x 4000 RA/Dec epochs: lean astrometry loop, > 32 epochs per warp (many 32-record staging rounds)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    spec, x = workloads.one_planet(4000, 0, 4096, seed=31)
    gx, gy, block, slice_ = _check(spec, x, oracle_lib, _subset(4096))
    assert slice_ > 32, slice_
    if env:
        assert gx * gy > RESIDENT_SLOTS, (gx, gy)


@pytest.mark.parametrize("env", [THROUGHPUT, {}], ids=["throughput-multiwave", "default"])
def test_multiwave_throughput_regime_rv_jitter(oracle_lib, monkeypatch, env):
    """4096 chains x 4000 star-RV epochs with offset and free jitter (per-pair variance, log per pair)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    spec, x = workloads.one_planet(0, 4000, 4096, seed=32)
    gx, gy, block, slice_ = _check(spec, x, oracle_lib, _subset(4096))
    assert slice_ > 32, slice_
    if env:
        assert gx * gy > RESIDENT_SLOTS, (gx, gy)


@pytest.mark.parametrize("env", [THROUGHPUT, {}], ids=["throughput-multiwave", "default"])
def test_mixed_tables_multiwave(oracle_lib, monkeypatch, env):
    """astrometry + RV in one model at 4096 x (1500 + 1500): a warp's range crosses the table boundary mid-slice."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    spec, x = workloads.one_planet(1500, 1500, 4096, seed=33)
    gx, gy, block, slice_ = _check(spec, x, oracle_lib, _subset(4096, per=10))
    if env:
        assert gx * gy > RESIDENT_SLOTS


def test_c5_right_end_geometry_is_the_multiwave_throughput_one():
    """4096 chains x 1e5 epochs (the C5 point the headline roofline fraction is quoted on) runs multi-wave on the
    throughput instantiation by default: no test above forces what the sweep does not actually do."""
    spec, x = workloads.one_planet(100000, 0, 64, seed=5)
    model = octo.LogDensityModel(spec)
    gx, gy, block, slice_, sub, lat = model.launch_geometry_full(4096)
    model.close()
    assert lat == 0 and sub == 1 and gx * gy > 4 * RESIDENT_SLOTS and slice_ >= 300


def test_many_chains_four_wave_branch(oracle_lib):
    """40 000 chains x 600 epochs: chain groups alone exceed four waves of resident CTAs (the `>= 4 waves` branch of
    the launch geometry); ragged last group (40 000 = 1250 x 32 exactly, so use 40 003)."""
    spec, x = workloads.one_planet(600, 0, 40003, seed=34)
    gx, gy, block, slice_ = _check(spec, x, oracle_lib, _subset(40003, per=8))
    assert gx >= 4 * RESIDENT_SLOTS and gx * gy >= 4 * RESIDENT_SLOTS


def test_c2_all_chains_against_oracle(oracle_lib):
    """Every one of C2's 1024 chains against the oracle (the config test checks the first 96 only)."""
    spec, x = workloads.config("C2")
    _check(spec, x, oracle_lib, np.arange(x.shape[0]))
