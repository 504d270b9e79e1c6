"""GPU parity tests: libocto_b200.so (through its C ABI) against the CPU oracle and the golden vectors.

Tolerances are the north_star's: 1e-10 relative on logp, 1e-8 on ∇logp (relative to the largest
component of each gradient row)."""
import os

import numpy as np
import pytest

import octofitter_jl_b200 as octo
import workloads
from helpers import golden_cases, grad_err, load_golden, rel_err

pytestmark = pytest.mark.gpu
LOGP_RTOL, GRAD_RTOL = 1e-10, 1e-8


class RawModel:
    """Drive the C ABI directly from packed structs (what the Julia ccall shim does)."""

    def __init__(self, packed, consts, device=0):
        import ctypes as C
        self.lib = octo.load_library()
        self.packed, self.consts, self.n_in = packed, consts, packed.layout.n_in
        h = C.c_void_p()
        rc = self.lib.octo_create(C.byref(consts), C.byref(packed.layout), packed.blocks, packed.n_blocks, device,
                                  C.byref(h))
        assert rc == 0, self.lib.octo_last_error()
        self.h = h

    def logp_grad(self, x):
        x = np.asfortranarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
        n = x.shape[0]
        ll, g = np.empty(n), np.empty((n, self.n_in), order="F")
        rc = self.lib.octo_logp_grad(self.h, x.ctypes.data, n, n, ll.ctypes.data, g.ctypes.data)
        assert rc == 0, self.lib.octo_last_error()
        return ll, g

    def logp(self, x):
        x = np.asfortranarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
        n = x.shape[0]
        ll = np.empty(n)
        rc = self.lib.octo_logp(self.h, x.ctypes.data, n, n, ll.ctypes.data)
        assert rc == 0, self.lib.octo_last_error()
        return ll

    def close(self):
        self.lib.octo_destroy(self.h)


@pytest.mark.parametrize("name", golden_cases())
def test_cuda_matches_golden(name):
    d, packed, consts = load_golden(name)
    m = RawModel(packed, consts)
    ll, g = m.logp_grad(d["x"])
    llv = m.logp(d["x"])
    m.close()
    assert rel_err(ll[0], d["ll"]) < LOGP_RTOL
    assert rel_err(llv[0], d["ll"]) < LOGP_RTOL
    assert grad_err(g, d["grad"]).max() < GRAD_RTOL


@pytest.mark.parametrize("name", golden_cases())
def test_cuda_matches_oracle_batch(oracle_lib, name):
    """Perturbed batches of each golden model, odd batch size (ragged last chain group)."""
    d, packed, consts = load_golden(name)
    rng = np.random.default_rng(5)
    x0 = np.array(d["x"])
    x = x0[None, :] * (1 + 0.01 * rng.standard_normal((77, len(x0))))
    for k, nme in enumerate(d["input_names"]):
        if nme.endswith(".e"):
            x[:, k] = np.clip(x[:, k], 0, 0.97)
    orc = oracle_lib.Oracle(packed, consts)
    ll_o, g_o = orc.logp_grad(x, threads=4)
    m = RawModel(packed, consts)
    ll, g = m.logp_grad(x)
    llv = m.logp(x)
    m.close()
    assert rel_err(ll, ll_o).max() < LOGP_RTOL
    assert rel_err(llv, ll_o).max() < LOGP_RTOL
    assert grad_err(g, g_o).max() < GRAD_RTOL


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4"])
def test_baseline_configs_match_oracle(oracle_lib, cfg):
    spec, x = workloads.config(cfg)
    consts = octo.default_constants()
    n = min(x.shape[0], 96)
    orc = oracle_lib.Oracle(spec.packed, consts)
    ll_o, g_o = orc.logp_grad(x[:n], threads=8)
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    llv = model.ln_like(x)
    assert ll.shape == (x.shape[0],) and g.shape == x.shape
    assert rel_err(ll[:n], ll_o).max() < LOGP_RTOL
    assert rel_err(llv[:n], ll_o).max() < LOGP_RTOL
    assert grad_err(g[:n], g_o).max() < GRAD_RTOL
    # run-to-run bit reproducibility (fixed reduction order)
    ll2, g2 = model.ln_like_and_gradient(x)
    assert np.array_equal(ll, ll2) and np.array_equal(g, g2)
    assert model.kernel_launches >= 3
    model.close()


@pytest.mark.parametrize("n_planets", [3, 4])
def test_many_planets_every_kind(oracle_lib, n_planets):
    """3 and 4 planets (the NPT=4 kernel instantiation): reflex astrometry from several interior companions in
    RA/Dec and PA/sep tables, relative RV with interior companions, star RV, marginalised RV."""
    spec, x = workloads.many_planets(n_planets, 45, seed=21)
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    llv = model.ln_like(x)
    model.close()
    orc = oracle_lib.Oracle(spec.packed, octo.default_constants())
    ll_o, g_o = orc.logp_grad(x, threads=8)
    assert np.all(np.isfinite(ll_o))
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and rel_err(llv, ll_o).max() < LOGP_RTOL
    assert grad_err(g, g_o).max() < GRAD_RTOL


def test_degenerate_shapes(oracle_lib):
    """Empty table next to a populated one, a single epoch, a model with no tables at all, zero chains."""
    import ctypes as C
    t1 = octo.Table(epoch=[50000.0], ra=[100.0], dec=[-50.0], σ_ra=[2.0], σ_dec=[3.0])
    t0 = octo.Table(epoch=[], ra=[], dec=[], σ_ra=[], σ_dec=[])
    for tabs in ([t1], [t0, t1], [t0], []):
        obs = [octo.PlanetRelAstromObs(t, name=f"o{k}") for k, t in enumerate(tabs)]
        b = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"], observations=obs)
        spec = octo.ModelSpec(octo.System(name="s", variables=["M", "plx"], companions=[b]))
        x = np.array([[1.2, 50.0, 10.0, 0.3, 1.0, 0.5, 2.0, 50000.0], [1.1, 49.0, 9.0, 0.2, 0.9, 0.4, 1.9, 50100.0]])
        model = octo.LogDensityModel(spec)
        ll, g = model.ln_like_and_gradient(x)
        orc = oracle_lib.Oracle(spec.packed, octo.default_constants())
        ll_o, g_o = orc.logp_grad(x)
        if spec.total_epochs == 0:
            assert np.all(ll == 0.0) and np.all(g == 0.0) and np.all(ll_o == 0.0)
        else:
            assert rel_err(ll, ll_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
        # zero chains: a no-op that succeeds
        assert model._lib.octo_logp_grad(model._h, None, 0, 0, None, None) == 0
        model.close()


def test_invalid_chains():
    spec, x = workloads.config("C2")
    x = np.array(x[:40])
    names = spec.input_names
    x[3, names.index("b.e")] = 1.0
    x[5, names.index("b.a")] = -2.0
    x[7, names.index("M")] = np.nan
    x[9, names.index("plx")] = 0.0
    x[11, names.index("rv.offset")] = np.inf
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    bad = [3, 5, 7, 9, 11]
    good = [k for k in range(40) if k not in bad]
    assert np.all(ll[bad] == -np.inf) and np.all(g[bad] == 0.0)
    assert np.all(np.isfinite(ll[good])) and np.all(np.isfinite(g[good]))
    model.close()


def test_leading_dimension_and_single_chain(oracle_lib):
    spec, x = workloads.config("C1")
    model = octo.LogDensityModel(spec)
    ll1, g1 = model.ln_like_and_gradient(x[0])
    assert np.isscalar(ll1) or ll1.shape == ()
    orc = oracle_lib.Oracle(spec.packed, octo.default_constants())
    ll_o, g_o = orc.logp_grad(x[:1])
    assert rel_err(ll1, ll_o[0]) < LOGP_RTOL and grad_err(g1, g_o[0]).max() < GRAD_RTOL
    # ld > n_chains through the raw ABI
    import ctypes as C
    n, ld = 5, 9
    xs = np.zeros((ld, spec.n_in), order="F"); xs[:n] = np.tile(x[0], (n, 1)) * (1 + 1e-3 * np.arange(n)[:, None])
    ll = np.full(ld, 123.0); g = np.full((ld, spec.n_in), 7.0, order="F")
    rc = model._lib.octo_logp_grad(model._h, xs.ctypes.data, n, ld, ll.ctypes.data, g.ctypes.data)
    assert rc == 0
    ll_o, g_o = orc.logp_grad(xs[:n])
    assert rel_err(ll[:n], ll_o).max() < LOGP_RTOL and grad_err(g[:n], g_o).max() < GRAD_RTOL
    assert np.all(ll[n:] == 123.0) and np.all(g[n:] == 7.0)
    model.close()


def test_epoch_split_geometry_consistent(oracle_lib, monkeypatch):
    """Many epochs, few chains, lane = chain mapping: the grid splits epochs across CTAs (K2 path) and still matches.
    (By default such a batch now splits its epochs over sub-lanes inside the warps: tests/test_gpu_sublanes.py.)"""
    monkeypatch.setenv("OCTO_B200_SUBLANES", "1")
    spec, x = workloads.one_planet(3000, 0, 40, seed=9)
    model = octo.LogDensityModel(spec)
    gx, gy, block, slice_ = model.launch_geometry(40)
    assert gy > 1
    ll, g = model.ln_like_and_gradient(x)
    orc = oracle_lib.Oracle(spec.packed, octo.default_constants())
    ll_o, g_o = orc.logp_grad(x, threads=8)
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    model.close()


@pytest.mark.parametrize("env", [{}, {"OCTO_B200_SUBLANES": "1"}, {"OCTO_B200_SUBLANES": "1", "OCTO_B200_SLICE": "1"},
                                 {"OCTO_B200_SUBLANES": "1", "OCTO_B200_SLICE": "40"},
                                 {"OCTO_B200_SUBLANES": "1", "OCTO_B200_CTAS_PER_SM": "1"}, {"OCTO_B200_FORCE": "2,0,3"}])
def test_launch_geometry_knobs_do_not_change_results(oracle_lib, env, monkeypatch):
    """Different epoch-split geometries (slice length, resident-CTA target) on C2 and the 2-planet C3."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for cfg in ("C2", "C3"):
        spec, x = workloads.config(cfg)
        x = x[:70]
        model = octo.LogDensityModel(spec)
        ll, g = model.ln_like_and_gradient(x)
        llv = model.ln_like(x)
        model.close()
        orc = oracle_lib.Oracle(spec.packed, octo.default_constants())
        ll_o, g_o = orc.logp_grad(x, threads=8)
        assert rel_err(ll, ll_o).max() < LOGP_RTOL and rel_err(llv, ll_o).max() < LOGP_RTOL
        assert grad_err(g, g_o).max() < GRAD_RTOL


def test_concurrent_calls_threads(oracle_lib):
    import threading
    spec, x = workloads.config("C4")
    model = octo.LogDensityModel(spec)
    ref = model.ln_like_and_gradient(x)
    out = [None] * 8

    def work(i):
        for _ in range(20):
            out[i] = model.ln_like_and_gradient(x)
    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
    for o in out:
        assert np.array_equal(o[0], ref[0]) and np.array_equal(o[1], ref[1])
    model.close()


def test_device_kepler_edge_cases(oracle_lib):
    """The device solve (FP32 Markley starter + FP64 fifth-order correction) against the oracle's Markley
    solve and Kepler's equation itself, including e -> 1, M -> 0, |M| = pi and many revolutions."""
    rng = np.random.default_rng(17)
    n = 200_000
    MA = np.concatenate([rng.uniform(-np.pi, np.pi, n), rng.uniform(-300, 300, n // 4),
                         10.0 ** rng.uniform(-30, -2, n // 4) * rng.choice([-1, 1], n // 4),
                         [0.0, np.pi, -np.pi, np.nextafter(np.pi, 4), 1e-300, 2 * np.pi, 12345.678]])
    e = np.concatenate([rng.uniform(0, 1, n) ** 0.3, rng.uniform(0, 0.99, n // 4),
                        1 - 10.0 ** rng.uniform(-9, -1, n // 4), [0.0, 0.5, 0.5, 0.5, 0.999, 0.3, 0.3]])
    e = np.clip(e, 0, 1 - 1e-9)
    s, c = np.empty_like(MA), np.empty_like(MA)
    lib = octo.load_library()
    rc = lib.octo_selftest_kepler(0, MA.ctypes.data, e.ctypes.data, len(MA), s.ctypes.data, c.ctypes.data)
    assert rc == 0, lib.octo_last_error()
    assert np.all(np.isfinite(s)) and np.all(np.isfinite(c))
    assert np.abs(s * s + c * c - 1).max() < 1e-15
    E = np.arctan2(s, c)
    M = np.array([oracle_lib.lib().octo_oracle_rem2pi(float(x)) for x in MA[-(n // 2 + 7):]])
    # Kepler's equation residual, modulo 2 pi (|M| = pi may come back as the other branch)
    res = E - e * s - np.concatenate([np.zeros(len(MA) - len(M)), M])
    res[: len(MA) - len(M)] -= MA[: len(MA) - len(M)]
    res = res - 2 * np.pi * np.round(res / (2 * np.pi))
    assert np.abs(res).max() < 3e-15
    # against the oracle's Markley solve where the problem is well conditioned
    idx = rng.choice(len(MA), 5000, replace=False)
    Eo = np.array([oracle_lib.kepler(float(MA[i]), float(e[i])) for i in idx])
    ok = e[idx] < 0.99
    assert np.abs(np.sin(Eo[ok]) - s[idx][ok]).max() < 2e-14
    assert np.abs(np.cos(Eo[ok]) - c[idx][ok]).max() < 2e-14


def test_create_destroy_cycles_and_workspace_growth():
    """Contexts are independent and reclaim everything; a workspace grows with the batch and shrinking is a no-op."""
    import torch
    spec, x = workloads.config("C2")
    free0 = torch.cuda.mem_get_info()[0]
    ref = None
    for it in range(30):
        model = octo.LogDensityModel(spec)
        for n in (1, 700, 33, 1024):
            ll, g = model.ln_like_and_gradient(x[:n])
            if n == 1024:
                if ref is None:
                    ref = (ll.copy(), g.copy())
                assert np.array_equal(ll, ref[0]) and np.array_equal(g, ref[1])
        model.close()
    torch.cuda.synchronize()
    assert free0 - torch.cuda.mem_get_info()[0] < 64 << 20        # nothing substantial left behind


def test_device_buffer_entry_point_matches_host_entry_point():
    import torch
    spec, x = workloads.config("C3")
    model = octo.LogDensityModel(spec)
    n, n_in = x.shape
    ll_h, g_h = model.ln_like_and_gradient(x)
    d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    d_ll = torch.empty(n, dtype=torch.float64, device="cuda")
    d_g = torch.empty((n_in, n), dtype=torch.float64, device="cuda")
    for stream in (torch.cuda.current_stream(), torch.cuda.Stream()):
        with torch.cuda.stream(stream):
            d_ll.zero_(); d_g.zero_()
            model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), d_g.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        assert np.array_equal(d_ll.cpu().numpy(), ll_h) and np.array_equal(d_g.cpu().numpy().T, g_h)
    # value-only through the device entry point
    d_ll.zero_()
    model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.max(np.abs(d_ll.cpu().numpy() - ll_h) / np.abs(ll_h)) < 1e-13
    model.close()


def test_extreme_orbits_stress(oracle_lib):
    """Random orbits across the whole Keplerian domain — e up to 0.995, a from 0.05 to 500 AU (periods of days to
    millennia, i.e. |mean anomaly| up to ~1e5 rad), face-on to edge-on — on an astrometry + RV model."""
    spec, x0 = workloads.one_planet(60, 60, 1, seed=31)
    names = list(spec.input_names)
    rng = np.random.default_rng(32)
    n = 600
    x = np.tile(x0[0], (n, 1))
    x[:, names.index("b.a")] = 10 ** rng.uniform(np.log10(0.05), np.log10(500), n)
    x[:, names.index("b.e")] = np.concatenate([rng.uniform(0, 0.9, n // 2), rng.uniform(0.9, 0.995, n - n // 2)])
    x[:, names.index("b.i")] = rng.uniform(0, np.pi, n)
    x[:, names.index("b.ω")] = rng.uniform(-2 * np.pi, 4 * np.pi, n)
    x[:, names.index("b.Ω")] = rng.uniform(-2 * np.pi, 4 * np.pi, n)
    x[:, names.index("b.tp")] = rng.uniform(20000, 80000, n)
    x[:, names.index("M")] = rng.uniform(0.1, 5, n)
    x[:, names.index("b.mass")] = rng.uniform(0, 80, n)
    x[:, names.index("rv.jitter")] = 10 ** rng.uniform(-2, 2, n)
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    model.close()
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=8)
    assert np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
    assert rel_err(ll, ll_o).max() < LOGP_RTOL
    # the Kepler problem is ill-conditioned as e -> 1 (dE/dM = 1/(1 - e cosE)): both sides lose the same digits, so the
    # bound is the north star's 1e-8 away from the singular corner and scaled by 1/(1-e) inside it
    cond = 1.0 / (1.0 - x[:, names.index("b.e")])
    assert (grad_err(g, g_o).max(axis=1) / np.maximum(1.0, cond / 10)).max() < GRAD_RTOL


def test_plain_c_consumer(oracle_lib, tmp_path):
    """The boundary is a C ABI: a plain C program (gcc, dlopen — what Julia's ccall amounts to) builds the model from
    the structs of include/octo_b200.h and gets the same numbers as the oracle and the mpmath golden vector."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "c_abi_smoke"
    subprocess.run(["gcc", "-O1", "-o", str(exe), os.path.join(root, "tests", "c_abi_smoke.c"), "-ldl"], check=True)
    out = subprocess.run([str(exe), octo.LIB_PATH], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    vals = np.array([[float.fromhex(t) for t in ln.split()] for ln in out])
    d, packed, consts = load_golden("case_fixture8")
    assert rel_err(vals[0, 0], d["ll"]) < LOGP_RTOL and grad_err(vals[0, 1:], d["grad"]).max() < GRAD_RTOL
    x = np.array([d["x"], [v * (1.0 if k == 7 else 1.01) for k, v in enumerate(d["x"])]])
    ll_o, g_o = oracle_lib.Oracle(packed, consts).logp_grad(x)
    assert rel_err(vals[:, 0], ll_o).max() < LOGP_RTOL and grad_err(vals[:, 1:], g_o).max() < GRAD_RTOL


def _single_epoch_model(spec, e):
    """The packed model reduced to epoch e of the concatenated tables (other tables empty), as
    generate_system_per_epoch (src/cross-validation.jl:453-497) does with likeobj_from_epoch_subset."""
    blocks, start = [], 0
    for b in spec.block_dicts:
        n = len(b["epoch"])
        sel = [e - start] if start <= e < start + n else []
        nb = dict(b)
        for k in ("epoch", "y1", "y2", "s1", "s2", "cor"):
            nb[k] = None if b.get(k) is None else np.asarray(b[k])[sel]
        blocks.append(nb)
        start += n
    blocks = [b for b in blocks if len(b["epoch"])]
    return octo.pack(spec.layout_dict, blocks)


@pytest.mark.parametrize("which", ["two_planet", "many_planets"])
def test_pointwise_like_matches_single_epoch_models(oracle_lib, which):
    """octo_logp_pointwise: column e equals the oracle's ln_like of the model reduced to epoch e alone, for every
    observation kind (incl. the marginalised RV formula on one epoch); the columns of additive kinds sum to ln_like."""
    import workloads
    if which == "two_planet":
        spec, x = workloads.two_planet(24, seed=3, n_b=9, n_c=7, n_rv=8)
    else:
        spec, x = workloads.many_planets(3, 24, seed=5, n_ep=4)
    model = octo.LogDensityModel(spec)
    x[2, spec.layout_dict["planets"][0]["e"]] = 1.5          # invalid chain: -Inf in every column
    E = model.total_epochs
    out = np.empty((x.shape[0], E), order="F")
    xf = np.asfortranarray(x)
    rc = model._lib.octo_logp_pointwise(model._h, xf.ctypes.data, x.shape[0], x.shape[0], out.ctypes.data, x.shape[0])
    assert rc == 0, model._lib.octo_last_error()
    consts = octo.default_constants()
    for e in range(E):
        ll_o = oracle_lib.Oracle(_single_epoch_model(spec, e), consts).logp(x)
        assert np.isneginf(out[2, e]) and np.isneginf(ll_o[2])
        ok = np.arange(x.shape[0]) != 2
        assert rel_err(out[ok, e], ll_o[ok]).max() < LOGP_RTOL, (which, e)
    has_margin = any(b["kind"] == octo.KIND_RV_STAR_MARGIN for b in spec.block_dicts)
    if not has_margin:
        assert rel_err(out[ok].sum(axis=1), model.ln_like(x)[ok]).max() < 1e-12
    # host mirror: reference column order (system-level tables first) and the matching epoch vector
    LL, epochs = model.pointwise_like(x)
    assert LL.shape == (x.shape[0], E) and epochs.shape == (E,)
    n_sys = sum(len(b["epoch"]) for b in spec.block_dicts if b["planet"] < 0)
    assert np.array_equal(LL[:, :n_sys], out[:, E - n_sys:]) and np.array_equal(LL[:, n_sys:], out[:, :E - n_sys])


def test_observable_prior_many_chains(oracle_lib):
    """ObsPriorAstromONeil2019-wrapped tables (RA/Dec next to its own wrapped copy as in the reference's docstring,
    and a wrapped sep/PA table with jitter / northangle and the reflex of an inner massive planet): value, gradient
    and the per-epoch pointwise values against the oracle over a cloud of chains."""
    d, packed, consts = load_golden("case_obsprior")
    import ctypes as C
    lib = octo.load_library()
    h = C.c_void_p()
    assert lib.octo_create(C.byref(consts), C.byref(packed.layout), packed.blocks, packed.n_blocks, 0, C.byref(h)) == 0, lib.octo_last_error()
    rng = np.random.default_rng(77)
    x0 = np.array(d["x"])
    n, n_in = 300, len(x0)
    x = np.asfortranarray(x0[None, :] * (1.0 + 0.01 * rng.standard_normal((n, n_in))))
    x[:, d["input_names"].index("b.e")] = rng.uniform(0.0, 0.9, n)
    x[1, d["input_names"].index("b.e")] = 0.0
    x[0] = x0
    ll = np.empty(n); g = np.empty((n, n_in), order="F"); llv = np.empty(n)
    assert lib.octo_logp_grad(h, x.ctypes.data, n, n, ll.ctypes.data, g.ctypes.data) == 0, lib.octo_last_error()
    assert lib.octo_logp(h, x.ctypes.data, n, n, llv.ctypes.data) == 0
    assert rel_err(ll[0], d["ll"]) < LOGP_RTOL and grad_err(g[0], d["grad"]).max() < GRAD_RTOL
    ora = oracle_lib.Oracle(packed, consts)
    ll_o, g_o = ora.logp_grad(x, threads=4)
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and rel_err(llv, ll_o).max() < LOGP_RTOL
    assert grad_err(g, g_o).max() < GRAD_RTOL
    lib.octo_destroy(h)


def test_hgca_many_chains_and_split_geometry(oracle_lib):
    """HGCAInstantaneousObs (evaluated by the last CTA of a chain group after the epoch sums are complete) next to an
    astrometry table: value and gradient against the oracle over a cloud of chains, with and without epoch splits,
    through the host mirror's class; HGCA alone (no epoch table at all) as well."""
    d, packed, consts = load_golden("case_hgca")
    import ctypes as C
    lib = octo.load_library()
    rng = np.random.default_rng(78)
    x0 = np.array(d["x"])
    n, n_in = 150, len(x0)
    x = np.asfortranarray(x0[None, :] * (1.0 + 0.02 * rng.standard_normal((n, n_in))))
    x[3, d["input_names"].index("b.e")] = -0.1                  # invalid chain
    ora = oracle_lib.Oracle(packed, consts)
    ll_o, g_o = ora.logp_grad(x, threads=4)
    # default geometry; lane = chain with epoch splits across CTAs (the L2 combine before the HGCA tail); sub-lanes with splits
    for env in ({}, {"OCTO_B200_SUBLANES": "1", "OCTO_B200_SLICE": "2"}, {"OCTO_B200_FORCE": "1,1,3"}, {"OCTO_B200_FORCE": "4,0,2"}):
        os.environ.update(env)
        try:
            h = C.c_void_p()
            assert lib.octo_create(C.byref(consts), C.byref(packed.layout), packed.blocks, packed.n_blocks, 0, C.byref(h)) == 0, lib.octo_last_error()
        finally:
            for k in env:
                os.environ.pop(k, None)
        assert lib.octo_total_epochs(h) == 6                        # the astrometry table only
        ll = np.empty(n); g = np.empty((n, n_in), order="F"); llv = np.empty(n)
        assert lib.octo_logp_grad(h, x.ctypes.data, n, n, ll.ctypes.data, g.ctypes.data) == 0, lib.octo_last_error()
        assert lib.octo_logp(h, x.ctypes.data, n, n, llv.ctypes.data) == 0
        ok = np.arange(n) != 3
        assert np.isneginf(ll[3]) and (g[3] == 0).all()
        assert rel_err(ll[ok], ll_o[ok]).max() < LOGP_RTOL and rel_err(llv[ok], ll_o[ok]).max() < LOGP_RTOL
        assert grad_err(g[ok], g_o[ok]).max() < GRAD_RTOL
        lib.octo_destroy(h)
    # HGCA only
    blocks = [b for b in d["blocks"] if b["kind"] == octo.KIND_HGCA_INSTANT]
    packed2 = octo.pack(d["layout"], blocks)
    h = C.c_void_p()
    assert lib.octo_create(C.byref(consts), C.byref(packed2.layout), packed2.blocks, packed2.n_blocks, 0, C.byref(h)) == 0, lib.octo_last_error()
    ll = np.empty(n); g = np.empty((n, n_in), order="F")
    assert lib.octo_logp_grad(h, x.ctypes.data, n, n, ll.ctypes.data, g.ctypes.data) == 0, lib.octo_last_error()
    ll_o2, g_o2 = oracle_lib.Oracle(packed2, consts).logp_grad(x, threads=4)
    assert rel_err(ll[ok], ll_o2[ok]).max() < LOGP_RTOL and grad_err(g[ok], g_o2[ok]).max() < GRAD_RTOL
    lib.octo_destroy(h)


@pytest.mark.parametrize("n_planets", [2, 4])
def test_every_observation_kind_in_one_model(oracle_lib, n_planets):
    """All kinds at once — RA/Dec (+cor, +jitter), PA/sep (+platescale, northangle), relative RV, star RV, marginalised
    RV, observable-prior wrappers and HGCA — with up to 4 planets (the CTA then runs with fewer warps to fit shared
    memory): value, gradient and value-only launch against the oracle."""
    import workloads
    spec, x = workloads.many_planets(n_planets, 70, seed=11, n_ep=12, extras=True)
    assert spec.n_in <= 48
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=4)
    assert np.isfinite(ll_o).all()
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    assert rel_err(model.ln_like(x), ll_o).max() < LOGP_RTOL
    geom = model.launch_geometry(70)
    assert geom[2] in (64, 128, 256, 384)          # 384: the latency-tuned instantiation (12 warps)


def test_thiele_innes_planets_with_hgca_and_observable_prior(oracle_lib):
    """Three Thiele-Innes planets (two massive): astrometry with reflex terms, an observable-prior wrapper and HGCA; the
    gradient w.r.t. A, B, F, G and plx runs through the virtual a-column of each planet."""
    d, packed, consts = load_golden("case_thiele_innes")
    import workloads
    rng = np.random.default_rng(91)
    base = dict(zip(d["input_names"], d["x"]))
    astrom_b = octo.PlanetRelAstromObs(octo.Table(epoch=d["blocks"][0]["epoch"], ra=d["blocks"][0]["y1"], dec=d["blocks"][0]["y2"],
                                                  σ_ra=d["blocks"][0]["s1"], σ_dec=d["blocks"][0]["s2"]), name="relastrom")
    ti = ["A", "B", "F", "G", "e", "tp"]
    pb = octo.Planet(name="b", basis="ThieleInnesOrbit", variables=ti + ["mass"], observations=[astrom_b, octo.ObsPriorAstromONeil2019(astrom_b)])
    pc = octo.Planet(name="c", basis="ThieleInnesOrbit", variables=ti + ["mass"])
    pd = octo.Planet(name="d", basis="ThieleInnesOrbit", variables=ti + ["mass"])
    hg = octo.HGCAInstantaneousObs(workloads.HGCA_ROW, N_ave=2)
    system = octo.System(name="ti3", variables=["M", "plx", "pmra", "pmdec"], companions=[pb, pc, pd], observations=[hg])
    spec = octo.ModelSpec(system)
    x0 = {"M": 1.21, "plx": 50.01, "pmra": 10.6, "pmdec": -4.9}
    for nm in ("b", "c"):
        x0.update({f"{nm}.{k}": base[f"{nm}.{k}"] for k in ti})
    x0.update({"b.mass": 3.0, "c.mass": 25.0, "d.mass": 8.0})
    x0.update({f"d.{k}": 0.4 * base[f"c.{k}"] if k in "ABFG" else base[f"c.{k}"] for k in ti})
    x0["d.A"], x0["d.e"] = x0["d.A"] + 30.0, 0.3
    xv = np.array([x0[nm] for nm in spec.input_names])
    x = xv[None, :] * (1.0 + 0.01 * rng.standard_normal((60, len(xv))))
    model = octo.LogDensityModel(spec)
    ll, g = model.ln_like_and_gradient(x)
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=4)
    assert np.isfinite(ll_o).all()
    assert rel_err(ll, ll_o).max() < LOGP_RTOL and grad_err(g, g_o).max() < GRAD_RTOL
    with pytest.raises(octo.OctoError):
        rv = octo.StarAbsoluteRVObs(octo.Table(epoch=[5e4], rv=[1.0], σ_rv=[1.0]), name="rv", variables=["offset", "jitter"])
        octo.ModelSpec(octo.System(name="bad", variables=["M", "plx"], companions=[pc], observations=[rv]))


def test_latency_and_throughput_instantiations_agree(oracle_lib, monkeypatch):
    """Small grids run the latency-tuned instantiation of the kernel (no register cap, one CTA per SM, fewer epoch
    splits), large ones the throughput instantiation; OCTO_B200_LATENCY=0 forces the latter.  Same results (different
    summation trees: agreement to rounding), both against the oracle."""
    import workloads
    spec, x = workloads.config("C2")
    x = x[:512]
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("OCTO_B200_LATENCY", mode)
        model = octo.LogDensityModel(spec)
        out[mode] = model.ln_like_and_gradient(x) + (model.launch_geometry(512),)
        model.close()
    assert out["1"][2][:2] != out["0"][2][:2] and out["1"][2][0] * out["1"][2][1] <= 148     # fewer, fatter CTAs
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x, threads=4)
    for mode in ("1", "0"):
        assert rel_err(out[mode][0], ll_o).max() < LOGP_RTOL and grad_err(out[mode][1], g_o).max() < GRAD_RTOL
    assert rel_err(out["1"][0], out["0"][0]).max() < 1e-13


@pytest.mark.parametrize("post", [False, True], ids=["likelihood", "fused-posterior"])
def test_lean_and_full_kernel_families_agree(oracle_lib, monkeypatch, post):
    """A model made of plain RA/Dec astrometry and RV tables runs the LEAN kernel family (compiled without the code of
    the other tables: round 2); OCTO_B200_NO_LEAN=1 sends it through the full kernels.  Same sums in
    the same order — agreement to rounding (the compiler contracts per instantiation) — and both against the oracle;
    in both launch regimes (C2-sized: latency instantiation with sub-lanes; 4096 x 400: throughput, multi-wave)."""
    import workloads
    cases = [workloads.one_planet_with_priors(100, 100, 600, seed=3)] if post else \
            [workloads.config("C2"), workloads.one_planet(200, 200, 4096, seed=9)]
    for spec, x in cases:
        out = {}
        for mode in ("0", "1"):
            if mode == "1":
                monkeypatch.setenv("OCTO_B200_NO_LEAN", "1")
            else:
                monkeypatch.delenv("OCTO_B200_NO_LEAN", raising=False)
            model = octo.LogDensityModel(spec)
            out[mode] = model.ℓπcallback_grad(x) if post else model.ln_like_and_gradient(x)
            model.close()
        assert rel_err(out["0"][0], out["1"][0]).max() < 1e-13
        assert grad_err(out["0"][1], out["1"][1]).max() < 1e-11
        pick = np.arange(0, x.shape[0], max(1, x.shape[0] // 64))
        if post:
            lp_o, g_o = oracle_lib.logpost(spec, octo.default_constants(), x[pick], threads=4)
        else:
            lp_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(np.asfortranarray(x[pick]), threads=4)
        for mode in ("0", "1"):
            assert rel_err(out[mode][0][pick], lp_o).max() < LOGP_RTOL and grad_err(out[mode][1][pick], g_o).max() < GRAD_RTOL


@pytest.mark.parametrize("k", [2, 3, 4])
def test_lean_multi_planet_kernels(oracle_lib, k):
    """2, 3 and 4 planets with lean tables only (plain RA/Dec astrometry with the reflex of the inner planets, star RV
    with offset and jitter): the multi-planet instantiations of the LEAN kernel family (one pair in flight per lane), in
    the latency regime (70 chains: sub-lanes) and in the throughput regime (6000 chains: lane = chain, two CTAs per SM)."""
    import workloads
    for n in (70, 6000):
        spec, x = workloads.k_planets_lean(k, n, seed=20 + k)
        model = octo.LogDensityModel(spec)
        ll, g = model.ln_like_and_gradient(x)
        llv = model.ln_like(x)
        geom = model.launch_geometry_full(n)
        model.close()
        assert geom[5] == (1 if n == 70 else 0), geom
        pick = np.arange(n) if n == 70 else np.r_[0:40, n // 2:n // 2 + 40, n - 40:n]
        ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(np.asfortranarray(x[pick]), threads=4)
        assert np.isfinite(ll_o).all()
        assert rel_err(ll[pick], ll_o).max() < LOGP_RTOL and grad_err(g[pick], g_o).max() < GRAD_RTOL
        assert rel_err(llv[pick], ll_o).max() < LOGP_RTOL


def test_pinned_inputs_read_in_place_match_the_copy_path(monkeypatch):
    """Page-locked inputs up to OCTO_B200_ZEROCOPY_MAX bytes are read by the kernel in place (no H2D copy launch);
    larger ones, or with the knob at 0, go through the copy engine.  Same kernel, same inputs: identical bits — blocking
    call, asynchronous halves, log posterior, a leading dimension larger than the batch."""
    import workloads
    spec, th = workloads.one_planet_with_priors(60, 40, 300, seed=8)
    out = {}
    for zmax in ("524288", "0", "1000"):          # in place / always copy / in place only for tiny inputs (here: copy)
        monkeypatch.setenv("OCTO_B200_ZEROCOPY_MAX", zmax)
        model = octo.LogDensityModel(spec)
        thp = model.pinned_empty(th.shape)
        thp[...] = th
        o1 = (model.pinned_empty(300), model.pinned_empty(th.shape))
        lp, g = model.ℓπcallback_grad(thp, out=o1)
        lp, g = lp.copy(), g.copy()
        hs = [model.ℓπcallback_grad_begin(thp, out=(model.pinned_empty(300), model.pinned_empty(th.shape))) for _ in range(3)]
        rs = [h.wait() for h in hs]
        for r in rs:
            assert np.array_equal(r[0], lp) and np.array_equal(r[1], g)
        out[zmax] = (lp, g)
        model.close()
    for z in ("0", "1000"):
        assert np.array_equal(out["524288"][0], out[z][0]) and np.array_equal(out["524288"][1], out[z][1])
    # likelihood entry point with ld > n through the C ABI
    spec2, x2 = workloads.config("C2")
    res = {}
    for zmax in ("524288", "0"):
        monkeypatch.setenv("OCTO_B200_ZEROCOPY_MAX", zmax)
        model = octo.LogDensityModel(spec2)
        n, ld, k = 200, 256, spec2.n_in
        xin = model.pinned_empty((ld, k)); xin[:n] = x2[:n]; xin[n:] = np.nan
        ll = model.pinned_empty(n); gg = model.pinned_empty((ld, k))
        assert model._lib.octo_logp_grad(model._h, xin.ctypes.data, n, ld, ll.ctypes.data, gg.ctypes.data) == 0
        res[zmax] = (ll.copy(), gg[:n].copy())
        model.close()
    assert np.array_equal(res["0"][0], res["524288"][0]) and np.array_equal(res["0"][1], res["524288"][1])
    assert np.isfinite(res["0"][0]).all()


def test_pointwise_like_more_epochs_than_one_grid(oracle_lib):
    """More than 65535 epochs: octo_logp_pointwise walks the epoch list in chunks.  The columns sum to ln_like and
    spot-checked columns equal the oracle on the one-epoch model."""
    import workloads
    spec, x = workloads.one_planet(70_000, 0, 6, seed=8)
    model = octo.LogDensityModel(spec)
    LL, epochs = model.pointwise_like(x)
    assert LL.shape == (6, 70_000) and np.all(np.diff(epochs) >= 0)
    assert rel_err(LL.sum(axis=1), model.ln_like(x)).max() < 1e-11
    consts = octo.default_constants()
    for e in (0, 65_534, 65_535, 65_536, 69_999):
        ll_o = oracle_lib.Oracle(_single_epoch_model(spec, e), consts).logp(x)
        assert rel_err(LL[:, e], ll_o).max() < LOGP_RTOL, e


def test_asynchronous_halves_of_the_host_call(oracle_lib):
    """octo_logp_grad_begin / octo_ready / octo_wait (and the log-posterior pair): several evaluations in flight on
    their own streams give the bits of the blocking calls; pageable and pinned buffers; an empty batch."""
    import ctypes as C
    spec, x = workloads.config("C2")
    model = octo.LogDensityModel(spec)
    ref = model.ln_like_and_gradient(x)
    xs = [np.asfortranarray(x[: 1024 - 37 * k]) for k in range(5)]
    pend = [model.ln_like_and_gradient_begin(xi) for xi in xs]
    for xi, h in zip(xs, pend):
        ll, g = h.wait()
        assert h.ready()
        assert np.array_equal(ll, ref[0][: xi.shape[0]]) and np.array_equal(g, ref[1][: xi.shape[0]])
    xp = model.pinned_empty(x.shape); xp[...] = x
    out = (model.pinned_empty(x.shape[0]), model.pinned_empty(x.shape))
    h = model.ln_like_and_gradient_begin(xp, out=out)
    ll, g = h.wait()
    assert np.array_equal(ll, ref[0]) and np.array_equal(g, ref[1])
    t = C.c_void_p()
    assert model._lib.octo_logp_grad_begin(model._h, None, 0, 0, None, None, C.byref(t)) == 0
    assert model._lib.octo_ready(t) == 1 and model._lib.octo_wait(t) == 0
    assert model._lib.octo_logp_grad_begin(model._h, None, 5, 5, None, None, C.byref(t)) != 0      # bad buffers
    model.close()
    import helpers
    spec_p = octo.ModelSpec(helpers.reference_test_system())
    mp_ = octo.LogDensityModel(spec_p)
    rng = np.random.default_rng(2)
    th = rng.normal(0, 0.8, (65, 11)); th[:, 1] = np.log(50.0 - 0.1) + 1e-3 * rng.standard_normal(65)
    lp0, g0 = mp_.ℓπcallback_grad(th)
    hs = [mp_.ℓπcallback_grad_begin(th) for _ in range(3)]
    for h in hs:
        lp, g = h.wait()
        assert np.array_equal(lp, lp0) and np.array_equal(g, g0)
    mp_.close()


def test_tables_are_validated_at_create():
    """A zero / negative / non-finite uncertainty, a non-finite datum or epoch, |cor| > 1 - 1e-5 (the reference ctor's own
    check, relative-astrometry.jl:69-71) are refused with OCTO_ERR_ARG instead of turning every chain into NaN."""
    import ctypes as C
    lib = octo.load_library()
    consts = octo.default_constants()
    base = dict(epoch=[50000.0, 50100.0], ra=[100.0, 90.0], dec=[-50.0, -40.0], σ_ra=[2.0, 2.0], σ_dec=[3.0, 3.0])
    for col, bad in (("σ_ra", 0.0), ("σ_dec", -1.0), ("σ_ra", np.nan), ("ra", np.inf), ("epoch", np.nan), ("cor", 0.999995)):
        cols = dict(base); cols.setdefault("cor", [0.0, 0.0])
        cols[col] = [cols[col][0], bad]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tab = {k: np.asarray(v, dtype=np.float64) for k, v in cols.items()}
        layout = {"n_in": 8, "planets": [{"M": 0, "plx": 1, "a": 2, "e": 3, "i": 4, "w": 5, "W": 6, "tp": 7, "mass": -1}]}
        blk = {"kind": 0, "planet": 0, "epoch": tab["epoch"], "y1": tab["ra"], "y2": tab["dec"], "s1": tab["σ_ra"], "s2": tab["σ_dec"], "cor": tab["cor"]}
        packed = octo.pack(layout, [blk])
        h = C.c_void_p()
        rc = lib.octo_create(C.byref(consts), C.byref(packed.layout), packed.blocks, packed.n_blocks, 0, C.byref(h))
        assert rc == 1 and not h.value, (col, bad)
        assert b"table 0" in lib.octo_last_error() or b"cor" in lib.octo_last_error()
    # the RV uncertainty too
    blk = {"kind": 2, "planet": -1, "epoch": [50000.0], "y1": [3.0], "s1": [0.0], "idx_offset": -1, "idx_jitter": -1}
    layout = {"n_in": 9, "planets": [{"M": 0, "plx": 1, "a": 2, "e": 3, "i": 4, "w": 5, "W": 6, "tp": 7, "mass": 8}]}
    packed = octo.pack(layout, [blk])
    h = C.c_void_p()
    assert lib.octo_create(C.byref(consts), C.byref(packed.layout), packed.blocks, packed.n_blocks, 0, C.byref(h)) == 1


def test_release_stream_and_threads_sharing_a_caller_stream():
    """Device-buffer entry point: several host threads enqueue on the SAME caller stream (serialised per stream by the
    library); octo_release_stream drops the stream's workspace."""
    import threading
    import torch
    spec, x = workloads.one_planet(600, 0, 40, seed=9)        # few chains, many epochs: uses the per-stream partial buffer
    model = octo.LogDensityModel(spec)
    ref = model.ln_like_and_gradient(x)
    n, n_in = x.shape
    d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    st = torch.cuda.Stream()
    outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty((n_in, n), dtype=torch.float64, device="cuda")) for _ in range(6)]

    def work(k):
        for _ in range(25):
            model.enqueue_device(d_in.data_ptr(), n, n, outs[k][0].data_ptr(), outs[k][1].data_ptr(), st.cuda_stream)
    th = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    [t.start() for t in th]; [t.join() for t in th]
    st.synchronize()
    for ll, g in outs:
        assert np.array_equal(ll.cpu().numpy(), ref[0]) and np.array_equal(g.cpu().numpy().T, ref[1])
    assert model._lib.octo_release_stream(model._h, st.cuda_stream) == 0
    assert model._lib.octo_release_stream(model._h, st.cuda_stream) == 0          # idempotent
    model.enqueue_device(d_in.data_ptr(), n, n, outs[0][0].data_ptr(), outs[0][1].data_ptr(), st.cuda_stream)
    st.synchronize()
    assert np.array_equal(outs[0][0].cpu().numpy(), ref[0])
    model.close()
