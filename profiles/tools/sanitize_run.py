"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): C2- and C3-shaped models,
ragged batch sizes, value and gradient kernels, epoch-split path; the fused parameterisation stage (and its
three-launch fallback), the likelihood-of-theta mode, pointwise mode, an observable-prior table, HGCA, Thiele-Innes
planets and the device-resident HMC explorer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import octofitter_jl_b200 as octo
import workloads
for cfg, n in (("C2", 37), ("C3", 33), ("C1", 1)):
    spec, x = workloads.config(cfg)
    m = octo.LogDensityModel(spec)
    ll, g = m.ln_like_and_gradient(x[:n]); v = m.ln_like(x[:n])
    pw, _ = m.pointwise_like(x[:min(n, 5)])
    print(cfg, n, m.launch_geometry(n), float(np.ravel(ll)[0]), np.isfinite(g).all(), np.isfinite(pw).all())
    m.close()
for fuse in ("1", "0"):
    os.environ["OCTO_B200_FUSE_PARAM"] = fuse
    spec, th = workloads.one_planet_with_priors(40, 30, 45, seed=2)
    m = octo.LogDensityModel(spec)
    lp, g = m.ℓπcallback_grad(th); v = m.ℓπcallback(th); l = m.ln_like_of_theta(th)
    print("logpost fuse=" + fuse, float(lp[0]), np.isfinite(g).all(), float(l[0]))
    m.close()
from helpers import load_golden
import ctypes as C
d, packed, consts = load_golden("case_obsprior")
lib = octo.load_library(); h = C.c_void_p()
assert lib.octo_create(C.byref(consts), C.byref(packed.layout), packed.blocks, packed.n_blocks, 0, C.byref(h)) == 0
x = np.asfortranarray(np.tile(np.array(d["x"]), (35, 1)))
ll = np.empty(35); g = np.empty((35, x.shape[1]), order="F")
assert lib.octo_logp_grad(h, x.ctypes.data, 35, 35, ll.ctypes.data, g.ctypes.data) == 0
print("obsprior", ll[0], d["ll"])
lib.octo_destroy(h)

for name in ("case_hgca", "case_thiele_innes"):
    d, packed, consts = load_golden(name)
    h = C.c_void_p()
    assert lib.octo_create(C.byref(consts), C.byref(packed.layout), packed.blocks, packed.n_blocks, 0, C.byref(h)) == 0
    x = np.asfortranarray(np.tile(np.array(d["x"]), (35, 1)))
    ll = np.empty(35); g = np.empty((35, x.shape[1]), order="F")
    assert lib.octo_logp_grad(h, x.ctypes.data, 35, 35, ll.ctypes.data, g.ctypes.data) == 0
    print(name, ll[0], d["ll"])
    lib.octo_destroy(h)
os.environ["OCTO_B200_FUSE_PARAM"] = "1"
spec, th = workloads.one_planet_with_priors(40, 30, 45, seed=2)
m = octo.LogDensityModel(spec)
r = octo.device_hmc(m, th, 3, step_size=1e-3, n_leapfrog=4, inv_mass=np.full(spec.D, 1e-4), seed=3)
print("hmc", r["accept_rate"], float(r["logpost_final"][0]))
pt = octo.device_parallel_tempering(m, th[:16], np.linspace(0, 1, 16), 3, n_iter=1, n_leapfrog=3, step_size=1e-3, inv_mass=np.full(spec.D, 1e-4), seed=5)
print("pt", sorted(pt["rung"]) == list(range(16)), float(pt["swap_accept"].mean()))
m.close()
# ---- round 2: sub-lane geometries (with and without epoch splits across CTAs), the launch-per-leapfrog explorer next to the
#      trajectory-resident kernel, the sharded-ladder entry point on one rank, the asynchronous halves, a linear trend
for force in ("4,1,1", "8,0,2", "32,1,1", "1,0,3"):
    os.environ["OCTO_B200_FORCE"] = force
    spec, x = workloads.config("C3")
    m = octo.LogDensityModel(spec)
    ll, g = m.ln_like_and_gradient(x[:21]); v = m.ln_like(x[:21])
    print("force", force, m.launch_geometry_full(21), float(ll[0]), np.isfinite(g).all(), float(v[0]) == float(ll[0]))
    m.close()
os.environ.pop("OCTO_B200_FORCE")
spec, th = workloads.one_planet_with_priors(40, 30, 45, seed=2)
m = octo.LogDensityModel(spec)
im = np.full(spec.D, 1e-4)
os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"] = "1"
r = octo.device_hmc(m, th, 2, step_size=1e-3, n_leapfrog=3, inv_mass=im, seed=3)
os.environ.pop("OCTO_B200_HMC_LAUNCH_PER_LEAPFROG")
print("hmc launch-per-leapfrog", r["accept_rate"])
ptl = octo.ParallelTempering(16, seed=5, beta=np.linspace(0, 1, 16), backend="local", model=m)
pd = octo.device_parallel_tempering_dist(m, ptl, th[:16], np.linspace(0, 1, 16), 3, n_iter=1, n_leapfrog=3, step_size=1e-3, inv_mass=im, seed=5)
print("pt sharded entry (1 rank)", sorted(pd["rung"]) == list(range(16)), np.array_equal(pd["swap_counts"], pt["swap_counts"]))
hs = [m.ℓπcallback_grad_begin(th) for _ in range(3)]
print("async", [float(h.wait()[0][0]) for h in hs])
ptl.close(); m.close()
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_trend as T
for kind, nt in (("star", 3), ("margin", 2), ("planet", 1)):
    system, prefix = T._system(kind, nt, n_ep=30)
    spec = octo.ModelSpec(system)
    x = T._inputs(spec, prefix, 19, seed=1)
    m = octo.LogDensityModel(spec)
    ll, g = m.ln_like_and_gradient(x)
    print("trend", kind, nt, float(ll[0]), np.isfinite(g).all())
    m.close()
# ---- page-locked buffers: inputs read by the kernel in place (zero-copy), outputs written in place; a 3-planet lean model
#      (the 3-planet instantiation) and the copy path for comparison
import workloads as W2
for name, mk in (("C2", lambda: W2.config("C2")), ("lean3", lambda: W2.k_planets_lean(3, 40, seed=7))):
    spec, x = mk()
    x = x[:40]
    res = {}
    for zmax in ("524288", "0"):
        os.environ["OCTO_B200_ZEROCOPY_MAX"] = zmax
        m = octo.LogDensityModel(spec)
        xp = m.pinned_empty(x.shape); xp[...] = x
        out = (m.pinned_empty(x.shape[0]), m.pinned_empty(x.shape))
        ll, g = m.ln_like_and_gradient(xp, out=out)
        hs = [m.ln_like_and_gradient_begin(xp, out=out) for _ in range(2)]
        for h in hs: h.wait()
        res[zmax] = (ll.copy(), g.copy())
        m.close()
    os.environ.pop("OCTO_B200_ZEROCOPY_MAX")
    print("pinned", name, float(res["0"][0][0]), np.array_equal(res["0"][0], res["524288"][0]), np.array_equal(res["0"][1], res["524288"][1]))
