"""Break down the end-to-end cost of one C2 evaluation through the C ABI (pinned host buffers)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import octofitter_jl_b200 as octo
import workloads
spec, x = workloads.config("C2")
model = octo.LogDensityModel(spec)
n, n_in = x.shape
xp = model.pinned_empty(x.shape); xp[...] = x
ll = model.pinned_empty(n); g = model.pinned_empty((n, n_in))
lib, h = model._lib, model._h
f = lib.octo_logp_grad
args = (h, xp.ctypes.data, n, n, ll.ctypes.data, g.ctypes.data)
for _ in range(50): f(*args)
K = 2000
t0 = time.perf_counter()
for _ in range(K): f(*args)
t_raw = (time.perf_counter() - t0) / K
t0 = time.perf_counter()
for _ in range(K): model.ln_like_and_gradient(xp, out=(ll, g))
t_py = (time.perf_counter() - t0) / K
fv = lib.octo_logp
t0 = time.perf_counter()
for _ in range(K): fv(h, xp.ctypes.data, n, n, ll.ctypes.data)
t_val = (time.perf_counter() - t0) / K
# tiny batch: latency floor of the ABI path
x1 = model.pinned_empty((1, n_in)); x1[...] = x[:1]; l1 = model.pinned_empty(1); g1 = model.pinned_empty((1, n_in))
for _ in range(50): f(h, x1.ctypes.data, 1, 1, l1.ctypes.data, g1.ctypes.data)
t0 = time.perf_counter()
for _ in range(K): f(h, x1.ctypes.data, 1, 1, l1.ctypes.data, g1.ctypes.data)
t_one = (time.perf_counter() - t0) / K
print(f"raw ctypes octo_logp_grad (1024 chains): {t_raw*1e6:.1f} us; via LogDensityModel: {t_py*1e6:.1f} us; "
      f"value-only octo_logp: {t_val*1e6:.1f} us; 1 chain x 200 epochs grad: {t_one*1e6:.1f} us")
