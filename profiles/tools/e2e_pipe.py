"""Host-side cost of the end-to-end C2 step (pinned host buffers): the Python mirror vs raw ctypes calls, pipelined
(3 in flight) and blocking.   OCTO_B200_ZEROCOPY_IN=0|1|2 python profiles/tools/e2e_pipe.py"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import octofitter_jl_b200 as octo
import workloads
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
if "x" in name:
    nn, E = name.split("x"); spec, x = workloads.one_planet(int(E), 0, int(nn), seed=5)
else:
    spec, x = workloads.config(name)
model = octo.LogDensityModel(spec)
n, n_in = x.shape
DEPTH = 3
slots = [(model.pinned_empty(x.shape), (model.pinned_empty(n), model.pinned_empty((n, n_in)))) for _ in range(DEPTH)]
for xi, _ in slots: xi[...] = x
lib, h = model._lib, model._h
def api(K):
    pend = []
    for k in range(K):
        xi, oi = slots[k % DEPTH]
        if len(pend) == DEPTH: pend.pop(0).wait()
        pend.append(model.ln_like_and_gradient_begin(xi, out=oi))
    for p in pend: p.wait()
ptrs = [(xi.ctypes.data, o[0].ctypes.data, o[1].ctypes.data) for xi, o in slots]
def raw(K):
    pend = []
    for k in range(K):
        a = ptrs[k % DEPTH]
        if len(pend) == DEPTH: lib.octo_wait(pend.pop(0))
        t = C.c_void_p(); lib.octo_logp_grad_begin(h, a[0], n, n, a[1], a[2], C.byref(t)); pend.append(t)
    for t in pend: lib.octo_wait(t)
def sync(K):
    a = ptrs[0]
    for k in range(K): lib.octo_logp_grad(h, a[0], n, n, a[1], a[2])
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
res = {}
for name, f in (("api pipelined", api), ("raw ctypes pipelined", raw), ("raw ctypes blocking", sync)):
    f(min(200, K))
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); f(K); best = min(best, (time.perf_counter() - t0) / K)
    res[name] = best * 1e6
ref = slots[0][1][0].copy()
print(name + " zerocopy_in=%s: " % os.environ.get("OCTO_B200_ZEROCOPY_IN", "0") + ", ".join(f"{k} {v:.2f} us/step" for k, v in res.items()), " ll[0]=%.6f" % ref[0])
model.close()
