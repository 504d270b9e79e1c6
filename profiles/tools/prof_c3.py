"""2-planet (C3-shaped) launches at a compute-bound size: 4096 chains, astrometry on both planets + star RV."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import octofitter_jl_b200 as octo
import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec, x = workloads.two_planet(n, seed=3, n_b=200 * scale, n_c=150 * scale, n_rv=150 * scale)
model = octo.LogDensityModel(spec)
d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
d_ll = torch.empty(n, dtype=torch.float64, device="cuda")
d_g = torch.empty((spec.n_in, n), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
ev = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), d_g.data_ptr(), st.cuda_stream); b.record(st)
    ev.append((a, b))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
E = spec.total_epochs
print(f"two-planet: chains={n} epochs={E} geometry={model.launch_geometry(n)} ms={ms} evals/s={n * E / (min(ms) * 1e-3):.3e}")
