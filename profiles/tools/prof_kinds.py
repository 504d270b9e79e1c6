"""Launch one gradient kernel per observation kind (astrometry-only, RV-only) on a compute-bound size, for
ncu instruction counting (FP64 flop per pair) and --set full captures.  Usage: python prof_kinds.py [chains] [epochs]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import octofitter_jl_b200 as octo  # noqa: E402
import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for name, (na, nr) in (("astrom", (E, 0)), ("rv", (0, E))):
    spec, x = workloads.one_planet(na, nr, n, seed=11)
    model = octo.LogDensityModel(spec)
    d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    d_ll = torch.empty(n, dtype=torch.float64, device="cuda")
    d_g = torch.empty((spec.n_in, n), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), d_g.data_ptr(), st.cuda_stream)
        b.record(st)
        ev.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    print(f"{name}: pairs={n * E} geometry={model.launch_geometry(n)} ms={ms} evals/s={n * E / (min(ms) * 1e-3):.3e}")
    model.close()
