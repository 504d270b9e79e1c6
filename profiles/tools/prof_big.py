import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import octofitter_jl_b200 as octo, workloads
for E, kind in ((20000, "astrom"), (20000, "rv")):
    spec, x = workloads.one_planet(E if kind == "astrom" else 0, E if kind == "rv" else 0, 4096, seed=5)
    model = octo.LogDensityModel(spec)
    n, n_in = x.shape
    d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda(); d_ll = torch.empty(n, dtype=torch.float64, device="cuda"); d_g = torch.empty((n_in, n), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(3): model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), d_g.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(10): model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), d_g.data_ptr(), st.cuda_stream)
    b.record(st); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(kind, E, model.launch_geometry(n), f"{ms:.4f} ms", f"{n*E/ms/1e-3:.3e} evals/s")
