"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): C2- and C3-shaped models,
ragged batch sizes, value and gradient kernels, epoch-split path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import octofitter_jl_b200 as octo
import workloads
for cfg, n in (("C2", 37), ("C3", 33), ("C1", 1)):
    spec, x = workloads.config(cfg)
    m = octo.LogDensityModel(spec)
    ll, g = m.ln_like_and_gradient(x[:n]); v = m.ln_like(x[:n])
    print(cfg, n, m.launch_geometry(n), float(ll[0]) if n > 1 else float(ll), np.isfinite(g).all())
    m.close()
