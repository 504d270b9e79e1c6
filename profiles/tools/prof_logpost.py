"""Launch the K0 forward / K1 / K0 backward pipeline a few times (for ncu launch timing)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import octofitter_jl_b200 as octo
import workloads
spec, th = workloads.one_planet_with_priors(100, 100, 1024, seed=2)
model = octo.LogDensityModel(spec)
for _ in range(6):
    model.ℓπcallback_grad(th)
