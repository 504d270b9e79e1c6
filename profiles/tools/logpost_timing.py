"""Phase timeline of the fused log-posterior launch (needs a library built with -DOCTO_TIMING)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import octofitter_jl_b200 as octo
import workloads
spec, th = workloads.one_planet_with_priors(100, 100, 1024, seed=2)
model = octo.LogDensityModel(spec)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(4):
    if it < 2:
        flush.zero_()
    print('flushed' if it < 2 else 'warm L2')
    model.ℓπcallback_grad(th)
    torch.cuda.synchronize()
    print("----")
