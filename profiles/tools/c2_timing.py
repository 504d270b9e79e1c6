"""Phase timing of the C2 launch (needs a library built with -DOCTO_TIMING; prints SM-clock deltas per phase)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import octofitter_jl_b200 as octo
import workloads
spec, x = workloads.config("C2")
model = octo.LogDensityModel(spec)
for _ in range(2):
    model.ln_like_and_gradient(x)
    torch.cuda.synchronize()
    print("----")
