"""A few device-resident launches of the bench workload (C2: 1024 chains x 200 epochs) and of the same tables with priors
(fused log posterior), for ncu captures.  Usage: python prof_c2.py [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import octofitter_jl_b200 as octo  # noqa: E402
import workloads  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
spec, x = workloads.config("C2")
model = octo.LogDensityModel(spec)
n, n_in = x.shape
d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
d_ll = torch.empty(n, dtype=torch.float64, device="cuda"); d_g = torch.empty((n_in, n), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
for _ in range(reps):
    model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), d_g.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
print("C2 geometry", model.launch_geometry_full(n))
model.close()
