"""Back-to-back C2 gradient steps on device-resident inputs (what bench.py's `value` times), for the library selected by
OCTO_B200_LIB: median over 11 regions of 200 steps.   python profiles/tools/c2_steps.py [config]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import octofitter_jl_b200 as octo
import workloads
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
spec, x = workloads.k_planets_lean(int(cfg[1]), 1024, seed=7) if cfg[0] == "L" else workloads.config(cfg)      # "L3": 3 planets, lean tables
n = x.shape[0]
model = octo.LogDensityModel(spec)
sets = [torch.from_numpy(np.ascontiguousarray((x * (1 + 1e-9 * k)).T)).cuda() for k in range(8)]
d_ll = torch.empty(n, dtype=torch.float64, device="cuda")
d_g = torch.empty((spec.n_in, n), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
def region(k):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(k):
        model.enqueue_device(sets[i % 8].data_ptr(), n, n, d_ll.data_ptr(), d_g.data_ptr(), st.cuda_stream)
    b.record(st); b.synchronize()
    return a.elapsed_time(b) * 1e3 / k
region(50)
t = sorted(region(200) for _ in range(11))
print(f"lib={os.environ.get('OCTO_B200_LIB','default')} {cfg} geometry={model.launch_geometry_full(n)}: median {t[5]:.2f} us/step (min {t[0]:.2f}, max {t[-1]:.2f})")
model.close()
