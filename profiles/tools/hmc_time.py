"""Per-leapfrog time of the device-resident explorers on the C2 tables with priors (what bench.py reports as hmc_device /
pt_device), for the library selected by OCTO_B200_LIB.   python profiles/tools/hmc_time.py [n_chains] [n_iter] [n_leapfrog]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import octofitter_jl_b200 as octo, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
leap = int(sys.argv[3]) if len(sys.argv) > 3 else 10
spec_p, th_p = workloads.one_planet_with_priors(100, 100, n, seed=2)
model = octo.LogDensityModel(spec_p)
im = np.full(spec_p.D, 1e-4)
octo.device_hmc(model, th_p, 2, step_size=1e-3, n_leapfrog=leap, inv_mass=im, seed=1, keep_samples=False)
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    r = octo.device_hmc(model, th_p, iters, step_size=1e-3, n_leapfrog=leap, inv_mass=im, seed=2, keep_samples=False)
    best = min(best, time.perf_counter() - t0)
# a short run of the same call measures its fixed cost (allocation, copies, launch, sync)
t0 = time.perf_counter(); octo.device_hmc(model, th_p, 1, step_size=1e-3, n_leapfrog=1, inv_mass=im, seed=2, keep_samples=False); t_fix = time.perf_counter() - t0
print(f"lib={os.environ.get('OCTO_B200_LIB','default')} force={os.environ.get('OCTO_B200_FORCE')} n={n}: {best/(iters*leap)*1e6:.2f} us/leapfrog wall "
      f"({(best-t_fix)/(iters*leap-1)*1e6:.2f} without the call's fixed {t_fix*1e6:.0f} us), accept {r['accept_rate']:.2f}")
pt_n, rounds = 64, 50
lad = np.linspace(0.0, 1.0, pt_n) ** 3
octo.device_parallel_tempering(model, th_p[:pt_n], lad, 2, n_iter=1, n_leapfrog=8, step_size=1e-3, inv_mass=im, seed=3)
t0 = time.perf_counter()
octo.device_parallel_tempering(model, th_p[:pt_n], lad, rounds, n_iter=1, n_leapfrog=8, step_size=1e-3, inv_mass=im, seed=4)
print(f"  pt 64 replicas: {(time.perf_counter()-t0)/rounds*1e6:.1f} us/round (1 transition x 8 leapfrogs + swap)")
model.close()
