"""Where the resident explorer and the launch-per-leapfrog explorer differ (debug aid for tests/test_gpu_resident.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import octofitter_jl_b200 as octo, workloads
for force in ("1,1,1", "4,1,1"):
    os.environ["OCTO_B200_FORCE"] = force
    for n in (200, 64):
        spec_p, th_p = workloads.one_planet_with_priors(100, 100, n, seed=2)
        model = octo.LogDensityModel(spec_p)
        kw = dict(step_size=1e-3, n_leapfrog=6, inv_mass=np.full(spec_p.D, 1e-4), seed=5)
        res = octo.device_hmc(model, th_p, 4, **kw)
        os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"] = "1"
        ref = octo.device_hmc(model, th_p, 4, **kw)
        del os.environ["OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"]
        for k in ("theta", "logpost", "theta_final", "logpost_final", "accept"):
            a, b = np.asarray(res[k], dtype=float), np.asarray(ref[k], dtype=float)
            d = np.abs(a - b)
            bad = np.argwhere(d > 0)
            print(force, n, k, "max abs diff", d.max(), "rel", (d / (np.abs(b) + 1e-300)).max(), "n_diff", len(bad), "first", bad[:3].tolist())
        model.close()
