"""In-kernel timeline of one evaluation inside the trajectory-resident HMC kernel (-DOCTO_TIMING build, %globaltimer).
    python profiles/tools/resident_timeline.py      (builds octofitter.jl_b200/lib/libocto_timing.so when missing)"""
import os, sys, importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
lib = os.environ.get("OCTO_TIMING_LIB") or os.path.join(ROOT, "octofitter.jl_b200", "lib", "libocto_timing.so")
if not os.path.exists(lib):
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "octofitter.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    mod.build_variant(lib, ["-DOCTO_TIMING"] + os.environ.get("OCTO_TIMING_DEFS", "").split())
if len(sys.argv) > 1 and sys.argv[1] == "build":
    sys.exit(0)
os.environ["OCTO_B200_LIB"] = lib
import numpy as np
import octofitter_jl_b200 as octo, workloads
spec_p, th_p = workloads.one_planet_with_priors(100, 100, 1024, seed=2)
model = octo.LogDensityModel(spec_p)
im = np.full(spec_p.D, 1e-4)
for k in range(3):
    octo.device_hmc(model, th_p, 2, step_size=1e-3, n_leapfrog=5, inv_mass=im, seed=k, keep_samples=False)
model.close()
