"""The boundary as the Julia glue drives it: tests/c_abi_julia_replay.c replays julia/OctofitterB200.jl's call sequence
(hand-declared structs, dlopen, no Python) on the reference's 11-parameter test model; its numbers are compared with the
oracle and with the Python mirror's."""
import os
import subprocess

import numpy as np
import pytest

import octofitter_jl_b200 as octo
from helpers import grad_err, reference_test_system, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_replay_of_the_julia_glue_call_sequence(oracle_lib, tmp_path):
    exe = tmp_path / "replay"
    subprocess.run(["gcc", "-O1", "-o", str(exe), os.path.join(ROOT, "tests", "c_abi_julia_replay.c"), "-ldl"], check=True)
    out = subprocess.run([str(exe), octo.LIB_PATH], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    rows = [[float.fromhex(t) if "p" in t or "x" in t else float(t) for t in ln.split()] for ln in out]
    # natural-space evaluation: inputs in the glue's order (plx, M, a, e, i, ω, Ω, θ, tp) -> oracle in the mirror's order
    spec = octo.ModelSpec(reference_test_system())
    names = {"plx": "plx", "M": "M", "a": "b.a", "e": "b.e", "i": "b.i", "ω": "b.ω", "Ω": "b.Ω", "θ": "b.θ", "tp": "b.tp"}
    glue_order = ["plx", "M", "a", "e", "i", "ω", "Ω", "θ", "tp"]
    xg = np.array([50.01, 1.21, 12.1, 0.12, 0.72, 0.65, 0.29, 1.7, 41500.0])
    x = np.zeros((1, spec.n_in))
    for k, nm in enumerate(glue_order):
        x[0, spec.column(names[nm])] = xg[k]
    ll_o, g_o = oracle_lib.Oracle(spec.packed, octo.default_constants()).logp_grad(x)
    ll1, ll2, g = rows[0][0], rows[0][1], np.array(rows[0][2:])
    assert ll1 == ll2 and rel_err(ll1, ll_o[0]) < 1e-10
    g_expect = np.array([g_o[0, spec.column(names[nm])] for nm in glue_order])
    assert grad_err(g[None, :], g_expect[None, :]).max() < 1e-8 and g[7] == 0.0          # θ itself is not a kernel input
    assert len(rows[1]) == 3 and rows[1][0] == ll1
    # device posterior: same θ_t through the Python mirror (same parameter order: M, plx, a, e, i, ωx, ωy, Ωx, Ωy, θx, θy)
    model = octo.LogDensityModel(spec)
    th0 = np.array([0.0953, 3.91, -1.99, -2.0, -0.5, 0.8, 0.6, 0.95, 0.28, -0.99, -0.13])
    th = np.array([[th0[j] + 0.01 * ((c * 7 + j * 3) % 11 - 5) for j in range(11)] for c in range(16)])
    lp, gt = model.ℓπcallback_grad(th)
    assert rel_err(rows[2][0], lp[0]) < 1e-12 and abs(rows[2][1] - gt[0, 0]) <= 1e-9 * np.abs(gt[0]).max()
    lp_o, _ = oracle_lib.logpost(spec, octo.default_constants(), th, threads=2)
    assert rel_err(rows[2][0], lp_o[0]) < 1e-10
    im = np.full(11, 1e-3)
    r = octo.device_hmc(model, th, 3, step_size=0.05, n_leapfrog=4, inv_mass=im, seed=42)
    assert rel_err(rows[3][0], r["theta_final"][0, 0]) < 1e-9 and rel_err(rows[3][1], r["logpost_final"][0]) < 1e-9
    assert rows[4][0] == sum(range(16))                                                  # the rungs are a permutation
    model.close()
