"""Shared helpers for the test-suite: golden loading, seeded synthetic systems."""
import glob
import json
import os

import numpy as np

import octofitter_jl_b200 as octo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "case_*.json")))


def load_golden(name):
    d = json.load(open(os.path.join(GOLDEN, name + ".json")))
    packed = octo.pack(d["layout"], d["blocks"])
    consts = octo.OctoConstants(*[d["constants"][k] for k in
                                  ("kepler_year_days", "year2day", "rad2as", "pc2au", "au2m", "sec2year", "mjup2msol")])
    return d, packed, consts


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else np.finfo(float).tiny)


def grad_err(g, g_ref):
    """Gradient error relative to the largest component of the reference row (∇ tolerance 1e-8)."""
    g, g_ref = np.atleast_2d(g), np.atleast_2d(g_ref)
    scale = np.maximum(np.abs(g_ref).max(axis=1, keepdims=True), 1e-300)
    return np.abs(g - g_ref) / scale


class GoldenSpec:
    """A parameterised model rebuilt from a tests/golden/post_*.json file (quacks like octo.ModelSpec)."""

    def __init__(self, d):
        import ctypes as C
        self.packed = octo.pack(d["layout"], d["blocks"])
        self.D = len(d["theta_t"])
        self.n_in = d["layout"]["n_in"]
        self.priors = (octo.OctoPrior * self.D)(*[octo.OctoPrior(p[0], 0, (C.c_double * 4)(*p[1:])) for p in d["priors"]])
        self.defs = (octo.OctoInputDef * len(d["defs"]))(*[octo.OctoInputDef(q[0], (C.c_int32 * 8)(*(list(q[1]) + [0] * (8 - len(q[1])))), q[2]) for q in d["defs"]])
        self.theta_names = tuple(d["theta_names"])


def post_cases():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "post_*.json")))


def load_post(name):
    d = json.load(open(os.path.join(GOLDEN, name + ".json")))
    consts = octo.OctoConstants(*[d["constants"][k] for k in
                                  ("kepler_year_days", "year2day", "rad2as", "pc2au", "au2m", "sec2year", "mjup2msol")])
    return d, GoldenSpec(d), consts


def reference_test_system():
    """The 11-parameter model of the reference's own tests (test/integration/sampling.jl:29-71): 8-epoch astrometry,
    Uniform/Sine/UniformCircular priors, tp from θ_at_epoch_to_tperi, truncated-normal M and plx."""
    d = json.load(open(os.path.join(GOLDEN, "fixture8_pin.json")))
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=d["epoch"], ra=d["ra"], dec=d["dec"], σ_ra=[10.] * 8,
                                                σ_dec=[10.] * 8, cor=[0.] * 8), name="relastrom")
    b = octo.Planet(name="b", observations=[astrom], variables={
        "a": octo.Uniform(0, 100), "e": octo.Uniform(0.0, 0.99), "i": octo.Sine(), "ω": octo.UniformCircular(),
        "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 50000)})
    return octo.System(name="TestSys", companions=[b], variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1)})
