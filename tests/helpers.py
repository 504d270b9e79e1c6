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
