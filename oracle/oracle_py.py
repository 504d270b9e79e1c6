"""ctypes wrapper around oracle/_build/libocto_oracle.so.  TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (octofitter.jl_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libocto_oracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("octo_oracle.cpp", "octo_oracle.hpp", "octo_oracle_param.hpp")] + \
           [os.path.join(HERE, "..", "include", "octo_b200.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.octo_oracle_kepler.restype = C.c_double
        _lib.octo_oracle_kepler.argtypes = [C.c_double, C.c_double]
        _lib.octo_oracle_rem2pi.restype = C.c_double
        _lib.octo_oracle_rem2pi.argtypes = [C.c_double]
        _lib.octo_oracle_last_error.restype = C.c_char_p
        _lib.octo_oracle_orbit_radecrv.argtypes = [C.c_void_p] + [C.c_double] * 8 + \
            [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.octo_oracle_logp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        _lib.octo_oracle_logp_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                               C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    return _lib


def max_threads():
    return int(lib().octo_oracle_max_threads())


def kepler(MA, e):
    return lib().octo_oracle_kepler(float(MA), float(e))


def orbit_radecrv(consts, a, e, i, w, W, tp, M, plx, t):
    t = np.ascontiguousarray(t, dtype=np.float64)
    ra, dec, rv = np.empty_like(t), np.empty_like(t), np.empty_like(t)
    lib().octo_oracle_orbit_radecrv(C.addressof(consts), a, e, i, w, W, tp, M, plx, t.ctypes.data, len(t),
                                    ra.ctypes.data, dec.ctypes.data, rv.ctypes.data)
    return ra, dec, rv


def logpost(spec, consts, theta_t, grad=True, threads=1):
    """ℓπ(θ_t) (and ∇) of a parameterised ModelSpec on the CPU oracle."""
    L = lib()
    L.octo_oracle_logpost.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    th = np.asfortranarray(np.atleast_2d(np.asarray(theta_t, dtype=np.float64)))
    n, D = th.shape
    assert D == spec.D
    lp = np.empty(n)
    g = np.empty((n, D), order="F") if grad else None
    rc = L.octo_oracle_logpost(C.addressof(consts), C.addressof(spec.packed.layout), spec.packed.blocks, spec.packed.n_blocks,
                               spec.priors, D, spec.defs, th.ctypes.data, n, n, lp.ctypes.data,
                               g.ctypes.data if grad else None, threads)
    if rc:
        raise RuntimeError(L.octo_oracle_last_error().decode())
    return (lp, g) if grad else lp


def loglike_theta(spec, consts, theta_t, threads=1):
    """ln_like(system, arr2nt(invlink(θ_t))) of a parameterised ModelSpec on the CPU oracle (rejection sampler)."""
    L = lib()
    L.octo_oracle_loglike_theta.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
    th = np.asfortranarray(np.atleast_2d(np.asarray(theta_t, dtype=np.float64)))
    n, D = th.shape
    assert D == spec.D
    ll = np.empty(n)
    rc = L.octo_oracle_loglike_theta(C.addressof(consts), C.addressof(spec.packed.layout), spec.packed.blocks,
                                     spec.packed.n_blocks, spec.priors, D, spec.defs, th.ctypes.data, n, n, ll.ctypes.data, threads)
    if rc:
        raise RuntimeError(L.octo_oracle_last_error().decode())
    return ll


def invlink(spec, theta_t):
    L = lib()
    L.octo_oracle_invlink.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    th = np.asfortranarray(np.atleast_2d(np.asarray(theta_t, dtype=np.float64)))
    out = np.empty_like(th, order="F")
    L.octo_oracle_invlink(spec.priors, spec.D, th.ctypes.data, th.shape[0], th.shape[0], out.ctypes.data)
    return out


class Oracle:
    """Evaluate a packed model (octofitter.jl_b200._abi.PackedModel) on the CPU oracle."""

    def __init__(self, packed, consts):
        self.packed, self.consts = packed, consts
        self.n_in = packed.layout.n_in

    def _in(self, theta):
        th = np.asarray(theta, dtype=np.float64)
        if th.ndim == 1:
            th = th[None, :]
        assert th.shape[1] == self.n_in
        return np.asfortranarray(th)

    def logp(self, theta, threads=1):
        x = self._in(theta)
        n = x.shape[0]
        ll = np.empty(n)
        rc = lib().octo_oracle_logp(C.addressof(self.consts), C.addressof(self.packed.layout), self.packed.blocks,
                                    self.packed.n_blocks, x.ctypes.data, n, n, ll.ctypes.data, threads)
        if rc:
            raise RuntimeError(lib().octo_oracle_last_error().decode())
        return ll

    def logp_grad(self, theta, threads=1):
        x = self._in(theta)
        n = x.shape[0]
        ll = np.empty(n)
        g = np.empty((n, self.n_in), order="F")
        rc = lib().octo_oracle_logp_grad(C.addressof(self.consts), C.addressof(self.packed.layout),
                                         self.packed.blocks, self.packed.n_blocks, x.ctypes.data, n, n,
                                         ll.ctypes.data, g.ctypes.data, threads)
        if rc:
            raise RuntimeError(lib().octo_oracle_last_error().decode())
        return ll, g
