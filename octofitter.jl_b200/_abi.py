"""ctypes view of include/octo_b200.h and the loader for libocto_b200.so.

The structs mirror the C header field for field.  `load_library()` fails loudly when the
CUDA extension has not been built: there is no CPU fallback in the product path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

OCTO_MAX_PLANETS = 4
KIND_ASTROM_RADEC, KIND_ASTROM_PASEP, KIND_RV_STAR_ABS, KIND_RV_STAR_MARGIN, KIND_RV_PLANET_REL, KIND_HGCA_INSTANT = range(6)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libocto_b200.so")


class OctoConstants(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("kepler_year_days", "year2day", "rad2as", "pc2au", "au2m", "sec2year", "mjup2msol")]


def default_constants() -> OctoConstants:
    """Recollected PlanetOrbits.jl 0.11 constants (see include/octo_b200.h)."""
    return OctoConstants(365.2568983840419, 365.25, 206265.0, 206265.0, 1.495978707e11,
                         1.0 / 31557600.0, 0.0009545942339693249)


_pd = C.POINTER(C.c_double)


class OctoObsBlock(C.Structure):
    _fields_ = [("kind", C.c_int32), ("planet", C.c_int32), ("n_epochs", C.c_int32), ("has_cor", C.c_int32),
                ("epoch", _pd), ("y1", _pd), ("y2", _pd), ("s1", _pd), ("s2", _pd), ("cor", _pd),
                ("idx_jitter", C.c_int32), ("idx_platescale", C.c_int32),
                ("idx_northangle", C.c_int32), ("idx_offset", C.c_int32), ("obs_prior", C.c_int32), ("idx_pmra", C.c_int32),
                ("idx_pmdec", C.c_int32), ("reserved", C.c_int32), ("aux", _pd),
                ("n_trend", C.c_int32), ("idx_trend", C.c_int32 * 3), ("trend_basis", _pd), ("trend_const", _pd)]


_i4 = C.c_int32 * OCTO_MAX_PLANETS


class OctoLayout(C.Structure):
    _fields_ = [("n_planets", C.c_int32), ("n_in", C.c_int32),
                ("idx_plx", _i4), ("idx_a", _i4), ("idx_e", _i4), ("idx_i", _i4), ("idx_w", _i4),
                ("idx_W", _i4), ("idx_tp", _i4), ("idx_M", _i4), ("idx_mass", _i4),
                ("basis", _i4), ("idx_A", _i4), ("idx_B", _i4), ("idx_F", _i4), ("idx_G", _i4)]


PRIOR_NORMAL, PRIOR_UNIFORM, PRIOR_LOGUNIFORM, PRIOR_SINE, PRIOR_TRUNCNORMAL = range(5)
IN_PARAM, IN_CONST, IN_CIRC, IN_TPERI, IN_TPERI_TI = range(5)


class OctoPrior(C.Structure):
    _fields_ = [("family", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 4)]


class OctoInputDef(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32 * 8), ("value", C.c_double)]


def _dptr(a):
    return a.ctypes.data_as(_pd) if a is not None else _pd()


class PackedModel:
    """Keeps the numpy arrays alive next to the ctypes structs that point into them."""

    def __init__(self, layout: OctoLayout, blocks, keep):
        self.layout = layout
        self.n_blocks = len(blocks)
        self.blocks = (OctoObsBlock * max(1, len(blocks)))(*blocks)
        self._keep = keep


def pack(layout_dict: dict, block_dicts: list) -> PackedModel:
    """Build the C structs from plain dictionaries.

    layout_dict: {"n_in": int, "planets": [{"plx":k,"a":k,"e":k,"i":k,"w":k,"W":k,"tp":k,"M":k,"mass":k|-1}, ...]}
    block_dicts: [{"kind":int,"planet":int,"epoch":[..],"y1":[..],"y2":..,"s1":..,"s2":..,"cor":..|None,
                   "idx_jitter":k|-1,"idx_platescale":..,"idx_northangle":..,"idx_offset":..}, ...]
    """
    L = OctoLayout()
    planets = layout_dict["planets"]
    if not 1 <= len(planets) <= OCTO_MAX_PLANETS:
        raise ValueError(f"1..{OCTO_MAX_PLANETS} planets supported, got {len(planets)}")
    L.n_planets = len(planets)
    L.n_in = int(layout_dict["n_in"])
    for f in ("plx", "a", "e", "i", "w", "W", "tp", "M", "mass", "A", "B", "F", "G"):
        arr = getattr(L, "idx_" + f)
        for p in range(OCTO_MAX_PLANETS):
            arr[p] = int(planets[p].get(f, -1)) if p < len(planets) else -1
    for p in range(OCTO_MAX_PLANETS):
        L.basis[p] = int(planets[p].get("basis", 0)) if p < len(planets) else 0
    keep, blocks = [], []
    for bd in block_dicts:
        cols = {}
        for k in ("epoch", "y1", "y2", "s1", "s2", "cor", "aux", "trend_basis", "trend_const"):
            v = bd.get(k)
            cols[k] = None if v is None else np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        keep.append(cols)
        n = len(cols["epoch"])
        idx_trend = [int(v) for v in bd.get("idx_trend", [])]
        if cols["trend_basis"] is not None and cols["trend_basis"].shape != (len(idx_trend), n):
            raise ValueError("trend_basis must be [n_trend x n_epochs]")
        for k, v in cols.items():
            if k not in ("aux", "trend_basis") and v is not None and len(v) != n:
                raise ValueError("The columns in the input data do not all have the same length")
        blocks.append(OctoObsBlock(
            int(bd["kind"]), int(bd.get("planet", -1)), n, 0 if cols["cor"] is None else 1,
            _dptr(cols["epoch"]), _dptr(cols["y1"]), _dptr(cols["y2"]), _dptr(cols["s1"]), _dptr(cols["s2"]),
            _dptr(cols["cor"]),
            int(bd.get("idx_jitter", -1)), int(bd.get("idx_platescale", -1)),
            int(bd.get("idx_northangle", -1)), int(bd.get("idx_offset", -1)), int(bd.get("obs_prior", 0)),
            int(bd.get("idx_pmra", -1)), int(bd.get("idx_pmdec", -1)), 0, _dptr(cols["aux"]),
            len(idx_trend), (C.c_int32 * 3)(*(idx_trend + [-1] * (3 - len(idx_trend)))), _dptr(cols["trend_basis"]),
            _dptr(cols["trend_const"])))
    return PackedModel(L, blocks, keep)


_lib = None


def load_library(path: str | None = None):
    """dlopen libocto_b200.so and declare the prototypes of every exported symbol."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("OCTO_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"libocto_b200.so not found at {p}: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  There is no CPU fallback for this path.")
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.octo_default_constants.argtypes = [C.POINTER(OctoConstants)]
    lib.octo_default_constants.restype = None
    lib.octo_abi_version.restype = C.c_int
    lib.octo_create.argtypes = [C.POINTER(OctoConstants), C.POINTER(OctoLayout), C.POINTER(OctoObsBlock),
                                i32, i32, C.POINTER(vp)]
    lib.octo_destroy.argtypes = [vp]
    lib.octo_destroy.restype = None
    lib.octo_logp.argtypes = [vp, vp, i64, i64, vp]
    lib.octo_logp_grad.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.octo_release_stream.argtypes = [vp, vp]
    lib.octo_logp_grad_begin.argtypes = [vp, vp, i64, i64, vp, vp, C.POINTER(vp)]
    lib.octo_logpost_grad_begin.argtypes = [vp, vp, i64, i64, vp, vp, C.POINTER(vp)]
    lib.octo_ready.argtypes = [vp]
    lib.octo_wait.argtypes = [vp]
    lib.octo_logp_grad_device.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    lib.octo_set_parameterization.argtypes = [vp, C.POINTER(OctoPrior), i32, C.POINTER(OctoInputDef)]
    lib.octo_logpost_grad.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.octo_loglike_theta.argtypes = [vp, vp, i64, i64, vp]
    lib.octo_logp_pointwise.argtypes = [vp, vp, i64, i64, vp, i64]
    lib.octo_hmc_run.argtypes = [vp, vp, i64, i64, i32, i32, C.c_double, vp, C.c_uint64, vp, vp, vp, vp, vp]
    lib.octo_pt_hmc_run.argtypes = [vp, vp, i64, i64, vp, i32, i32, i32, C.c_double, vp, C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.octo_hmc_random.argtypes = [C.c_uint64, i64, i64, i32, vp, vp]
    lib.octo_hmc_random.restype = None
    lib.octo_logpost_workspace.argtypes = [vp, i64]
    lib.octo_logpost_workspace.restype = i64
    lib.octo_logpost_grad_device.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp]
    lib.octo_invlink.argtypes = [vp, vp, i64, i64, vp]
    lib.octo_alloc_pinned.argtypes = [C.c_size_t]
    lib.octo_alloc_pinned.restype = vp
    lib.octo_free_pinned.argtypes = [vp]
    lib.octo_free_pinned.restype = None
    lib.octo_selftest_kepler.argtypes = [i32, vp, vp, i64, vp, vp]
    lib.octo_n_in.argtypes = [vp]
    lib.octo_n_in.restype = i32
    lib.octo_n_planets.argtypes = [vp]
    lib.octo_n_planets.restype = i32
    lib.octo_total_epochs.argtypes = [vp]
    lib.octo_total_epochs.restype = i64
    lib.octo_device.argtypes = [vp]
    lib.octo_device.restype = i32
    lib.octo_kernel_launches.argtypes = [vp]
    lib.octo_kernel_launches.restype = i64
    lib.octo_launch_geometry.argtypes = [vp, i64, C.POINTER(i32 * 6)]
    lib.octo_pt_unique_id.argtypes = [vp]
    lib.octo_pt_init.argtypes = [vp, vp, i32, i32, i32, C.c_uint64]
    lib.octo_pt_swap_round.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.octo_pt_swap_round_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.octo_pt_hmc_run_dist.argtypes = [vp, vp, i64, i64, vp, i32, i32, i32, C.c_double, vp, C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.octo_pt_decide.argtypes = [vp, vp, vp, i32, i64, C.c_uint64, vp]
    lib.octo_pt_finalize.argtypes = [vp]
    lib.octo_pt_finalize.restype = None
    lib.octo_last_error.restype = C.c_char_p
    if path is None:
        _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "octo_default_constants", "octo_abi_version", "octo_create", "octo_destroy", "octo_logp", "octo_logp_grad",
    "octo_logp_grad_device", "octo_set_parameterization", "octo_logpost_grad", "octo_logpost_workspace",
    "octo_logpost_grad_device", "octo_loglike_theta", "octo_logp_pointwise", "octo_hmc_run", "octo_pt_hmc_run", "octo_hmc_random", "octo_invlink", "octo_alloc_pinned", "octo_free_pinned", "octo_selftest_kepler", "octo_n_in", "octo_n_planets", "octo_total_epochs", "octo_device",
    "octo_kernel_launches", "octo_launch_geometry", "octo_pt_unique_id", "octo_pt_init", "octo_pt_swap_round",
    "octo_pt_decide", "octo_pt_finalize", "octo_pt_swap_round_device", "octo_pt_hmc_run_dist", "octo_logp_grad_begin", "octo_logpost_grad_begin", "octo_ready", "octo_wait", "octo_release_stream", "octo_last_error")
