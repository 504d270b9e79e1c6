"""CPU tests of the boundary: the library loads without a GPU, exports every symbol of
include/octo_b200.h, fails loudly (no CPU fallback) and the host-side mirror validates like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import octofitter_jl_b200 as octo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(octo.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("octo_build", os.path.join(ROOT, "octofitter.jl_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        mod.build()
    return octo.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "octo_b200.h")).read()
    declared = set(re.findall(r"\b(octo_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(octo.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    hdr_version = int(re.search(r"#define OCTO_ABI_VERSION (\d+)", hdr).group(1))
    assert lib.octo_abi_version() == hdr_version == 4


def test_struct_sizes_match_header():
    assert C.sizeof(octo.OctoConstants) == 7 * 8
    assert C.sizeof(octo.OctoLayout) == 4 * (2 + 14 * 4)
    assert C.sizeof(octo.OctoObsBlock) == 4 * 4 + 6 * 8 + 8 * 4 + 8 + 4 * 4 + 2 * 8


def test_default_constants(lib):
    c = octo.OctoConstants()
    lib.octo_default_constants(C.byref(c))
    d = octo.default_constants()
    for f, _ in octo.OctoConstants._fields_:
        assert getattr(c, f) == getattr(d, f)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback_without_gpu(lib):
    t = octo.Table(epoch=[50000., 50100.], ra=[1., 2.], dec=[3., 4.], σ_ra=[1., 1.], σ_dec=[1., 1.])
    b = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"],
                    observations=[octo.PlanetRelAstromObs(t, name="x")])
    s = octo.System(name="s", variables=["M", "plx"], companions=[b])
    with pytest.raises(octo.OctoError, match="no CUDA device|CUDA"):
        octo.LogDensityModel(s)


def test_create_rejects_bad_layout(lib):
    packed = octo.pack({"n_in": 3, "planets": [dict(plx=0, a=1, e=2, i=3, w=0, W=0, tp=0, M=0)]}, [])
    h = C.c_void_p()
    c = octo.default_constants()
    rc = lib.octo_create(C.byref(c), C.byref(packed.layout), packed.blocks, 0, 0, C.byref(h))
    assert rc == 1 and b"column" in lib.octo_last_error()


def test_model_mirror_validation_and_order():
    t = octo.Table(epoch=[50100., 50000.], ra=[1., 2.], dec=[3., 4.], sigma_ra=[1., 1.], sigma_dec=[1., 1.])
    a = octo.PlanetRelAstromObs(t, name="GPI astrom", variables=["jitter"])
    assert list(a.table["epoch"]) == [50000., 50100.] and list(a.table["ra"]) == [2., 1.]   # sorted by epoch
    with pytest.raises(ValueError, match="Expected columns"):
        octo.PlanetRelAstromObs(octo.Table(epoch=[1.], ra=[1.]), name="bad")
    with pytest.raises(ValueError, match="same length"):
        octo.PlanetRelAstromObs({"epoch": [50000., 50001.], "ra": [1.], "dec": [1.], "σ_ra": [1.], "σ_dec": [1.]}, name="bad")
    with pytest.raises(ValueError, match="Correlation"):
        octo.PlanetRelAstromObs(octo.Table(epoch=[50000.], ra=[1.], dec=[1.], σ_ra=[1.], σ_dec=[1.], cor=[0.999999]), name="bad")
    with pytest.raises(ValueError, match="gaussian_process"):
        octo.StarAbsoluteRVObs(octo.Table(epoch=[50000.], rv=[1.], σ_rv=[1.]), name="gp", gaussian_process=object())
    with pytest.raises(ValueError, match="jitter"):
        octo.MarginalizedStarAbsoluteRVObs(octo.Table(epoch=[50000.], rv=[1.], σ_rv=[1.]), name="m", variables=[])
    b = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[a])
    rv = octo.StarAbsoluteRVObs(octo.Table(epoch=[50000.], rv=[1.], σ_rv=[1.]), name="harps")
    s = octo.System(name="s", variables=["M", "plx"], companions=[b], observations=[rv])
    spec = octo.ModelSpec(s)
    # reference parameter order: system, system-obs, planet, planet-obs (src/variables.jl:691-730)
    assert spec.input_names == ("M", "plx", "harps.offset", "harps.jitter", "b.a", "b.e", "b.i", "b.ω", "b.Ω",
                                "b.tp", "b.mass", "b.GPI_astrom.jitter")
    # reference summation order: planet observations, then system observations (system.jl:223-236)
    assert [bd["kind"] for bd in spec.block_dicts] == [0, 2]
    with pytest.raises(octo.OctoError, match="missing orbital variables"):
        octo.ModelSpec(octo.System(name="s", variables=["M"], companions=[b]))


def test_workload_generators_are_seeded():
    import workloads
    s1, x1 = workloads.config("C2")
    s2, x2 = workloads.config("C2")
    assert np.array_equal(x1, x2) and s1.total_epochs == 200 and x1.shape == (1024, s1.n_in)
    s3, x3 = workloads.config("C3")
    assert s3.total_epochs == 500 and x3.shape[0] == 256
