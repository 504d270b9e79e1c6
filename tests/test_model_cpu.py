"""CPU tests of the host mirror's newer classes (no GPU, no CUDA library calls): table construction follows the
reference constructors, the packed layout carries what the kernel expects, and the oracle agrees with itself across
orbit bases."""
import numpy as np
import pytest

import octofitter_jl_b200 as octo
from helpers import load_golden, rel_err

HGCA_ROW = dict(pmra_hip=10.1, pmdec_hip=-5.2, pmra_hip_error=0.9, pmdec_hip_error=0.8, pmra_pmdec_hip=0.2,
                pmra_hg=10.5, pmdec_hg=-5.0, pmra_hg_error=0.05, pmdec_hg_error=0.04, pmra_pmdec_hg=-0.1,
                pmra_gaia=11.2, pmdec_gaia=-4.6, pmra_gaia_error=0.12, pmdec_gaia_error=0.1, pmra_pmdec_gaia=0.35,
                epoch_ra_hip=1991.1, epoch_dec_hip=1991.3, epoch_ra_gaia=2016.0, epoch_dec_gaia=2016.2)


def test_hgca_rows_follow_the_reference_constructor():
    """src/likelihoods/hgca.jl:78-110: Julian-year epochs -> MJD, N_ave points over 4 yr (Hipparcos) / 1038 d (Gaia),
    RA and Dec rows interleaved, Hipparcos first."""
    h = octo.HGCAInstantaneousObs(HGCA_ROW, N_ave=3, factor=1.5)
    ep, code = h.table["epoch"], h.table["code"]
    assert len(ep) == 12 and list(code) == [0, 1] * 3 + [2, 3] * 3
    mjd = lambda y: (y - 2000.0) * 365.25 + 51544.5
    assert np.allclose(ep[0::2][:3], mjd(1991.1) + np.array([-730.5, 0.0, 730.5]))
    assert np.allclose(ep[1::2][3:], mjd(2016.2) + np.array([-519.0, 0.0, 519.0]))
    assert h.aux.shape == (15,) and h.aux[2] == pytest.approx(0.9 * 1.5) and h.aux[4] == 0.2 and h.aux[5] == 10.5
    one = octo.HGCAInstantaneousObs(HGCA_ROW)                       # N_ave = 1: the four catalogue epochs themselves
    assert list(one.table["code"]) == [0, 1, 2, 3] and one.table["epoch"][2] == pytest.approx(mjd(2016.0))
    b = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"])
    with pytest.raises(octo.OctoError, match="pmra and pmdec"):
        octo.ModelSpec(octo.System(name="s", variables=["M", "plx"], companions=[b], observations=[h]))
    spec = octo.ModelSpec(octo.System(name="s", variables=["M", "plx", "pmra", "pmdec"], companions=[b], observations=[h]))
    blk = spec.block_dicts[-1]
    assert blk["kind"] == octo.KIND_HGCA_INSTANT and blk["idx_pmra"] == 2 and blk["idx_pmdec"] == 3
    assert spec.total_epochs == 0                                    # HGCA rows are not part of the epoch list


def test_observable_prior_wrapper_and_orbit_bases():
    tab = octo.Table(epoch=[50000.0, 50100.0], ra=[1.0, 2.0], dec=[3.0, 4.0], σ_ra=[1.0, 1.0], σ_dec=[1.0, 1.0])
    astrom = octo.PlanetRelAstromObs(tab, name="cam", variables=["jitter"])
    wrapped = octo.ObsPriorAstromONeil2019(astrom)
    assert wrapped.name == "obspri_cam" and wrapped.variables == ("jitter",) and wrapped.kind == astrom.kind
    b = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"], observations=[astrom, wrapped])
    spec = octo.ModelSpec(octo.System(name="s", variables=["M", "plx"], companions=[b]))
    assert [blk.get("obs_prior", 0) for blk in spec.block_dicts] == [0, 1]
    assert "b.cam.jitter" in spec.input_names and "b.obspri_cam.jitter" in spec.input_names      # its own copy of the variables
    with pytest.raises(ValueError):
        octo.ObsPriorAstromONeil2019(octo.StarAbsoluteRVObs(octo.Table(epoch=[5e4], rv=[1.0], σ_rv=[1.0]), name="rv"))
    # Thiele-Innes: the layout switches basis and carries A, B, F, G; a, i, ω, Ω are absent
    ti = octo.Planet(name="b", basis="ThieleInnesOrbit", variables=["A", "B", "F", "G", "e", "tp"], observations=[astrom])
    spec = octo.ModelSpec(octo.System(name="s", variables=["M", "plx"], companions=[ti]))
    L = spec.packed.layout
    assert L.basis[0] == 1 and L.idx_a[0] == -1 and [L.idx_A[0], L.idx_B[0], L.idx_F[0], L.idx_G[0]] == [2, 3, 4, 5]
    with pytest.raises(octo.OctoError, match="missing orbital variables"):
        octo.ModelSpec(octo.System(name="s", variables=["M", "plx"], companions=[
            octo.Planet(name="b", basis="ThieleInnesOrbit", variables=["A", "B", "e", "tp"], observations=[astrom])]))
    # RV-only orbit: i = π/2, Ω = 0, plx injected as constants
    rvo = octo.Planet(name="b", basis="RadialVelocityOrbit", variables={"a": 1.0, "e": 0.1, "ω": 0.3, "tp": 5e4, "mass": 1.0})
    consts = dict(rvo.var_specs)
    assert consts["i"] == pytest.approx(np.pi / 2) and consts["Ω"] == 0.0 and "plx" in consts
    with pytest.raises(ValueError, match="dict"):
        octo.Planet(name="b", basis="RadialVelocityOrbit", variables=["a", "e", "ω", "tp"])


def test_oracle_thiele_innes_equals_campbell(oracle_lib):
    """The same physical orbits in the two bases give the same likelihood on the CPU oracle (the mpmath generator
    asserts the same at 60 digits)."""
    d, packed, consts = load_golden("case_thiele_innes")
    x = dict(zip(d["input_names"], d["x"]))
    plx = x["plx"]

    def to_campbell(A, B, F, G):
        u = 0.5 * (A * A + B * B + F * F + G * G); v = A * G - B * F
        alpha = np.sqrt(u + np.sqrt((u + v) * (u - v)))
        wpW = np.arctan2(B - F, A + G); wmW = np.arctan2(-B - F, A - G)
        w, W = 0.5 * (wpW + wmW), 0.5 * (wpW - wmW)
        if W < 0:
            w, W = w + np.pi, W + np.pi
        d1, d2 = abs((A + G) * np.cos(wmW)), abs((F - B) * np.sin(wmW))
        i = 2 * np.arctan(np.sqrt(abs((A - G) * np.cos(wpW)) / d1)) if d1 >= d2 else 2 * np.arctan(np.sqrt(abs((B + F) * np.sin(wpW)) / d2))
        return alpha / plx, i, w, W
    tabs = {b["name"]: b for b in d["blocks"]}
    mk = lambda b, **kw: octo.PlanetRelAstromObs(octo.Table(epoch=b["epoch"], ra=b["y1"], dec=b["y2"], σ_ra=b["s1"], σ_dec=b["s2"],
                                                            **({"cor": b["cor"]} if b.get("cor") is not None else {})), **kw)
    pb = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"], observations=[mk(tabs["relastrom"], name="relastrom")])
    pc = octo.Planet(name="c", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[mk(tabs["SPHERE"], name="SPHERE", variables=["jitter"])])
    spec = octo.ModelSpec(octo.System(name="chk", variables=["M", "plx"], companions=[pb, pc]))
    xc = {"M": x["M"], "plx": plx, "c.mass": x["c.mass"], "c.SPHERE.jitter": x["c.SPHERE.jitter"]}
    for nm in "bc":
        a, i, w, W = to_campbell(x[f"{nm}.A"], x[f"{nm}.B"], x[f"{nm}.F"], x[f"{nm}.G"])
        xc.update({f"{nm}.a": a, f"{nm}.i": i, f"{nm}.ω": w, f"{nm}.Ω": W, f"{nm}.e": x[f"{nm}.e"], f"{nm}.tp": x[f"{nm}.tp"]})
    ll_c = oracle_lib.Oracle(spec.packed, consts).logp(np.array([[xc[n] for n in spec.input_names]]))[0]
    ll_t = oracle_lib.Oracle(packed, consts).logp(np.array([d["x"]]))[0]
    assert rel_err(ll_t, d["ll"]) < 1e-12 and rel_err(ll_c, ll_t) < 1e-10
