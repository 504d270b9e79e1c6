"""Generate the committed golden vectors (tests/golden/*.json).

Run from the repo root:  python tests/golden/make_golden.py
Needs mpmath + scipy; does NOT need a GPU or the reference.  Two kinds of vectors:

1. `fixture8_pin.json` — the reference's own 8-epoch relative-astrometry fixture
   (/root/reference/test/integration-tests.jl:8-15; same numbers in docs/src/thiele-innes.md:22)
   together with the orbit that generated it.  The orbit was identified by least squares:
   a=12 AU, e=0.11, i=41°, ω=38°, Ω=16°, M=1.2, plx=50 reproduce all 16 numbers to <1e-12 mas
   when the Kepler-year constant is 365.2422 d (the fit returns 365.2422000000 with the
   constant free), i.e. the table was produced by an earlier PlanetOrbits/DirectOrbits release
   whose year constant was the tropical year.  Only tp is a fitted value.  This pins the
   oracle's orbit geometry and Kepler solve against numbers the reference
   ships, with the year constant injected through OctoConstants.
2. `case_*.json` — full log-likelihood + gradient vectors computed by the independent
   60-digit mpmath restatement (oracle/mp_reference.py) for models assembled from the fixture
   orbits the reference's tests/docs use.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpmath as mp  # noqa: E402
import octofitter_jl_b200 as octo  # noqa: E402
from oracle import mp_reference as mpr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CONSTS = {"kepler_year_days": 365.2568983840419, "year2day": 365.25, "rad2as": 206265.0, "pc2au": 206265.0,
          "au2m": 1.495978707e11, "sec2year": 1.0 / 31557600.0, "mjup2msol": 0.0009545942339693249}

FIX_EPOCH = [50000, 50120, 50240, 50360, 50480, 50600, 50720, 50840.]
FIX_RA = [-505.7637580573554, -502.570356287689, -498.2089148883798, -492.67768482682357,
          -485.9770335870402, -478.1095526888573, -469.0801731788123, -458.89628893460525]
FIX_DEC = [-66.92982418533026, -37.47217527025044, -7.927548139010479, 21.63557115669823,
           51.147204404903704, 80.53589069730698, 109.72870493064629, 138.65128697876773]


def mp_states(el, epochs, consts=CONSTS):
    c = {k: mp.mpf(v) for k, v in consts.items()}
    elm = {k: mp.mpf(v) for k, v in el.items()}
    return [mpr.planet_state(c, elm, mp.mpf(float(t))) for t in epochs]


def fixture_pin():
    from scipy.optimize import least_squares
    consts = dict(CONSTS, kepler_year_days=365.2422)
    el = dict(a=12.0, e=0.11, i=np.radians(41.0), w=np.radians(38.0), W=np.radians(16.0), M=1.2, plx=50.0)

    def res(p):
        st = mp_states(dict(el, tp=p[0]), FIX_EPOCH, consts)
        return np.array([float(s[0] - mp.mpf(r)) for s, r in zip(st, FIX_RA)] +
                        [float(s[1] - mp.mpf(d)) for s, d in zip(st, FIX_DEC)])
    sol = least_squares(res, [41479.1485], xtol=1e-15, ftol=1e-15, gtol=1e-15)
    tp = float(sol.x[0])
    r = res([tp])
    print("fixture pin: tp =", repr(tp), "max |resid| [mas] =", np.abs(r).max())
    assert np.abs(r).max() < 5e-12
    json.dump({"source": "/root/reference/test/integration-tests.jl:8-15",
               "constants": consts, "orbit": dict(el, tp=tp),
               "epoch": FIX_EPOCH, "ra": FIX_RA, "dec": FIX_DEC,
               "max_abs_resid_mas_mpmath": float(np.abs(r).max())},
              open(os.path.join(OUT, "fixture8_pin.json"), "w"), indent=1)


def noisy(vals, sigma, rng):
    return [float(v) + float(sigma * rng.standard_normal()) for v in vals]


def build_cases():
    rng = np.random.default_rng(20261017)
    cases = {}

    # --- case 1: the reference's 8-epoch fixture as data, evaluated near the generating orbit
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=FIX_EPOCH, ra=FIX_RA, dec=FIX_DEC, σ_ra=[10.] * 8,
                                                σ_dec=[10.] * 8, cor=[0.] * 8), name="relastrom")
    b = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"], observations=[astrom])
    sys1 = octo.System(name="TestSys", variables=["M", "plx"], companions=[b])
    x1 = {"M": 1.21, "plx": 50.01, "b.a": 12.1, "b.e": 0.12, "b.i": 0.72, "b.ω": 0.65, "b.Ω": 0.29,
          "b.tp": 41500.0}
    cases["case_fixture8"] = (sys1, x1)

    # --- case 2/3: the northangle regression orbit (test/unit/likelihoods.jl:38-39), both table formats,
    #     with jitter / platescale / northangle sampled and a cor column
    el = dict(plx=50.0, M=1.2, a=15.0, e=0.2, i=0.6, w=0.3, W=1.1, tp=50000.0)
    epochs = [50000.0, 50300.0, 50600.0, 50900.0, 51200.0]
    st = mp_states(el, epochs)
    ra_m = [float(s[0]) for s in st]; dec_m = [float(s[1]) for s in st]
    pa_m = np.arctan2(ra_m, dec_m); sep_m = np.hypot(ra_m, dec_m)
    eps = 0.05
    tab_seppa = octo.Table(epoch=epochs, sep=noisy(sep_m, 1.0, rng), pa=noisy(pa_m + eps, 0.001, rng),
                           σ_sep=[1.0, 1.5, 0.8, 1.2, 1.0], σ_pa=[0.001, 0.002, 0.0015, 0.001, 0.003],
                           cor=[0.0, 0.3, -0.5, 0.1, 0.8])
    tab_radec = octo.Table(epoch=epochs, ra=noisy(sep_m * np.sin(pa_m + eps), 1.0, rng),
                           dec=noisy(sep_m * np.cos(pa_m + eps), 1.0, rng), σ_ra=[1.0, 1.5, 0.8, 1.2, 1.0],
                           σ_dec=[1.1, 0.9, 1.0, 1.3, 0.7], cor=[0.2, -0.3, 0.0, 0.6, -0.7])
    for nm, tab in (("case_northangle_seppa", tab_seppa), ("case_northangle_radec", tab_radec)):
        obs = octo.PlanetRelAstromObs(tab, name="inst", variables=["jitter", "platescale", "northangle"])
        pl = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"], observations=[obs])
        sy = octo.System(name="northangle_test", variables=["M", "plx"], companions=[pl])
        x = {"M": 1.19, "plx": 50.2, "b.a": 15.1, "b.e": 0.21, "b.i": 0.61, "b.ω": 0.31, "b.Ω": 1.09,
             "b.tp": 50010.0, "b.inst.jitter": 0.4 if "radec" in nm else 0.0007,
             "b.inst.platescale": 1.003, "b.inst.northangle": -0.045}
        cases[nm] = (sy, x)
    # same RA/Dec table without any obs variable: the jitter == 0 precomputed-distribution path
    obs = octo.PlanetRelAstromObs(tab_radec, name="inst")
    pl = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"], observations=[obs])
    sy = octo.System(name="plain_cor", variables=["M", "plx"], companions=[pl])
    cases["case_radec_cor_nojitter"] = (sy, {"M": 1.19, "plx": 50.2, "b.a": 15.1, "b.e": 0.21, "b.i": 0.61,
                                            "b.ω": 0.31, "b.Ω": 1.09, "b.tp": 50010.0})

    # --- case 4: the RV+astrometry tutorial orbit (docs/src/fit-rv-astrom.md:20-30): e = 0.7
    el = dict(a=1.0, e=0.7, i=np.pi / 4, W=0.1, w=np.pi / 4, M=1.0, plx=100.0, tp=58829.0 - 40)
    ep_a = [58849., 58852., 58858., 58890.]
    st = mp_states(el, ep_a)
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=ep_a, ra=noisy([s[0] for s in st], 1.0, rng),
                                                dec=noisy([s[1] for s in st], 1.0, rng), σ_ra=[1.0] * 4,
                                                σ_dec=[1.0] * 4, cor=[0.0] * 4), name="simulated")
    ep_rv = list(58849.0 + np.sort(rng.uniform(0, 365, 24)))
    mass = 30.0
    mu = mass * CONSTS["mjup2msol"] / el["M"]
    rv_true = [-mu * float(s[2]) for s in mp_states(el, ep_rv)]
    rv1 = octo.StarAbsoluteRVObs(octo.Table(epoch=ep_rv[:12], rv=noisy(np.array(rv_true[:12]) + 150.0, 5.0, rng),
                                            σ_rv=list(rng.uniform(3, 8, 12))), name="HARPS")
    rv2 = octo.MarginalizedStarAbsoluteRVObs(octo.Table(epoch=ep_rv[12:], rv=noisy(np.array(rv_true[12:]) - 40.0, 5.0, rng),
                                                        σ_rv=list(rng.uniform(3, 8, 12))), name="HIRES")
    pl = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[astrom])
    sy = octo.System(name="rvastrom", variables=["M", "plx"], companions=[pl], observations=[rv1, rv2])
    cases["case_rv_astrom"] = (sy, {"M": 1.02, "plx": 100.3, "HARPS.offset": 148.0, "HARPS.jitter": 3.0,
                                    "HIRES.jitter": 2.5, "b.a": 1.01, "b.e": 0.69, "b.i": 0.8, "b.ω": 0.77,
                                    "b.Ω": 0.12, "b.tp": 58790.0, "b.mass": 29.0})

    # --- case 5: two planets (OFTI example orbit + an inner companion), reflex astrometry,
    #     relative RV on the outer planet, star RV; hierarchical masses (each planet its own M)
    elb = dict(a=10.0, e=0.3, i=1.0, w=0.5, W=2.0, tp=50000.0, M=1.2, plx=50.0)
    elc = dict(a=4.0, e=0.1, i=1.0, w=1.3, W=2.0, tp=50400.0, M=1.195, plx=50.0)
    ep_b = list(np.linspace(50000, 50000 + 9000, 7)); ep_c = list(np.linspace(50100, 52800, 6))
    stb, stc = mp_states(elb, ep_b), mp_states(elc, ep_c)
    ab = octo.PlanetRelAstromObs(octo.Table(epoch=ep_b, ra=noisy([s[0] for s in stb], 2.0, rng),
                                            dec=noisy([s[1] for s in stb], 2.0, rng), σ_ra=[2.0] * 7,
                                            σ_dec=[2.5] * 7, cor=list(rng.uniform(-0.9, 0.9, 7))), name="GPI")
    ac = octo.PlanetRelAstromObs(octo.Table(epoch=ep_c, ra=noisy([s[0] for s in stc], 2.0, rng),
                                            dec=noisy([s[1] for s in stc], 2.0, rng), σ_ra=[2.0] * 6,
                                            σ_dec=[2.0] * 6), name="SPHERE", variables=["jitter"])
    ep_r = list(np.linspace(50050, 53000, 5))
    rvb = octo.PlanetRelativeRVObs(octo.Table(epoch=ep_r, rv=noisy([float(s[2]) for s in mp_states(elb, ep_r)], 300.0, rng),
                                              σ_rv=[300.0] * 5), name="CRIRES", variables=["jitter"])
    ep_s = list(np.linspace(50010, 53500, 9))
    rvs = octo.StarAbsoluteRVObs(octo.Table(epoch=ep_s, rv=noisy(np.zeros(9) + 20.0, 30.0, rng), σ_rv=[8.0] * 9),
                                 name="HARPS")
    pb = octo.Planet(name="b", variables=["M", "a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[ab, rvb])
    pc = octo.Planet(name="c", variables=["M", "a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[ac])
    sy = octo.System(name="two", variables=["plx"], companions=[pb, pc], observations=[rvs])
    cases["case_two_planet"] = (sy, {"plx": 50.1, "HARPS.offset": 18.0, "HARPS.jitter": 12.0,
                                     "b.M": 1.21, "b.a": 10.2, "b.e": 0.31, "b.i": 1.02, "b.ω": 0.52, "b.Ω": 1.97,
                                     "b.tp": 50030.0, "b.mass": 9.0, "b.CRIRES.jitter": 100.0,
                                     "c.M": 1.2, "c.a": 4.1, "c.e": 0.12, "c.i": 0.98, "c.ω": 1.28, "c.Ω": 2.03,
                                     "c.tp": 50390.0, "c.mass": 25.0, "c.SPHERE.jitter": 1.5})

    # --- case 6: the observable-based prior of O'Neil (2019) exactly as the reference's docstring attaches it
    #     (src/likelihoods/prior-observable.jl:26-52: `end astrom_like obs_prior`), on the 8-epoch fixture, plus a
    #     wrapped sep/PA table with sampled jitter / northangle on a second, outer planet (reflex of the inner, massive one)
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=FIX_EPOCH, ra=FIX_RA, dec=FIX_DEC, σ_ra=[10.] * 8,
                                                σ_dec=[10.] * 8, cor=[0.] * 8), name="relastrom")
    obs_seppa = octo.PlanetRelAstromObs(tab_seppa, name="inst", variables=["jitter", "northangle"])
    pb = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"],
                     observations=[astrom, octo.ObsPriorAstromONeil2019(astrom)])
    pc = octo.Planet(name="c", variables=["a", "e", "i", "ω", "Ω", "tp"],
                     observations=[octo.ObsPriorAstromONeil2019(obs_seppa)])
    sy = octo.System(name="obsprior", variables=["M", "plx"], companions=[pb, pc])
    cases["case_obsprior"] = (sy, {"M": 1.21, "plx": 50.01, "b.a": 12.1, "b.e": 0.12, "b.i": 0.72, "b.ω": 0.65,
                                   "b.Ω": 0.29, "b.tp": 41500.0, "b.mass": 20.0, "c.a": 15.1, "c.e": 0.21, "c.i": 0.61, "c.ω": 0.31,
                                   "c.Ω": 1.09, "c.tp": 50010.0, "c.obspri_inst.jitter": 0.0007,
                                   "c.obspri_inst.northangle": -0.045})

    # --- case 7: Hipparcos-Gaia proper-motion anomaly (HGCAInstantaneousObs, src/likelihoods/hgca.jl), N_ave = 3, two
    #     massive planets (the reference's planets-times-rows averaging shows with more than one), next to
    #     relative astrometry of the outer planet
    hgca_row = dict(pmra_hip=10.1, pmdec_hip=-5.2, pmra_hip_error=0.9, pmdec_hip_error=0.8, pmra_pmdec_hip=0.2,
                    pmra_hg=10.5, pmdec_hg=-5.0, pmra_hg_error=0.05, pmdec_hg_error=0.04, pmra_pmdec_hg=-0.1,
                    pmra_gaia=11.2, pmdec_gaia=-4.6, pmra_gaia_error=0.12, pmdec_gaia_error=0.1, pmra_pmdec_gaia=0.35,
                    epoch_ra_hip=1991.1, epoch_dec_hip=1991.3, epoch_ra_gaia=2016.0, epoch_dec_gaia=2016.2)
    hg = octo.HGCAInstantaneousObs(hgca_row, N_ave=3, factor=1.2)
    pb = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"])
    pc = octo.Planet(name="c", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[ac])
    sy = octo.System(name="pma", variables=["M", "plx", "pmra", "pmdec"], companions=[pb, pc], observations=[hg])
    cases["case_hgca"] = (sy, {"M": 1.2, "plx": 50.1, "pmra": 10.6, "pmdec": -4.9,
                               "b.a": 4.1, "b.e": 0.12, "b.i": 0.98, "b.ω": 1.28, "b.Ω": 2.03, "b.tp": 50390.0, "b.mass": 25.0,
                               "c.a": 10.2, "c.e": 0.31, "c.i": 1.02, "c.ω": 0.52, "c.Ω": 1.97, "c.tp": 50030.0, "c.mass": 9.0,
                               "c.SPHERE.jitter": 1.5})

    # --- case 8: ThieleInnesOrbit basis (docs/src/thiele-innes.md): the 8-epoch fixture observed on planet b, an inner
    #     massive Thiele-Innes planet c (reflex terms) with its own jitter table
    def campbell_to_ti(a, i, w, W, plx):
        s = a * plx
        return (s * (np.cos(W) * np.cos(w) - np.sin(W) * np.sin(w) * np.cos(i)), s * (np.sin(W) * np.cos(w) + np.cos(W) * np.sin(w) * np.cos(i)),
                s * (-np.cos(W) * np.sin(w) - np.sin(W) * np.cos(w) * np.cos(i)), s * (-np.sin(W) * np.sin(w) + np.cos(W) * np.cos(w) * np.cos(i)))
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=FIX_EPOCH, ra=FIX_RA, dec=FIX_DEC, σ_ra=[10.] * 8,
                                                σ_dec=[10.] * 8, cor=[0.] * 8), name="relastrom")
    pb = octo.Planet(name="b", basis="ThieleInnesOrbit", variables=["A", "B", "F", "G", "e", "tp"], observations=[astrom])
    pc = octo.Planet(name="c", basis="ThieleInnesOrbit", variables=["A", "B", "F", "G", "e", "tp", "mass"], observations=[ac])
    sy = octo.System(name="ti", variables=["M", "plx"], companions=[pb, pc])
    Ab, Bb, Fb, Gb = campbell_to_ti(12.1, 0.72, 0.65, 0.29, 50.01)
    Ac, Bc, Fc, Gc = campbell_to_ti(4.1, 0.98, 1.28, 2.03, 50.01)
    xti = {"M": 1.21, "plx": 50.01, "b.A": Ab, "b.B": Bb, "b.F": Fb, "b.G": Gb, "b.e": 0.12, "b.tp": 41500.0,
           "c.A": Ac, "c.B": Bc, "c.F": Fc, "c.G": Gc, "c.e": 0.12, "c.tp": 50390.0, "c.mass": 25.0, "c.SPHERE.jitter": 1.5}
    cases["case_thiele_innes"] = (sy, xti)
    # self-check of the basis: the same physical orbits in Campbell elements give the same likelihood
    pb2 = octo.Planet(name="b", variables=["a", "e", "i", "ω", "Ω", "tp"], observations=[astrom])
    pc2 = octo.Planet(name="c", variables=["a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[ac])
    sy2 = octo.System(name="ti_check", variables=["M", "plx"], companions=[pb2, pc2])
    xc = {"M": 1.21, "plx": 50.01, "b.a": 12.1, "b.e": 0.12, "b.i": 0.72, "b.ω": 0.65, "b.Ω": 0.29, "b.tp": 41500.0,
          "c.a": 4.1, "c.e": 0.12, "c.i": 0.98, "c.ω": 1.28, "c.Ω": 2.03, "c.tp": 50390.0, "c.mass": 25.0, "c.SPHERE.jitter": 1.5}
    cases["_check_ti_campbell"] = (sy2, xc)
    return cases


def build_post_cases():
    """Parameterised models (priors + UniformCircular + θ_at_epoch_to_tperi): the reference's 11-D test model
    (test/integration/sampling.jl:29-71, test/integration-tests.jl:17-40) and an RV+astrometry model."""
    cases = {}
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=FIX_EPOCH, ra=FIX_RA, dec=FIX_DEC, σ_ra=[10.] * 8,
                                                σ_dec=[10.] * 8, cor=[0.] * 8), name="relastrom")
    b = octo.Planet(name="b", observations=[astrom], variables={
        "a": octo.Uniform(0, 100), "e": octo.Uniform(0.0, 0.99), "i": octo.Sine(), "ω": octo.UniformCircular(),
        "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 50000)})
    sy = octo.System(name="TestSys", companions=[b], variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1)})
    # θ_t chosen so that the natural parameters sit near the fixture's generating orbit
    cases["post_fixture8"] = (sy, [np.log(1.2 - 0.1) + 0.01, np.log(50.0 - 0.1) + 1e-4, -1.99, -2.0, -0.5,
                                   0.8, 0.6, 0.95, 0.28, -0.99, -0.13])
    rng = np.random.default_rng(99)
    el = dict(a=1.0, e=0.7, i=np.pi / 4, W=0.1, w=np.pi / 4, M=1.0, plx=100.0, tp=58829.0 - 40)
    ep_a = [58849., 58852., 58858., 58890.]
    st = mp_states(el, ep_a)
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=ep_a, ra=noisy([s[0] for s in st], 1.0, rng),
                                                dec=noisy([s[1] for s in st], 1.0, rng), σ_ra=[1.0] * 4, σ_dec=[1.0] * 4),
                                     name="sim", variables={"jitter": octo.LogUniform(0.01, 10.0)})
    ep_rv = list(58849.0 + np.sort(rng.uniform(0, 365, 10)))
    mu = 30.0 * CONSTS["mjup2msol"] / el["M"]
    rv_true = [-mu * float(s[2]) for s in mp_states(el, ep_rv)]
    rv = octo.StarAbsoluteRVObs(octo.Table(epoch=ep_rv, rv=noisy(np.array(rv_true) + 150.0, 5.0, rng), σ_rv=[5.0] * 10),
                                name="HARPS", variables={"offset": octo.Normal(150, 100), "jitter": octo.LogUniform(0.1, 100.0)})
    pl = octo.Planet(name="b", observations=[astrom], variables={
        "a": octo.LogUniform(0.1, 10), "e": octo.Uniform(0, 0.999), "i": octo.Sine(), "ω": octo.UniformCircular(),
        "Ω": octo.UniformCircular(np.pi), "tp": octo.Uniform(58000, 59000), "mass": octo.truncated(octo.Normal(30, 20), lower=0, upper=200)})
    sy2 = octo.System(name="rvastrom", companions=[pl], observations=[rv],
                      variables={"M": octo.truncated(octo.Normal(1.0, 0.05), lower=0.1), "plx": 100.0})
    cases["post_rv_astrom"] = (sy2, [0.02 + np.log(0.9), 148.0, 0.1, -0.1, 0.8, -0.5, 0.7, 0.72, 0.99, 0.2, 1.32, -1.73, 0.05])
    # the reference's Thiele-Innes tutorial model (docs/src/thiele-innes.md:43-73)
    astrom = octo.PlanetRelAstromObs(octo.Table(epoch=FIX_EPOCH, ra=FIX_RA, dec=FIX_DEC, σ_ra=[10.] * 8,
                                                σ_dec=[10.] * 8, cor=[0.] * 8), name="GPI")
    bti = octo.Planet(name="b", basis="ThieleInnesOrbit", observations=[astrom], variables={
        "e": octo.Uniform(0.0, 0.5), "A": octo.Normal(0, 1000), "B": octo.Normal(0, 1000), "F": octo.Normal(0, 1000),
        "G": octo.Normal(0, 1000), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 50000.0)})
    sy3 = octo.System(name="TutoriaPrime", companions=[bti], variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1)})
    cases["post_thiele_innes"] = (sy3, [np.log(1.2 - 0.1) + 0.01, np.log(50.0 - 0.1) + 1e-4, -0.9, 250.0, 330.0, -420.0, 180.0, -0.99, -0.13])
    return cases


def main():
    only = set(sys.argv[1:])          # optional: regenerate just the named cases
    if not only:
        fixture_pin()
    for name, (system, th) in build_post_cases().items():
        if only and name not in only:
            continue
        spec = octo.ModelSpec(system)
        assert len(th) == spec.D, (name, len(th), spec.D, spec.theta_names)
        blocks = [{k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in b.items()} for b in spec.block_dicts]
        priors = [[int(p.family)] + [float(v) for v in p.p] for p in spec.priors]
        defs = [[int(d.op), [int(v) for v in d.a], float(d.value)] for d in spec.defs]
        lp, g = mpr.logpost_grad(CONSTS, spec.layout_dict, blocks, priors, defs, th)
        out = {"constants": CONSTS, "input_names": list(spec.input_names), "theta_names": list(spec.theta_names),
               "layout": spec.layout_dict, "blocks": blocks, "priors": priors, "defs": defs, "theta_t": [float(v) for v in th],
               "lp": float(lp), "grad": [float(v) for v in g], "lp_str": mp.nstr(lp, 30),
               "how": "oracle/mp_reference.py logpost_grad, mp.dps=60, central differences h=1e-18"}
        json.dump(out, open(os.path.join(OUT, name + ".json"), "w"), indent=1)
        print(name, "lp =", mp.nstr(lp, 20), "D =", spec.D)
    for name, (system, xd) in build_cases().items():
        if only and name not in only:
            continue
        spec = octo.ModelSpec(system)
        x = [xd[n] for n in spec.input_names]
        blocks = [{k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in b.items()} for b in spec.block_dicts]
        ll, g = mpr.ln_like_grad(CONSTS, spec.layout_dict, blocks, x)
        out = {"constants": CONSTS, "input_names": list(spec.input_names), "layout": spec.layout_dict,
               "blocks": blocks, "x": x, "ll": float(ll), "grad": [float(v) for v in g],
               "ll_str": mp.nstr(ll, 30), "grad_str": [mp.nstr(v, 30) for v in g],
               "how": "oracle/mp_reference.py, mp.dps=60, central differences h=1e-18"}
        if not name.startswith("_"):
            json.dump(out, open(os.path.join(OUT, name + ".json"), "w"), indent=1)
        print(name, "ll =", mp.nstr(ll, 20), "n_in =", len(x))


if __name__ == "__main__":
    main()
