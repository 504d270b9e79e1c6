"""Seeded synthetic workloads for BASELINE.json's configs.

Data synthesis only (numpy Newton solve for the noise-free model values); no likelihood is
evaluated here.  Used by tests/ and bench.py; nothing reads /root/reference.
"""
from __future__ import annotations

import numpy as np

import octofitter_jl_b200 as octo

KYJ = 365.2568983840419
MJUP2MSOL = 0.0009545942339693249
TRUTH_B = dict(a=10.0, e=0.3, i=1.0, w=0.5, W=2.0, tp=50000.0, M=1.2, plx=50.0, mass=10.0)
TRUTH_C = dict(a=4.0, e=0.1, i=1.0, w=1.3, W=2.0, tp=50400.0, M=1.2, plx=50.0, mass=5.0)


def _state(el, t):
    a, e, i, w, W, tp, M, plx = (el[k] for k in ("a", "e", "i", "w", "W", "tp", "M", "plx"))
    P = np.sqrt(a ** 3 / M) * KYJ
    MA = 2 * np.pi * (np.asarray(t, float) - tp) / P
    MA = MA - 2 * np.pi * np.round(MA / (2 * np.pi))
    E = MA + e * np.sin(MA)
    for _ in range(50):
        E = E - (E - e * np.sin(E) - MA) / (1 - e * np.cos(E))
    s = np.sqrt(1 - e * e)
    X, Y = np.cos(E) - e, s * np.sin(E)
    A = np.cos(W) * np.cos(w) - np.sin(W) * np.sin(w) * np.cos(i)
    B = np.sin(W) * np.cos(w) + np.cos(W) * np.sin(w) * np.cos(i)
    F = -np.cos(W) * np.sin(w) - np.sin(W) * np.cos(w) * np.cos(i)
    G = -np.sin(W) * np.sin(w) + np.cos(W) * np.cos(w) * np.cos(i)
    ra, dec = a * plx * (X * B + Y * G), a * plx * (X * A + Y * F)
    D = 1 - e * np.cos(E)
    K = (2 * np.pi * a / (P / 365.25)) / s * 1.495978707e11 / 31557600.0 * np.sin(i)
    rv = K * ((X / D) * np.cos(w) - (Y / D) * np.sin(w) + e * np.cos(w))
    return ra, dec, rv, P


def _astrom_table(el, epochs, rng, sigma=10.0, cor_frac=0.0, others=()):
    ra, dec, _, _ = _state(el, epochs)
    for o in others:        # interior companions pull the star
        r2, d2, _, _ = _state(o, epochs)
        mu = o["mass"] * MJUP2MSOL / o["M"]
        ra, dec = ra + mu * r2, dec + mu * d2
    n = len(epochs)
    tab = dict(epoch=epochs, ra=ra + sigma * rng.standard_normal(n), dec=dec + sigma * rng.standard_normal(n),
               σ_ra=np.full(n, sigma), σ_dec=np.full(n, sigma))
    if cor_frac > 0:
        cor = np.zeros(n)
        pick = rng.random(n) < cor_frac
        cor[pick] = rng.uniform(-0.9, 0.9, pick.sum())
        tab["cor"] = cor
    return octo.Table(**tab)


def _chains(spec, truth_by_name, n_chains, rng, rel=0.02):
    x0 = np.array([truth_by_name[n] for n in spec.input_names])
    x = x0[None, :] * (1.0 + rel * rng.standard_normal((n_chains, len(x0))))
    x[0] = x0
    for k, n in enumerate(spec.input_names):
        if n.endswith(".e"):
            x[:, k] = np.clip(x[:, k], 0.0, 0.95)
        if n.endswith("jitter"):
            x[:, k] = np.abs(x[:, k])
    return np.asfortranarray(x)


def one_planet(n_astrom, n_rv, n_chains, seed, span=0.9, with_mass=None):
    """1 planet; n_astrom RA/Dec epochs (+ n_rv star-RV epochs with offset & jitter, interleaved)."""
    rng = np.random.default_rng(seed)
    P = _state(TRUTH_B, [0.0])[3]
    obs, sysobs = [], []
    truth = {"M": 1.2, "plx": 50.0, "b.a": 10.0, "b.e": 0.3, "b.i": 1.0, "b.ω": 0.5, "b.Ω": 2.0, "b.tp": 50000.0}
    pvars = ["a", "e", "i", "ω", "Ω", "tp"]
    n_tot = n_astrom + n_rv
    grid = np.linspace(50000.0, 50000.0 + span * P, max(n_tot, 1))
    if n_astrom:
        ep = grid[::2][:n_astrom] if n_rv else grid
        ep = np.linspace(50000.0, 50000.0 + span * P, n_astrom) if len(ep) != n_astrom else ep
        obs.append(octo.PlanetRelAstromObs(_astrom_table(TRUTH_B, ep, rng), name="astrom"))
    if n_rv or with_mass:
        pvars.append("mass"); truth["b.mass"] = 10.0
    if n_rv:
        ep = grid[1::2][:n_rv]
        ep = np.linspace(50010.0, 50000.0 + span * P, n_rv) if len(ep) != n_rv else ep
        _, _, rv, _ = _state(TRUTH_B, ep)
        mu = TRUTH_B["mass"] * MJUP2MSOL / TRUTH_B["M"]
        rvd = 150.0 - mu * rv + np.hypot(5.0, 3.0) * rng.standard_normal(n_rv)
        sysobs.append(octo.StarAbsoluteRVObs(octo.Table(epoch=ep, rv=rvd, σ_rv=np.full(n_rv, 5.0)), name="rv"))
        truth["rv.offset"] = 150.0; truth["rv.jitter"] = 3.0
    b = octo.Planet(name="b", variables=pvars, observations=obs)
    system = octo.System(name="synthetic", variables=["M", "plx"], companions=[b], observations=sysobs)
    spec = octo.ModelSpec(system)
    return spec, _chains(spec, truth, n_chains, rng)


def two_planet(n_chains, seed, n_b=200, n_c=150, n_rv=150):
    """C3: hierarchical 2-planet system, 500 epochs: astrometry on both (cor on 20% of b), star RV."""
    rng = np.random.default_rng(seed)
    Pb = _state(TRUTH_B, [0.0])[3]
    ep_b = np.linspace(50000.0, 50000.0 + 0.9 * Pb, n_b)
    ep_c = np.linspace(50020.0, 50000.0 + 0.5 * Pb, n_c)
    ep_r = np.linspace(50005.0, 50000.0 + 0.4 * Pb, n_rv)
    ab = octo.PlanetRelAstromObs(_astrom_table(TRUTH_B, ep_b, rng, cor_frac=0.2, others=[TRUTH_C]), name="astrom_b")
    ac = octo.PlanetRelAstromObs(_astrom_table(TRUTH_C, ep_c, rng), name="astrom_c", variables=["jitter"])
    rv = sum(-el["mass"] * MJUP2MSOL / el["M"] * _state(el, ep_r)[2] for el in (TRUTH_B, TRUTH_C))
    rvo = octo.StarAbsoluteRVObs(octo.Table(epoch=ep_r, rv=150.0 + rv + np.hypot(5, 3) * rng.standard_normal(n_rv),
                                            σ_rv=np.full(n_rv, 5.0)), name="rv")
    pb = octo.Planet(name="b", variables=["M", "a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[ab])
    pc = octo.Planet(name="c", variables=["M", "a", "e", "i", "ω", "Ω", "tp", "mass"], observations=[ac])
    system = octo.System(name="two", variables=["plx"], companions=[pb, pc], observations=[rvo])
    spec = octo.ModelSpec(system)
    truth = {"plx": 50.0, "rv.offset": 150.0, "rv.jitter": 3.0, "c.astrom_c.jitter": 2.0}
    for nm, el in (("b", TRUTH_B), ("c", TRUTH_C)):
        truth.update({f"{nm}.M": el["M"], f"{nm}.a": el["a"], f"{nm}.e": el["e"], f"{nm}.i": el["i"],
                      f"{nm}.ω": el["w"], f"{nm}.Ω": el["W"], f"{nm}.tp": el["tp"], f"{nm}.mass": el["mass"]})
    return spec, _chains(spec, truth, n_chains, rng)


def k_planets_lean(k, n_chains, seed, n_ep=60, n_rv=80):
    """k planets (2..4), only lean tables: plain RA/Dec astrometry on every planet (with the reflex of the inner massive
    ones) and star RV with offset and jitter — the multi-planet instantiations of the lean kernel family."""
    rng = np.random.default_rng(seed)
    els = [dict(a=3.0 * 1.9 ** j, e=0.05 + 0.07 * j, i=0.9 + 0.05 * j, w=0.4 + 0.9 * j, W=1.8 + 0.1 * j,
                tp=50000.0 + 300.0 * j, M=1.2, plx=40.0, mass=(3.0 + 2 * j)) for j in range(k)]
    truth = {"M": 1.2, "plx": 40.0, "rv.offset": 20.0, "rv.jitter": 3.0}
    planets, names = [], "bcde"
    for j, el in enumerate(els):
        P = _state(el, [0.0])[3]
        ep = np.linspace(50000.0 + 7 * j, 50000.0 + min(0.8 * P, 9000.0), n_ep + 3 * j)
        tab = _astrom_table(el, ep, rng, sigma=3.0, others=[o for o in els if o["a"] < el["a"]])
        planets.append(octo.Planet(name=names[j], variables=["a", "e", "i", "ω", "Ω", "tp", "mass"],
                                   observations=[octo.PlanetRelAstromObs(tab, name=f"cam{j}")]))
        truth.update({f"{names[j]}.{q}": el[v] for q, v in (("a", "a"), ("e", "e"), ("i", "i"), ("ω", "w"), ("Ω", "W"), ("tp", "tp"), ("mass", "mass"))})
    ep_r = np.linspace(50005.0, 52000.0, n_rv)
    rv = sum(-el["mass"] * MJUP2MSOL / el["M"] * _state(el, ep_r)[2] for el in els)
    rvo = octo.StarAbsoluteRVObs(octo.Table(epoch=ep_r, rv=20.0 + rv + np.hypot(5, 3) * rng.standard_normal(n_rv),
                                            σ_rv=np.full(n_rv, 5.0)), name="rv")
    system = octo.System(name="lean%d" % k, variables=["M", "plx"], companions=planets, observations=[rvo])
    spec = octo.ModelSpec(system)
    return spec, _chains(spec, truth, n_chains, rng)


def config(name):
    """BASELINE.json configs by name: C1..C4; C5 via one_planet(E, 0, 4096, 5 + k)."""
    if name == "C1":
        return one_planet(50, 0, 1, seed=1)
    if name == "C2":
        return one_planet(100, 100, 1024, seed=2)
    if name == "C3":
        return two_planet(256, seed=3)
    if name == "C4":
        return one_planet(100, 0, 64, seed=4)
    raise KeyError(name)


HGCA_ROW = dict(pmra_hip=10.1, pmdec_hip=-5.2, pmra_hip_error=0.9, pmdec_hip_error=0.8, pmra_pmdec_hip=0.2,
                pmra_hg=10.5, pmdec_hg=-5.0, pmra_hg_error=0.5, pmdec_hg_error=0.4, pmra_pmdec_hg=-0.1,
                pmra_gaia=11.2, pmdec_gaia=-4.6, pmra_gaia_error=0.3, pmdec_gaia_error=0.25, pmra_pmdec_gaia=0.35,
                epoch_ra_hip=1991.1, epoch_dec_hip=1991.3, epoch_ra_gaia=2016.0, epoch_dec_gaia=2016.2)


def many_planets(n_planets, n_chains, seed, n_ep=40, extras=False):
    """3-4 planet system touching every observation kind: RA/Dec (+cor, +jitter), PA/sep (+platescale,
    northangle), relative RV, star RV and marginalised star RV; masses on all but the outermost planet."""
    rng = np.random.default_rng(seed)
    els = [dict(a=3.0 * 1.9 ** k, e=0.05 + 0.07 * k, i=0.9 + 0.05 * k, w=0.4 + 0.9 * k, W=1.8 + 0.1 * k,
                tp=50000.0 + 300.0 * k, M=1.2, plx=40.0, mass=(3.0 + 2 * k)) for k in range(n_planets)]
    truth = {"M": 1.2, "plx": 40.0}
    planets, names = [], "bcde"
    for k, el in enumerate(els):
        P = _state(el, [0.0])[3]
        ep = np.linspace(50000.0 + 7 * k, 50000.0 + min(0.8 * P, 9000.0), n_ep + 3 * k)
        inner = [o for o in els if o["a"] < el["a"]]
        obs = []
        if k % 2 == 0:
            tab = _astrom_table(el, ep, rng, sigma=3.0, cor_frac=0.3, others=inner)
            obs.append(octo.PlanetRelAstromObs(tab, name=f"cam{k}", variables=["jitter"] if k == 2 else []))
            if k == 2:
                truth[f"{names[k]}.cam{k}.jitter"] = 1.5
            if extras:          # the observable-based prior wrapped around the same table (its own copy of the variables)
                obs.append(octo.ObsPriorAstromONeil2019(obs[-1]))
                if k == 2:
                    truth[f"{names[k]}.obspri_cam{k}.jitter"] = 1.2
        else:
            ra, dec, _, _ = _state(el, ep)
            for o in inner:
                r2, d2, _, _ = _state(o, ep); mu = o["mass"] * MJUP2MSOL / o["M"]; ra, dec = ra + mu * r2, dec + mu * d2
            sep, pa = np.hypot(ra, dec), np.arctan2(ra, dec)
            tab = octo.Table(epoch=ep, sep=sep + 2.0 * rng.standard_normal(len(ep)), pa=pa + 0.002 * rng.standard_normal(len(ep)),
                             σ_sep=np.full(len(ep), 2.0), σ_pa=np.full(len(ep), 0.002), cor=rng.uniform(-0.5, 0.5, len(ep)))
            obs.append(octo.PlanetRelAstromObs(tab, name=f"ifs{k}", variables=["platescale", "northangle"]))
            truth[f"{names[k]}.ifs{k}.platescale"] = 1.002; truth[f"{names[k]}.ifs{k}.northangle"] = 0.003
            epr = np.linspace(50100.0, 50900.0, 11)
            obs.append(octo.PlanetRelativeRVObs(octo.Table(epoch=epr, rv=_state(el, epr)[2] + 200 * rng.standard_normal(11),
                                                           σ_rv=np.full(11, 200.0)), name=f"crires{k}", variables=["jitter"]))
            truth[f"{names[k]}.crires{k}.jitter"] = 80.0
        pv = ["a", "e", "i", "ω", "Ω", "tp"] + ([] if k == n_planets - 1 and False else ["mass"])
        planets.append(octo.Planet(name=names[k], variables=pv, observations=obs))
        truth.update({f"{names[k]}.a": el["a"], f"{names[k]}.e": el["e"], f"{names[k]}.i": el["i"], f"{names[k]}.ω": el["w"],
                      f"{names[k]}.Ω": el["W"], f"{names[k]}.tp": el["tp"], f"{names[k]}.mass": el["mass"]})
    eps = np.linspace(50020.0, 53000.0, 25)
    rv = sum(-el["mass"] * MJUP2MSOL / el["M"] * _state(el, eps)[2] for el in els)
    s1 = octo.StarAbsoluteRVObs(octo.Table(epoch=eps[:13], rv=30 + rv[:13] + 4 * rng.standard_normal(13), σ_rv=np.full(13, 4.0)),
                                name="harps", variables=["offset"])
    s2 = octo.MarginalizedStarAbsoluteRVObs(octo.Table(epoch=eps[13:], rv=-12 + rv[13:] + 4 * rng.standard_normal(12),
                                                       σ_rv=np.full(12, 4.0)), name="hires")
    truth.update({"harps.offset": 30.0, "hires.jitter": 2.0})
    sysobs, sysvars = [s1, s2], ["M", "plx"]
    if extras:                  # Hipparcos-Gaia proper-motion anomaly next to everything else
        sysobs.append(octo.HGCAInstantaneousObs(HGCA_ROW, N_ave=4)); sysvars += ["pmra", "pmdec"]
        truth.update({"pmra": 10.6, "pmdec": -4.9})
    system = octo.System(name="many", variables=sysvars, companions=planets, observations=sysobs)
    spec = octo.ModelSpec(system)
    return spec, _chains(spec, truth, n_chains, rng, rel=0.01)


def one_planet_with_priors(n_astrom, n_rv, n_chains, seed):
    """The C2 tables with the reference's standard priors attached (device-side parameterisation, N1):
    returns (spec, θ_t [n_chains x D]) with θ_t scattered around the truth in unconstrained space."""
    raw_spec, _ = one_planet(n_astrom, n_rv, 1, seed)
    sys0 = raw_spec.system
    planet0 = sys0.planets[0]
    astrom = [octo.PlanetRelAstromObs(o.table, name=o.name) for o in planet0.observations]
    rvs = [octo.StarAbsoluteRVObs({"epoch": o.table["epoch"], "rv": o.table["rv"], "σ_rv": o.table["σ_rv"]}, name=o.name,
                                  variables={"offset": octo.Normal(150, 100), "jitter": octo.LogUniform(0.1, 100.0)})
           for o in sys0.observations]
    pv = {"a": octo.LogUniform(1, 100), "e": octo.Uniform(0, 0.99), "i": octo.Sine(), "ω": octo.UniformCircular(),
          "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 50000.0)}
    if rvs:
        pv["mass"] = octo.LogUniform(0.1, 100)
    b = octo.Planet(name="b", variables=pv, observations=astrom)
    system = octo.System(name="synthetic", companions=[b], observations=rvs, variables={
        "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1)})
    spec = octo.ModelSpec(system)
    rng = np.random.default_rng(seed + 77)
    # truth in natural space -> θ_t by hand (logit / log-shift / identity), then scatter
    ra, dec, _, _ = _state(TRUTH_B, [50000.0])
    theta_pa = float(np.arctan2(ra[0], dec[0]))
    nat = {"M": 1.2, "plx": 50.0, "rv.offset": 150.0, "rv.jitter": 3.0, "b.a": 10.0, "b.e": 0.3, "b.i": 1.0, "b.mass": 10.0}
    ang = {"b.ω": 0.5, "b.Ω": 2.0, "b.θ": theta_pa}
    th0 = np.zeros(spec.D)
    for j, (name, pr) in enumerate(zip(spec.theta_names, spec.priors)):
        if name[:-1] in ang and name[-1] in "xy":
            th0[j] = np.cos(ang[name[:-1]]) if name[-1] == "x" else np.sin(ang[name[:-1]])
            continue
        x = nat[name]
        lo, hi = -np.inf, np.inf
        if pr.family in (1, 2):
            lo, hi = pr.p[0], pr.p[1]
        elif pr.family == 3:
            lo, hi = 0.0, np.pi
        elif pr.family == 4:
            lo, hi = pr.p[2], pr.p[3]
        if np.isfinite(lo) and np.isfinite(hi):
            u = (x - lo) / (hi - lo); th0[j] = np.log(u / (1 - u))
        elif np.isfinite(lo):
            th0[j] = np.log(x - lo)
        else:
            th0[j] = x
    th = th0[None, :] + 0.01 * rng.standard_normal((n_chains, spec.D)) * np.maximum(1.0, 0.0)
    th[0] = th0
    return spec, np.asfortranarray(th)
