"""Host-side mirror of the reference interface for the hot path.

Same names and argument meaning as the Julia types the path sits behind
(src/variables.jl:468-594 `Planet`/`System`; src/likelihoods/relative-astrometry.jl:19-93
`PlanetRelAstromObs`; OctofitterRadialVelocity/src/rv-*.jl; src/likelihoods/prior-observable.jl;
src/likelihoods/hgca.jl; src/logdensitymodel.jl `LogDensityModel`), reduced to what the path needs: the
observation tables, which variables exist, and where they sit in the kernel's input matrix.

Two ways to declare variables.  A list of names: the caller passes NATURAL-space values of every variable
(`ln_like`, `ln_like_and_gradient`) and keeps priors/bijectors to itself (SURVEY.md §8b "split").  A dict
name -> prior | UniformCircular | constant | θ_at_epoch_to_tperi: the standard parameterisation then runs on the
device as well and the model offers the sampler-facing surface of the reference (`ℓπcallback`, `ℓπcallback_grad`,
`link`/`invlink`, `sample_priors`, ...).

The compute is exclusively libocto_b200.so (hand-written sm_100a kernels).  Nothing here
evaluates a likelihood on the CPU.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _abi

astrom_cols1 = ("epoch", "ra", "dec", "σ_ra", "σ_dec")
astrom_cols3 = ("epoch", "pa", "sep", "σ_pa", "σ_sep")
rv_cols = ("epoch", "rv", "σ_rv")
_ALIASES = {"sigma_ra": "σ_ra", "sigma_dec": "σ_dec", "sigma_pa": "σ_pa", "sigma_sep": "σ_sep",
            "sigma_rv": "σ_rv", "omega": "ω", "Omega": "Ω", "w": "ω", "W": "Ω"}
_MJD_1950, _MJD_2050 = 33282.0, 69807.0


class OctoError(RuntimeError):
    pass


def Table(rows=None, **cols):
    """TypedTables-like constructor: Table(epoch=[...], ra=[...]) or Table([{...}, {...}])."""
    if rows is not None:
        if isinstance(rows, dict):
            cols = dict(rows)
        else:
            keys = list(rows[0].keys())
            cols = {k: [r[k] for r in rows] for k in keys}
    out = {}
    for k, v in cols.items():
        out[_ALIASES.get(k, k)] = np.atleast_1d(np.asarray(v, dtype=np.float64)).copy()
    return out


def _check_table(table, name):
    table = Table(table) if not all(isinstance(v, np.ndarray) for v in table.values()) else dict(table)
    n = {len(v) for v in table.values()}
    if len(n) > 1:
        raise ValueError("The columns in the input data do not all have the same length")
    ep = table.get("epoch")
    if ep is not None and len(ep) and (np.any(ep >= _MJD_2050) or np.any(ep <= _MJD_1950)):
        warnings.warn("The data you entered fell outside the range year 1950 to year 2050. "
                      "The expected input format is MJD (modified julian date).")
    return table


# ---------------------------------------------------------------------------------------------------------
# The `@variables` vocabulary of the reference that the device-side parameterisation understands:
# priors (Distributions.jl families used by the reference's docs/tests), UniformCircular, constants, and the
# derived variable tp = θ_at_epoch_to_tperi(θ, t_ref; M, e, a, i, ω, Ω).
# ---------------------------------------------------------------------------------------------------------
class Prior:
    family, p = -1, (0.0, 0.0, 0.0, 0.0)


class Normal(Prior):
    def __init__(self, mu, sigma):
        if not sigma > 0:
            raise ValueError("Normal: the condition σ > 0 is not satisfied")
        self.family, self.p = _abi.PRIOR_NORMAL, (float(mu), float(sigma), 0.0, 0.0)


class Uniform(Prior):
    def __init__(self, a, b):
        if not a < b:
            raise ValueError("Uniform: the condition a < b is not satisfied")
        self.family, self.p = _abi.PRIOR_UNIFORM, (float(a), float(b), 0.0, 0.0)


class LogUniform(Prior):
    def __init__(self, a, b):
        if not 0 < a < b:
            raise ValueError("LogUniform: the condition 0 < a < b is not satisfied")
        self.family, self.p = _abi.PRIOR_LOGUNIFORM, (float(a), float(b), 0.0, 0.0)


class Sine(Prior):
    """src/distributions.jl:15-40: pdf sin(x)/2 on (0, π)."""
    def __init__(self):
        self.family, self.p = _abi.PRIOR_SINE, (0.0, 0.0, 0.0, 0.0)


def truncated(d, lower=None, upper=None):
    """truncated(Normal(μ, σ), lower=, upper=) as used throughout the reference's docs."""
    if not isinstance(d, Normal):
        raise ValueError("only truncated(Normal(...)) is offloaded")
    out = Prior()
    out.family = _abi.PRIOR_TRUNCNORMAL
    out.p = (d.p[0], d.p[1], -np.inf if lower is None else float(lower), np.inf if upper is None else float(upper))
    if not out.p[2] < out.p[3]:
        raise ValueError("truncated: lower must be below upper")
    return out


class UniformCircular:
    """src/variables.jl:260-299: expands to `<v>x, <v>y ~ Normal(0, 1)`, `<v> = atan(<v>y, <v>x)/2π*domain` and a
    UnitLengthPrior on the vector length."""
    def __init__(self, domain=2 * np.pi):
        self.domain = float(domain)


class θ_at_epoch_to_tperi:
    """Derived `tp = θ_at_epoch_to_tperi(θ, t_ref; M, e, a, i, ω, Ω)` (src/parameterizations.jl:6-69); `theta` names
    the position-angle variable of the same planet, the orbital elements are taken from merge(system, planet)."""
    def __init__(self, theta, theta_epoch):
        self.theta, self.theta_epoch = _ALIASES.get(theta, theta), float(theta_epoch)


theta_at_epoch_to_tperi = θ_at_epoch_to_tperi


def _norm_variables(variables):
    """list of names (raw natural-space inputs) or {name: prior | UniformCircular | number | θ_at_epoch_to_tperi}"""
    if isinstance(variables, dict):
        return tuple((_ALIASES.get(k, k), v) for k, v in variables.items())
    return tuple((_ALIASES.get(v, v), None) for v in variables)


class AbstractObs:
    kind = -1
    allowed_variables = ()

    def _set_variables(self, variables):
        self.var_specs = _norm_variables(variables)
        self.variables = tuple(n for n, _ in self.var_specs)
        for v in self.variables:
            if v not in self.allowed_variables:
                raise ValueError(f"{type(self).__name__} '{self.name}': variable '{v}' is not offloadable "
                                 f"(supported: {self.allowed_variables}); keep such terms in the host model")


class PlanetRelAstromObs(AbstractObs):
    """Relative astrometry of a planet (relative-astrometry.jl:19-93).

    Columns: epoch plus (ra, dec, σ_ra, σ_dec) or (pa, sep, σ_pa, σ_sep), optional cor.
    Units: mas / rad.  The table is sorted by epoch, as the reference ctor does (:46-47).
    variables: any of "jitter", "platescale", "northangle" that are sampled/defined.
    """
    allowed_variables = ("jitter", "platescale", "northangle")

    def __init__(self, observations, *, name, variables=()):
        self.name = str(name)
        table = _check_table(observations, name)
        c1 = set(astrom_cols1) <= set(table)
        c3 = set(astrom_cols3) <= set(table)
        if not (c1 or c3):
            raise ValueError(f"Expected columns {astrom_cols1} or {astrom_cols3}")
        ii = np.argsort(table["epoch"], kind="stable")
        table = {k: v[ii] for k, v in table.items()}
        if "pa" in table and "sep" in table:
            self.kind = _abi.KIND_ASTROM_PASEP
            if np.any(table["pa"] >= 2 * np.pi) or np.any(table["pa"] <= -2 * np.pi):
                warnings.warn("The data you entered fell outside the range [-2pi, +2pi]. "
                              "The expected input format is radians.")
        else:
            self.kind = _abi.KIND_ASTROM_RADEC
        if "cor" in table and np.any(np.abs(table["cor"]) > 1 - 1e-5):
            raise ValueError(f"Correlation values may not be well-specified: {table['cor']}")
        self.table = table
        self._set_variables(variables)

    def _columns(self):
        t = self.table
        if self.kind == _abi.KIND_ASTROM_PASEP:
            return t["epoch"], t["pa"], t["sep"], t["σ_pa"], t["σ_sep"], t.get("cor")
        return t["epoch"], t["ra"], t["dec"], t["σ_ra"], t["σ_dec"], t.get("cor")


PlanetRelAstromLikelihood = PlanetRelAstromObs


class ObsPriorAstromONeil2019(AbstractObs):
    """`ObsPriorAstromONeil2019(astrometry_likelihood)` (src/likelihoods/prior-observable.jl:12-137): the observable-based
    prior of O'Neil (2019) for relative astrometry.  Like the reference object it wraps the table (its ln_like is the
    wrapped likelihood's plus 2 log of the summed Jacobian term over the table's epochs) and carries its own copy of
    the wrapped likelihood's variables, under the name "obspri_<name>"."""

    def __init__(self, obs):
        if not isinstance(obs, PlanetRelAstromObs):
            raise ValueError("the observable-based prior is offloaded for PlanetRelAstromObs only")
        self.wrapped_like = obs
        self.name = "obspri_" + obs.name
        self.kind, self.table = obs.kind, obs.table
        self.allowed_variables = obs.allowed_variables
        self.var_specs, self.variables = obs.var_specs, obs.variables
        self.obs_prior = 1

    def _columns(self):
        return self.wrapped_like._columns()


class HGCAInstantaneousObs(AbstractObs):
    """`HGCAInstantaneousObs(; gaia_id, N_ave=1, factor=1)` (src/likelihoods/hgca.jl:29-153): Hipparcos-Gaia Catalog of
    Accelerations proper-motion anomaly, instantaneous approximation.  The reference reads the star's row from the
    HGCA FITS file; here the caller passes that row as `hgca` (a mapping with the catalogue's column names:
    pmra_hip, pmdec_hip, pmra_hip_error, pmdec_hip_error, pmra_pmdec_hip, the same for _hg and _gaia, and
    epoch_ra_hip, epoch_dec_hip, epoch_ra_gaia, epoch_dec_gaia in Julian years).  System-level; the system needs
    `pmra` and `pmdec` variables and every planet a `mass`."""
    kind = _abi.KIND_HGCA_INSTANT
    allowed_variables = ()

    def __init__(self, hgca, *, N_ave=1, factor=1.0, name="hgca", variables=()):
        self.name = str(name)
        self.hgca = dict(hgca)
        h = self.hgca
        julian_year, J2000_mjd = 365.25, 51544.5
        mjd = lambda key: (float(h[key]) - 2000.0) * julian_year + J2000_mjd          # hgca.jl:78-84
        e_ra_hip, e_dec_hip, e_ra_gaia, e_dec_gaia = (mjd(k) for k in ("epoch_ra_hip", "epoch_dec_hip", "epoch_ra_gaia", "epoch_dec_gaia"))
        dt_gaia, dt_hip = 1038.0, 4 * 365.25                                            # hgca.jl:86-88
        if N_ave == 1:
            d_hip = d_gaia = [0.0]
        else:
            d_hip = np.linspace(-dt_hip / 2, dt_hip / 2, N_ave); d_gaia = np.linspace(-dt_gaia / 2, dt_gaia / 2, N_ave)
        epoch, code = [], []
        for d in d_hip:
            epoch += [e_ra_hip + d, e_dec_hip + d]; code += [0.0, 1.0]
        for d in d_gaia:
            epoch += [e_ra_gaia + d, e_dec_gaia + d]; code += [2.0, 3.0]
        self.table = {"epoch": np.asarray(epoch), "code": np.asarray(code)}
        f = float(factor)
        self.aux = np.asarray([v for tag in ("hip", "hg", "gaia") for v in (
            float(h[f"pmra_{tag}"]), float(h[f"pmdec_{tag}"]), float(h[f"pmra_{tag}_error"]) * f,
            float(h[f"pmdec_{tag}_error"]) * f, float(h[f"pmra_pmdec_{tag}"]))], dtype=np.float64)
        self._set_variables(variables)

    def _columns(self):
        return self.table["epoch"], self.table["code"], None, None, None, None


class _RVObs(AbstractObs):
    """`trend_function(θ_obs, epoch)` as in the reference (rv-absolute.jl:69,143; rv-absolute-margin.jl:52,111;
    rv-relative.jl:64,131): any callable that is LINEAR in the observation variables — the docs' example
    `θ_obs.trend_slope * (epoch - 57000)`, polynomials with coefficient variables, fixed-period sinusoids with amplitude
    variables.  Its coefficient variables (at most three) are listed in `variables` next to offset / jitter.  The closure
    is probed here (unit vectors in θ_obs, then a linearity check at random points) and handed to the kernel as per-epoch
    basis values; anything non-linear is refused — in Julia such an observation simply stays on the host."""

    def __init__(self, observations, *, name, variables=None, trend_function=None, gaussian_process=None):
        self.name = str(name)
        if gaussian_process is not None:
            raise ValueError("gaussian_process likelihoods are out of scope for the offloaded path "
                             "(SURVEY.md §2: celerite GP branch stays in the reference)")
        table = _check_table(observations, name)
        if not set(rv_cols) <= set(table):
            raise ValueError(f"Expected columns {rv_cols}")
        self.table = table
        self.trend = None
        specs = _norm_variables(self.default_variables if variables is None else variables)
        if trend_function is not None:
            self.trend = self._probe_trend(trend_function, [n for n, _ in specs])
            self.allowed_variables = tuple(self.allowed_variables) + tuple(self.trend[0])
        self._set_variables(self.default_variables if variables is None else variables)

    def _probe_trend(self, f, names):
        from types import SimpleNamespace
        ep = self.table["epoch"]
        ev = lambda th: np.array([float(f(SimpleNamespace(**th), t)) for t in ep], dtype=np.float64)
        zero = {n: 0.0 for n in names}
        b0 = ev(zero)
        basis = {}
        for n in names:
            bn = ev({**zero, n: 1.0}) - b0
            if np.any(bn != 0.0):
                basis[n] = bn
        rng = np.random.default_rng(0)
        for _ in range(4):
            th = {n: 10.0 * rng.standard_normal() for n in names}
            lin = b0 + sum(th[n] * basis[n] for n in basis)
            got = ev(th)
            if not np.allclose(got, lin, rtol=1e-11, atol=1e-11 * max(1.0, float(np.abs(lin).max(initial=0.0)))):
                raise ValueError("trend_function is not linear in the observation variables: not offloadable "
                                 "(the reference evaluates such a trend on the host)")
        if len(basis) > 3:
            raise ValueError("at most three trend coefficient variables are offloaded")
        if any(n in ("offset", "jitter") for n in basis):
            raise ValueError("a trend_function that reads offset / jitter is not offloaded")
        return list(basis), (np.stack([basis[n] for n in basis]) if basis else np.zeros((0, len(ep)))), (b0 if np.any(b0 != 0.0) else None)

    def _columns(self):
        t = self.table
        return t["epoch"], t["rv"], None, t["σ_rv"], None, None


class StarAbsoluteRVObs(_RVObs):
    """Absolute RV of the star, no GP (rv-absolute.jl:55-112, 135-204)."""
    kind = _abi.KIND_RV_STAR_ABS
    allowed_variables = ("offset", "jitter")
    default_variables = ("offset", "jitter")      # rv-absolute.jl:74-79 default priors


class MarginalizedStarAbsoluteRVObs(_RVObs):
    """Absolute RV with the zero point marginalised analytically (rv-absolute-margin.jl:140-185)."""
    kind = _abi.KIND_RV_STAR_MARGIN
    allowed_variables = ("jitter",)
    default_variables = ("jitter",)

    def __init__(self, observations, *, name, variables=None, trend_function=None):
        super().__init__(observations, name=name, variables=variables, trend_function=trend_function)
        if "jitter" not in self.variables:
            raise ValueError("MarginalizedStarAbsoluteRVObs requires a `jitter` variable")


class PlanetRelativeRVObs(_RVObs):
    """RV of a planet relative to its star, no GP (rv-relative.jl:121-211)."""
    kind = _abi.KIND_RV_PLANET_REL
    allowed_variables = ("offset", "jitter")
    default_variables = ("jitter",)


StarAbsoluteRVLikelihood = StarAbsoluteRVObs
MarginalizedStarAbsoluteRVLikelihood = MarginalizedStarAbsoluteRVObs
PlanetRelativeRVLikelihood = PlanetRelativeRVObs

_ELEMS = ("a", "e", "i", "ω", "Ω", "tp", "M", "plx")


class Planet:
    """Planet(name=, basis=, variables=, observations=) — src/variables.jl:468-508.

    variables: names of this planet's natural-space variables (e.g. "a","e","i","ω","Ω","tp","mass").
    Bases: Visual{KepOrbit}, and RadialVelocityOrbit (PlanetOrbits: a, e, ω, tp, M only — the orbit as radial
    velocities see it, K without the sin i factor).  The latter is the same kernel with i = π/2 (sin i = 1 exactly),
    Ω = 0 and an arbitrary parallax injected as constants, so it needs the dict form of `variables`; astrometry
    cannot be attached to it.
    """

    def __init__(self, *, name, basis="Visual{KepOrbit}", variables, observations=()):
        if basis not in ("Visual{KepOrbit}", "VisualKepOrbit", "RadialVelocityOrbit", "ThieleInnesOrbit"):
            raise ValueError(f"basis {basis!r} is not offloaded; only Visual{{KepOrbit}}, ThieleInnesOrbit and RadialVelocityOrbit")
        self.name = str(name)
        self.basis = basis if basis in ("RadialVelocityOrbit", "ThieleInnesOrbit") else "Visual{KepOrbit}"
        if self.basis == "ThieleInnesOrbit" and any(o.kind == _abi.KIND_RV_PLANET_REL for o in observations):
            raise ValueError("radial velocities of a ThieleInnesOrbit planet are not offloaded")
        if self.basis == "RadialVelocityOrbit":
            if not isinstance(variables, dict):
                raise ValueError("RadialVelocityOrbit needs variables as a dict (i, Ω and plx are injected as constants)")
            if any(o.kind in (_abi.KIND_ASTROM_RADEC, _abi.KIND_ASTROM_PASEP) for o in observations):
                raise ValueError("a RadialVelocityOrbit has no sky-plane projection: astrometry cannot be attached to it")
            variables = dict(variables)
            for k, v in (("i", np.pi / 2), ("Ω", 0.0), ("plx", 1.0)):
                if _ALIASES.get(k, k) not in {_ALIASES.get(n, n) for n in variables}:
                    variables[k] = v
        self.var_specs = _norm_variables(variables)
        self.variables = tuple(n for n, _ in self.var_specs)
        self.observations = list(observations)
        for o in self.observations:
            if o.kind in (_abi.KIND_RV_STAR_ABS, _abi.KIND_RV_STAR_MARGIN):
                raise ValueError(f"{type(o).__name__} is a system-level observation")


class System:
    """System(name=, variables=, companions=, observations=) — src/variables.jl:544-594."""

    def __init__(self, *, name, variables, companions, observations=()):
        self.name = str(name)
        self.var_specs = _norm_variables(variables)
        self.variables = tuple(n for n, _ in self.var_specs)
        self.planets = list(companions)
        self.observations = list(observations)
        names = [p.name for p in self.planets]
        if len(set(names)) != len(names):
            raise ValueError("planet names must be unique")
        for o in self.observations:
            if o.kind not in (_abi.KIND_RV_STAR_ABS, _abi.KIND_RV_STAR_MARGIN, _abi.KIND_HGCA_INSTANT):
                raise ValueError(f"{type(o).__name__} must be attached to a planet")


def _normalizename(s):
    # src/variables.jl:1068-1073: non-identifier characters become underscores
    out = "".join(ch if (ch.isalnum() or ch == "_") else "_" for ch in s)
    return out


class ModelSpec:
    """What make_ln_like (system.jl:21-110) derives from a System: the epoch tables in summation
    order and the position of every variable in the kernel-input matrix.  Pure host bookkeeping.

    `input_names` lists the kernel-input columns in the reference's parameter order: system
    variables, system-observation variables, then per planet its variables followed by its
    observations' variables (src/variables.jl:691-730).
    Observation blocks are listed in the order the reference sums them: planet observations
    first, then system observations (system.jl:223-236).
    """

    def __init__(self, system: System):
        self.system = system
        names = list(system.variables)
        col = {n: k for k, n in enumerate(names)}
        obs_cols = {}
        for o in system.observations:
            for v in o.variables:
                key = f"{_normalizename(o.name)}.{v}"
                obs_cols[(id(o), v)] = len(names)
                names.append(key)
        planets_layout = []
        for p in system.planets:
            pcol = {}
            for v in p.variables:
                pcol[v] = len(names)
                names.append(f"{p.name}.{v}")
            for o in p.observations:
                for v in o.variables:
                    obs_cols[(id(o), v)] = len(names)
                    names.append(f"{p.name}.{_normalizename(o.name)}.{v}")
            merged = dict(col)
            merged.update(pcol)          # merge(θ_system, θ_planet): planet wins (system.jl:117)
            if p.basis == "ThieleInnesOrbit":
                missing = [e for e in ("A", "B", "F", "G", "e", "tp", "M", "plx") if e not in merged]
                if missing:
                    raise OctoError(f"planet {p.name}: missing orbital variables {missing} for ThieleInnesOrbit")
                if any(o.kind in (_abi.KIND_RV_STAR_ABS, _abi.KIND_RV_STAR_MARGIN) for o in system.observations):
                    raise OctoError("star radial velocities with a ThieleInnesOrbit planet are not offloaded")
                planets_layout.append({"basis": 1, "A": merged["A"], "B": merged["B"], "F": merged["F"], "G": merged["G"],
                                       "e": merged["e"], "tp": merged["tp"], "M": merged["M"], "plx": merged["plx"],
                                       "a": -1, "i": -1, "w": -1, "W": -1, "mass": pcol.get("mass", -1)})
                continue
            missing = [e for e in _ELEMS if e not in merged]
            if missing:
                raise OctoError(f"planet {p.name}: missing orbital variables {missing} for Visual{{KepOrbit}}")
            planets_layout.append({"plx": merged["plx"], "a": merged["a"], "e": merged["e"], "i": merged["i"],
                                   "w": merged["ω"], "W": merged["Ω"], "tp": merged["tp"], "M": merged["M"],
                                   "mass": pcol.get("mass", -1)})
        self.input_names = tuple(names)
        self.n_in = len(names)
        self._build_parameterization(system, col)
        blocks = []
        for ip, p in enumerate(system.planets):
            for o in p.observations:
                blocks.append(self._block(o, ip, obs_cols))
        for o in system.observations:
            blk = self._block(o, -1, obs_cols)
            if o.kind == _abi.KIND_HGCA_INSTANT:
                if "pmra" not in col or "pmdec" not in col:
                    raise OctoError("HGCAInstantaneousObs needs system variables pmra and pmdec (hgca.jl:298-299)")
                blk.update(idx_pmra=col["pmra"], idx_pmdec=col["pmdec"], aux=o.aux)
            blocks.append(blk)
        self.layout_dict = {"n_in": self.n_in, "planets": planets_layout}
        self.block_dicts = blocks
        self.packed = _abi.pack(self.layout_dict, blocks)
        self.total_epochs = sum(len(b["epoch"]) for b in blocks if b["kind"] != _abi.KIND_HGCA_INSTANT)

    def column(self, name):
        return self.input_names.index(name)

    def _build_parameterization(self, system, syscol):
        """θ_t layout (priors in the reference's order, UniformCircular expanded to x, y) and one OctoInputDef per
        kernel input.  `self.priors is None` when the variables were given as bare names (raw-input mode)."""
        scopes = [("", system.var_specs, None)]
        for o in system.observations:
            scopes.append((f"{_normalizename(o.name)}.", o.var_specs, None))
        for p in system.planets:
            scopes.append((f"{p.name}.", p.var_specs, p))
            for o in p.observations:
                scopes.append((f"{p.name}.{_normalizename(o.name)}.", o.var_specs, None))
        if any(spec is None for _, specs, _ in scopes for _, spec in specs):
            self.priors = self.defs = self.theta_names = None
            self.D = None
            return
        priors, tnames, defs = [], [], [None] * self.n_in
        incol = {n: k for k, n in enumerate(self.input_names)}
        for prefix, specs, planet in scopes:
            for name, spec in specs:
                k = incol[prefix + name]
                d = _abi.OctoInputDef()
                if isinstance(spec, Prior):
                    d.op, d.a[0] = _abi.IN_PARAM, len(priors)
                    priors.append(spec); tnames.append(prefix + name)
                elif isinstance(spec, UniformCircular):
                    d.op, d.a[0], d.a[1], d.value = _abi.IN_CIRC, len(priors), len(priors) + 1, spec.domain
                    priors += [Normal(0, 1), Normal(0, 1)]; tnames += [prefix + name + "x", prefix + name + "y"]
                elif isinstance(spec, θ_at_epoch_to_tperi):
                    if planet is None:
                        raise OctoError("θ_at_epoch_to_tperi belongs in a planet's variables")
                    look = lambda v: incol.get(prefix + v, incol.get(v))
                    ti = planet.basis == "ThieleInnesOrbit"
                    needs = ("M", "e", "plx", "A", "B", "F", "G") if ti else ("M", "e", "a", "i", "ω", "Ω")
                    args = [look(spec.theta)] + [look(v) for v in needs]
                    if any(a is None for a in args):
                        raise OctoError(f"{prefix}{name}: θ_at_epoch_to_tperi needs θ, " + ", ".join(needs))
                    if any(a >= k for a in args):
                        raise OctoError(f"{prefix}{name}: define it after the variables it depends on")
                    d.op, d.value = (_abi.IN_TPERI_TI if ti else _abi.IN_TPERI), spec.theta_epoch
                    for q, a in enumerate(args):
                        d.a[q] = a
                elif np.isscalar(spec):
                    d.op, d.value = _abi.IN_CONST, float(spec)
                else:
                    raise OctoError(f"{prefix}{name}: unsupported variable definition {spec!r}")
                defs[k] = d
        self.priors = (_abi.OctoPrior * len(priors))(*[_abi.OctoPrior(p.family, 0, (C.c_double * 4)(*p.p)) for p in priors])
        self.defs = (_abi.OctoInputDef * self.n_in)(*defs)
        self.theta_names = tuple(tnames)
        self.D = len(priors)

    @staticmethod
    def _block(o, ip, obs_cols):
        ep, y1, y2, s1, s2, cor = o._columns()
        g = lambda v: obs_cols.get((id(o), v), -1)
        blk = {"kind": o.kind, "planet": ip, "epoch": ep, "y1": y1, "y2": y2, "s1": s1, "s2": s2, "cor": cor,
               "idx_jitter": g("jitter"), "idx_platescale": g("platescale"),
               "idx_northangle": g("northangle"), "idx_offset": g("offset"), "name": o.name,
               "obs_prior": int(getattr(o, "obs_prior", 0))}
        trend = getattr(o, "trend", None)
        if trend is not None:
            names, basis, const = trend
            blk.update(idx_trend=[g(n) for n in names], trend_basis=basis if len(names) else None, trend_const=const)
        return blk


class _PinnedArray(np.ndarray):
    """ndarray over page-locked memory from octo_alloc_pinned; remembers its address so that the per-call ctypes
    pointer extraction (about 1 µs per array) is skipped.  Views and copies are plain results without the attribute."""
    _octo_ptr = None

    def __array_finalize__(self, obj):
        self._octo_ptr = None


def _ptr(a):
    return getattr(a, "_octo_ptr", None) or a.ctypes.data


class _Pending:
    """An evaluation in flight (octo_*_begin): keeps the buffers alive; `ready()` polls, `wait()` blocks and returns."""

    def __init__(self, model, ticket, bufs, single):
        self._model, self._t, self._bufs, self._single = model, ticket, bufs, single

    def ready(self):
        return self._t is None or self._model._lib.octo_ready(self._t) != 0

    def wait(self):
        if self._t is not None:
            t, self._t = self._t, None
            self._model._check(self._model._lib.octo_wait(t))
        _, a, g = self._bufs
        return (a[0], g[0]) if self._single else (a, g)


class LogDensityModel:
    """The sampler-facing surface for the offloaded terms (src/logdensitymodel.jl:5-24, 252-256),
    backed by libocto_b200.so on one CUDA device."""

    def __init__(self, system, *, device: int = 0, constants=None, lib=None):
        self.spec = system if isinstance(system, ModelSpec) else ModelSpec(system)
        self.system = self.spec.system
        self.input_names = self.spec.input_names
        self.D = self.n_in = self.spec.n_in
        self.packed = self.spec.packed
        self.constants = constants if constants is not None else _abi.default_constants()
        self._lib = lib if lib is not None else _abi.load_library()
        h = C.c_void_p()
        rc = self._lib.octo_create(C.byref(self.constants), C.byref(self.packed.layout), self.packed.blocks,
                                   self.packed.n_blocks, int(device), C.byref(h))
        if rc != 0:
            raise OctoError(f"octo_create failed ({rc}): {self._lib.octo_last_error().decode()}")
        self._h = h
        self._pinned = []
        self.device = int(device)
        if self.spec.priors is not None:        # device-side standard parameterisation (N1)
            self.D = self.spec.D
            self.theta_names = self.spec.theta_names
            self._check(self._lib.octo_set_parameterization(self._h, self.spec.priors, self.spec.D, self.spec.defs))

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.octo_destroy(self._h)
            self._h = None
            for p in self._pinned:
                self._lib.octo_free_pinned(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- evaluation -------------------------------------------------------------------
    @property
    def total_epochs(self) -> int:
        return int(self._lib.octo_total_epochs(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.octo_kernel_launches(self._h))

    def launch_geometry(self, n_chains):
        """(grid.x = chain groups, grid.y = epoch splits, block, epochs per warp) — see octo_launch_geometry."""
        return self.launch_geometry_full(n_chains)[:4]

    def launch_geometry_full(self, n_chains):
        """(grid.x, grid.y, block, epochs per unit, sub-lanes per chain, latency instantiation?)"""
        out = (C.c_int32 * 6)()
        self._lib.octo_launch_geometry(self._h, int(n_chains), C.byref(out))
        return tuple(out)

    def _as_in(self, theta):
        if (type(theta) in (np.ndarray, _PinnedArray) and theta.ndim == 2 and theta.dtype == np.float64
                and theta.flags.f_contiguous and theta.shape[1] == self.n_in):
            return theta, False                                  # fast path: already column-major float64
        th = np.asarray(theta, dtype=np.float64)
        single = th.ndim == 1
        if single:
            th = th[None, :]
        if th.ndim != 2 or th.shape[1] != self.n_in:
            raise ValueError(f"expected (n_chains, {self.n_in}) natural-space inputs, got {th.shape}")
        return np.asfortranarray(th), single     # column-major [n_chains x n_in]: chain index fastest

    def _check(self, rc):
        if rc != 0:
            raise OctoError(f"libocto_b200 error {rc}: {self._lib.octo_last_error().decode()}")

    def pinned_empty(self, shape):
        """Column-major float64 array in page-locked host memory (octo_alloc_pinned): inputs/outputs placed
        here are copied to/from the device without a staging copy.  Freed with the model."""
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        nbytes = int(np.prod(shape)) * 8
        p = self._lib.octo_alloc_pinned(nbytes)
        if not p:
            raise OctoError(f"octo_alloc_pinned failed: {self._lib.octo_last_error().decode()}")
        self._pinned.append(p)
        buf = (C.c_double * (nbytes // 8)).from_address(p)
        out = np.frombuffer(buf, dtype=np.float64).reshape(shape, order="F").view(_PinnedArray)
        out._octo_ptr = int(p)
        return out

    def ln_like(self, theta, out=None):
        """Epoch-summed log-likelihood per chain (value-only kernel K1v)."""
        x, single = self._as_in(theta)
        n = x.shape[0]
        ll = np.empty(n) if out is None else out
        self._check(self._lib.octo_logp(self._h, _ptr(x), n, n, _ptr(ll)))
        return ll[0] if single else ll

    def ln_like_and_gradient(self, theta, out=None):
        """(ll, ∂ll/∂inputs) per chain (fused kernel K1).  out=(ll[n], g[n, n_in] column-major) reuses buffers."""
        x, single = self._as_in(theta)
        n = x.shape[0]
        if out is None:
            ll, g = np.empty(n), np.empty((n, self.n_in), order="F")
        else:
            ll, g = out
            if ll.shape != (n,) or g.shape != (n, self.n_in) or not g.flags.f_contiguous:
                raise ValueError("out must be (ll[n], g[n, n_in]) with g column-major")
        self._check(self._lib.octo_logp_grad(self._h, _ptr(x), n, n, _ptr(ll), _ptr(g)))
        return (ll[0], g[0]) if single else (ll, g)

    def ln_like_and_gradient_begin(self, theta, out=None):
        """Asynchronous half of `ln_like_and_gradient` (C ABI `octo_logp_grad_begin`): enqueues copy-in, kernel and
        copy-out on a stream of its own and returns a handle at once; `handle.wait()` returns (ll, g).  Several handles
        may be in flight: a batch of independent evaluations overlaps the copies of one with the kernel of another, and a
        sampler can do its host-side work (priors, Jacobians) while the GPU evaluates."""
        x, single = self._as_in(theta)
        n = x.shape[0]
        if out is None:
            ll, g = np.empty(n), np.empty((n, self.n_in), order="F")
        else:
            ll, g = out
        t = C.c_void_p()
        self._check(self._lib.octo_logp_grad_begin(self._h, _ptr(x), n, n, _ptr(ll), _ptr(g), C.byref(t)))
        return _Pending(self, t, (x, ll, g), single)

    def ℓπcallback_grad_begin(self, theta_t, out=None):
        """Asynchronous half of `ℓπcallback_grad` (C ABI `octo_logpost_grad_begin`)."""
        th, single = self._as_theta(theta_t)
        n = th.shape[0]
        if out is None:
            lp, g = np.empty(n), np.empty((n, self.D), order="F")
        else:
            lp, g = out
        t = C.c_void_p()
        self._check(self._lib.octo_logpost_grad_begin(self._h, _ptr(th), n, n, _ptr(lp), _ptr(g), C.byref(t)))
        return _Pending(self, t, (th, lp, g), single)

    # -- sampler-facing surface when the model was given priors (src/logdensitymodel.jl:110-146, 169-177, 252-256)
    def _as_theta(self, theta_t):
        if self.spec.priors is None:
            raise OctoError("this model was built from bare variable names: no priors/bijectors to evaluate "
                            "(pass natural-space inputs to ln_like / ln_like_and_gradient)")
        if (type(theta_t) in (np.ndarray, _PinnedArray) and theta_t.ndim == 2 and theta_t.dtype == np.float64
                and theta_t.flags.f_contiguous and theta_t.shape[1] == self.D):
            return theta_t, False                                # fast path: already column-major float64
        th = np.asarray(theta_t, dtype=np.float64)
        single = th.ndim == 1
        if single:
            th = th[None, :]
        if th.ndim != 2 or th.shape[1] != self.D:
            raise ValueError(f"expected (n_chains, {self.D}) unconstrained parameters, got {th.shape}")
        return np.asfortranarray(th), single

    # -- non-epoch terms that stay on the host: UserLikelihood / DirectLLObs (src/variables.jl:332-451),
    #    PlanetOrderPrior / NonCrossingPrior / HillStabilityPrior (src/likelihoods/prior-*.jl) are arbitrary user code in
    #    the reference; here they are host callbacks added to the device posterior
    def add_host_term(self, fn):
        """Register `fn(theta_nat) -> (value[n], grad_nat[n, D])`: a per-chain log-density term of the NATURAL-space
        parameters (rows of `invlink(θ_t)`, the order of `theta_names`) evaluated on the host, e.g. a `UserLikelihood`
        or a planet-order prior; `grad_nat` may be None for terms only used value-only.  `ℓπcallback[_grad]` add it (and
        its gradient through the bijectors) to the device result; with the asynchronous halves of the call it is evaluated
        while the GPU works."""
        if self.spec.priors is None:
            raise OctoError("host terms are added to the device posterior: build the model with priors")
        self._host_terms = getattr(self, "_host_terms", []) + [fn]

    def _host_value_grad(self, th, grad):
        """Σ host terms at θ_t (and ∇ w.r.t. θ_t: the natural-space gradient times d invlink / d θ_t, by central
        differences of the monotone scalar bijectors — they are elementwise)."""
        nat = self.invlink(th)
        val = np.zeros(th.shape[0]); g = np.zeros(th.shape) if grad else None
        for fn in self._host_terms:
            v, gn = fn(nat)
            val += np.asarray(v, dtype=np.float64)
            if grad:
                if gn is None:
                    raise OctoError("a host term without gradient cannot be used by ℓπcallback_grad")
                g += np.asarray(gn, dtype=np.float64)
        if grad:
            h = 1e-6 * np.maximum(1.0, np.abs(th))
            dxdy = (self.invlink(th + h) - self.invlink(th - h)) / (2 * h)
            g = g * dxdy
        return val, g

    def ℓπcallback(self, theta_t):
        """log-posterior of the unconstrained vector(s) θ_t: priors + bijector Jacobians + likelihood, on device."""
        th, single = self._as_theta(theta_t)
        n = th.shape[0]
        lp = np.empty(n)
        self._check(self._lib.octo_logpost_grad(self._h, th.ctypes.data, n, n, lp.ctypes.data, None))
        if getattr(self, "_host_terms", None):
            lp = lp + self._host_value_grad(th, False)[0]
        return lp[0] if single else lp

    def ℓπcallback_grad(self, theta_t, out=None):
        """(ℓπ, ∇ℓπ) of the unconstrained vector(s) θ_t, one fused launch.  out=(lp[n], g[n, D] column-major) reuses
        buffers; with buffers from pinned_empty() the copies need no staging."""
        th, single = self._as_theta(theta_t)
        n = th.shape[0]
        if out is None:
            lp, g = np.empty(n), np.empty((n, self.D), order="F")
        else:
            lp, g = out
            if lp.shape != (n,) or g.shape != (n, self.D) or not g.flags.f_contiguous:
                raise ValueError("out must be (lp[n], g[n, D]) with g column-major")
        if getattr(self, "_host_terms", None):
            # device evaluation in flight while the host evaluates its own terms (octo_logpost_grad_begin / octo_wait)
            t = C.c_void_p()
            self._check(self._lib.octo_logpost_grad_begin(self._h, _ptr(th), n, n, _ptr(lp), _ptr(g), C.byref(t)))
            try:
                hv, hg = self._host_value_grad(th, True)
            finally:
                self._check(self._lib.octo_wait(t))
            ok = np.isfinite(lp)
            lp += np.where(ok, hv, 0.0); g += np.where(ok[:, None], hg, 0.0)
            return (lp[0], g[0]) if single else (lp, g)
        self._check(self._lib.octo_logpost_grad(self._h, _ptr(th), n, n, _ptr(lp), _ptr(g)))
        return (lp[0], g[0]) if single else (lp, g)

    def invlink(self, theta_t):
        th, single = self._as_theta(theta_t)
        n = th.shape[0]
        out = np.empty((n, self.D), order="F")
        self._check(self._lib.octo_invlink(self._h, th.ctypes.data, n, n, out.ctypes.data))
        return out[0] if single else out

    logpost = ℓπcallback
    logpost_and_gradient = ℓπcallback_grad

    # -- batched value-only consumers -------------------------------------------------
    def sample_priors(self, rng, n=1):
        """n draws from the priors, natural space, shape (n, D) (model.sample_priors, src/variables.jl:1385-1440)."""
        if self.spec.priors is None:
            raise OctoError("sample_priors needs a model built with priors")
        out = np.empty((n, self.D))
        for j, pr in enumerate(self.spec.priors):
            p = pr.p
            if pr.family == _abi.PRIOR_NORMAL:
                out[:, j] = rng.normal(p[0], p[1], n)
            elif pr.family == _abi.PRIOR_UNIFORM:
                out[:, j] = rng.uniform(p[0], p[1], n)
            elif pr.family == _abi.PRIOR_LOGUNIFORM:
                out[:, j] = np.exp(rng.uniform(np.log(p[0]), np.log(p[1]), n))
            elif pr.family == _abi.PRIOR_SINE:
                out[:, j] = np.arccos(1.0 - 2.0 * rng.uniform(0, 1, n))          # quantile, src/distributions.jl:40
            else:                                                                # truncated normal: inverse-cdf
                from scipy.stats import truncnorm
                a, b = (p[2] - p[0]) / p[1], (p[3] - p[0]) / p[1]
                out[:, j] = truncnorm.rvs(a, b, loc=p[0], scale=p[1], size=n, random_state=rng)
        return out

    def link(self, theta_nat):
        """Natural -> unconstrained parameters (Bijectors.link per prior; inverse of `invlink`)."""
        x = np.atleast_2d(np.asarray(theta_nat, dtype=np.float64))
        y = np.empty_like(x)
        for j, pr in enumerate(self.spec.priors):
            lo, hi = -np.inf, np.inf
            if pr.family in (_abi.PRIOR_UNIFORM, _abi.PRIOR_LOGUNIFORM):
                lo, hi = pr.p[0], pr.p[1]
            elif pr.family == _abi.PRIOR_SINE:
                lo, hi = np.finfo(float).eps, np.pi - np.finfo(float).eps
            elif pr.family == _abi.PRIOR_TRUNCNORMAL:
                lo, hi = pr.p[2], pr.p[3]
            if np.isfinite(lo) and np.isfinite(hi):
                u = (x[:, j] - lo) / (hi - lo)
                y[:, j] = np.log(u) - np.log1p(-u)
            elif np.isfinite(lo):
                y[:, j] = np.log(x[:, j] - lo)
            elif np.isfinite(hi):
                y[:, j] = np.log(hi - x[:, j])
            else:
                y[:, j] = x[:, j]
        return y if np.ndim(theta_nat) == 2 else y[0]

    def guess_starting_position(self, rng, N=500_000, batch=65536):
        """Sample IID from the prior N times and return the highest-posterior draw and its log-posterior
        (src/initialization.jl:14-66) — the N ℓπ evaluations run as value-only device batches (K1v)."""
        best, best_lp = None, -np.inf
        done = 0
        while done < N:
            m = min(batch, N - done)
            params = self.sample_priors(rng, m)
            lp = self.ℓπcallback(self.link(params))
            k = int(np.argmax(lp))
            if lp[k] > best_lp:
                best, best_lp = params[k].copy(), float(lp[k])
            done += m
        return best, best_lp

    # LogDensityProblems-style names (src/logdensitymodel.jl:252-256)
    def logdensity(self, theta):
        return self.ℓπcallback(theta) if self.spec.priors is not None else self.ln_like(theta)

    def logdensity_and_gradient(self, theta):
        return self.ℓπcallback_grad(theta) if self.spec.priors is not None else self.ln_like_and_gradient(theta)

    def dimension(self):
        return self.D

    def __call__(self, theta):           # Pigeons calls the model (ext/OctofitterPigeonsExt:10-12)
        return self.logdensity(theta)

    def enqueue_device(self, d_in, n_chains, ld, d_ll, d_g, stream=0):
        """Asynchronous launch on device-resident buffers (raw pointers, cudaStream_t handle)."""
        self._check(self._lib.octo_logp_grad_device(self._h, int(d_in), int(n_chains), int(ld), int(d_ll),
                                                    int(d_g) if d_g else None, int(stream) if stream else None))

    def ln_like_of_theta(self, theta_t):
        """ln_like(system, arr2nt(invlink(θ_t))) alone (UnitLengthPrior terms included), -Inf where not finite: what
        `octofit_rejection` evaluates per prior draw (src/sampling.jl:261-270).  One fused launch, value only."""
        th, single = self._as_theta(theta_t)
        n = th.shape[0]
        ll = np.empty(n)
        self._check(self._lib.octo_loglike_theta(self._h, th.ctypes.data, n, n, ll.ctypes.data))
        return ll[0] if single else ll

    def pointwise_like(self, theta, batch=4096):
        """(LL_out[n_samples, n_epochs], epochs) as `pointwise_like(model, chain)` builds them
        (src/cross-validation.jl:6-49): column e is ln_like of the system reduced to its e-th epoch alone, epochs
        ordered like `generate_system_per_epoch` (:453-497) — system-level tables first, then the planets' tables.
        `theta` are natural-space kernel inputs, one row per posterior sample."""
        x, single = self._as_in(theta)
        n, E = x.shape[0], self.total_epochs
        out = np.empty((n, E), order="F")
        for lo in range(0, n, batch):
            xb = np.asfortranarray(x[lo:lo + batch])
            ob = np.empty((xb.shape[0], E), order="F")
            self._check(self._lib.octo_logp_pointwise(self._h, xb.ctypes.data, xb.shape[0], xb.shape[0], ob.ctypes.data, xb.shape[0]))
            out[lo:lo + batch] = ob
        # kernel order = planet tables then system tables (the order ln_like sums them); reference order = system first
        blocks, start, order, epochs = self.spec.block_dicts, 0, [], []
        spans = []
        for b in blocks:
            if b["kind"] == _abi.KIND_HGCA_INSTANT:
                continue            # not part of the epoch list (a one-row subset is 0/0 in the reference as well)
            spans.append((b["planet"] < 0, start, len(b["epoch"]), b["epoch"]))
            start += len(b["epoch"])
        for system_level in (True, False):
            for is_sys, s0, cnt, ep in spans:
                if is_sys == system_level:
                    order.extend(range(s0, s0 + cnt)); epochs.extend(ep)
        out = out[:, order]
        return (out[0] if single else out), np.asarray(epochs, dtype=np.float64)

    def logpost_workspace_bytes(self, n_chains):
        """Scratch the device entry point needs for n_chains (0 when the parameterisation is fused into the kernel)."""
        return int(self._lib.octo_logpost_workspace(self._h, int(n_chains)))

    def enqueue_logpost_device(self, d_theta_t, n_chains, ld, d_lp, d_g_t, d_work=0, stream=0):
        """Asynchronous ℓπ(θ_t), ∇ℓπ(θ_t) on device-resident buffers (raw pointers, cudaStream_t handle)."""
        self._check(self._lib.octo_logpost_grad_device(self._h, int(d_theta_t), int(n_chains), int(ld), int(d_lp),
                                                       int(d_g_t) if d_g_t else None, int(d_work) if d_work else None,
                                                       int(stream) if stream else None))

