"""Independent 40+-digit restatement of the hot path in mpmath.  TEST INFRASTRUCTURE ONLY.

Deliberately written along a different route from oracle/octo_oracle.hpp so the two can pin
each other: Kepler's equation by Newton iteration at working precision (not Markley), positions
through Thiele-Innes constants and (cosE - e, sqrt(1-e^2) sinE) (the algebra of
src/parameterizations.jl:34-37, 337-353), radial velocity through cos/sin of the true anomaly
from E, gradients by mpmath's high-order numerical differentiation at 60 digits.

Used only by tests/golden/make_golden.py to produce the committed golden vectors.
"""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 60


def _consts(c):
    return {k: mp.mpf(v) for k, v in c.items()}


def kepler_E(M, e):
    M = M - 2 * mp.pi * mp.nint(M / (2 * mp.pi))
    E = M + e * mp.sin(M) if e < mp.mpf("0.8") else mp.pi * mp.sign(M) if M != 0 else mp.mpf(0)
    for _ in range(200):
        dE = (E - e * mp.sin(E) - M) / (1 - e * mp.cos(E))
        E -= dE
        if abs(dE) < mp.mpf(10) ** (-mp.mp.dps + 5):
            break
    return E


def ti_semimajor(A, B, F, G, plx):
    """a [AU] of a Thiele-Innes orbit (src/parameterizations.jl:14-18)."""
    u = (A * A + B * B + F * F + G * G) / 2
    v = A * G - B * F
    return mp.sqrt(u + mp.sqrt((u + v) * (u - v))) / plx


def planet_state(c, el, t):
    """(ra [mas], dec [mas], rv [m/s]) of the planet relative to the star at epoch t."""
    if "A" in el:        # ThieleInnesOrbit: ra = xB + yG, dec = xA + yF [mas] (src/parameterizations.jl:346-353)
        e, tp, M, plx = el["e"], el["tp"], el["M"], el["plx"]
        a = ti_semimajor(el["A"], el["B"], el["F"], el["G"], plx)
        P_days = mp.sqrt(a ** 3 / M) * c["kepler_year_days"]
        E = kepler_E(2 * mp.pi * (t - tp) / P_days, e)
        X, Y = mp.cos(E) - e, mp.sqrt(1 - e * e) * mp.sin(E)
        return X * el["B"] + Y * el["G"], X * el["A"] + Y * el["F"], None
    a, e, i, w, W, tp, M, plx = (el[k] for k in ("a", "e", "i", "w", "W", "tp", "M", "plx"))
    P_days = mp.sqrt(a ** 3 / M) * c["kepler_year_days"]
    E = kepler_E(2 * mp.pi * (t - tp) / P_days, e)
    s = mp.sqrt(1 - e * e)
    X, Y = mp.cos(E) - e, s * mp.sin(E)
    A = mp.cos(W) * mp.cos(w) - mp.sin(W) * mp.sin(w) * mp.cos(i)
    B = mp.sin(W) * mp.cos(w) + mp.cos(W) * mp.sin(w) * mp.cos(i)
    F = -mp.cos(W) * mp.sin(w) - mp.sin(W) * mp.cos(w) * mp.cos(i)
    G = -mp.sin(W) * mp.sin(w) + mp.cos(W) * mp.cos(w) * mp.cos(i)
    dist = 1000 / plx * c["pc2au"]
    c2a = c["rad2as"] * 1000 / dist
    ra = a * c2a * (X * B + Y * G)
    dec = a * c2a * (X * A + Y * F)
    D = 1 - e * mp.cos(E)
    cosnu, sinnu = X / D, Y / D
    P_yr = P_days / c["year2day"]
    K = (2 * mp.pi * a / P_yr) / s * c["au2m"] * c["sec2year"] * mp.sin(i)
    rv = K * (cosnu * mp.cos(w) - sinnu * mp.sin(w) + e * mp.cos(w))
    return ra, dec, rv


def _mvn2(s1, s2, cor, r1, r2):
    det = s1 ** 2 * s2 ** 2 * (1 - cor ** 2)
    maha = (r1 ** 2 / s1 ** 2 - 2 * cor * r1 * r2 / (s1 * s2) + r2 ** 2 / s2 ** 2) / (1 - cor ** 2)
    return -mp.log(2 * mp.pi) - mp.log(det) / 2 - maha / 2


def ln_like(consts, layout, blocks, x):
    """Scalar log-likelihood for one chain.  layout/blocks are the dictionaries of
    octofitter.jl_b200._abi.pack; x is a sequence of mp numbers (natural-space inputs)."""
    c = _consts(consts)
    x = [mp.mpf(v) for v in x]
    els = []
    for p in layout["planets"]:
        if p.get("basis", 0) == 1:
            el = {k: x[p[k]] for k in ("A", "B", "F", "G", "e", "tp", "M", "plx")}
            el["a"] = ti_semimajor(el["A"], el["B"], el["F"], el["G"], el["plx"])
        else:
            el = {k: x[p[k]] for k in ("a", "e", "i", "w", "W", "tp", "M", "plx")}
        el["mu"] = x[p["mass"]] * c["mjup2msol"] / el["M"] if p.get("mass", -1) >= 0 else None
        els.append(el)
    ll = mp.mpf(0)
    two_pi = 2 * mp.pi
    for b in blocks:
        kind = b["kind"]
        jit = x[b["idx_jitter"]] if b.get("idx_jitter", -1) >= 0 else mp.mpf(0)
        off = x[b["idx_offset"]] if b.get("idx_offset", -1) >= 0 else mp.mpf(0)
        n = len(b["epoch"])
        if kind in (0, 1):
            ps = x[b["idx_platescale"]] if b.get("idx_platescale", -1) >= 0 else mp.mpf(1)
            na = x[b["idx_northangle"]] if b.get("idx_northangle", -1) >= 0 else mp.mpf(0)
            ip = b["planet"]
            for k in range(n):
                t = mp.mpf(float(b["epoch"][k]))
                ra, dec, _ = planet_state(c, els[ip], t)
                for j, el in enumerate(els):
                    if j != ip and el["a"] < els[ip]["a"] and el["mu"] is not None:
                        rj, dj, _ = planet_state(c, el, t)
                        ra += el["mu"] * rj          # minus (star reflex = -mu * planet offset)
                        dec += el["mu"] * dj
                y1, y2 = mp.mpf(float(b["y1"][k])), mp.mpf(float(b["y2"][k]))
                s1, s2 = mp.mpf(float(b["s1"][k])), mp.mpf(float(b["s2"][k]))
                cor = mp.mpf(float(b["cor"][k])) if b.get("cor") is not None else mp.mpf(0)
                if kind == 1:
                    rho = mp.hypot(ra, dec)
                    pa = mp.atan2(ra, dec)
                    d = (y1 + na) - pa
                    d = d - two_pi * mp.nint(d / two_pi)        # wrapped to [-pi, pi]
                    r1, r2 = d, y2 * ps - rho
                else:
                    # data rotated by -northangle (East through North) and scaled by platescale
                    ra_d = ps * (y1 * mp.cos(na) + y2 * mp.sin(na))
                    dec_d = ps * (y2 * mp.cos(na) - y1 * mp.sin(na))
                    r1, r2 = ra_d - ra, dec_d - dec
                s1, s2 = mp.sqrt(s1 ** 2 + jit ** 2), mp.sqrt(s2 ** 2 + jit ** 2)
                ll += _mvn2(s1, s2, cor, r1, r2)
            if b.get("obs_prior", 0):
                # ObsPriorAstromONeil2019 (src/likelihoods/prior-observable.jl:78-137)
                el = els[ip]
                P_days = mp.sqrt(el["a"] ** 3 / el["M"]) * c["kepler_year_days"]
                jac = mp.mpf(0)
                for k in range(n):
                    t = mp.mpf(float(b["epoch"][k]))
                    E = kepler_E(2 * mp.pi * (t - el["tp"]) / P_days, el["e"])
                    Mm = E - el["e"] * mp.sin(E)
                    jac += abs(3 * Mm * (el["e"] + mp.cos(E)) + 2 * (-2 + el["e"] ** 2 + el["e"] * mp.cos(E)) * mp.sin(E))
                jac *= mp.cbrt(P_days / mp.mpf("365.25")) / mp.sqrt(1 - el["e"] ** 2)
                ll += 2 * mp.log(jac)
        elif kind in (2, 3):
            A = Bq = Cq = mp.mpf(0)
            for k in range(n):
                t = mp.mpf(float(b["epoch"][k]))
                m = mp.mpf(0) if kind == 3 else off
                for el in els:
                    m += -el["mu"] * planet_state(c, el, t)[2]
                r = mp.mpf(float(b["y1"][k])) - m
                var = mp.mpf(float(b["s1"][k])) ** 2 + jit ** 2
                if kind == 3:
                    A += 1 / var; Bq -= 2 * r / var; Cq += r * r / var
                    ll -= mp.log(two_pi * var)
                else:
                    ll += -(mp.log(two_pi) + mp.log(var) + r * r / var) / 2
            if kind == 3:
                ll -= -Bq ** 2 / (4 * A) + Cq + mp.log(A)
        elif kind == 5:
            # HGCAInstantaneousObs (src/likelihoods/hgca.jl:155-417): proper motions as numerical time derivatives
            # of the star's reflex position (independent of the closed-form velocity formula the oracle uses)
            pm_sys = (x[b["idx_pmra"]], x[b["idx_pmdec"]])
            P = len(els)
            def star_pos(t, axis):
                tot = mp.mpf(0)
                for el in els:
                    st = planet_state(c, el, t)
                    tot += -el["mu"] * st[axis]
                return tot
            pos, pm, ep = {}, {}, {}
            for inst in (0, 1):
                for axis in (0, 1):
                    ts = [mp.mpf(float(b["epoch"][k])) for k in range(n) if int(b["y1"][k]) == 2 * inst + axis]
                    # the reference divides the sum over planets AND rows by (planets x rows) (hgca.jl:247-290)
                    pos[inst, axis] = sum(star_pos(t, axis) for t in ts) / (P * len(ts))
                    pm[inst, axis] = sum(mp.diff(lambda tt: star_pos(tt, axis), t) * mp.mpf("365.25") for t in ts) / (P * len(ts)) + pm_sys[axis]
                    ep[inst, axis] = sum(ts) / len(ts)
            hg = [(pos[1, ax] - pos[0, ax]) / (ep[1, ax] - ep[0, ax]) * mp.mpf("365.25") + pm_sys[ax] for ax in (0, 1)]
            q = [mp.mpf(float(v)) for v in b["aux"]]
            ll += _mvn2(q[2], q[3], q[4], pm[0, 0] - q[0], pm[0, 1] - q[1])
            ll += _mvn2(q[7], q[8], q[9], hg[0] - q[5], hg[1] - q[6])
            ll += _mvn2(q[12], q[13], q[14], pm[1, 0] - q[10], pm[1, 1] - q[11])
        elif kind == 4:
            ip = b["planet"]
            for k in range(n):
                t = mp.mpf(float(b["epoch"][k]))
                m = off + planet_state(c, els[ip], t)[2]
                for j, el in enumerate(els):
                    if j != ip and el["a"] < els[ip]["a"] and el["mu"] is not None:
                        m += -el["mu"] * planet_state(c, el, t)[2]
                r = mp.mpf(float(b["y1"][k])) - m
                var = mp.mpf(float(b["s1"][k])) ** 2 + jit ** 2
                ll += -(mp.log(two_pi) + mp.log(var) + r * r / var) / 2
        else:
            raise ValueError(kind)
    return ll


def ln_like_grad(consts, layout, blocks, x):
    x = [mp.mpf(v) for v in x]
    f0 = ln_like(consts, layout, blocks, x)
    g = []
    for k in range(len(x)):
        def f(v, k=k):
            y = list(x); y[k] = v
            return ln_like(consts, layout, blocks, y)
        h = max(abs(x[k]), mp.mpf(1)) * mp.mpf(10) ** (-18)
        g.append((f(x[k] + h) - f(x[k] - h)) / (2 * h))      # central difference, error O(h^2) ~ 1e-36
    return f0, g


# ---------------------------------------------------------------------------------------------------------
# Standard parameterisation (priors + bijectors + derived inputs): independent high-precision restatement
# of ℓπcallback for the prior families of octofitter.jl_b200.model.
# priors: list of (family, p0, p1, p2, p3); defs: list of (op, [a0..a6], value) as in include/octo_b200.h
# ---------------------------------------------------------------------------------------------------------
def _bounds(fam, p):
    if fam in (1, 2):
        return mp.mpf(p[0]), mp.mpf(p[1])
    if fam == 3:
        eps = mp.mpf(2) ** -52
        return eps, mp.mpf(float(mp.pi)) - eps       # π as the Float64 the reference uses
    if fam == 4:
        return (mp.mpf(p[2]) if p[2] != float("-inf") else None), (mp.mpf(p[3]) if p[3] != float("inf") else None)
    return None, None


def _invlink(fam, p, y):
    lo, hi = _bounds(fam, p)
    if lo is not None and hi is not None:
        return lo + (hi - lo) / (1 + mp.e ** (-y))
    if lo is not None:
        return lo + mp.e ** y
    if hi is not None:
        return hi - mp.e ** y
    return y


def _logpdf_with_trans(fam, p, x):
    lo, hi = _bounds(fam, p)
    if fam == 0:
        lp = mp.log(mp.npdf(x, p[0], p[1]))
    elif fam == 1:
        lp = -mp.log(mp.mpf(p[1]) - mp.mpf(p[0]))
    elif fam == 2:
        lp = -mp.log(x) - mp.log(mp.log(mp.mpf(p[1]) / mp.mpf(p[0])))
    elif fam == 3:
        lp = mp.log(mp.sin(x) / 2)
    elif fam == 4:
        a = mp.ncdf(lo, p[0], p[1]) if lo is not None else mp.mpf(0)
        b = mp.ncdf(hi, p[0], p[1]) if hi is not None else mp.mpf(1)
        lp = mp.log(mp.npdf(x, p[0], p[1])) - mp.log(b - a)
    if lo is not None and hi is not None:
        lp += mp.log((x - lo) * (hi - x) / (hi - lo))
    elif lo is not None:
        lp += mp.log(x - lo)
    elif hi is not None:
        lp += mp.log(hi - x)
    return lp


def _tperi(c, theta, t_ref, M, e, a, i, w, W):
    """tp such that the position angle at t_ref is theta: solve the geometry independently of the reference's
    matrix route — rotate the sky-plane direction back into the orbital plane."""
    # sky-plane unit direction: (x, y) = (cosθ, sinθ) in the reference's (dec-like, ra-like) convention of
    # [A F; B G] (x/r, y/r) = (cosθ, sinθ)
    A = mp.cos(W) * mp.cos(w) - mp.sin(W) * mp.sin(w) * mp.cos(i)
    B = mp.sin(W) * mp.cos(w) + mp.cos(W) * mp.sin(w) * mp.cos(i)
    F = -mp.cos(W) * mp.sin(w) - mp.sin(W) * mp.cos(w) * mp.cos(i)
    G = -mp.sin(W) * mp.sin(w) + mp.cos(W) * mp.cos(w) * mp.cos(i)
    sol = mp.lu_solve(mp.matrix([[A, F], [B, G]]), mp.matrix([mp.cos(theta), mp.sin(theta)]))
    nu = mp.atan2(sol[1], sol[0])
    # eccentric anomaly from the true anomaly, then Kepler's equation forwards
    E = 2 * mp.atan(mp.sqrt((1 - e) / (1 + e)) * mp.tan(nu / 2))
    MA = E - e * mp.sin(E)
    MA = MA % (2 * mp.pi)                      # the reference's atan(...)+π form lands in [0, 2π)
    P_days = mp.sqrt(a ** 3 / M) * c["kepler_year_days"]
    return t_ref - MA * P_days / (2 * mp.pi)


def _tperi_ti(c, theta, t_ref, M, e, plx, A, B, F, G):
    """The same for a Thiele-Innes orbit (src/parameterizations.jl:9-19): the constants are given, a follows from them."""
    sol = mp.lu_solve(mp.matrix([[A, F], [B, G]]), mp.matrix([mp.cos(theta), mp.sin(theta)]))
    nu = mp.atan2(sol[1], sol[0])
    E = 2 * mp.atan(mp.sqrt((1 - e) / (1 + e)) * mp.tan(nu / 2))
    MA = (E - e * mp.sin(E)) % (2 * mp.pi)
    a = ti_semimajor(A, B, F, G, plx)
    P_days = mp.sqrt(a ** 3 / M) * c["kepler_year_days"]
    return t_ref - MA * P_days / (2 * mp.pi)


def logpost(consts, layout, blocks, priors, defs, theta_t):
    c = _consts(consts)
    y = [mp.mpf(v) for v in theta_t]
    th = [_invlink(f, p, yy) for (f, *p), yy in zip(priors, y)]
    lp = sum(_logpdf_with_trans(f, p, x) for (f, *p), x in zip(priors, th))
    inp, extra = [], mp.mpf(0)
    for op, a, val in defs:
        if op == 0:
            inp.append(th[a[0]])
        elif op == 1:
            inp.append(mp.mpf(val))
        elif op == 2:
            x, yv = th[a[0]], th[a[1]]
            inp.append(mp.atan2(yv, x) / (2 * mp.pi) * mp.mpf(val))
            r = mp.sqrt(x * x + yv * yv)
            extra += mp.log(mp.npdf(mp.log(r), 0, mp.mpf("0.1")) / r)          # LogNormal(0, 0.1) density at r
        elif op == 3:
            inp.append(_tperi(c, inp[a[0]], mp.mpf(val), *[inp[k] for k in a[1:7]]))
        elif op == 4:
            inp.append(_tperi_ti(c, inp[a[0]], mp.mpf(val), *[inp[k] for k in a[1:8]]))
    return lp + extra + ln_like(consts, layout, blocks, inp)


def logpost_grad(consts, layout, blocks, priors, defs, theta_t):
    x = [mp.mpf(v) for v in theta_t]
    f0 = logpost(consts, layout, blocks, priors, defs, x)
    g = []
    for k in range(len(x)):
        h = mp.mpf(10) ** (-18)
        xp = list(x); xp[k] += h
        xm = list(x); xm[k] -= h
        g.append((logpost(consts, layout, blocks, priors, defs, xp) - logpost(consts, layout, blocks, priors, defs, xm)) / (2 * h))
    return f0, g
