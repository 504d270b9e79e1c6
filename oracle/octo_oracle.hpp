// octo_oracle.hpp — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
//
// PARITY STATUS: "parity partially pinned".  The reference is Julia and cannot run in the
// build container or on the GPU box, and the arithmetic core (PlanetOrbits.jl, compat
// "0.11.1", Project.toml:112) is not vendored in /root/reference.  What IS pinned:
//   * orbit geometry + Kepler solve (rows a3-a6 below) reproduce the reference's own
//     8-epoch astrometry fixture (test/integration-tests.jl:8-15) to <= 2e-12 mas once the
//     generating orbit is identified (tests/golden/make_golden.py, tests/test_oracle_golden.py);
//   * every formula is cross-checked against an independent 40-digit mpmath restatement
//     (oracle/mp_reference.py -> tests/golden/*.json).
// What is NOT pinned: the current values of PlanetOrbits' constants (injected through
// OctoConstants for that reason) and the Distributions.jl normalisation, for which no
// reference test holds a number.
//
// Nothing in the product path (octofitter.jl_b200/, libocto_b200.so) may include, link or
// call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker / CPU baseline.
//
// Reference rows restated here:
//   a1  make_ln_like generated body             src/likelihoods/system.jl:21-242
//   a2  _kepsolve_all!                          src/likelihoods/system.jl:244-269
//   a3  KepOrbit / Visual{KepOrbit} ctor        PlanetOrbits.jl 0.11 (call site system.jl:117;
//                                               same algebra in src/parameterizations.jl:62-67,215-216)
//   a4  orbitsolve                              PlanetOrbits.jl (call sites system.jl:165,259;
//                                               twin at src/parameterizations.jl:337-345)
//   a5  kepler_solver(MA, e, Markley)           PlanetOrbits.jl ("tweaked copy of AstroLib's",
//                                               docs/src/kepler.md:15-19; Markley 1995 eqs 5-28)
//   a6  raoff/decoff                            PlanetOrbits.jl (restated src/parameterizations.jl:244-245)
//   a7  radvel                                  PlanetOrbits.jl (call sites rv-absolute.jl:150-153)
//   a8  simulate!(::PlanetRelAstromObs)         src/likelihoods/relative-astrometry.jl:104-142
//   a9  ln_like(::PlanetRelAstromObs)           src/likelihoods/relative-astrometry.jl:166-253
//   a10 StarAbsoluteRVObs                       OctofitterRadialVelocity/src/rv-absolute.jl:135-204
//   a11 MarginalizedStarAbsoluteRVObs           OctofitterRadialVelocity/src/rv-absolute-margin.jl:106-185
//   a12 PlanetRelativeRVObs                     OctofitterRadialVelocity/src/rv-relative.jl:121-211
//   a13 ForwardDiff pass                        src/logdensitymodel.jl:43-45,169-177 — forward-mode
//       dual numbers with chunk = n_in; kepler_solver on Duals uses PlanetOrbits' ForwardDiff
//       extension (implicit differentiation: dE = (dM + sinE de)/(1 - e cosE)).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include "../include/octo_b200.h"

namespace octo_oracle {

// ---------------------------------------------------------------------------------------
// Forward-mode dual number with a compile-time number of partials (ForwardDiff.Dual, a13).
// ---------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N > 0 ? N : 1];
    Dual() : v(0.0) { for (int k = 0; k < N; ++k) d[k] = 0.0; }
    Dual(double x) : v(x) { for (int k = 0; k < N; ++k) d[k] = 0.0; }
};

template <int N> inline Dual<N> seed(double x, int k) { Dual<N> r(x); if (k >= 0 && k < N) r.d[k] = 1.0; return r; }

template <int N> inline Dual<N> unary(const Dual<N>& a, double f, double df) {
    Dual<N> r; r.v = f; for (int k = 0; k < N; ++k) { r.d[k] = df * a.d[k]; } return r;
}
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int k = 0; k < N; ++k) { r.d[k] = a.d[k] + b.d[k]; } return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int k = 0; k < N; ++k) { r.d[k] = a.d[k] - b.d[k]; } return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int k = 0; k < N; ++k) { r.d[k] = -a.d[k]; } return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int k = 0; k < N; ++k) { r.d[k] = a.d[k] * b.v + a.v * b.d[k]; } return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v / b.v; const double ib = 1.0 / b.v;
    for (int k = 0; k < N; ++k) { r.d[k] = (a.d[k] - r.v * b.d[k]) * ib; } return r;
}
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return b + a; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { return (-b) + a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int k = 0; k < N; ++k) { r.d[k] = a.d[k] * b; } return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v / b; for (int k = 0; k < N; ++k) r.d[k] = a.d[k] / b; return r; }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a = a + b; return a; }
template <int N> inline Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a = a - b; return a; }

inline double value(double x) { return x; }
template <int N> inline double value(const Dual<N>& x) { return x.v; }

using std::sin; using std::cos; using std::tan; using std::atan; using std::atan2; using std::sqrt;
using std::log; using std::hypot; using std::fabs; using std::cbrt;
template <int N> inline Dual<N> sin(const Dual<N>& a) { return unary(a, std::sin(a.v), std::cos(a.v)); }
template <int N> inline Dual<N> cos(const Dual<N>& a) { return unary(a, std::cos(a.v), -std::sin(a.v)); }
template <int N> inline Dual<N> tan(const Dual<N>& a) { double t = std::tan(a.v); return unary(a, t, 1.0 + t * t); }
template <int N> inline Dual<N> atan(const Dual<N>& a) { return unary(a, std::atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) { double s = std::sqrt(a.v); return unary(a, s, 0.5 / s); }
template <int N> inline Dual<N> log(const Dual<N>& a) { return unary(a, std::log(a.v), 1.0 / a.v); }
template <int N> inline Dual<N> cbrt(const Dual<N>& a) { double r = std::cbrt(a.v); return unary(a, r, r / (3.0 * a.v)); }
template <int N> inline Dual<N> atan2(const Dual<N>& y, const Dual<N>& x) {
    Dual<N> r; r.v = std::atan2(y.v, x.v); const double h = x.v * x.v + y.v * y.v;
    for (int k = 0; k < N; ++k) { r.d[k] = (x.v * y.d[k] - y.v * x.d[k]) / h; } return r;
}
template <int N> inline Dual<N> hypot(const Dual<N>& x, const Dual<N>& y) {
    Dual<N> r; r.v = std::hypot(x.v, y.v);
    for (int k = 0; k < N; ++k) { r.d[k] = (x.v * x.d[k] + y.v * y.d[k]) / r.v; } return r;
}
template <int N> inline Dual<N> hypot(double x, const Dual<N>& y) { return hypot(Dual<N>(x), y); }

// ---------------------------------------------------------------------------------------
// a5: rem2pi(x, RoundNearest) + Markley (1995) non-iterative solver for e < 1.
// Julia's rem2pi reduces exactly (double-double 2π, Payne-Hanek beyond); binary128 with a
// 159-bit 2π (three doubles) keeps the error below |k| * 2^-111.
// ---------------------------------------------------------------------------------------
inline double rem2pi_nearest(double x) {
    const __float128 TWO_PI_Q = (__float128)0x1.921fb54442d18p+2 + (__float128)0x1.1a62633145c07p-52 +
                                (__float128)(-0x1.f1976b7ed8fbcp-108);
    const __float128 PI_Q = TWO_PI_Q / 2;
    double k = std::nearbyint(x / 6.283185307179586);
    __float128 r = (__float128)x - (__float128)k * TWO_PI_Q;
    if (r > PI_Q) r -= TWO_PI_Q;
    else if (r < -PI_Q) r += TWO_PI_Q;
    return (double)r;
}

inline double kepler_markley(double MA, double e) {
    const double M = rem2pi_nearest(MA);
    if (M == 0.0 || e == 0.0) return M;
    const double pi = 3.14159265358979323846;
    const double pi2 = pi * pi;
    const double alpha = (3.0 * pi2 + 8.0 * (pi2 - pi * std::fabs(M)) / (5.0 * (1.0 + e))) / (pi2 - 6.0);  // eq 20
    const double d = 3.0 * (1.0 - e) + alpha * e;                                                          // eq 5
    const double q = 2.0 * alpha * d * (1.0 - e) - M * M;                                                  // eq 9
    const double r = 3.0 * alpha * d * (d - 1.0 + e) * M + M * M * M;                                      // eq 10
    const double t = std::fabs(r) + std::sqrt(q * q * q + r * r);
    const double w = std::cbrt(t * t);                                                                     // eq 14
    const double E1 = (2.0 * r * w / (w * w + w * q + q * q) + M) / d;                                     // eq 15
    const double f2 = e * std::sin(E1), f3 = e * std::cos(E1);                                             // eqs 26, 27
    const double f0 = E1 - f2 - M;                                                                         // eq 21
    const double f1 = 1.0 - f3;                                                                            // eq 25
    const double d3 = -f0 / (f1 - f0 * f2 / (2.0 * f1));                                                   // eq 22
    const double d4 = -f0 / (f1 + f2 * d3 / 2.0 + d3 * d3 * f3 / 6.0);                                     // eq 23
    const double d5 = -f0 / (f1 + d4 * f2 / 2.0 + d4 * d4 * f3 / 6.0 - d4 * d4 * d4 * f2 / 24.0);          // eqs 24, 28
    return E1 + d5;
}
inline double kepler_solver(double MA, double e) { return kepler_markley(MA, e); }
// ForwardDiff extension of PlanetOrbits: primal from the solver, partials by the implicit
// function theorem on E - e sinE = M.
template <int N> inline Dual<N> kepler_solver(const Dual<N>& MA, const Dual<N>& e) {
    const double EA = kepler_markley(MA.v, e.v);
    const double sea = std::sin(EA), cea = std::cos(EA);
    const double inv = 1.0 / (1.0 - e.v * cea);
    Dual<N> r; r.v = EA;
    for (int k = 0; k < N; ++k) r.d[k] = MA.d[k] * inv + e.d[k] * sea * inv;
    return r;
}

// ---------------------------------------------------------------------------------------
// a3: Visual{KepOrbit} constructor caches.
// ---------------------------------------------------------------------------------------
template <class T>
struct Orbit {
    bool ti = false;              // ThieleInnesOrbit: tiA..tiG [mas] replace the Campbell angles
    T tiA, tiB, tiF, tiG;
    T a, e, i, w, W, tp, M, plx;
    T n, nu_fact, p, cosi, sini, cosW, sinW, ecosw, esinw, cosi_cosW, cosi_sinW, J, K, dist;
};

template <class T>
inline Orbit<T> make_orbit(const OctoConstants& c, T a, T e, T i, T w, T W, T tp, T M, T plx) {
    Orbit<T> o; o.a = a; o.e = e; o.i = i; o.w = w; o.W = W; o.tp = tp; o.M = M; o.plx = plx;
    const double two_pi = 6.283185307179586;
    T period_days = sqrt(a * a * a / M) * c.kepler_year_days;
    T period_yrs = period_days / c.year2day;
    o.n = two_pi / period_yrs;                         // mean motion [rad/yr]
    o.nu_fact = sqrt((1.0 + e) / (1.0 - e));           // true-anomaly prefactor
    T oneminusesq = 1.0 - e * e;
    o.p = a * oneminusesq;                             // semi-latus rectum [AU]
    o.sini = sin(i); o.cosi = cos(i);
    T sinw = sin(w), cosw = cos(w);
    o.sinW = sin(W); o.cosW = cos(W);
    o.ecosw = e * cosw; o.esinw = e * sinw;
    o.cosi_cosW = o.cosi * o.cosW; o.cosi_sinW = o.cosi * o.sinW;
    o.J = ((two_pi * a) / period_yrs) / sqrt(oneminusesq);   // [AU/yr]
    o.K = o.J * c.au2m * c.sec2year * o.sini;                 // [m/s]
    o.dist = 1000.0 / plx * c.pc2au;                           // [AU]
    return o;
}

// ThieleInnesOrbit(; e, tp, M, plx, A, B, F, G) (PlanetOrbits; the in-repo restatement of its semi-major axis is
// src/parameterizations.jl:14-18, of its projection :346-353)
template <class T>
inline Orbit<T> make_orbit_ti(const OctoConstants& c, T A, T B, T F, T G, T e, T tp, T M, T plx) {
    Orbit<T> o; o.ti = true; o.tiA = A; o.tiB = B; o.tiF = F; o.tiG = G;
    o.e = e; o.tp = tp; o.M = M; o.plx = plx;
    T u = (A * A + B * B + F * F + G * G) / 2.0;
    T v = A * G - B * F;
    T alpha = sqrt(u + sqrt((u + v) * (u - v)));
    o.a = alpha / plx;
    const double two_pi = 6.283185307179586;
    T period_days = sqrt(o.a * o.a * o.a / M) * c.kepler_year_days;
    T period_yrs = period_days / c.year2day;
    o.n = two_pi / period_yrs;
    o.nu_fact = sqrt((1.0 + e) / (1.0 - e));
    o.p = o.a * (1.0 - e * e);
    o.i = T(0.0); o.w = T(0.0); o.W = T(0.0);
    o.sini = T(0.0); o.cosi = T(1.0); o.sinW = T(0.0); o.cosW = T(1.0); o.ecosw = e; o.esinw = T(0.0);
    o.cosi_cosW = T(1.0); o.cosi_sinW = T(0.0); o.J = T(0.0); o.K = T(0.0);
    o.dist = 1000.0 / plx * c.pc2au;
    return o;
}

// a4: solution at one epoch
template <class T>
struct Solution { T nu, EA, sinnu_w, cosnu_w, ecosnu, r, cart2angle; double t; };

template <class T>
inline Solution<T> orbitsolve(const OctoConstants& c, const Orbit<T>& o, double t) {
    Solution<T> s; s.t = t;
    T MA = o.n / c.year2day * (t - o.tp);
    s.EA = kepler_solver(MA, o.e);
    s.nu = 2.0 * atan(o.nu_fact * tan(s.EA / 2.0));
    T arg = o.w + s.nu;
    s.sinnu_w = sin(arg); s.cosnu_w = cos(arg);
    s.ecosnu = o.e * cos(s.nu);
    s.r = o.p / (1.0 + s.ecosnu);
    s.cart2angle = c.rad2as * 1e3 / o.dist;
    return s;
}

// a6, a7
// Thiele-Innes projection: x = cos E - e, y = sqrt(1 - e²) sin E; ra = xB + yG, dec = xA + yF [mas]
template <class T> inline void ti_xy(const Orbit<T>& o, const Solution<T>& s, T& x, T& y, T& xdot, T& ydot) {
    T sE = sin(s.EA), cE = cos(s.EA), rt = sqrt(1.0 - o.e * o.e);
    x = cE - o.e; y = rt * sE;
    T Edot = o.n / (1.0 - o.e * cE);          // [rad/yr]
    xdot = -sE * Edot; ydot = rt * cE * Edot;
}
template <class T> inline T raoff(const Orbit<T>& o, const Solution<T>& s) {
    if (o.ti) { T x, y, xd, yd; ti_xy(o, s, x, y, xd, yd); return x * o.tiB + y * o.tiG; }
    T xcart = s.r * (s.cosnu_w * o.sinW + s.sinnu_w * o.cosi * o.cosW);   // [AU]
    return xcart * s.cart2angle;                                           // [mas]
}
template <class T> inline T decoff(const Orbit<T>& o, const Solution<T>& s) {
    if (o.ti) { T x, y, xd, yd; ti_xy(o, s, x, y, xd, yd); return x * o.tiA + y * o.tiF; }
    T ycart = s.r * (s.cosnu_w * o.cosW - s.sinnu_w * o.cosi * o.sinW);
    return ycart * s.cart2angle;
}
template <class T> inline T radvel(const Orbit<T>& o, const Solution<T>& s) { return o.K * (s.cosnu_w + o.ecosw); }
// star reflex given the companion mass [Msol]
template <class T> inline T raoff(const Orbit<T>& o, const Solution<T>& s, const T& m) { return -m / o.M * raoff(o, s); }
template <class T> inline T decoff(const Orbit<T>& o, const Solution<T>& s, const T& m) { return -m / o.M * decoff(o, s); }
template <class T> inline T radvel(const Orbit<T>& o, const Solution<T>& s, const T& m) { return -m / o.M * radvel(o, s); }
// proper motion of the relative orbit [mas/yr] (PlanetOrbits pmra/pmdec: d(raoff)/dt, d(decoff)/dt):
// with u = ν + ω:  d(r cos u)/dt = -J (sin u + e sin ω),  d(r sin u)/dt = J (cos u + e cos ω),  J = n a / sqrt(1 - e²) [AU/yr]
template <class T> inline T pmra(const Orbit<T>& o, const Solution<T>& s) {
    if (o.ti) { T x, y, xd, yd; ti_xy(o, s, x, y, xd, yd); return xd * o.tiB + yd * o.tiG; }
    T xdot = o.J * (o.cosi_cosW * (s.cosnu_w + o.ecosw) - o.sinW * (s.sinnu_w + o.esinw));
    return xdot * s.cart2angle;
}
template <class T> inline T pmdec(const Orbit<T>& o, const Solution<T>& s) {
    if (o.ti) { T x, y, xd, yd; ti_xy(o, s, x, y, xd, yd); return xd * o.tiA + yd * o.tiF; }
    T ydot = -o.J * (o.cosi_sinW * (s.cosnu_w + o.ecosw) + o.cosW * (s.sinnu_w + o.esinw));
    return ydot * s.cart2angle;
}
template <class T> inline T pmra(const Orbit<T>& o, const Solution<T>& s, const T& m) { return -m / o.M * pmra(o, s); }
template <class T> inline T pmdec(const Orbit<T>& o, const Solution<T>& s, const T& m) { return -m / o.M * pmdec(o, s); }

// 2-D zero-mean normal logpdf, Σ = [v1 c√(v1v2); c√(v1v2) v2]  (Distributions.MvNormal)
template <class T>
inline T logpdf_mvnormal2(const T& s1, const T& s2, double cor, const T& r1, const T& r2) {
    const double log2pi = 1.8378770664093453;
    T v1 = s1 * s1, v2 = s2 * s2;
    double omc = 1.0 - cor * cor;
    T det = v1 * v2 * omc;
    T maha = (r1 * r1 / v1 - 2.0 * cor * r1 * r2 / (s1 * s2) + r2 * r2 / v2) / omc;
    return -log2pi - 0.5 * log(det) - 0.5 * maha;
}

inline double rem_trunc(double x, double m) { return std::fmod(x, m); }
template <int N> inline Dual<N> rem_trunc(const Dual<N>& x, double m) { Dual<N> r = x; r.v = std::fmod(x.v, m); return r; }

inline bool chain_valid(const OctoLayout& L, const double* in, int64_t ld, int64_t c) {
    for (int k = 0; k < L.n_in; ++k) if (!std::isfinite(in[c + k * ld])) return false;
    for (int p = 0; p < L.n_planets; ++p) {
        double e = in[c + L.idx_e[p] * ld], a;
        double M = in[c + L.idx_M[p] * ld], plx = in[c + L.idx_plx[p] * ld];
        if (L.basis[p] == OCTO_BASIS_THIELE_INNES) {
            const double A = in[c + L.idx_A[p] * ld], B = in[c + L.idx_B[p] * ld], F = in[c + L.idx_F[p] * ld], G = in[c + L.idx_G[p] * ld];
            a = (A * A + B * B + F * F + G * G) > 0.0 ? 1.0 : 0.0;
        } else a = in[c + L.idx_a[p] * ld];
        if (!(e >= 0.0 && e < 1.0) || !(a > 0.0) || !(M > 0.0) || !(plx > 0.0)) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// a1: one chain of ln_like_generated.  T = double (value) or Dual<N> (value + gradient).
// `X` holds the n_in inputs already lifted to T.
// ---------------------------------------------------------------------------------------
// trend_function(θ_obs, epoch_k) of an RV observation, for trends linear in the observation variables (the ABI's
// representation of the closure; rv-absolute.jl:143, rv-absolute-margin.jl:111, rv-relative.jl:131)
template <class T, class XT>
inline T trend(const OctoObsBlock& B, const XT& X, int k, int n) {
    T tr(B.trend_const ? B.trend_const[k] : 0.0);
    for (int v = 0; v < B.n_trend; ++v) tr += X[B.idx_trend[v]] * B.trend_basis[(size_t)v * n + k];
    return tr;
}

template <class T>
inline T ln_like_chain(const OctoConstants& c, const OctoLayout& L, const OctoObsBlock* blocks, int n_blocks,
                       const T* X) {
    const int P = L.n_planets;
    // epoch list: concatenation in block order (system.jl:35-54), start index per block
    std::vector<int64_t> start(n_blocks);
    int64_t E = 0;
    for (int b = 0; b < n_blocks; ++b) { start[b] = E; E += blocks[b].n_epochs; }

    // orbits (system.jl:116-118)
    Orbit<T> orb[OCTO_MAX_PLANETS];
    for (int p = 0; p < P; ++p) {
        if (L.basis[p] == OCTO_BASIS_THIELE_INNES)
            orb[p] = make_orbit_ti<T>(c, X[L.idx_A[p]], X[L.idx_B[p]], X[L.idx_F[p]], X[L.idx_G[p]], X[L.idx_e[p]],
                                      X[L.idx_tp[p]], X[L.idx_M[p]], X[L.idx_plx[p]]);
        else
            orb[p] = make_orbit<T>(c, X[L.idx_a[p]], X[L.idx_e[p]], X[L.idx_i[p]], X[L.idx_w[p]], X[L.idx_W[p]],
                                   X[L.idx_tp[p]], X[L.idx_M[p]], X[L.idx_plx[p]]);
    }
    // HOT LOOP 1: every planet at every epoch (system.jl:156-170, 257-262)
    std::vector<Solution<T>> sols((size_t)P * (size_t)E);
    for (int p = 0; p < P; ++p)
        for (int b = 0; b < n_blocks; ++b)
            for (int k = 0; k < blocks[b].n_epochs; ++k)
                sols[(size_t)p * E + start[b] + k] = orbitsolve(c, orb[p], blocks[b].epoch[k]);

    // HOT LOOP 2: the likelihood objects
    T ll(0.0);
    const double log2pi = 1.8378770664093453, two_pi = 6.283185307179586, pi = 3.141592653589793;
    for (int b = 0; b < n_blocks; ++b) {
        const OctoObsBlock& B = blocks[b];
        const int n = B.n_epochs;
        T jitter = B.idx_jitter >= 0 ? X[B.idx_jitter] : T(0.0);
        T offset = B.idx_offset >= 0 ? X[B.idx_offset] : T(0.0);
        if (B.kind == OCTO_KIND_ASTROM_RADEC || B.kind == OCTO_KIND_ASTROM_PASEP) {
            T platescale = B.idx_platescale >= 0 ? X[B.idx_platescale] : T(1.0);
            T northangle = B.idx_northangle >= 0 ? X[B.idx_northangle] : T(0.0);
            const int ip = B.planet;
            for (int k = 0; k < n; ++k) {
                const Solution<T>& sol = sols[(size_t)ip * E + start[b] + k];
                // a8: reflex of the star due to interior companions with a mass variable
                T ra_pert(0.0), dec_pert(0.0);
                for (int j = 0; j < P; ++j) {
                    if (value(orb[j].a) < value(orb[ip].a)) {
                        if (L.idx_mass[j] < 0) continue;
                        T m = X[L.idx_mass[j]] * c.mjup2msol;
                        const Solution<T>& s2 = sols[(size_t)j * E + start[b] + k];
                        ra_pert += raoff(orb[j], s2, m);
                        dec_pert += decoff(orb[j], s2, m);
                    }
                }
                T ra_model = raoff(orb[ip], sol) - ra_pert;
                T dec_model = decoff(orb[ip], sol) - dec_pert;
                T resid1, resid2;
                if (B.kind == OCTO_KIND_ASTROM_PASEP) {
                    T rho = hypot(ra_model, dec_model);
                    T pa = atan2(ra_model, dec_model);
                    T pa_dat = B.y1[k] + northangle;
                    // Julia `%` is rem (sign of the dividend); derivative w.r.t. the dividend is 1
                    T pa_diff = rem_trunc(pa_dat - pa + pi, two_pi) - pi;
                    if (value(pa_diff) < -pi) pa_diff = pa_diff + two_pi;
                    resid1 = pa_diff;
                    resid2 = B.y2[k] * platescale - rho;
                } else {
                    T pa_dat = std::atan2(B.y2[k], B.y1[k]) - northangle;
                    T sep_dat = std::hypot(B.y2[k], B.y1[k]) * platescale;
                    T ra_dat = sep_dat * cos(pa_dat);
                    T dec_dat = sep_dat * sin(pa_dat);
                    resid1 = ra_dat - ra_model;
                    resid2 = dec_dat - dec_model;
                }
                const double cor = B.has_cor ? B.cor[k] : 0.0;
                if (value(jitter) == 0.0) {
                    ll += logpdf_mvnormal2(T(B.s1[k]), T(B.s2[k]), cor, resid1, resid2);
                } else {
                    T s1 = hypot(B.s1[k], jitter), s2 = hypot(B.s2[k], jitter);
                    ll += logpdf_mvnormal2(s1, s2, cor, resid1, resid2);
                }
            }
            if (B.obs_prior) {
                // ObsPriorAstromONeil2019 (src/likelihoods/prior-observable.jl:78-137): on top of the wrapped table's
                // ln_like.  meananom(sol) / eccanom(sol) are PlanetOrbits accessors (third-party, absent):
                // eccanom = the solver's E, meananom = E - e sin E.
                const Orbit<T>& o = orb[ip];
                T P = sqrt(o.a * o.a * o.a / o.M) * c.kepler_year_days / 365.25;     // period(orbit) / 365.25
                T jac(0.0);
                for (int k = 0; k < n; ++k) {
                    const Solution<T>& sol = sols[(size_t)ip * E + start[b] + k];
                    T EA = sol.EA;
                    T Mm = EA - o.e * sin(EA);
                    T f = 3.0 * Mm * (o.e + cos(EA)) + 2.0 * (-2.0 + o.e * o.e + o.e * cos(EA)) * sin(EA);
                    jac += value(f) < 0.0 ? -f : f;
                }
                T sqrt_eccen = sqrt(1.0 - o.e * o.e);
                jac = jac * cbrt(P) / sqrt_eccen;
                ll += 2.0 * log(jac);
            }
        } else if (B.kind == OCTO_KIND_RV_STAR_ABS || B.kind == OCTO_KIND_RV_STAR_MARGIN) {
            const bool margin = (B.kind == OCTO_KIND_RV_STAR_MARGIN);
            T A(0.0), Bq(0.0), C(0.0), acc(0.0);
            for (int k = 0; k < n; ++k) {
                T rv_model = margin ? T(0.0) : offset;
                rv_model += trend<T>(B, X, k, n);               // rv-absolute.jl:143, rv-absolute-margin.jl:111
                for (int p = 0; p < P; ++p) {
                    T m = X[L.idx_mass[p]] * c.mjup2msol;
                    rv_model += radvel(orb[p], sols[(size_t)p * E + start[b] + k], m);
                }
                T resid = B.y1[k] - rv_model;
                T var = B.s1[k] * B.s1[k] + jitter * jitter;
                if (margin) {
                    A += 1.0 / var; Bq -= 2.0 * resid / var; C += resid * resid / var;
                    acc -= log(two_pi * var);
                } else {
                    acc += -0.5 * (log2pi + log(var) + resid * resid / var);  // MvNormal(Diagonal(var))
                }
            }
            if (margin) acc -= -(Bq * Bq) / (4.0 * A) + C + log(A);   // rv-absolute-margin.jl:181, verbatim
            ll += acc;
        } else if (B.kind == OCTO_KIND_RV_PLANET_REL) {
            const int ip = B.planet;
            T acc(0.0);
            for (int k = 0; k < n; ++k) {
                T rv_model = offset;
                rv_model += trend<T>(B, X, k, n);               // rv-relative.jl:131
                rv_model += radvel(orb[ip], sols[(size_t)ip * E + start[b] + k]);
                for (int j = 0; j < P; ++j) {
                    if (value(orb[j].a) < value(orb[ip].a)) {
                        if (L.idx_mass[j] < 0) continue;
                        T m = X[L.idx_mass[j]] * c.mjup2msol;
                        rv_model += radvel(orb[j], sols[(size_t)j * E + start[b] + k], m);
                    }
                }
                T resid = B.y1[k] - rv_model;
                T var = B.s1[k] * B.s1[k] + jitter * jitter;
                acc += -0.5 * (log2pi + log(var) + resid * resid / var);
            }
            ll += acc;
        } else if (B.kind == OCTO_KIND_HGCA_INSTANT) {
            // HGCAInstantaneousObs: simulate (hgca.jl:219-417, absolute_orbits = false) + ln_like (:155-216).
            // The averaging counters and epoch sums advance inside the planet loop, exactly as written there.
            T pmra_sys = X[B.idx_pmra], pmdec_sys = X[B.idx_pmdec];
            T ra_m[2] = {T(0.0), T(0.0)}, dec_m[2] = {T(0.0), T(0.0)}, pmra_m[2] = {T(0.0), T(0.0)}, pmdec_m[2] = {T(0.0), T(0.0)};
            double ep_ra[2] = {0, 0}, ep_dec[2] = {0, 0};
            int N_ra[2] = {0, 0}, N_dec[2] = {0, 0};
            for (int inst = 0; inst < 2; ++inst)
                for (int p = 0; p < P; ++p)
                    for (int k = 0; k < n; ++k) {
                        const int code = (int)B.y1[k];
                        if (code / 2 != inst) continue;
                        const Solution<T>& sol = sols[(size_t)p * E + start[b] + k];
                        T m = X[L.idx_mass[p]] * c.mjup2msol;
                        if (code % 2 == 0) {
                            N_ra[inst] += 1; ep_ra[inst] += B.epoch[k];
                            ra_m[inst] += raoff(orb[p], sol, m); pmra_m[inst] += pmra(orb[p], sol, m);
                        } else {
                            N_dec[inst] += 1; ep_dec[inst] += B.epoch[k];
                            dec_m[inst] += decoff(orb[p], sol, m); pmdec_m[inst] += pmdec(orb[p], sol, m);
                        }
                    }
            for (int inst = 0; inst < 2; ++inst) {
                ra_m[inst] = ra_m[inst] / (double)N_ra[inst]; dec_m[inst] = dec_m[inst] / (double)N_dec[inst];
                pmra_m[inst] = pmra_m[inst] / (double)N_ra[inst] + pmra_sys;
                pmdec_m[inst] = pmdec_m[inst] / (double)N_dec[inst] + pmdec_sys;
                ep_ra[inst] /= N_ra[inst]; ep_dec[inst] /= N_dec[inst];
            }
            const double julian_year = 365.25;
            T pmra_hg = (ra_m[1] - ra_m[0]) / (ep_ra[1] - ep_ra[0]) * julian_year + pmra_sys;
            T pmdec_hg = (dec_m[1] - dec_m[0]) / (ep_dec[1] - ep_dec[0]) * julian_year + pmdec_sys;
            const double* q = B.aux;       // hip, hg, gaia: pmra, pmdec, σ_pmra, σ_pmdec, correlation
            ll += logpdf_mvnormal2(T(q[2]), T(q[3]), q[4], pmra_m[0] - q[0], pmdec_m[0] - q[1]);
            ll += logpdf_mvnormal2(T(q[7]), T(q[8]), q[9], pmra_hg - q[5], pmdec_hg - q[6]);
            ll += logpdf_mvnormal2(T(q[12]), T(q[13]), q[14], pmra_m[1] - q[10], pmdec_m[1] - q[11]);
        }
    }
    return ll;
}

}  // namespace octo_oracle
