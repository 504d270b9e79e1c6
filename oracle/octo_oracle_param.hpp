// octo_oracle_param.hpp — CPU restatement of the standard-parameterisation layer around the hot path
// (SURVEY.md §8f N1).  TEST INFRASTRUCTURE ONLY (same rules as octo_oracle.hpp).
//
// PARITY STATUS: unpinned.  This layer is third-party arithmetic that is not in the reference tree:
//   Bijectors.jl 0.14-0.16 (`invlink`, `logpdf_with_trans`, TruncatedBijector) and Distributions.jl 0.25
//   (`logpdf` of Normal / Uniform / LogUniform / truncated Normal / LogNormal) — published formulas restated
//   (SURVEY.md Appendix B); call sites: src/variables.jl:1225,1260,1295,1330 (logpdf_with_trans),
//   :1464-1469 (invlink), src/logdensitymodel.jl:126-134 (ℓπcallback order of operations).
// In-tree pieces restated line by line:
//   Sine                        src/distributions.jl:15-40   (support [eps, π-eps])
//   UniformCircular / UnitLengthPrior   src/variables.jl:260-323
//   θ_at_epoch_to_tperi         src/parameterizations.jl:6-69
//   "healing" of a non-finite prior term   src/variables.jl:1229-1236
#pragma once
#include "octo_oracle.hpp"

namespace octo_oracle {

using std::exp;
template <int N> inline Dual<N> exp(const Dual<N>& a) { double e = std::exp(a.v); return unary(a, e, e); }

template <class T> inline T logistic(const T& y) { return 1.0 / (1.0 + exp(-y)); }

inline void prior_bounds(const OctoPrior& pr, double& lo, double& hi) {
    const double inf = std::numeric_limits<double>::infinity();
    switch (pr.family) {
        case OCTO_PRIOR_UNIFORM: case OCTO_PRIOR_LOGUNIFORM: lo = pr.p[0]; hi = pr.p[1]; break;
        case OCTO_PRIOR_SINE: lo = 0.0 + 2.220446049250313e-16; hi = 3.141592653589793 - 2.220446049250313e-16; break;
        case OCTO_PRIOR_TRUNCNORMAL: lo = pr.p[2]; hi = pr.p[3]; break;
        default: lo = -inf; hi = inf;
    }
}

// Bijectors.invlink(d, y): TruncatedBijector on [lo, hi]
template <class T> inline T prior_invlink(const OctoPrior& pr, const T& y) {
    double lo, hi; prior_bounds(pr, lo, hi);
    const bool lb = std::isfinite(lo), ub = std::isfinite(hi);
    if (lb && ub) {
        T x = (hi - lo) * logistic(y) + lo;
        if (value(x) < lo) return T(lo);
        if (value(x) > hi) return T(hi);
        return x;
    }
    if (lb) return exp(y) + lo;
    if (ub) return hi - exp(y);
    return y;
}

inline double normal_logcdf_diff(double mu, double sigma, double lo, double hi) {   // log(Φ(β) - Φ(α))
    const double is2 = 0.7071067811865476;
    const double a = std::isfinite(lo) ? (lo - mu) / sigma : -std::numeric_limits<double>::infinity();
    const double b = std::isfinite(hi) ? (hi - mu) / sigma : std::numeric_limits<double>::infinity();
    // Φ(b) - Φ(a) = ½ (erfc(-b/√2) - erfc(-a/√2)); evaluated on the side that avoids cancellation
    double tp;
    if (a > 0) tp = 0.5 * (std::erfc(a * is2) - std::erfc(b * is2));
    else tp = 0.5 * (std::erfc(-b * is2) - std::erfc(-a * is2));
    return std::log(tp);
}

// Bijectors.logpdf_with_trans(d, x, true) = logpdf(d, x) - logabsdetjac(bijector(d), x)
template <class T> inline T logpdf_with_trans(const OctoPrior& pr, const T& x) {
    const double half_log2pi = 0.9189385332046727;
    double lo, hi; prior_bounds(pr, lo, hi);
    T lp(0.0);
    switch (pr.family) {
        case OCTO_PRIOR_NORMAL: { T z = (x - pr.p[0]) / pr.p[1]; lp = -0.5 * z * z - std::log(pr.p[1]) - half_log2pi; break; }
        case OCTO_PRIOR_UNIFORM: lp = T(-std::log(pr.p[1] - pr.p[0])); break;
        case OCTO_PRIOR_LOGUNIFORM: lp = -log(x) - std::log(std::log(pr.p[1] / pr.p[0])); break;
        case OCTO_PRIOR_SINE: lp = log(sin(x) / 2.0); break;
        case OCTO_PRIOR_TRUNCNORMAL: {
            T z = (x - pr.p[0]) / pr.p[1];
            lp = -0.5 * z * z - std::log(pr.p[1]) - half_log2pi - normal_logcdf_diff(pr.p[0], pr.p[1], lo, hi); break;
        }
    }
    const bool lb = std::isfinite(lo), ub = std::isfinite(hi);
    if (lb && ub) lp = lp + log((x - lo) * (hi - x) / (hi - lo));
    else if (lb) lp = lp + log(x - lo);
    else if (ub) lp = lp + log(hi - x);
    return lp;
}

// src/parameterizations.jl:6-69 (Campbell branch)
template <class T>
inline T theta_at_epoch_to_tperi(const OctoConstants& c, const T& theta, double theta_epoch, const T& M, const T& e,
                                 const T& a, const T& i, const T& w, const T& W) {
    const double pi = 3.141592653589793, two_pi = 6.283185307179586;
    T A = (cos(W) * cos(w) - sin(W) * sin(w) * cos(i));
    T B = (sin(W) * cos(w) + cos(W) * sin(w) * cos(i));
    T F = (-cos(W) * sin(w) - sin(W) * cos(w) * cos(i));
    T G = (-sin(W) * sin(w) + cos(W) * cos(w) * cos(i));
    // [A F; B G] \ [cosθ; sinθ]
    T ct = cos(theta), st = sin(theta);
    T det = A * G - F * B;
    T x_over_r = (G * ct - F * st) / det;
    T y_over_r = (A * st - B * ct) / det;
    T nu = atan2(y_over_r, x_over_r);
    T s = sqrt(1.0 - e * e);
    T MA = atan2(-s * sin(nu), -e - cos(nu)) + pi - e * s * sin(nu) / (1.0 + e * cos(nu));
    T period_days = sqrt(a * a * a / M) * c.kepler_year_days;
    T period_yrs = period_days / c.year2day;
    T n = two_pi / period_yrs;
    return theta_epoch - MA / n * c.year2day;
}

// src/parameterizations.jl:6-69, Thiele-Innes branch (:9-19): a from the constants, T = [A F; B G] as given
template <class T>
inline T theta_at_epoch_to_tperi_ti(const OctoConstants& c, const T& theta, double theta_epoch, const T& M, const T& e,
                                    const T& plx, const T& A, const T& B, const T& F, const T& G) {
    const double pi = 3.141592653589793, two_pi = 6.283185307179586;
    T u = (A * A + B * B + F * F + G * G) / 2.0;
    T v = A * G - B * F;
    T alpha = sqrt(u + sqrt((u + v) * (u - v)));
    T a = alpha / plx;
    T ct = cos(theta), st = sin(theta);
    T det = A * G - F * B;
    T x_over_r = (G * ct - F * st) / det;
    T y_over_r = (A * st - B * ct) / det;
    T nu = atan2(y_over_r, x_over_r);
    T s = sqrt(1.0 - e * e);
    T MA = atan2(-s * sin(nu), -e - cos(nu)) + pi - e * s * sin(nu) / (1.0 + e * cos(nu));
    T period_days = sqrt(a * a * a / M) * c.kepler_year_days;
    T period_yrs = period_days / c.year2day;
    T n = two_pi / period_yrs;
    return theta_epoch - MA / n * c.year2day;
}

// ℓπcallback(θ_t) for the standard model families (src/logdensitymodel.jl:110-146)
template <class T>
inline T logpost_chain(const OctoConstants& c, const OctoLayout& L, const OctoObsBlock* blocks, int n_blocks,
                       const OctoPrior* priors, int D, const OctoInputDef* defs, const T* theta_t, bool like_only = false) {
    const double ninf = -std::numeric_limits<double>::infinity();
    for (int j = 0; j < D; ++j) if (!std::isfinite(value(theta_t[j]))) return T(ninf);      // :120-124
    std::vector<T> th(D), in(L.n_in);
    for (int j = 0; j < D; ++j) th[j] = prior_invlink(priors[j], theta_t[j]);                // :126
    // arr2nt: derived variables (:127)
    T extra(0.0);
    for (int k = 0; k < L.n_in; ++k) {
        const OctoInputDef& d = defs[k];
        switch (d.op) {
            case OCTO_IN_PARAM: in[k] = th[d.a[0]]; break;
            case OCTO_IN_CONST: in[k] = T(d.value); break;
            case OCTO_IN_CIRC: {
                const T& x = th[d.a[0]]; const T& y = th[d.a[1]];
                in[k] = atan2(y, x) / 6.283185307179586 * d.value;
                // UnitLengthPrior: logpdf(LogNormal(log(1), 0.1), sqrt(x^2 + y^2)) — part of ln_like in the reference
                T r = sqrt(x * x + y * y);
                T lr = log(r);
                extra += -lr - std::log(0.1) - 0.9189385332046727 - lr * lr / (2.0 * 0.1 * 0.1);
                break;
            }
            case OCTO_IN_TPERI:
                in[k] = theta_at_epoch_to_tperi(c, in[d.a[0]], d.value, in[d.a[1]], in[d.a[2]], in[d.a[3]], in[d.a[4]],
                                                in[d.a[5]], in[d.a[6]]);
                break;
            case OCTO_IN_TPERI_TI:
                in[k] = theta_at_epoch_to_tperi_ti(c, in[d.a[0]], d.value, in[d.a[1]], in[d.a[2]], in[d.a[3]], in[d.a[4]],
                                                   in[d.a[5]], in[d.a[6]], in[d.a[7]]);
                break;
        }
    }
    // ln_prior_transformed (:128), with the reference's "healing" of a non-finite term (variables.jl:1229-1236)
    // like_only: what octofit_rejection evaluates per prior draw — ln_like(system, arr2nt(θ)) alone
    // (src/sampling.jl:261-270); the UnitLengthPrior terms are part of ln_like in the reference
    T lp(0.0);
    for (int j = 0; j < D && !like_only; ++j) {
        T p = logpdf_with_trans(priors[j], th[j]);
        if (!std::isfinite(value(p))) { lp = T(-std::numeric_limits<double>::max()); goto like; }
        lp += p;
    }
like:
    if (!std::isfinite(value(lp))) return lp;                                              // :130-133
    // orbit-constructor failure => -Inf (system.jl:214-221); here: elements outside the Keplerian domain
    for (int p = 0; p < L.n_planets; ++p) {
        const double e = value(in[L.idx_e[p]]), M = value(in[L.idx_M[p]]), plx = value(in[L.idx_plx[p]]);
        double a;
        if (L.basis[p] == OCTO_BASIS_THIELE_INNES) {
            const double A = value(in[L.idx_A[p]]), B = value(in[L.idx_B[p]]), F = value(in[L.idx_F[p]]), G = value(in[L.idx_G[p]]);
            a = (A * A + B * B + F * F + G * G) > 0.0 ? 1.0 : 0.0;
        } else a = value(in[L.idx_a[p]]);
        if (!(e >= 0.0 && e < 1.0) || !(a > 0.0) || !(M > 0.0) || !(plx > 0.0)) return T(ninf);
    }
    for (int k = 0; k < L.n_in; ++k) if (!std::isfinite(value(in[k]))) return T(ninf);
    return lp + extra + ln_like_chain<T>(c, L, blocks, n_blocks, in.data());                // :134
}

}  // namespace octo_oracle
