// octo_param_dev.cuh — device functions of the standard parameterisation, shared by the fused
// path inside K1 (octo_kernels.cu) and the stand-alone K0 kernels (octo_param.cu): one source for both (they agree to
// rounding — every kernel inlines these functions and the compiler contracts a*b+c per instantiation).
//
// Reference semantics: invlink (src/variables.jl:1449-1493), logpdf_with_trans / ln_prior_transformed
// (src/variables.jl:1205-1369), UniformCircular + UnitLengthPrior (src/variables.jl:279-323),
// θ_at_epoch_to_tperi (src/parameterizations.jl:6-69).  Bijectors/Distributions formulas: SURVEY.md Appendix B.
#pragma once
#include <cfloat>
#include <math_constants.h>

#include "octo_internal.h"

namespace octo_param_dev {

constexpr double kTwoPi = 6.283185307179586477, kPi = 3.14159265358979323846, kHalfLog2Pi = 0.91893853320467274178;

// libm wrappers (one place to change how the parameterisation stage calls exp / log / atan2 / sincos).  Inline: round 1
// kept one out-of-line copy for all ~15 call sites (smaller code, no difference then); with the stage itself inline the
// calls cost 3.5 % of a leapfrog (register spills around each call, see octo_kernels.cu "INLINE OR OUT OF LINE").
static __device__ __forceinline__ double p_exp(double x) { return exp(x); }
static __device__ __forceinline__ double p_log(double x) { return log(x); }
static __device__ __forceinline__ double p_atan2(double y, double x) { return atan2(y, x); }
static __device__ __forceinline__ void p_sincos(double x, double* s, double* c) { sincos(x, s, c); }

struct PriorEval { double x, dxdy, L, dLdx; };

// x = invlink(y); L = logpdf_with_trans(prior, x); derivatives for the chain rule.
// pc = per-prior constants prepared by octo_set_parameterization (DevParam::pc):
//   [0] lower bound, [1] upper bound (±Inf when open), [2] the constant part of the log density
//   (Normal: -log σ - ½log 2π - log(Φ(β)-Φ(α)); Uniform: -log(b-a); LogUniform: -log(log(b/a))), [3] 1/(hi-lo), [4] 1/σ
static __device__ __forceinline__ PriorEval prior_eval(int family, double mu, const double* __restrict__ pc, double y) {
    PriorEval r;
    const double lo = pc[0], hi = pc[1];
    const bool lb = isfinite(lo), ub = isfinite(hi);
    double J = 0.0, dJ = 0.0;
    if (lb && ub) {                                   // scaled logit, clamped (Bijectors TruncatedBijector)
        const double s = 1.0 / (1.0 + p_exp(-y));
        double x = (hi - lo) * s + lo;
        r.dxdy = (hi - lo) * s * (1.0 - s);
        if (x < lo) { x = lo; r.dxdy = 0.0; }
        if (x > hi) { x = hi; r.dxdy = 0.0; }
        r.x = x;
        const double a = x - lo, b = hi - x, ab = a * b;
        J = p_log(ab * pc[3]); dJ = (b - a) / ab;
    } else if (lb) {
        const double ex = p_exp(y);
        r.x = ex + lo; r.dxdy = ex;
        const double a = r.x - lo;
        J = p_log(a); dJ = 1.0 / a;
    } else if (ub) {
        const double ex = p_exp(y);
        r.x = hi - ex; r.dxdy = -ex;
        const double b = hi - r.x;
        J = p_log(b); dJ = -1.0 / b;
    } else { r.x = y; r.dxdy = 1.0; }
    const double x = r.x;
    double lp = pc[2], dlp = 0.0;
    switch (family) {
        case OCTO_PRIOR_NORMAL: case OCTO_PRIOR_TRUNCNORMAL: {
            const double z = (x - mu) * pc[4];
            lp = fma(-0.5 * z, z, pc[2]); dlp = -z * pc[4]; break;       // (explicit: see circ_forward)
        }
        case OCTO_PRIOR_LOGUNIFORM: lp = -p_log(x) + pc[2]; dlp = -1.0 / x; break;
        case OCTO_PRIOR_SINE: { double sn, cs; p_sincos(x, &sn, &cs); lp = p_log(sn * 0.5); dlp = cs / sn; break; }
        default: break;                               // Uniform: the constant
    }
    r.L = lp + J; r.dLdx = dlp + dJ;
    return r;
}

// UniformCircular (src/variables.jl:279-299): angle = atan(y, x) / 2π · domain, plus the UnitLengthPrior term
// LogNormal(0, 0.1) on √(x² + y²) (src/variables.jl:301-323)
// Also hands out what the reverse pass needs of r² = x² + y² (its reciprocal and d(UnitLengthPrior)/d log r), so that
// the fused stage does not evaluate the logarithm and the division a second time per parameter of the pair.
__device__ __forceinline__ void circ_forward(double x, double y, double domain, double& v, double& ext, double* ir2_out = nullptr,
                                             double* dfdlr_out = nullptr) {
    v = p_atan2(y, x) * (domain / kTwoPi);
    // explicit fma / rounded products wherever a*b+c appears on a value-only path: the stage is inlined into several
    // kernels, and left to the compiler each instantiation may contract differently (a last-bit difference in lp between
    // the resident explorer and the launch-per-leapfrog one)
    const double r2 = fma(x, x, y * y), l2 = p_log(r2);
    const double lr = 0.5 * l2;
    ext = fma(-lr * lr, 50.0, -lr - (-2.302585092994045684 /* log 0.1 */) - kHalfLog2Pi);
    if (ir2_out) { *ir2_out = 1.0 / r2; *dfdlr_out = -1.0 - 0.5 * l2 * 100.0; }
}
// gk = ∂/∂angle; returns the contributions to ∂/∂x and ∂/∂y (angle and UnitLengthPrior)
__device__ __forceinline__ void circ_backward_pre(double x, double y, double domain, double gk, double ir2, double dfdlr,
                                                  double& gx, double& gy) {
    const double sc = gk * (domain / kTwoPi);
    gx = (dfdlr * x - sc * y) * ir2;
    gy = (dfdlr * y + sc * x) * ir2;
}
__device__ __forceinline__ void circ_backward(double x, double y, double domain, double gk, double& gx, double& gy) {
    const double r2 = x * x + y * y;
    circ_backward_pre(x, y, domain, gk, 1.0 / r2, -1.0 - 0.5 * p_log(r2) * 100.0, gx, gy);
}

// ordered sum of the prior terms with the reference's "healing" of a non-finite term (variables.jl:1229-1236: the
// sum stops at the first non-finite term and becomes -floatmax), the UnitLengthPrior terms, validity of the inputs.
// flags: 1 = every θ_t finite, 2 = healed, 4 = valid.  `stride` = distance between consecutive entries.
// skip_derived: the θ_at_epoch_to_tperi inputs are still being computed (by another warp, which checks them itself)
__device__ __forceinline__ int prior_sums(const double* L, const double* aux, const double* in, int D, int n_in, int stride,
                                          bool finite_in, double& lp, double& extra, const DevParam* skip_derived = nullptr) {
    double sum = 0.0, ex = 0.0;
    bool bad = false, valid = true;
#pragma unroll 4
    for (int j = 0; j < D; ++j) { const double v = L[j * stride]; bad = bad || !isfinite(v); sum += v; }
#pragma unroll 4
    for (int k = 0; k < n_in; ++k) {
        ex += aux[k * stride];
        const bool derived = skip_derived && (skip_derived->defs[k].op == OCTO_IN_TPERI || skip_derived->defs[k].op == OCTO_IN_TPERI_TI);
        valid = valid && (derived || isfinite(in[k * stride]));
    }
    lp = bad ? -DBL_MAX : sum; extra = ex;
    return (finite_in ? 1 : 0) | (bad ? 2 : 0) | ((valid && finite_in) ? 4 : 0);
}

// d lp / dθ_j before the invlink factor: the prior term (0 when healed) plus, last input first, every input that
// reads parameter j (DevParam::gat).  aux = ∂ll/∂inputs after the θ_at_epoch_to_tperi contributions were folded in.
// cir != nullptr: [2][n_in] (1/r², d prior/d log r) of the UniformCircular inputs, saved by the forward stage
__device__ __forceinline__ double param_gather(const DevParam& P, int j, double g0, const double* th, const double* aux,
                                               int stride, const double* cir = nullptr) {
    double g = g0;
#pragma unroll 1
    for (int it = P.gat_start[j]; it < P.gat_start[j + 1]; ++it) {
        const int k = P.gat[it] & 255, role = P.gat[it] >> 8;
        if (role == 0) g += aux[k * stride];
        else {
            const OctoInputDef& d = P.defs[k];
            double gx, gy;
            if (cir) circ_backward_pre(th[d.a[0] * stride], th[d.a[1] * stride], d.value, aux[k * stride], cir[k * stride],
                                       cir[(P.n_in + k) * stride], gx, gy);
            else circ_backward(th[d.a[0] * stride], th[d.a[1] * stride], d.value, aux[k * stride], gx, gy);
            if (role & 1) g += gx;
            if (role & 2) g += gy;
        }
    }
    return g;
}

// θ_at_epoch_to_tperi, src/parameterizations.jl:6-69.
//   Campbell branch: arguments in definition order arg[0..6] = (θ, M, e, a, i, ω, Ω); trig[0..7] = sin, cos of θ, i, ω, Ω
//   (computed by the caller, possibly on other warps).  Thiele-Innes branch (`ti`, :9-19): arg[0..7] = (θ, M, e, plx, A,
//   B, F, G), a = sqrt(u + sqrt((u+v)(u-v))) / plx; only trig[0..1] is used.
//   sin/cos of the true anomaly come from (xr, yr) / r instead of sincos(atan2(yr, xr)).  Returns tp and the mean
//   anomaly MA (the one transcendental the reverse pass needs).
struct TperiMid { double A, B, F, G, idet, xr, yr, ir, snu, cnu, s, u, v, iw2, q, p, a, tu, tv, tw, alpha; };
// keep / reuse: the reciprocals and roots (7 values per lane, stride `ks` apart): the forward pass of the fused stage
// stores them, its reverse pass reads them back instead of evaluating the divisions and square roots again
__device__ __forceinline__ TperiMid tperi_mid(const OctoConstants& c, const double* arg, const double* trig, bool ti,
                                              double* keep = nullptr, const double* reuse = nullptr, int ks = 0) {
    const double st = trig[0], ct = trig[1];
    const double M = arg[1], e = arg[2];
    TperiMid m;
    if (ti) {
        m.A = arg[4]; m.B = arg[5]; m.F = arg[6]; m.G = arg[7];
        m.tu = 0.5 * (m.A * m.A + m.B * m.B + m.F * m.F + m.G * m.G); m.tv = m.A * m.G - m.B * m.F;
        if (reuse) { m.tw = reuse[5 * ks]; m.alpha = reuse[6 * ks]; }
        else { m.tw = sqrt((m.tu + m.tv) * (m.tu - m.tv)); m.alpha = sqrt(m.tu + m.tw); }
        m.a = m.alpha / arg[3];
    } else {
        const double ci = trig[3], sw = trig[4], cw = trig[5], sW = trig[6], cW = trig[7];
        m.A = cW * cw - sW * sw * ci; m.B = sW * cw + cW * sw * ci;
        m.F = -(cW * sw) - sW * cw * ci; m.G = -(sW * sw) + cW * cw * ci;
        m.a = arg[3]; m.tu = m.tv = m.tw = m.alpha = 0.0;
    }
    m.idet = reuse ? reuse[0] : 1.0 / (m.A * m.G - m.F * m.B);
    m.xr = (m.G * ct - m.F * st) * m.idet; m.yr = (m.A * st - m.B * ct) * m.idet;
    m.ir = reuse ? reuse[ks] : rsqrt(m.xr * m.xr + m.yr * m.yr);
    m.snu = m.yr * m.ir; m.cnu = m.xr * m.ir;
    m.s = reuse ? reuse[2 * ks] : sqrt(1.0 - e * e);
    m.u = -(m.s * m.snu); m.v = -e - m.cnu;
    m.iw2 = reuse ? reuse[3 * ks] : 1.0 / (e * m.cnu + 1.0);
    m.q = e * m.s * m.snu * m.iw2;
    m.p = reuse ? reuse[4 * ks] : sqrt(m.a * m.a * m.a / M) * (c.kepler_year_days / c.year2day);      // period [yr]
    if (keep) {
        keep[0] = m.idet; keep[ks] = m.ir; keep[2 * ks] = m.s; keep[3 * ks] = m.iw2; keep[4 * ks] = m.p;
        keep[5 * ks] = m.tw; keep[6 * ks] = m.alpha;
    }
    return m;
}
static __device__ __forceinline__ double tperi_value(const OctoConstants& c, double t_ref, const double* arg, const double* trig,
                                              double* MA_out, bool ti, double* keep = nullptr, int ks = 0) {
    const TperiMid m = tperi_mid(c, arg, trig, ti, keep, nullptr, ks);
    const double MA = p_atan2(m.u, m.v) + kPi - m.q;
    *MA_out = MA;
    // n = 2π / period_yrs;  tp = t_ref - MA / n * year2day
    return t_ref - MA * m.p * (c.year2day / kTwoPi);
}
// hand-derived reverse pass: grad[q] = ∂tp/∂arg[q] (7 entries, or 8 for the Thiele-Innes branch).  Cheap arithmetic only
// (the forward intermediates are recomputed, MA comes from the forward pass); this replaced forward-mode dual
// evaluations whose code size made the once-per-CTA reverse stage instruction-fetch bound.
static __device__ __forceinline__ void tperi_reverse(const OctoConstants& c, const double* arg, const double* trig, double MA,
                                              double* grad, bool ti, const double* reuse = nullptr, int ks = 0) {
    const double st = trig[0], ct = trig[1];
    const double M = arg[1], e = arg[2];
    const TperiMid m = tperi_mid(c, arg, trig, ti, nullptr, reuse, ks);
    const double cc = c.year2day / kTwoPi;
    const double MAb = -m.p * cc, pb = -MA * cc;                 // tp = t_ref - MA p cc
    const double g_a = pb * 1.5 * m.p / m.a, g_M = -pb * 0.5 * m.p / M;
    // MA = atan2(u, v) + π - q
    const double qb = -MAb;
    const double w2b = -qb * m.q * m.iw2;                        // q = e s snu / w2, w2 = e cnu + 1
    const double ih = 1.0 / (m.u * m.u + m.v * m.v);
    const double ub = MAb * m.v * ih, vb = -MAb * m.u * ih;
    double eb = qb * m.s * m.snu * m.iw2 + w2b * m.cnu - vb;
    double sb = qb * e * m.snu * m.iw2 - ub * m.snu;
    const double snub = qb * e * m.s * m.iw2 - ub * m.s;
    const double cnub = w2b * e - vb;
    eb += sb * (-e / m.s);                                       // s = sqrt(1 - e²)
    // snu = yr / r, cnu = xr / r
    const double dot = (snub * m.yr + cnub * m.xr) * m.ir * m.ir * m.ir;
    const double xrb = cnub * m.ir - dot * m.xr, yrb = snub * m.ir - dot * m.yr;
    // xr = (G ct - F st) / det, yr = (A st - B ct) / det, det = A G - F B
    const double xd = xrb * m.idet, yd = yrb * m.idet;
    const double detb = -(xd * m.xr + yd * m.yr);
    const double Ab = yd * st + detb * m.G, Bb = -yd * ct - detb * m.F;
    const double Fb = -xd * st - detb * m.B, Gb = xd * ct + detb * m.A;
    const double stb = -xd * m.F + yd * m.A, ctb = xd * m.G - yd * m.B;
    grad[0] = stb * ct - ctb * st;                               // θ
    grad[1] = g_M; grad[2] = eb;
    if (ti) {
        // a = alpha / plx, alpha² = u + w, w² = (u+v)(u-v)
        const double plx = arg[3], k = g_a / (2.0 * m.alpha * plx), iw = 1.0 / m.tw;
        grad[3] = -g_a * m.a / plx;
        grad[4] = Ab + k * (m.A + (m.tu * m.A - m.tv * m.G) * iw);
        grad[5] = Bb + k * (m.B + (m.tu * m.B + m.tv * m.F) * iw);
        grad[6] = Fb + k * (m.F + (m.tu * m.F + m.tv * m.B) * iw);
        grad[7] = Gb + k * (m.G + (m.tu * m.G - m.tv * m.A) * iw);
        return;
    }
    const double si = trig[2], ci = trig[3], sw = trig[4], cw = trig[5], sW = trig[6], cW = trig[7];
    // A = cW cw - sW sw ci, B = sW cw + cW sw ci, F = -cW sw - sW cw ci, G = -sW sw + cW cw ci
    const double cWb = Ab * cw + Bb * sw * ci - Fb * sw + Gb * cw * ci;
    const double sWb = -Ab * sw * ci + Bb * cw - Fb * cw * ci - Gb * sw;
    const double cwb = Ab * cW + Bb * sW - Fb * sW * ci + Gb * cW * ci;
    const double swb = -Ab * sW * ci + Bb * cW * ci - Fb * cW - Gb * sW;
    const double cib = -Ab * sW * sw + Bb * cW * sw - Fb * sW * cw + Gb * cW * cw;
    grad[3] = g_a;
    grad[4] = -cib * si;                                         // i
    grad[5] = swb * cw - cwb * sw;                               // ω
    grad[6] = sWb * cW - cWb * sW;                               // Ω
}

}  // namespace octo_param_dev
