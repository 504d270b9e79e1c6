// octo_param.cu — K0: the standard parameterisation around the hot path, on the device.
//
//   forward  (k_param_forward):  θ_t -> invlink -> natural parameters -> derived kernel inputs `in`
//                                + Σ logpdf_with_trans + UnitLengthPrior terms            (one thread per chain)
//   K1 / K1v (octo_kernels.cu):  ll(in), ∂ll/∂in
//   backward (k_param_backward): ∂ll/∂in -> ∂/∂θ (reverse through the input definitions; θ_at_epoch_to_tperi by
//                                forward-mode duals) -> + prior gradients -> × d invlink/dθ_t; lp = prior + ll
//
// Reference semantics: ℓπcallback (src/logdensitymodel.jl:110-146): non-finite θ_t => -Inf; invlink
// (src/variables.jl:1449-1493); arr2nt derived variables (UniformCircular src/variables.jl:279-299,
// θ_at_epoch_to_tperi src/parameterizations.jl:6-69); ln_prior_transformed with the "healing" of a non-finite
// term (src/variables.jl:1205-1369); then ln_like (which contains the UnitLengthPrior terms, :301-323).
// Bijectors/Distributions formulas: SURVEY.md Appendix B.  All per-chain work; performance is irrelevant next
// to K1 (two tiny launches), correctness is checked against oracle/octo_oracle_param.hpp.
#include <cfloat>
#include <math_constants.h>

#include "octo_internal.h"

namespace {

constexpr int MAXD = OCTO_PARAM_MAX, MAXIN = OCTO_PARAM_MAX;
constexpr double kTwoPi = 6.283185307179586477, kPi = 3.14159265358979323846, kHalfLog2Pi = 0.91893853320467274178;

struct PriorEval { double x, dxdy, L, dLdx; };

__device__ void prior_bounds(const OctoPrior& pr, double& lo, double& hi) {
    lo = -CUDART_INF; hi = CUDART_INF;
    switch (pr.family) {
        case OCTO_PRIOR_UNIFORM: case OCTO_PRIOR_LOGUNIFORM: lo = pr.p[0]; hi = pr.p[1]; break;
        case OCTO_PRIOR_SINE: lo = 2.220446049250313e-16; hi = kPi - 2.220446049250313e-16; break;
        case OCTO_PRIOR_TRUNCNORMAL: lo = pr.p[2]; hi = pr.p[3]; break;
        default: break;
    }
}

// x = invlink(y); L = logpdf_with_trans(prior, x); derivatives for the chain rule
__device__ PriorEval prior_eval(const OctoPrior& pr, double lognorm, double y) {
    PriorEval r;
    double lo, hi; prior_bounds(pr, lo, hi);
    const bool lb = isfinite(lo), ub = isfinite(hi);
    double J = 0.0, dJ = 0.0;
    if (lb && ub) {
        const double s = 1.0 / (1.0 + exp(-y));
        double x = (hi - lo) * s + lo;
        r.dxdy = (hi - lo) * s * (1.0 - s);
        if (x < lo) { x = lo; r.dxdy = 0.0; }
        if (x > hi) { x = hi; r.dxdy = 0.0; }
        r.x = x;
        J = log((x - lo) * (hi - x) / (hi - lo)); dJ = 1.0 / (x - lo) - 1.0 / (hi - x);
    } else if (lb) {
        const double ex = exp(y);
        r.x = ex + lo; r.dxdy = ex;
        J = log(r.x - lo); dJ = 1.0 / (r.x - lo);
    } else if (ub) {
        const double ex = exp(y);
        r.x = hi - ex; r.dxdy = -ex;
        J = log(hi - r.x); dJ = -1.0 / (hi - r.x);
    } else { r.x = y; r.dxdy = 1.0; }
    const double x = r.x;
    double lp = 0.0, dlp = 0.0;
    switch (pr.family) {
        case OCTO_PRIOR_NORMAL: case OCTO_PRIOR_TRUNCNORMAL: {
            const double z = (x - pr.p[0]) / pr.p[1];
            lp = -0.5 * z * z - log(pr.p[1]) - kHalfLog2Pi - lognorm; dlp = -z / pr.p[1]; break;
        }
        case OCTO_PRIOR_UNIFORM: lp = -log(pr.p[1] - pr.p[0]); break;
        case OCTO_PRIOR_LOGUNIFORM: lp = -log(x) - log(log(pr.p[1] / pr.p[0])); dlp = -1.0 / x; break;
        case OCTO_PRIOR_SINE: { double sn, cs; sincos(x, &sn, &cs); lp = log(sn / 2.0); dlp = cs / sn; break; }
        default: break;
    }
    r.L = lp + J; r.dLdx = dlp + dJ;
    return r;
}

// minimal forward-mode dual with 7 partials, for θ_at_epoch_to_tperi only
struct D7 { double v, d[7]; };
__device__ D7 mk(double v, int k) { D7 r; r.v = v; for (int q = 0; q < 7; ++q) r.d[q] = (q == k) ? 1.0 : 0.0; return r; }
__device__ D7 un(const D7& a, double f, double df) { D7 r; r.v = f; for (int q = 0; q < 7; ++q) r.d[q] = df * a.d[q]; return r; }
__device__ D7 operator+(const D7& a, const D7& b) { D7 r; r.v = a.v + b.v; for (int q = 0; q < 7; ++q) r.d[q] = a.d[q] + b.d[q]; return r; }
__device__ D7 operator-(const D7& a, const D7& b) { D7 r; r.v = a.v - b.v; for (int q = 0; q < 7; ++q) r.d[q] = a.d[q] - b.d[q]; return r; }
__device__ D7 operator-(const D7& a) { D7 r; r.v = -a.v; for (int q = 0; q < 7; ++q) r.d[q] = -a.d[q]; return r; }
__device__ D7 operator*(const D7& a, const D7& b) { D7 r; r.v = a.v * b.v; for (int q = 0; q < 7; ++q) r.d[q] = a.d[q] * b.v + a.v * b.d[q]; return r; }
__device__ D7 operator/(const D7& a, const D7& b) { D7 r; r.v = a.v / b.v; for (int q = 0; q < 7; ++q) r.d[q] = (a.d[q] - r.v * b.d[q]) / b.v; return r; }
__device__ D7 operator+(const D7& a, double b) { D7 r = a; r.v += b; return r; }
__device__ D7 operator-(double a, const D7& b) { D7 r = -b; r.v += a; return r; }
__device__ D7 operator*(const D7& a, double b) { D7 r; r.v = a.v * b; for (int q = 0; q < 7; ++q) r.d[q] = a.d[q] * b; return r; }
__device__ D7 dsin(const D7& a) { return un(a, sin(a.v), cos(a.v)); }
__device__ D7 dcos(const D7& a) { return un(a, cos(a.v), -sin(a.v)); }
__device__ D7 dsqrt(const D7& a) { const double s = sqrt(a.v); return un(a, s, 0.5 / s); }
__device__ D7 datan2(const D7& y, const D7& x) {
    D7 r; r.v = atan2(y.v, x.v); const double h = x.v * x.v + y.v * y.v;
    for (int q = 0; q < 7; ++q) r.d[q] = (x.v * y.d[q] - y.v * x.d[q]) / h;
    return r;
}
__device__ double psin(double a) { return sin(a); }
__device__ double pcos(double a) { return cos(a); }
__device__ double psqrt(double a) { return sqrt(a); }
__device__ double patan2(double y, double x) { return atan2(y, x); }
__device__ D7 psin(const D7& a) { return dsin(a); }
__device__ D7 pcos(const D7& a) { return dcos(a); }
__device__ D7 psqrt(const D7& a) { return dsqrt(a); }
__device__ D7 patan2(const D7& y, const D7& x) { return datan2(y, x); }

// src/parameterizations.jl:6-69, Campbell branch.  T = double (forward) or D7 (backward).
template <class T>
__device__ T tperi(const OctoConstants& c, const T& theta, double t_ref, const T& M, const T& e, const T& a, const T& i,
                   const T& w, const T& W) {
    const T cW = pcos(W), sW = psin(W), cw = pcos(w), sw = psin(w), ci = pcos(i);
    const T A = cW * cw - sW * sw * ci, B = sW * cw + cW * sw * ci;
    const T F = -(cW * sw) - sW * cw * ci, G = -(sW * sw) + cW * cw * ci;
    const T ct = pcos(theta), st = psin(theta);
    const T det = A * G - F * B;
    const T xr = (G * ct - F * st) / det, yr = (A * st - B * ct) / det;
    const T nu = patan2(yr, xr);
    const T s = psqrt(1.0 - e * e);
    const T snu = psin(nu), cnu = pcos(nu);
    const T MA = patan2(-(s * snu), -e - cnu) + kPi - e * s * snu / (e * cnu + 1.0);
    const T period_yrs = psqrt(a * a * a / M) * (c.kepler_year_days / c.year2day);
    // n = 2π / period_yrs;  tp = t_ref - MA / n * year2day
    return t_ref - MA * period_yrs * (c.year2day / kTwoPi);
}

struct ChainState {
    double th[MAXD], dxdy[MAXD], dLdx[MAXD], in[MAXIN];
    double lp_prior, extra;
    bool finite_in, healed, valid;
};

// shared by forward and backward: everything up to the kernel inputs
__device__ void chain_forward(const DevParam& P, const DevModel& m, const double* __restrict__ theta_t, int64_t c, int64_t ld,
                              ChainState& S) {
    S.finite_in = true; S.healed = false; S.lp_prior = 0.0; S.extra = 0.0;
    for (int j = 0; j < P.D; ++j) {
        const double y = theta_t[c + (int64_t)j * ld];
        if (!isfinite(y)) S.finite_in = false;
        const PriorEval r = prior_eval(P.priors[j], P.lognorm[j], isfinite(y) ? y : 0.0);
        S.th[j] = r.x; S.dxdy[j] = r.dxdy; S.dLdx[j] = r.dLdx;
        if (!S.healed) {
            if (!isfinite(r.L)) { S.healed = true; S.lp_prior = -DBL_MAX; }   // nextfloat(typemin(Float64)), then return
            else S.lp_prior += r.L;
        }
    }
    for (int k = 0; k < P.n_in; ++k) {
        const OctoInputDef& d = P.defs[k];
        double v = 0.0;
        switch (d.op) {
            case OCTO_IN_PARAM: v = S.th[d.a[0]]; break;
            case OCTO_IN_CONST: v = d.value; break;
            case OCTO_IN_CIRC: {
                const double x = S.th[d.a[0]], y = S.th[d.a[1]];
                v = atan2(y, x) / kTwoPi * d.value;
                const double lr = log(sqrt(x * x + y * y));
                S.extra += -lr - log(0.1) - kHalfLog2Pi - lr * lr / (2.0 * 0.1 * 0.1);
                break;
            }
            case OCTO_IN_TPERI:
                v = tperi<double>(m.c, S.in[d.a[0]], d.value, S.in[d.a[1]], S.in[d.a[2]], S.in[d.a[3]], S.in[d.a[4]],
                                  S.in[d.a[5]], S.in[d.a[6]]);
                break;
        }
        S.in[k] = v;
    }
    S.valid = S.finite_in;
    for (int k = 0; k < P.n_in; ++k) if (!isfinite(S.in[k])) S.valid = false;
}

__global__ void k_param_forward(const DevParam* __restrict__ Pp, const __grid_constant__ DevModel m,
                                const double* __restrict__ theta_t, int64_t n, int64_t ld, double* __restrict__ in_out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const DevParam& P = *Pp;
    ChainState S;
    chain_forward(P, m, theta_t, c, ld, S);
    // an invalid chain gets NaN inputs: K1 then returns -Inf / zero gradient for it
    for (int k = 0; k < P.n_in; ++k) in_out[c + (int64_t)k * n] = S.valid ? S.in[k] : CUDART_NAN;
}

__global__ void k_param_backward(const DevParam* __restrict__ Pp, const __grid_constant__ DevModel m,
                                 const double* __restrict__ theta_t, int64_t n, int64_t ld, const double* __restrict__ ll,
                                 const double* __restrict__ g_in, double* __restrict__ lp_out, double* __restrict__ g_t,
                                 int64_t ldg) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const DevParam& P = *Pp;
    ChainState S;
    chain_forward(P, m, theta_t, c, ld, S);
    const double llc = ll[c];
    const bool ok = S.finite_in && S.valid && isfinite(llc);
    lp_out[c] = !S.finite_in ? -CUDART_INF : ((S.valid && isfinite(llc)) ? S.lp_prior + (S.extra + llc) : -CUDART_INF);
    if (!g_t) return;
    if (!ok) { for (int j = 0; j < P.D; ++j) g_t[c + (int64_t)j * ldg] = 0.0; return; }
    double gin[MAXIN], gth[MAXD];
    for (int k = 0; k < P.n_in; ++k) gin[k] = g_in[c + (int64_t)k * n];
    for (int j = 0; j < P.D; ++j) gth[j] = S.healed ? 0.0 : S.dLdx[j];        // a healed prior is a constant
    for (int k = P.n_in - 1; k >= 0; --k) {
        const OctoInputDef& d = P.defs[k];
        const double gk = gin[k];
        switch (d.op) {
            case OCTO_IN_PARAM: gth[d.a[0]] += gk; break;
            case OCTO_IN_CIRC: {
                const double x = S.th[d.a[0]], y = S.th[d.a[1]], r2 = x * x + y * y, sc = d.value / kTwoPi;
                // angle, plus the UnitLengthPrior term  f(lr) = -lr - lr^2/(2*0.01), lr = ½ log r2
                const double lr = 0.5 * log(r2), dfdlr = -1.0 - lr / (0.1 * 0.1);
                gth[d.a[0]] += gk * sc * (-y / r2) + dfdlr * x / r2;
                gth[d.a[1]] += gk * sc * (x / r2) + dfdlr * y / r2;
                break;
            }
            case OCTO_IN_TPERI: {
                D7 a[7];
                for (int q = 0; q < 7; ++q) a[q] = mk(S.in[d.a[q]], q);
                const D7 t = tperi<D7>(m.c, a[0], d.value, a[1], a[2], a[3], a[4], a[5], a[6]);
                for (int q = 0; q < 7; ++q) gin[d.a[q]] += gk * t.d[q];
                break;
            }
            default: break;
        }
    }
    for (int j = 0; j < P.D; ++j) g_t[c + (int64_t)j * ldg] = gth[j] * S.dxdy[j];
}

__global__ void k_invlink(const DevParam* __restrict__ Pp, const double* __restrict__ theta_t, int64_t n, int64_t ld,
                          double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const DevParam& P = *Pp;
    for (int j = 0; j < P.D; ++j) out[c + (int64_t)j * ld] = prior_eval(P.priors[j], P.lognorm[j], theta_t[c + (int64_t)j * ld]).x;
}

}  // namespace

cudaError_t octo_param_forward(const DevParam* d_param, const DevModel& m, const double* d_theta, int64_t n, int64_t ld,
                               double* d_in, cudaStream_t st) {
    k_param_forward<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_param, m, d_theta, n, ld, d_in);
    return cudaGetLastError();
}
cudaError_t octo_param_backward(const DevParam* d_param, const DevModel& m, const double* d_theta, int64_t n, int64_t ld,
                                const double* d_ll, const double* d_g_in, double* d_lp, double* d_g_t, int64_t ldg,
                                cudaStream_t st) {
    k_param_backward<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_param, m, d_theta, n, ld, d_ll, d_g_in, d_lp, d_g_t, ldg);
    return cudaGetLastError();
}
cudaError_t octo_param_invlink(const DevParam* d_param, const double* d_theta, int64_t n, int64_t ld, double* d_out,
                               cudaStream_t st) {
    k_invlink<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_param, d_theta, n, ld, d_out);
    return cudaGetLastError();
}
