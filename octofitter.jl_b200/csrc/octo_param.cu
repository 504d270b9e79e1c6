// octo_param.cu — K0: the standard parameterisation around the hot path, on the device.
//
//   forward  (k_param_forward):  θ_t -> invlink -> natural parameters -> derived kernel inputs `in`
//                                + Σ logpdf_with_trans + UnitLengthPrior terms            (8 lanes per chain)
//   K1 / K1v (octo_kernels.cu):  ll(in), ∂ll/∂in
//   backward (k_param_backward): ∂ll/∂in -> ∂/∂θ (reverse through the input definitions; θ_at_epoch_to_tperi by
//                                forward-mode duals) -> + prior gradients -> × d invlink/dθ_t; lp = prior + ll
//
// Reference semantics: ℓπcallback (src/logdensitymodel.jl:110-146): non-finite θ_t => -Inf; invlink
// (src/variables.jl:1449-1493); arr2nt derived variables (UniformCircular src/variables.jl:279-299,
// θ_at_epoch_to_tperi src/parameterizations.jl:6-69); ln_prior_transformed with the "healing" of a non-finite
// term (src/variables.jl:1205-1369); then ln_like (which contains the UnitLengthPrior terms, :301-323).
// Bijectors/Distributions formulas: SURVEY.md Appendix B.  All per-chain work; performance is irrelevant next
// to K1 (two small launches, 8 lanes per chain), correctness is checked against oracle/octo_oracle_param.hpp.
#include <cfloat>
#include <math_constants.h>

#include "octo_internal.h"

namespace {

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

// minimal forward-mode dual with ONE partial: θ_at_epoch_to_tperi's 7 partial derivatives are computed by 7
// lanes, each seeding a different argument
struct D1 { double v, d; };
__device__ __forceinline__ D1 mk(double v, bool seed) { return D1{v, seed ? 1.0 : 0.0}; }
__device__ __forceinline__ D1 operator+(const D1& a, const D1& b) { return D1{a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ D1 operator-(const D1& a, const D1& b) { return D1{a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ D1 operator-(const D1& a) { return D1{-a.v, -a.d}; }
__device__ __forceinline__ D1 operator*(const D1& a, const D1& b) { return D1{a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ D1 operator/(const D1& a, const D1& b) { const double q = a.v / b.v; return D1{q, (a.d - q * b.d) / b.v}; }
__device__ __forceinline__ D1 operator+(const D1& a, double b) { return D1{a.v + b, a.d}; }
__device__ __forceinline__ D1 operator-(double a, const D1& b) { return D1{a - b.v, -b.d}; }
__device__ __forceinline__ D1 operator*(const D1& a, double b) { return D1{a.v * b, a.d * b}; }
__device__ __forceinline__ double psin(double a) { return sin(a); }
__device__ __forceinline__ double pcos(double a) { return cos(a); }
__device__ __forceinline__ double psqrt(double a) { return sqrt(a); }
__device__ __forceinline__ double patan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ D1 psin(const D1& a) { double s, c; sincos(a.v, &s, &c); return D1{s, c * a.d}; }
__device__ __forceinline__ D1 pcos(const D1& a) { double s, c; sincos(a.v, &s, &c); return D1{c, -s * a.d}; }
__device__ __forceinline__ D1 psqrt(const D1& a) { const double s = sqrt(a.v); return D1{s, 0.5 / s * a.d}; }
__device__ __forceinline__ D1 patan2(const D1& y, const D1& x) {
    return D1{atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / (x.v * x.v + y.v * y.v)};
}

// src/parameterizations.jl:6-69, Campbell branch.  T = double (forward) or D1 (backward, one partial per lane).
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

// ---------------------------------------------------------------------------------------------
// SUB lanes cooperate on one chain: prior evaluations and the 7 tperi partials run lane-parallel, the few
// order-dependent steps (sums, reverse pass) run on the chain's lane 0 in a fixed order.  Per-chain scratch
// lives in shared memory.  What backward needs from forward is saved in the caller's workspace.
// ---------------------------------------------------------------------------------------------
constexpr int SUB = 8, CHAINS = 16;                  // 128 threads per CTA
struct ChainSmem { double *th, *dxdy, *dLdx, *L, *in, *aux, *gth; };
__device__ __forceinline__ ChainSmem chain_smem(double* base, int local_chain, int D, int n_in) {
    const int Dp = D < SUB ? SUB : D;                  // every array has at least SUB slots (scratch use)
    double* p = base + (size_t)local_chain * (5 * Dp + 2 * n_in);
    ChainSmem s; s.th = p; s.dxdy = p + Dp; s.dLdx = p + 2 * Dp; s.L = p + 3 * Dp; s.gth = p + 4 * Dp;
    s.in = p + 5 * Dp; s.aux = p + 5 * Dp + n_in;
    return s;
}

// save layout per chain: [th(D) | dxdy(D) | dLdx(D) | lp_prior | extra | flags], column-major over chains
__global__ void __launch_bounds__(SUB * CHAINS)
k_param_forward(const DevParam* __restrict__ Pp, const __grid_constant__ DevModel m, const double* __restrict__ theta_t,
                int64_t n, int64_t ld, double* __restrict__ in_out, double* __restrict__ save) {
    extern __shared__ double sm[];
    const DevParam& P = *Pp;
    const int D = P.D, n_in = P.n_in;
    const int lc = threadIdx.x / SUB, s = threadIdx.x % SUB;
    const int64_t c_raw = (int64_t)blockIdx.x * CHAINS + lc;
    const bool active = c_raw < n;
    const int64_t c = active ? c_raw : n - 1;
    ChainSmem S = chain_smem(sm, lc, D, n_in);
    // phase 1: invlink + logpdf_with_trans, lane-parallel over parameters
    for (int j = s; j < D; j += SUB) {
        const double y = theta_t[c + (int64_t)j * ld];
        const bool fin = isfinite(y);
        const PriorEval r = prior_eval(P.priors[j], P.lognorm[j], fin ? y : 0.0);
        S.th[j] = r.x; S.dxdy[j] = r.dxdy; S.dLdx[j] = r.dLdx;
        S.L[j] = fin ? r.L : CUDART_NAN;                         // NaN marks a non-finite θ_t entry
    }
    __syncthreads();
    // phase 2a: derived inputs that depend on parameters only (arr2nt), lane-parallel
    for (int k = s; k < n_in; k += SUB) {
        const OctoInputDef& d = P.defs[k];
        double v = 0.0, ext = 0.0;
        if (d.op == OCTO_IN_PARAM) v = S.th[d.a[0]];
        else if (d.op == OCTO_IN_CONST) v = d.value;
        else if (d.op == OCTO_IN_CIRC) {
            const double x = S.th[d.a[0]], y = S.th[d.a[1]];
            v = atan2(y, x) / kTwoPi * d.value;
            const double lr = log(sqrt(x * x + y * y));
            ext = -lr - log(0.1) - kHalfLog2Pi - lr * lr / (2.0 * 0.1 * 0.1);      // UnitLengthPrior
        }
        S.in[k] = v; S.aux[k] = ext;
    }
    __syncthreads();
    // phase 2b + 3: the chain's lane 0 — θ_at_epoch_to_tperi in definition order, ordered sums, validity
    if (s == 0) {
        for (int k = 0; k < n_in; ++k) {
            const OctoInputDef& d = P.defs[k];
            if (d.op == OCTO_IN_TPERI)
                S.in[k] = tperi<double>(m.c, S.in[d.a[0]], d.value, S.in[d.a[1]], S.in[d.a[2]], S.in[d.a[3]], S.in[d.a[4]],
                                        S.in[d.a[5]], S.in[d.a[6]]);
        }
        bool finite_in = true, healed = false, valid = true;
        double lp = 0.0, extra = 0.0;
        for (int j = 0; j < D; ++j) {
            const double L = S.L[j];
            if (isnan(L) && !isfinite(theta_t[c + (int64_t)j * ld])) { finite_in = false; continue; }
            if (!healed) { if (!isfinite(L)) { healed = true; lp = -DBL_MAX; } else lp += L; }   // variables.jl:1229-1236
        }
        for (int k = 0; k < n_in; ++k) { extra += S.aux[k]; if (!isfinite(S.in[k])) valid = false; }
        valid = valid && finite_in;
        if (active) {
            double* sv = save + c;
            sv[(int64_t)(3 * D) * n] = lp; sv[(int64_t)(3 * D + 1) * n] = extra;
            sv[(int64_t)(3 * D + 2) * n] = (double)((finite_in ? 1 : 0) | (healed ? 2 : 0) | (valid ? 4 : 0));
        }
        S.aux[0] = valid ? 1.0 : 0.0;
    }
    __syncthreads();
    if (active) {
        const bool valid = S.aux[0] != 0.0;
        // an invalid chain gets NaN inputs: K1 then returns -Inf / zero gradient for it
        for (int k = s; k < n_in; k += SUB) in_out[c + (int64_t)k * n] = valid ? S.in[k] : CUDART_NAN;
        for (int j = s; j < D; j += SUB) {
            save[c + (int64_t)j * n] = S.th[j]; save[c + (int64_t)(D + j) * n] = S.dxdy[j];
            save[c + (int64_t)(2 * D + j) * n] = S.dLdx[j];
        }
    }
}

__global__ void __launch_bounds__(SUB * CHAINS)
k_param_backward(const DevParam* __restrict__ Pp, const __grid_constant__ DevModel m, int64_t n,
                 const double* __restrict__ in_vals, const double* __restrict__ save, const double* __restrict__ ll,
                 const double* __restrict__ g_in, double* __restrict__ lp_out, double* __restrict__ g_t, int64_t ldg) {
    extern __shared__ double sm[];
    const DevParam& P = *Pp;
    const int D = P.D, n_in = P.n_in;
    const int lc = threadIdx.x / SUB, s = threadIdx.x % SUB;
    const int64_t c_raw = (int64_t)blockIdx.x * CHAINS + lc;
    const bool active = c_raw < n;
    const int64_t c = active ? c_raw : n - 1;
    ChainSmem S = chain_smem(sm, lc, D, n_in);
    const double lp_prior = save[c + (int64_t)(3 * D) * n], extra = save[c + (int64_t)(3 * D + 1) * n];
    const int flags = (int)save[c + (int64_t)(3 * D + 2) * n];
    const bool finite_in = flags & 1, healed = flags & 2, valid = flags & 4;
    const double llc = ll[c];
    const bool ok = valid && isfinite(llc);
    if (active && s == 0) lp_out[c] = !finite_in ? -CUDART_INF : (ok ? lp_prior + (extra + llc) : -CUDART_INF);
    if (!g_t) return;
    for (int j = s; j < D; j += SUB) {
        S.th[j] = save[c + (int64_t)j * n]; S.dxdy[j] = save[c + (int64_t)(D + j) * n];
        S.gth[j] = healed ? 0.0 : save[c + (int64_t)(2 * D + j) * n];          // a healed prior is a constant
    }
    for (int k = s; k < n_in; k += SUB) { S.in[k] = ok ? in_vals[c + (int64_t)k * n] : 1.0; S.aux[k] = ok ? g_in[c + (int64_t)k * n] : 0.0; }
    __syncthreads();
    // θ_at_epoch_to_tperi inputs, last definition first: lanes 0..6 each produce one partial derivative
    for (int k = n_in - 1; k >= 0; --k) {
        const OctoInputDef& d = P.defs[k];
        if (d.op != OCTO_IN_TPERI) continue;
        double part = 0.0;
        if (s < 7) {
            D1 a[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) a[q] = mk(S.in[d.a[q]], q == s);
            part = tperi<D1>(m.c, a[0], d.value, a[1], a[2], a[3], a[4], a[5], a[6]).d;
        }
        S.L[s] = part;                                      // L is free scratch here (D >= 1; SUB slots needed)
        __syncthreads();
        if (s == 0) { const double gk = S.aux[k]; for (int q = 0; q < 7; ++q) S.aux[d.a[q]] += gk * S.L[q]; }
        __syncthreads();
    }
    // parameters and UniformCircular inputs: ordered accumulation on lane 0
    if (s == 0) {
        for (int k = n_in - 1; k >= 0; --k) {
            const OctoInputDef& d = P.defs[k];
            const double gk = S.aux[k];
            if (d.op == OCTO_IN_PARAM) S.gth[d.a[0]] += gk;
            else if (d.op == OCTO_IN_CIRC) {
                const double x = S.th[d.a[0]], y = S.th[d.a[1]], r2 = x * x + y * y, sc = d.value / kTwoPi;
                const double lr = 0.5 * log(r2), dfdlr = -1.0 - lr / (0.1 * 0.1);   // angle + UnitLengthPrior
                S.gth[d.a[0]] += gk * sc * (-y / r2) + dfdlr * x / r2;
                S.gth[d.a[1]] += gk * sc * (x / r2) + dfdlr * y / r2;
            }
        }
    }
    __syncthreads();
    if (active) for (int j = s; j < D; j += SUB) g_t[c + (int64_t)j * ldg] = ok ? S.gth[j] * S.dxdy[j] : 0.0;
}

__global__ void k_invlink(const DevParam* __restrict__ Pp, const double* __restrict__ theta_t, int64_t n, int64_t ld,
                          double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const DevParam& P = *Pp;
    for (int j = 0; j < P.D; ++j) out[c + (int64_t)j * ld] = prior_eval(P.priors[j], P.lognorm[j], theta_t[c + (int64_t)j * ld]).x;
}

}  // namespace

static size_t param_smem(int D, int n_in) { return (size_t)CHAINS * (5 * (size_t)(D < SUB ? SUB : D) + 2 * n_in) * sizeof(double); }

// opt both kernels in to the dynamic shared memory a (D, n_in) model needs (57 KB at the 64/64 limit)
cudaError_t octo_param_init(int D, int n_in) {
    cudaError_t e = cudaFuncSetAttribute(k_param_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)param_smem(D, n_in));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_param_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)param_smem(D, n_in));
}

cudaError_t octo_param_forward(const DevParam* d_param, int D, const DevModel& m, const double* d_theta, int64_t n,
                               int64_t ld, double* d_in, double* d_save, cudaStream_t st) {
    k_param_forward<<<(unsigned)((n + CHAINS - 1) / CHAINS), SUB * CHAINS, param_smem(D, m.n_in), st>>>(
        d_param, m, d_theta, n, ld, d_in, d_save);
    return cudaGetLastError();
}
cudaError_t octo_param_backward(const DevParam* d_param, int D, const DevModel& m, int64_t n, const double* d_in,
                                const double* d_save, const double* d_ll, const double* d_g_in, double* d_lp,
                                double* d_g_t, int64_t ldg, cudaStream_t st) {
    k_param_backward<<<(unsigned)((n + CHAINS - 1) / CHAINS), SUB * CHAINS, param_smem(D, m.n_in), st>>>(
        d_param, m, n, d_in, d_save, d_ll, d_g_in, d_lp, d_g_t, ldg);
    return cudaGetLastError();
}
cudaError_t octo_param_invlink(const DevParam* d_param, const double* d_theta, int64_t n, int64_t ld, double* d_out,
                               cudaStream_t st) {
    k_invlink<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_param, d_theta, n, ld, d_out);
    return cudaGetLastError();
}
