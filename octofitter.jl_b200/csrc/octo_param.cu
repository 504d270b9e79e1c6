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
#include "octo_param_dev.cuh"

namespace {

using namespace octo_param_dev;

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
        const PriorEval r = prior_eval(P.priors[j].family, P.priors[j].p[0], P.pc[j], fin ? y : 0.0);
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
        else if (d.op == OCTO_IN_CIRC) circ_forward(S.th[d.a[0]], S.th[d.a[1]], d.value, v, ext);
        S.in[k] = v; S.aux[k] = ext;
    }
    __syncthreads();
    // phase 2b + 3: the chain's lane 0 — θ_at_epoch_to_tperi in definition order, ordered sums, validity
    if (s == 0) {
        for (int k = 0; k < n_in; ++k) {
            const OctoInputDef& d = P.defs[k];
            if (d.op == OCTO_IN_TPERI || d.op == OCTO_IN_TPERI_TI) {
                const bool ti = d.op == OCTO_IN_TPERI_TI;
                double arg[8], trig[8];
                for (int q = 0; q < (ti ? 8 : 7); ++q) arg[q] = S.in[d.a[q]];
                p_sincos(arg[0], &trig[0], &trig[1]);
                if (!ti) { p_sincos(arg[4], &trig[2], &trig[3]); p_sincos(arg[5], &trig[4], &trig[5]); p_sincos(arg[6], &trig[6], &trig[7]); }
                double MA;
                S.in[k] = tperi_value(m.c, d.value, arg, trig, &MA, ti);
            }
        }
        bool finite_in = true;
        for (int j = 0; j < D; ++j) if (isnan(S.L[j]) && !isfinite(theta_t[c + (int64_t)j * ld])) finite_in = false;
        double lp, extra;
        const int flags = prior_sums(S.L, S.aux, S.in, D, n_in, 1, finite_in, lp, extra);
        const bool valid = flags & 4;
        if (active) {
            double* sv = save + c;
            sv[(int64_t)(3 * D) * n] = lp; sv[(int64_t)(3 * D + 1) * n] = extra;
            sv[(int64_t)(3 * D + 2) * n] = (double)flags;
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
                 const double* __restrict__ g_in, double* __restrict__ lp_out, double* __restrict__ g_t, int64_t ldg,
                 int post_mode) {
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
    if (active && s == 0) {
        const double like = extra + llc;               // post_mode 1: the likelihood part alone
        lp_out[c] = !finite_in ? -CUDART_INF : (ok ? (post_mode == 1 ? like : lp_prior + like) : -CUDART_INF);
    }
    if (!g_t) return;
    for (int j = s; j < D; j += SUB) {
        S.th[j] = save[c + (int64_t)j * n]; S.dxdy[j] = save[c + (int64_t)(D + j) * n];
        S.gth[j] = healed ? 0.0 : save[c + (int64_t)(2 * D + j) * n];          // a healed prior is a constant
    }
    for (int k = s; k < n_in; k += SUB) { S.in[k] = ok ? in_vals[c + (int64_t)k * n] : 1.0; S.aux[k] = ok ? g_in[c + (int64_t)k * n] : 0.0; }
    __syncthreads();
    // θ_at_epoch_to_tperi inputs, last definition first (hand-derived reverse pass on the chain's lane 0)
    if (s == 0) {
        for (int k = n_in - 1; k >= 0; --k) {
            const OctoInputDef& d = P.defs[k];
            if (d.op != OCTO_IN_TPERI && d.op != OCTO_IN_TPERI_TI) continue;
            const bool ti = d.op == OCTO_IN_TPERI_TI;
            const int na = ti ? 8 : 7;
            double arg[8], trig[8], part[8], MA;
            for (int q = 0; q < na; ++q) arg[q] = S.in[d.a[q]];
            p_sincos(arg[0], &trig[0], &trig[1]);
            if (!ti) { p_sincos(arg[4], &trig[2], &trig[3]); p_sincos(arg[5], &trig[4], &trig[5]); p_sincos(arg[6], &trig[6], &trig[7]); }
            tperi_value(m.c, d.value, arg, trig, &MA, ti);
            tperi_reverse(m.c, arg, trig, MA, part, ti);
            const double gk = S.aux[k];
            for (int q = 0; q < na; ++q) S.aux[d.a[q]] += gk * part[q];
        }
    }
    __syncthreads();
    // parameters and UniformCircular inputs: every parameter gathers its inputs, last input first
    if (active) for (int j = s; j < D; j += SUB) g_t[c + (int64_t)j * ldg] = ok ? param_gather(P, j, S.gth[j], S.th, S.aux, 1) * S.dxdy[j] : 0.0;
}

__global__ void k_invlink(const DevParam* __restrict__ Pp, const double* __restrict__ theta_t, int64_t n, int64_t ld,
                          double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const DevParam& P = *Pp;
    for (int j = 0; j < P.D; ++j) out[c + (int64_t)j * ld] = prior_eval(P.priors[j].family, P.priors[j].p[0], P.pc[j], theta_t[c + (int64_t)j * ld]).x;
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
                                double* d_g_t, int64_t ldg, int post_mode, cudaStream_t st) {
    k_param_backward<<<(unsigned)((n + CHAINS - 1) / CHAINS), SUB * CHAINS, param_smem(D, m.n_in), st>>>(
        d_param, m, n, d_in, d_save, d_ll, d_g_in, d_lp, d_g_t, ldg, post_mode);
    return cudaGetLastError();
}
cudaError_t octo_param_invlink(const DevParam* d_param, const double* d_theta, int64_t n, int64_t ld, double* d_out,
                               cudaStream_t st) {
    k_invlink<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_param, d_theta, n, ld, d_out);
    return cudaGetLastError();
}
