/* Replays, in plain C, the exact call sequence of julia/OctofitterB200.jl on the reference's own 11-parameter test
 * model (test/integration/sampling.jl:29-71): what `B200Model(system)` + one `ℓπcallback` / `∇ℓπcallback` evaluation
 * + `DevicePosterior(model)` + `hmc_run` + `pt_init` / `pt_hmc_run(sharded = true)` do through `ccall`.
 *
 * The structs are NOT taken from the header: they are declared here field by field in the order and with the C types
 * the Julia file declares them (Int32 -> int32_t, Cdouble -> double, Ptr -> pointer, NTuple{N,T} -> T[N]), and their
 * sizes / field offsets are checked against include/octo_b200.h at start-up.  A drift between the Julia glue's layout
 * and the ABI therefore fails here (and in tests/test_julia_glue_cpu.py, which parses the .jl file itself).
 * Prints hex floats; compiled and run by tests/test_gpu_boundary.py. */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/octo_b200.h"

/* ---- the Julia file's structs, restated */
typedef struct { double kepler_year_days, year2day, rad2as, pc2au, au2m, sec2year, mjup2msol; } JlConstants;
typedef struct {
    int32_t kind, planet, n_epochs, has_cor;
    const double *epoch, *y1, *y2, *s1, *s2, *cor;
    int32_t idx_jitter, idx_platescale, idx_northangle, idx_offset, obs_prior, idx_pmra, idx_pmdec, reserved;
    const double* aux;
    int32_t n_trend, idx_trend[3];
    const double *trend_basis, *trend_const;
} JlObsBlock;
typedef struct {
    int32_t n_planets, n_in;
    int32_t idx_plx[4], idx_a[4], idx_e[4], idx_i[4], idx_w[4], idx_W[4], idx_tp[4], idx_M[4], idx_mass[4];
    int32_t basis[4], idx_A[4], idx_B[4], idx_F[4], idx_G[4];
} JlLayout;
typedef struct { int32_t family, reserved; double p[4]; } JlPrior;
typedef struct { int32_t op; int32_t a[8]; double value; } JlInputDef;

#define SAME(T1, T2, f) (offsetof(T1, f) == offsetof(T2, f))
static int layout_ok(void) {
    int ok = sizeof(JlConstants) == sizeof(OctoConstants) && sizeof(JlObsBlock) == sizeof(OctoObsBlock) &&
             sizeof(JlLayout) == sizeof(OctoLayout) && sizeof(JlPrior) == sizeof(OctoPrior) && sizeof(JlInputDef) == sizeof(OctoInputDef);
    ok = ok && SAME(JlObsBlock, OctoObsBlock, epoch) && SAME(JlObsBlock, OctoObsBlock, cor) && SAME(JlObsBlock, OctoObsBlock, idx_jitter) &&
         SAME(JlObsBlock, OctoObsBlock, obs_prior) && SAME(JlObsBlock, OctoObsBlock, idx_pmra) && SAME(JlObsBlock, OctoObsBlock, aux) &&
         SAME(JlObsBlock, OctoObsBlock, n_trend) && SAME(JlObsBlock, OctoObsBlock, idx_trend) && SAME(JlObsBlock, OctoObsBlock, trend_const);
    ok = ok && SAME(JlLayout, OctoLayout, idx_plx) && SAME(JlLayout, OctoLayout, idx_mass) && SAME(JlLayout, OctoLayout, basis) &&
         SAME(JlLayout, OctoLayout, idx_G) && SAME(JlPrior, OctoPrior, p) && SAME(JlInputDef, OctoInputDef, a) && SAME(JlInputDef, OctoInputDef, value);
    return ok;
}

#define SYM(name) name##_t name = (name##_t)dlsym(h, "octo_" #name); if (!name) { fprintf(stderr, "missing octo_" #name "\n"); return 2; }
typedef int (*abi_version_t)(void);
typedef void (*default_constants_t)(void*);
typedef int (*create_t)(const void*, const void*, const void*, int32_t, int32_t, void**);
typedef void (*destroy_t)(void*);
typedef int (*logp_t)(void*, const double*, int64_t, int64_t, double*);
typedef int (*logp_grad_t)(void*, const double*, int64_t, int64_t, double*, double*);
typedef int (*logp_grad_begin_t)(void*, const double*, int64_t, int64_t, double*, double*, void**);
typedef int (*ready_t)(void*);
typedef int (*wait_t)(void*);
typedef int64_t (*total_epochs_t)(const void*);
typedef int (*set_parameterization_t)(void*, const void*, int32_t, const void*);
typedef int (*logpost_grad_t)(void*, const double*, int64_t, int64_t, double*, double*);
typedef int (*hmc_run_t)(void*, const double*, int64_t, int64_t, int32_t, int32_t, double, const double*, uint64_t, double*, double*, double*, double*, double*);
typedef int (*pt_init_t)(void*, const void*, int32_t, int32_t, int32_t, uint64_t);
typedef int (*pt_hmc_run_dist_t)(void*, const double*, int64_t, int64_t, const double*, int32_t, int32_t, int32_t, double, const double*,
                                 uint64_t, double*, double*, double*, double*, int32_t*, double*, double*, double*);
typedef const char* (*last_error_t)(void);

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, #call ": %s\n", last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s libocto_b200.so\n", argv[0]); return 2; }
    if (!layout_ok()) { fprintf(stderr, "struct layout of the Julia glue differs from include/octo_b200.h\n"); return 3; }
    void* h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    SYM(abi_version) SYM(default_constants) SYM(create) SYM(destroy) SYM(logp) SYM(logp_grad) SYM(logp_grad_begin) SYM(ready) SYM(wait)
    SYM(total_epochs) SYM(set_parameterization) SYM(logpost_grad) SYM(hmc_run) SYM(pt_init) SYM(pt_hmc_run_dist) SYM(last_error)
    if (abi_version() != OCTO_ABI_VERSION) { fprintf(stderr, "ABI version %d\n", abi_version()); return 3; }     /* __init__ */

    /* ---- build_context: kernel inputs in the order the glue registers them for planet b of the test model:
     *      plx, M (system), then a, e, i, ω, Ω, the position-angle variable θ (argument of θ_at_epoch_to_tperi), tp last */
    enum { PLX, M, A_, E_, I_, W_, OM, TH, TP, N_IN };
    double ep[8] = {50000, 50120, 50240, 50360, 50480, 50600, 50720, 50840};
    double ra[8] = {-505.7637580573554, -502.570356287689, -498.2089148883798, -492.67768482682357,
                    -485.9770335870402, -478.1095526888573, -469.0801731788123, -458.89628893460525};
    double dec[8] = {-66.92982418533026, -37.47217527025044, -7.927548139010479, 21.63557115669823,
                     51.147204404903704, 80.53589069730698, 109.72870493064629, 138.65128697876773};
    double sg[8] = {10, 10, 10, 10, 10, 10, 10, 10}, cr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    JlObsBlock B; memset(&B, 0, sizeof B);
    B.kind = 0; B.planet = 0; B.n_epochs = 8; B.has_cor = 1;
    B.epoch = ep; B.y1 = ra; B.y2 = dec; B.s1 = sg; B.s2 = sg; B.cor = cr;
    B.idx_jitter = B.idx_platescale = B.idx_northangle = B.idx_offset = -1; B.obs_prior = 0; B.idx_pmra = B.idx_pmdec = -1; B.n_trend = 0; B.idx_trend[0] = B.idx_trend[1] = B.idx_trend[2] = -1;
    JlLayout L; memset(&L, 0, sizeof L);
    L.n_planets = 1; L.n_in = N_IN;
    for (int p = 0; p < 4; ++p) {
        L.idx_plx[p] = L.idx_a[p] = L.idx_e[p] = L.idx_i[p] = L.idx_w[p] = L.idx_W[p] = L.idx_tp[p] = L.idx_M[p] = L.idx_mass[p] = -1;
        L.idx_A[p] = L.idx_B[p] = L.idx_F[p] = L.idx_G[p] = -1; L.basis[p] = 0;
    }
    L.idx_plx[0] = PLX; L.idx_M[0] = M; L.idx_a[0] = A_; L.idx_e[0] = E_; L.idx_i[0] = I_; L.idx_w[0] = W_; L.idx_W[0] = OM; L.idx_tp[0] = TP;
    JlConstants C; default_constants(&C);        /* the Julia glue fills this from PlanetOrbits' live constants instead */
    void* ctx = NULL;
    CHECK(create(&C, &L, &B, 1, 0, &ctx));
    if (total_epochs(ctx) != 8) return 4;

    /* ---- evaluate(::NTuple{K,Float64}) and evaluate(::NTuple{K,Dual}): single chain, n = 1, ld = 1 */
    double x[N_IN] = {50.01, 1.21, 12.1, 0.12, 0.72, 0.65, 0.29, 1.7, 41500.0}, ll1, ll2, g[N_IN];
    CHECK(logp(ctx, x, 1, 1, &ll1));
    CHECK(logp_grad(ctx, x, 1, 1, &ll2, g));
    printf("%a %a", ll1, ll2);
    for (int k = 0; k < N_IN; ++k) printf(" %a", g[k]);
    printf("\n");
    /* ---- logp_grad_begin / ready / wait_for: N x n_in column-major */
    enum { N = 3 };
    double X[N * N_IN], llN[N], G[N * N_IN];
    for (int c = 0; c < N; ++c) for (int k = 0; k < N_IN; ++k) X[c + k * N] = x[k] * (k == TP ? 1.0 : 1.0 + 0.003 * c);
    void* ticket = NULL;
    CHECK(logp_grad_begin(ctx, X, N, N, llN, G, &ticket));
    while (ready(ticket) == 0) { }
    CHECK(wait(ticket));
    for (int c = 0; c < N; ++c) printf("%a ", llN[c]);
    printf("\n");

    /* ---- DevicePosterior: priors in the reference's parameter order (system priors, then planet priors with
     *      UniformCircular expanded): M, plx, a, e, i, ωx, ωy, Ωx, Ωy, θx, θy  (D = 11) */
    enum { D = 11 };
    JlPrior pri[D]; memset(pri, 0, sizeof pri);
    pri[0].family = 4; pri[0].p[0] = 1.2;  pri[0].p[1] = 0.1;  pri[0].p[2] = 0.1; pri[0].p[3] = 1.0 / 0.0;     /* truncated(Normal(1.2, 0.1), lower = 0.1) */
    pri[1].family = 4; pri[1].p[0] = 50.0; pri[1].p[1] = 0.02; pri[1].p[2] = 0.1; pri[1].p[3] = 1.0 / 0.0;
    pri[2].family = 1; pri[2].p[0] = 0.0; pri[2].p[1] = 100.0;                                               /* a ~ Uniform(0, 100) */
    pri[3].family = 1; pri[3].p[0] = 0.0; pri[3].p[1] = 0.99;                                                /* e ~ Uniform(0, 0.99) */
    pri[4].family = 3;                                                                                       /* i ~ Sine() */
    for (int j = 5; j < D; ++j) { pri[j].family = 0; pri[j].p[0] = 0.0; pri[j].p[1] = 1.0; }                 /* Normal(0, 1) x 6 */
    JlInputDef def[N_IN]; memset(def, 0, sizeof def);
    def[PLX].op = 0; def[PLX].a[0] = 1;  def[M].op = 0; def[M].a[0] = 0;
    def[A_].op = 0; def[A_].a[0] = 2;    def[E_].op = 0; def[E_].a[0] = 3;   def[I_].op = 0; def[I_].a[0] = 4;
    const double two_pi = 6.283185307179586;
    def[W_].op = 2; def[W_].a[0] = 5; def[W_].a[1] = 6; def[W_].value = two_pi;                              /* UniformCircular: (x, y) */
    def[OM].op = 2; def[OM].a[0] = 7; def[OM].a[1] = 8; def[OM].value = two_pi;
    def[TH].op = 2; def[TH].a[0] = 9; def[TH].a[1] = 10; def[TH].value = two_pi;
    def[TP].op = 3; def[TP].value = 50000.0;                                                                 /* θ_at_epoch_to_tperi(θ, 50000; M, e, a, i, ω, Ω) */
    { int32_t args[7] = {TH, M, E_, A_, I_, W_, OM}; memcpy(def[TP].a, args, sizeof args); }
    CHECK(set_parameterization(ctx, pri, D, def));
    enum { NC = 16 };
    double th[NC * D], lp[NC], gt[NC * D];
    double th0[D] = {0.0953, 3.91, -1.99, -2.0, -0.5, 0.8, 0.6, 0.95, 0.28, -0.99, -0.13};
    for (int c = 0; c < NC; ++c) for (int j = 0; j < D; ++j) th[c + j * NC] = th0[j] + 0.01 * ((c * 7 + j * 3) % 11 - 5);
    CHECK(logpost_grad(ctx, th, NC, NC, lp, gt));
    printf("%a %a %a\n", lp[0], gt[0], gt[NC]);
    /* ---- hmc_run */
    double thf[NC * D], lpf[NC], acc[NC], im[D];
    for (int j = 0; j < D; ++j) im[j] = 1e-3;
    CHECK(hmc_run(ctx, th, NC, NC, 3, 4, 0.05, im, 42, NULL, NULL, thf, lpf, acc));
    printf("%a %a\n", thf[0], lpf[0]);
    /* ---- pt_init(world = 1) + pt_hmc_run(sharded = true) */
    double lad[NC], llf[NC], beta[NC], swaps[NC - 1], cold[5 * D];
    int32_t rung[NC];
    for (int c = 0; c < NC; ++c) lad[c] = (double)c / (NC - 1);
    CHECK(pt_init(ctx, NULL, 0, 1, NC, 7));
    CHECK(pt_hmc_run_dist(ctx, th, NC, NC, lad, 5, 1, 4, 0.05, im, 9, thf, lpf, llf, beta, rung, swaps, cold, acc));
    int perm = 0;
    for (int c = 0; c < NC; ++c) perm += rung[c];
    printf("%d %a\n", perm, cold[0]);
    destroy(ctx);
    return 0;
}
