/*
 * octo_b200.h — C ABI of libocto_b200.so
 *
 * B200-native (sm_100a) replacement for ONE path of Octofitter.jl: the batched
 * Keplerian orbit solve + epoch-vectorised Gaussian log-likelihood and its
 * gradient, i.e. what the reference evaluates inside
 *     ln_like_generated(system, θ)            src/likelihoods/system.jl:21-242
 * for the observation types
 *     PlanetRelAstromObs                       src/likelihoods/relative-astrometry.jl:104-253
 *     StarAbsoluteRVObs (non-GP)               OctofitterRadialVelocity/src/rv-absolute.jl:135-204
 *     MarginalizedStarAbsoluteRVObs            OctofitterRadialVelocity/src/rv-absolute-margin.jl:106-185
 *     PlanetRelativeRVObs (non-GP)             OctofitterRadialVelocity/src/rv-relative.jl:121-211
 * on Visual{KepOrbit} orbits (PlanetOrbits.jl 0.11: KepOrbit ctor, orbitsolve,
 * kepler_solver(Markley), raoff/decoff/radvel), and what ForwardDiff computes
 * for the same terms in ∇ℓπcallback (src/logdensitymodel.jl:169-177).
 *
 * The caller (Julia `ccall`, or the ctypes mirror in octofitter.jl_b200/) keeps
 * priors, bijectors, derived variables and samplers; it hands this library a
 * matrix of NATURAL-space kernel inputs, one row per chain, and gets back the
 * epoch-summed log-likelihood per chain and its gradient w.r.t. those inputs.
 *
 * Conventions
 *  - All numbers are IEEE float64. `in` and `g_in` are column-major
 *    [n_chains x n_in] with leading dimension `ld` (>= n_chains): element
 *    (chain c, input k) is at in[c + k*ld]. This is Julia's native layout for
 *    an N x n_in Matrix{Float64} and makes chain the coalesced index on device.
 *  - Every entry point returns 0 on success, non-zero on error; the message is
 *    in octo_last_error() (thread-local). No C++ exception crosses the ABI.
 *  - Tables are validated at octo_create: non-finite epochs / data, uncertainties that are not finite and > 0, or
 *    |cor| > 1 - 1e-5 (the reference ctor's own check) are OCTO_ERR_ARG — they would make every chain NaN.
 *  - Numerical invalidity is NOT an error: a chain whose inputs are non-finite
 *    or have e outside [0,1), a <= 0, M <= 0 or plx <= 0 gets ll = -Inf and a
 *    zero gradient row (reference: non-finite θ => -Inf, logdensitymodel.jl:120-124;
 *    orbit ctor failure => -Inf, system.jl:214-221).
 *  - The library never falls back to the CPU: without a CUDA device
 *    octo_create fails with OCTO_ERR_CUDA.
 *  - Entry points are re-entrant. Concurrent calls on one OctoCtx are allowed
 *    (Pigeons calls the target from many threads, ext/OctofitterPigeonsExt:118);
 *    each call leases a private stream + workspace from the context's pool.
 */
#ifndef OCTO_B200_H
#define OCTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCTO_ABI_VERSION 4
#define OCTO_MAX_PLANETS 4

/* error codes */
#define OCTO_OK            0
#define OCTO_ERR_ARG       1   /* bad argument / unsupported layout            */
#define OCTO_ERR_CUDA      2   /* CUDA runtime error or no device              */
#define OCTO_ERR_NCCL      3   /* NCCL error / NCCL not loadable               */
#define OCTO_ERR_STATE     4   /* call sequence error (e.g. pt round w/o init) */

/*
 * Physical constants of PlanetOrbits.jl. The source of PlanetOrbits is not part
 * of the reference tree, so the Julia glue injects the live values
 * (PlanetOrbits.kepler_year_to_julian_day_conversion_factor, .year2day_julian,
 * .rad2as, .pc2au, .au2m, .sec2year_julian, .mjup2msol_IAU; names confirmed by
 * in-tree use: src/parameterizations.jl:28,62-63,215-216; src/Octofitter.jl:43).
 * octo_default_constants() fills in the recollected defaults.
 */
typedef struct OctoConstants {
    double kepler_year_days;  /* 365.2568983840419                         */
    double year2day;          /* 365.25                                    */
    double rad2as;            /* 206265                                    */
    double pc2au;             /* 206265                                    */
    double au2m;              /* 1.495978707e11                            */
    double sec2year;          /* 1/31557600                                */
    double mjup2msol;         /* 0.0009545942339693249                     */
} OctoConstants;

/* observation kinds */
#define OCTO_KIND_ASTROM_RADEC   0  /* PlanetRelAstromObs, (ra,dec,σ_ra,σ_dec[,cor])     */
#define OCTO_KIND_ASTROM_PASEP   1  /* PlanetRelAstromObs, (pa,sep,σ_pa,σ_sep[,cor])     */
#define OCTO_KIND_RV_STAR_ABS    2  /* StarAbsoluteRVObs, gaussian_process = nothing      */
#define OCTO_KIND_RV_STAR_MARGIN 3  /* MarginalizedStarAbsoluteRVObs                      */
#define OCTO_KIND_RV_PLANET_REL  4  /* PlanetRelativeRVObs, gaussian_process = nothing    */
#define OCTO_KIND_HGCA_INSTANT   5  /* HGCAInstantaneousObs (src/likelihoods/hgca.jl:29-417), Visual{KepOrbit} planets */

/*
 * One observation table (SoA, host pointers; copied at octo_create, not retained).
 *   kind 0: y1=ra  y2=dec  s1=σ_ra s2=σ_dec  cor optional       [mas]
 *   kind 1: y1=pa  y2=sep  s1=σ_pa s2=σ_sep  cor optional       [rad, mas]
 *   kind 2,3,4: y1=rv s1=σ_rv; y2,s2,cor must be NULL            [m/s]
 *   kind 5 (system-level, planet = -1): one row per simulated position measurement, as the reference ctor lays them
 *           out (hgca.jl:94-110): epoch [MJD], y1 = 0 Hipparcos-RA | 1 Hipparcos-Dec | 2 Gaia-RA | 3 Gaia-Dec; every
 *           other column NULL.  aux[15] = for Hipparcos, Hipparcos-Gaia, Gaia in turn:
 *           { pmra, pmdec, pmra_error * factor, pmdec_error * factor, pmra_pmdec correlation }   [mas/yr].
 *           Every planet needs a mass variable (hgca.jl:271, 275).  The rows are not part of octo_total_epochs and
 *           octo_logp_pointwise ignores them (a one-row subset of this likelihood is 0/0 in the reference as well).
 * Rows must already be in the order the reference holds them (the astrometry
 * ctor sorts by epoch, relative-astrometry.jl:46-47).
 * idx_*: column of the per-observation variable in the input matrix, or -1 for
 * the reference default (jitter 0, platescale 1, northangle 0, offset 0;
 * relative-astrometry.jl:170-172, rv-absolute.jl:139,181). Kind 3 requires
 * idx_jitter >= 0 (rv-absolute-margin.jl:149 reads θ_obs.jitter unconditionally).
 * trend_function: identically zero by default; trends linear in the observation variables via n_trend / trend_basis below.
 */
typedef struct OctoObsBlock {
    int32_t kind;
    int32_t planet;        /* 0-based planet index for kinds 0,1,4; -1 for system-level kinds 2,3 */
    int32_t n_epochs;
    int32_t has_cor;
    const double* epoch;
    const double* y1;
    const double* y2;
    const double* s1;
    const double* s2;
    const double* cor;
    int32_t idx_jitter;
    int32_t idx_platescale;
    int32_t idx_northangle;
    int32_t idx_offset;
    int32_t obs_prior;     /* 1 (astrometry kinds only): the table is wrapped in ObsPriorAstromONeil2019 — on top of
                            * the table's ln_like add 2 log( Σ_epochs |3M(e+cosE) + 2(-2+e²+e cosE) sinE| · cbrt(P)/√(1-e²) )
                            * for the observed planet (src/likelihoods/prior-observable.jl:78-137) */
    int32_t idx_pmra;      /* kind 5 only: columns of the system proper motion (θ_system.pmra, .pmdec) [mas/yr] */
    int32_t idx_pmdec;
    int32_t reserved;
    const double* aux;     /* kind 5 only: the 15 catalogue numbers, see OCTO_KIND_HGCA_INSTANT */
    /* kinds 2, 3, 4: `trend_function(θ_obs, epoch)` (rv-absolute.jl:143, rv-absolute-margin.jl:111, rv-relative.jl:131)
     * when it is LINEAR in the observation variables — polynomials in time with coefficient variables, fixed-period
     * sinusoids with amplitude variables, the docs' `θ_obs.trend_slope * (epoch - 57000)`:
     *     trend(θ_obs, t_k) = trend_const[k] + Σ_v in[idx_trend[v]] * trend_basis[v * n_epochs + k],  v < n_trend <= 3
     * trend_const may be NULL (zero).  n_trend = 0 and trend_const = NULL: the default zero trend.  The caller obtains
     * the basis by probing the closure (unit vectors in θ_obs) and must keep anything non-linear in Julia. */
    int32_t n_trend;
    int32_t idx_trend[3];
    const double* trend_basis;
    const double* trend_const;
} OctoObsBlock;

/*
 * Where each orbital element of each planet lives in the input matrix.
 * Orbit kwargs are merge(θ_system, θ_planet) (system.jl:117), so plx and M are
 * per planet (they usually point at the same system-level column).
 * idx_mass = -1 means the planet has no `mass` variable: it contributes no
 * reflex motion (relative-astrometry.jl:121-123); kinds 2/3 then fail at create
 * (the reference would throw on θ_system.planets[i].mass, rv-absolute.jl:147).
 */
typedef struct OctoLayout {
    int32_t n_planets;
    int32_t n_in;
    int32_t idx_plx [OCTO_MAX_PLANETS];
    int32_t idx_a   [OCTO_MAX_PLANETS];
    int32_t idx_e   [OCTO_MAX_PLANETS];
    int32_t idx_i   [OCTO_MAX_PLANETS];
    int32_t idx_w   [OCTO_MAX_PLANETS];   /* ω  argument of periastron      */
    int32_t idx_W   [OCTO_MAX_PLANETS];   /* Ω  longitude of ascending node */
    int32_t idx_tp  [OCTO_MAX_PLANETS];
    int32_t idx_M   [OCTO_MAX_PLANETS];
    int32_t idx_mass[OCTO_MAX_PLANETS];   /* Mjup, or -1 */
    /* Orbit basis per planet: 0 = Visual{KepOrbit} (Campbell elements above), 1 = ThieleInnesOrbit
     * (docs/src/thiele-innes.md; PlanetOrbits): the planet's inputs are e, tp, M, plx and the Thiele-Innes
     * constants A, B, F, G in mas (columns idx_A.. below; idx_a, idx_i, idx_w, idx_W are ignored).  The semi-major
     * axis that sets the period is a = sqrt(u + sqrt((u+v)(u-v))) / plx, u = (A²+B²+F²+G²)/2, v = AG - BF
     * (src/parameterizations.jl:14-18); ra = xB + yG, dec = xA + yF with x = cos E - e, y = sqrt(1-e²) sin E
     * (:346-353).  Radial-velocity tables cannot be combined with a Thiele-Innes planet (not offloaded). */
    int32_t basis   [OCTO_MAX_PLANETS];
    int32_t idx_A   [OCTO_MAX_PLANETS];
    int32_t idx_B   [OCTO_MAX_PLANETS];
    int32_t idx_F   [OCTO_MAX_PLANETS];
    int32_t idx_G   [OCTO_MAX_PLANETS];
} OctoLayout;
#define OCTO_BASIS_CAMPBELL     0
#define OCTO_BASIS_THIELE_INNES 1

typedef struct OctoCtx OctoCtx;

void octo_default_constants(OctoConstants* out);
int  octo_abi_version(void);

/*
 * Build a context: validates the layout, copies the observation tables to
 * `device` (CUDA ordinal), precomputes per-epoch weights, sizes the launch.
 * Replaces the code generation of make_ln_like (system.jl:21-242): the epoch
 * list is the concatenation, in `blocks` order, of every table's epochs.
 */
int  octo_create(const OctoConstants* consts, const OctoLayout* layout,
                 const OctoObsBlock* blocks, int32_t n_blocks, int32_t device,
                 OctoCtx** out);
void octo_destroy(OctoCtx* ctx);

/* value only (K1v) — what ℓπcallback adds at logdensitymodel.jl:134; HOST buffers */
int  octo_logp(OctoCtx* ctx, const double* in, int64_t n_chains, int64_t ld, double* ll);
/* value + gradient w.r.t. the inputs (K1) — replaces the ForwardDiff pass over
 * the epoch loop (logdensitymodel.jl:169-177); HOST buffers */
int  octo_logp_grad(OctoCtx* ctx, const double* in, int64_t n_chains, int64_t ld,
                    double* ll, double* g_in);

/*
 * The same two calls split in two, so that the caller overlaps its own host work with the GPU: the Julia glue evaluates
 * the "rest" model (priors, user likelihoods: ForwardDiff on the CPU) between begin and wait, and a batch of independent
 * evaluations keeps several tickets in flight (copy-in of one overlaps the kernel of another; each ticket has a stream
 * of its own).  begin enqueues copy-in, kernel and copy-out and returns at once; `in` must stay untouched and `ll` /
 * `g_in` unread until octo_wait(ticket) returned (it also frees the ticket).  octo_ready: 1 finished, 0 still running,
 * negative on error — never blocks.  g_in may be NULL (value only).  Buffers from octo_alloc_pinned avoid staging.
 */
typedef struct OctoTicket OctoTicket;
int  octo_logp_grad_begin(OctoCtx* ctx, const double* in, int64_t n_chains, int64_t ld, double* ll, double* g_in,
                          OctoTicket** ticket);
int  octo_ready(OctoTicket* ticket);
int  octo_wait(OctoTicket* ticket);

/*
 * ---- Standard parameterisation on the device ------------------------------------------------
 * Optional: describe how the sampler's unconstrained vector θ_t (length D) maps to natural-space parameters, their
 * priors, and the kernel inputs, and the library evaluates the whole log-posterior of the standard model families
 *     ℓπ(θ_t) = Σ_j logpdf_with_trans(prior_j, invlink_j(θ_t[j]))      src/variables.jl:1205-1369, 1449-1493
 *             + Σ UnitLengthPrior terms of UniformCircular variables    src/variables.jl:267-323
 *             + ln_like(inputs(θ))                                      the path above
 * and its gradient w.r.t. θ_t on the device, with no per-chain host work.  Prior families and bijectors follow
 * Distributions.jl / Bijectors.jl (third-party, not in the reference tree; formulas in SURVEY.md Appendix B).
 */
#define OCTO_PRIOR_NORMAL      0   /* p = {mu, sigma}                     support R        : identity          */
#define OCTO_PRIOR_UNIFORM     1   /* p = {a, b}                          [a, b]           : scaled logit      */
#define OCTO_PRIOR_LOGUNIFORM  2   /* p = {a, b}                          [a, b]           : scaled logit      */
#define OCTO_PRIOR_SINE        3   /* src/distributions.jl:15-40          [eps, pi - eps]  : scaled logit      */
#define OCTO_PRIOR_TRUNCNORMAL 4   /* p = {mu, sigma, lower, upper} (+-Inf allowed): log-shift / scaled logit    */
typedef struct OctoPrior {
    int32_t family;
    int32_t reserved;
    double  p[4];
} OctoPrior;

#define OCTO_IN_PARAM 0   /* input = theta[a[0]]                                                                  */
#define OCTO_IN_CONST 1   /* input = value                                                                        */
#define OCTO_IN_CIRC  2   /* UniformCircular: input = atan(theta[a[1]], theta[a[0]]) / 2pi * value, and the
                           * UnitLengthPrior LogNormal(0, 0.1) on hypot(theta[a[0]], theta[a[1]]) is added         */
#define OCTO_IN_TPERI 3   /* tp = θ_at_epoch_to_tperi(in[a[0]], value; M=in[a[1]], e=in[a[2]], a=in[a[3]],
                           * i=in[a[4]], ω=in[a[5]], Ω=in[a[6]])  (src/parameterizations.jl:6-69); the arguments
                           * are EARLIER kernel inputs (their own definitions may be PARAM / CONST / CIRC)         */
#define OCTO_IN_TPERI_TI 4 /* the same for a ThieleInnesOrbit planet: tp = θ_at_epoch_to_tperi(in[a[0]], value; M=in[a[1]],
                           * e=in[a[2]], plx=in[a[3]], A=in[a[4]], B=in[a[5]], F=in[a[6]], G=in[a[7]])
                           * (src/parameterizations.jl:9-19)                                                      */
typedef struct OctoInputDef {
    int32_t op;
    int32_t a[8];
    double  value;
} OctoInputDef;

/* Attach the parameterisation: D priors (one per entry of θ_t, in the reference's parameter order) and one
 * definition per kernel input (n_in of them, evaluated in index order).  Copied; may be called again to replace. */
int  octo_set_parameterization(OctoCtx* ctx, const OctoPrior* priors, int32_t D, const OctoInputDef* defs);
/* θ_t: HOST column-major [n_chains x D] (leading dimension ld); lp[n_chains]; g_t [n_chains x D] or NULL.
 * What `ℓπcallback` / `∇ℓπcallback` return (src/logdensitymodel.jl:110-146, 169-177) for such a model. */
int  octo_logpost_grad(OctoCtx* ctx, const double* theta_t, int64_t n_chains, int64_t ld, double* lp, double* g_t);
/* asynchronous half of octo_logpost_grad (see octo_logp_grad_begin); finish with octo_wait */
int  octo_logpost_grad_begin(OctoCtx* ctx, const double* theta_t, int64_t n_chains, int64_t ld, double* lp, double* g_t,
                             OctoTicket** ticket);
/* Same on DEVICE buffers, enqueued on `stream`; d_work is caller-provided scratch of
 * octo_logpost_workspace(ctx, n_chains) bytes.  Normally the parameterisation runs inside the likelihood kernel
 * (one launch; the workspace size is 0 and d_work may be NULL); models the fused stage cannot hold (more than 8
 * θ_at_epoch_to_tperi definitions, a tperi that depends on another one, shared memory) use three launches that
 * hand their intermediates through d_work. */
int64_t octo_logpost_workspace(const OctoCtx* ctx, int64_t n_chains);
int  octo_logpost_grad_device(OctoCtx* ctx, const double* d_theta_t, int64_t n_chains, int64_t ld, double* d_lp,
                              double* d_g_t, void* d_work, void* stream);
/* The likelihood part alone for a batch of θ_t: ln_like(system, arr2nt(invlink(θ_t))), UnitLengthPrior terms
 * included, -Inf where it is not finite.  What `octofit_rejection` evaluates per prior draw
 * (src/sampling.jl:168-279, _rejection_evaluate_likelihoods :261-270).  HOST buffers, value only. */
int  octo_loglike_theta(OctoCtx* ctx, const double* theta_t, int64_t n_chains, int64_t ld, double* ll);
/* Per-epoch log-likelihoods of natural-space kernel inputs: out[chain + e * ldo] (e = 0 .. octo_total_epochs - 1, in
 * the order the tables were passed to octo_create) is ln_like of the model reduced to that single epoch, i.e. one
 * column per system of `generate_system_per_epoch` — the matrix `pointwise_like` builds
 * (src/cross-validation.jl:6-49, 453-497).  HOST buffers; n_chains * epochs <= 2e9 per call. */
int  octo_logp_pointwise(OctoCtx* ctx, const double* in, int64_t n_chains, int64_t ld, double* out, int64_t ldo);
/* A device-resident, chain-batched static-trajectory HMC explorer over the log-posterior launch: n_iter transitions of
 * n_leapfrog leapfrog steps (step_size, diagonal inverse mass inv_mass[D] or NULL = identity) for all n_chains in
 * lockstep, everything enqueued on one stream (one small update kernel + one log-posterior launch per leapfrog), one
 * synchronisation at the end.  The reference advances one chain at a time with AdvancedHMC (src/sampling.jl:412-423);
 * this is the batched counterpart of its leapfrog loop, not a NUTS implementation.  Randomness is counter-based
 * (octo_hmc_random reproduces it on the host): a run is a pure function of its arguments.
 * HOST buffers: theta0 / theta_final column-major [n_chains x D] (ld); optional stores theta_samples
 * [n_iter][D][n_chains], lp_samples [n_iter][n_chains]; lp_final[n_chains]; accept_rate[n_chains]. */
int  octo_hmc_run(OctoCtx* ctx, const double* theta0, int64_t n_chains, int64_t ld, int32_t n_iter, int32_t n_leapfrog,
                  double step_size, const double* inv_mass, uint64_t seed, double* theta_samples, double* lp_samples,
                  double* theta_final, double* lp_final, double* accept_rate);
/* Device-resident parallel tempering on top of the same explorer (the exploration phase of Pigeons' non-reversible PT,
 * ext/OctofitterPigeonsExt: reference = the prior-only model, i.e. the prior terms of this model; target = the model):
 * chain c starts on rung c of `ladder` (weights in [0, 1], ladder[n_chains-1] = 1 by convention); the tempered density
 * of a chain is prior terms + beta * ln_like.  n_rounds rounds of { n_iter tempered HMC transitions, one deterministic
 * even-odd swap round (the decisions of octo_pt_decide, same random stream), re-evaluation at the new weights }, all
 * on one stream.  Swaps exchange rungs, not states.  Outputs (HOST, each may be NULL): final states, their tempered
 * log posterior, raw ln_like, weight and rung per chain, swap acceptance counts per adjacent pair [n_chains - 1], the
 * state of the chain on the last rung after every round [n_rounds x D], HMC acceptance rate per chain. */
int  octo_pt_hmc_run(OctoCtx* ctx, const double* theta0, int64_t n_chains, int64_t ld, const double* ladder, int32_t n_rounds,
                     int32_t n_iter, int32_t n_leapfrog, double step_size, const double* inv_mass, uint64_t seed,
                     double* theta_final, double* lp_final, double* ll_final, double* beta_final, int32_t* rung_final,
                     double* swap_accept, double* cold_samples, double* accept_rate);
/* the random numbers of transition `it` of chain `chain`: D standard normals (momentum, before the mass scaling) and the
 * uniform of the accept step */
void octo_hmc_random(uint64_t seed, int64_t it, int64_t chain, int32_t D, double* z, double* u);
/* invlink only: natural-space parameters [n_chains x D] for a batch of θ_t (HOST buffers). */
int  octo_invlink(OctoCtx* ctx, const double* theta_t, int64_t n_chains, int64_t ld, double* theta_nat);

/* Page-locked host memory for the HOST-buffer entry points.  Buffers that come from octo_alloc_pinned
 * need no staging copy: outputs are written by the kernel straight into them, inputs of up to 512 KB (environment
 * OCTO_B200_ZEROCOPY_MAX, bytes) are read by the kernel in place over PCIe and larger ones are copied to the device by
 * the copy engine.  Like any asynchronous copy: do not touch the buffers of an octo_*_begin call before its octo_wait.
 * Any other host pointer works too and is staged through the context's own pinned buffers.  Returns NULL on failure
 * (see octo_last_error). */
void* octo_alloc_pinned(size_t bytes);
void  octo_free_pinned(void* p);

/* Same, DEVICE buffers on the context's device, enqueued on `stream`
 * (a cudaStream_t; NULL = the legacy default stream). Asynchronous: the caller
 * synchronises. g_in may be NULL for value only. */
int  octo_logp_grad_device(OctoCtx* ctx, const double* d_in, int64_t n_chains, int64_t ld,
                           double* d_ll, double* d_g_in, void* stream);

/* The device-buffer entry points keep a small workspace (cross-CTA partials, tickets) per caller stream; calls on one
 * stream may come from several host threads (enqueueing is serialised per stream).  Release it before destroying the
 * stream — a recycled stream handle would inherit it.  Waits for the stream's pending work. */
int  octo_release_stream(OctoCtx* ctx, void* stream);

/* Diagnostic: run only the device Kepler solve (rem2pi + Markley) on n (mean anomaly, e) pairs
 * and return sin E, cos E.  HOST buffers.  Used by the test-suite to probe e -> 1, M -> 0, |M| = pi. */
int  octo_selftest_kepler(int32_t device, const double* MA, const double* e, int64_t n,
                          double* sinE, double* cosE);

/* introspection */
int32_t octo_n_in(const OctoCtx* ctx);
int32_t octo_n_planets(const OctoCtx* ctx);
int64_t octo_total_epochs(const OctoCtx* ctx);   /* E: length of the concatenated epoch list      */
int32_t octo_device(const OctoCtx* ctx);
int64_t octo_kernel_launches(const OctoCtx* ctx); /* kernels launched so far through this context */
/* launch geometry chosen for a batch of n_chains: {grid.x = chain groups, grid.y = epoch splits across CTAs, block,
 * epochs per (warp, sub-lane) unit, sub-lanes per chain inside a warp (1 = lane is chain; a chain group has 32 / that
 * many chains), 1 if the latency-tuned instantiation (one CTA per SM) runs it} */
int  octo_launch_geometry(const OctoCtx* ctx, int64_t n_chains, int32_t out[6]);

/*
 * Parallel-tempering swap round (replaces the replica exchange Pigeons does over
 * threads / MPI, ext/OctofitterPigeonsExt/OctofitterPigeonsExt.jl:76-128).
 * Replicas are block-partitioned over `world` ranks (one process per GPU).
 * Each replica contributes (ℓ_ref, ℓ_target); one ncclAllGather of
 * n_replicas x 2 float64 per round; every rank then computes the same
 * deterministic even/odd swap decisions and swaps β INDICES, not states.
 * `nccl_unique_id` is the 128-byte ncclUniqueId (rank 0: octo_pt_unique_id).
 */
int  octo_pt_unique_id(void* out128);
int  octo_pt_init(OctoCtx* ctx, const void* nccl_unique_id, int32_t rank, int32_t world,
                  int32_t n_replicas_local, uint64_t seed);
/*
 * ll_pair: HOST [n_replicas_local x 2] row-major (l_ref, l_target) of this rank's replicas (l_target from
 *          octo_logp, l_ref from the caller's prior-only model, ext/OctofitterPigeonsExt:61-67).
 * beta: HOST [n_replicas_total] ladder, ascending.
 * chain_of_replica: HOST in/out [n_replicas_total] — which ladder rung each replica holds (a permutation).
 * round: swap round counter (even rounds pair rungs (0,1),(2,3).., odd rounds (1,2),(3,4)..).
 * accepted: HOST out [n_replicas_total-1] 0/1 per adjacent pair (may be NULL).
 * Collective: every rank calls it with the same beta / chain_of_replica / round.
 */
int  octo_pt_swap_round(OctoCtx* ctx, const double* ll_pair, const double* beta,
                        int32_t* chain_of_replica, int64_t round, int32_t* accepted);
/*
 * The same round in stream order, nothing returns to the host (the "K3 fused with ncclAllGather on the same stream" of
 * SURVEY.md §2): d_ll_pair_local [n_replicas_local x 2] is on the DEVICE (written by the caller's evaluation on
 * `stream`); one ncclAllGather on `stream`, then one kernel in which every rank takes every decision.  The rung
 * assignment lives on the device, replicated on every rank, and is updated in place: d_ladder [R] weights per rung
 * (ascending), d_chain_of_rung [R] / d_rung_of_chain [R] inverse permutations (start: identity), d_swap_count [R]
 * accepted swaps per adjacent pair (accumulated).  d_beta_local [n_replicas_local] (may be NULL) receives the new
 * weight of every local replica that swapped.  Same decisions as octo_pt_decide given the same values and seed.
 */
int  octo_pt_swap_round_device(OctoCtx* ctx, const double* d_ll_pair_local, const double* d_ladder,
                               int32_t* d_chain_of_rung, int32_t* d_rung_of_chain, double* d_swap_count,
                               double* d_beta_local, int64_t round, void* stream);
/*
 * octo_pt_hmc_run with the ladder SHARDED over the ranks of octo_pt_init (one process per GPU, replicas block-
 * partitioned: this rank's n_local chains are chains [rank n_local, (rank + 1) n_local) of R = world n_local).
 * Collective.  Per round and rank, all on one stream with no host involvement: one launch of the trajectory-resident
 * explorer on the local chains (it also packs their (l_ref, l_target)), ONE ncclAllGather of R x 2 float64, one
 * decision kernel (every rank decides every pair and re-weights its own chains), the record of the last rung.
 * ladder_all [R]; theta0 / outputs are this rank's chains, except swap_accept [R - 1] and cold_samples [n_rounds x D]
 * (identical on every rank; the rows are summed over ranks at the end of the run: one ncclAllReduce).  With the same
 * seed the swap history and every chain are bit-identical to octo_pt_hmc_run of all R replicas on one GPU (chain groups
 * are sized by R, and the result of a chain does not depend on its neighbours).
 */
int  octo_pt_hmc_run_dist(OctoCtx* ctx, const double* theta0_local, int64_t n_local, int64_t ld, const double* ladder_all,
                          int32_t n_rounds, int32_t n_iter, int32_t n_leapfrog, double step_size, const double* inv_mass,
                          uint64_t seed, double* theta_final, double* lp_final, double* ll_final, double* beta_final,
                          int32_t* rung_final, double* swap_accept, double* cold_samples, double* accept_rate);
/* The decision step alone (pure host code, no CUDA/NCCL): ll_pair_all is [n_replicas_total x 2]. */
int  octo_pt_decide(const double* ll_pair_all, const double* beta, int32_t* chain_of_replica,
                    int32_t n_replicas_total, int64_t round, uint64_t seed, int32_t* accepted);
void octo_pt_finalize(OctoCtx* ctx);

const char* octo_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* OCTO_B200_H */
