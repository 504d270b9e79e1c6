// octo_internal.h — structures shared by the C-ABI shim (octo_shim.cu) and the kernels
// (octo_kernels.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/octo_b200.h"

#define OCTO_MAX_BLOCKS 24        // observation tables per model
#ifndef OCTO_WARPS
#define OCTO_WARPS 8              // warps per CTA: each warp owns one contiguous epoch range
#endif
#ifndef OCTO_LAT_WARPS
#define OCTO_LAT_WARPS 12         // warps per CTA of the latency-tuned instantiation (one CTA per SM, <= 168 registers); 6 / 8 / 10 / 12 / 14 / 16 measured: C2 12.0 / 9.8 / 9.8 / 8.9 / 9.6 / 9.2 us per step
#endif
#define OCTO_LANES 32             // lanes = chains of one chain group
#define OCTO_MIN_SLICE 2          // fewest epochs worth giving a warp

// Per chain*planet constants staged in shared memory by the prologue ([slot][lane] doubles).
enum PlanetConst {
    PC_nd = 0,   // mean motion [rad/day]
    PC_tp, PC_e, PC_ome /*1-e*/, PC_ca1 /*Markley alpha slope*/, PC_s /*sqrt(1-e^2)*/,
    PC_Bh, PC_Gs, PC_Ah, PC_Fs,       // scaled Thiele-Innes: ra = X*Bh + sinE*Gs, dec = X*Ah + sinE*Fs [mas]
    PC_Pc, PC_Ps,                     // rv = (Pc cosE - Ps sinE)/(1 - e cosE)  [m/s]
    PC_mu,                            // mass*mjup2msol/M, 0 when the planet has no mass variable
    PC_a,
    // epilogue only
    PC_sinW, PC_cosW /* = PC_sinW + 1 */, PC_sinw, PC_cosw, PC_sini, PC_cosi, PC_M, PC_plx, PC_sc /*a*c2a*/, PC_c2a,
    PC_K /*RV semi-amplitude*/, PC_Kb /*K / sin i*/,
    PC_A, PC_B, PC_F, PC_G,
    PC_inv_s, PC_inv_a, PC_inv_M,
    PC_COUNT
};

// Raw (pre chain-rule) accumulators per planet: sums over epochs that the epilogue maps to
// gradients w.r.t. (a, e, i, ω, Ω, tp, M, plx, mass).
enum PlanetAcc { PA_Bh = 0, PA_Gs, PA_Ah, PA_Fs, PA_Pc, PA_Ps, PA_e, PA_S0, PA_S1, PA_mu, PA_COUNT };

// extra accumulators of one marginalised-RV table (rv-absolute-margin.jl:161-181)
enum MarginAcc { MA_A = 0, MA_S1, MA_C, MA_LG, MA_R2, MA_R1, MA_Q, MA_COUNT };   // then 5 V-slots per planet
enum { MV_Pc = 0, MV_Ps, MV_e, MV_S0, MV_S1, MV_mu, MV_COUNT };
// accumulators of an astrometry table wrapped in ObsPriorAstromONeil2019: S = Σ|f|, then Σ sign(f) ∂f/∂(MA, MA·dt, e)
enum ObsPriorAcc { OP_S = 0, OP_Q0, OP_Q1, OP_Qe, OP_COUNT };

struct DevBlock {
    int32_t kind, planet, start, n;          // epoch range [start, start+n) in the concatenated tables
    int32_t has_cor, jit;                     // jit: jitter is an input column (per-pair variance path)
    int32_t idx_jitter, idx_platescale, idx_northangle, idx_offset;
    int32_t slot_jitter, slot_platescale, slot_northangle, slot_offset;   // accumulator slots (or -1)
    int32_t slot_margin;                      // first of the margin accumulators (kind 3) or -1
    int32_t slot_obsprior;                    // first of the 4 observable-prior accumulators (OP_*) or -1
    // RV tables: trend linear in n_trend observation variables (coefficient columns idx_trend, per-epoch basis values in the
    // record's spare slots y2, c2, c3); accumulator slot per variable (marginalised RV: two consecutive slots, Σ g b and Σ b/var)
    int32_t n_trend, idx_trend[3], slot_trend[3], pad_trend;
    double wgt, cum;                          // relative cost of one epoch of this table; Σ n*wgt of the tables before it
    double wgt_lat, cum_lat;                  // the same for latency-bound launches (cost = length of the dependent chain of one pair)
};

// HGCAInstantaneousObs (kind 5): a handful of rows evaluated by the last CTA of a chain group (octo_kernels.cu, hgca_tail)
#define OCTO_MAX_HGCA 2
struct DevHg {
    int32_t n_rows, row_off;          // rows [t, code] at tab + row_off (doubles); code 0 hip-ra, 1 hip-dec, 2 gaia-ra, 3 gaia-dec
    int32_t idx_pmra, idx_pmdec;
    int32_t slot_pmra, slot_pmdec;    // accumulator slots that carry d ll / d pmra, d ll / d pmdec to the epilogue
    double inv_N[4];                  // 1 / (planets x rows of that code): the reference's averaging (hgca.jl:247-290)
    double k_ra, k_dec;               // 365.25 / (mean Gaia epoch - mean Hipparcos epoch) of the RA / Dec rows
    double cat[3][2];                 // catalogue pmra, pmdec: Hipparcos, Hipparcos-Gaia, Gaia
    double w[3][3];                   // precision matrices w11, w12, w22
};

struct DevModel {
    OctoConstants c;
    double kappa;            // 2π * year2day / kepler_year_days * au2m * sec2year  (K = kappa * sqrt(M/a) * sin i / s)
    double c2a_per_plx;      // rad2as*1e3 / (1000*pc2au): mas per AU per mas of parallax
    double wtot;             // Σ n*wgt over all tables: warps split this, not the raw epoch count
    double wtot_lat;         // Σ n*wgt_lat
    double const_ll;         // Σ of the chain-independent normalisation terms of tables without free jitter
    int32_t n_planets, n_in, n_blocks, n_acc;
    int32_t has_margin, lean;    // any marginalised-RV or observable-prior table (their epilogue fold needs an extra barrier);
                                 // lean: only lean tables and Campbell planets — the kernels compiled without the rest (octo_kernels.cu)
    int64_t n_epochs;
    int32_t idx_plx[OCTO_MAX_PLANETS], idx_a[OCTO_MAX_PLANETS], idx_e[OCTO_MAX_PLANETS], idx_i[OCTO_MAX_PLANETS],
            idx_w[OCTO_MAX_PLANETS], idx_W[OCTO_MAX_PLANETS], idx_tp[OCTO_MAX_PLANETS], idx_M[OCTO_MAX_PLANETS],
            idx_mass[OCTO_MAX_PLANETS];
    // Thiele-Innes planets (basis 1): columns of A, B, F, G; their idx_a is the VIRTUAL gradient column n_in + p (the
    // semi-major axis is an intermediate there: a = alpha(A,B,F,G) / plx), idx_i / idx_w / idx_W are -1
    int32_t basis[OCTO_MAX_PLANETS], idx_A[OCTO_MAX_PLANETS], idx_B[OCTO_MAX_PLANETS], idx_F[OCTO_MAX_PLANETS],
            idx_G[OCTO_MAX_PLANETS];
    DevBlock blocks[OCTO_MAX_BLOCKS];
    int32_t n_hg, any_ti;         // any_ti: some planet uses the Thiele-Innes basis
    DevHg hg[OCTO_MAX_HGCA];
    // device table, one 48-byte record per epoch of the concatenated list: [t, y1, c1, y2, c2, c3]
    //   astrometry: c1,c2,c3 = w11,w12,w22 (no jitter) | σ1², σ2², cor (jitter);   RV: c1 = 1/σ² | σ², (y2,c2,c3) = trend basis values
    const double* tab;
};

// slot 0 = ll; planets follow
__host__ __device__ inline int slot_planet(int p, int a) { return 1 + p * PA_COUNT + a; }

// ---- device-side standard parameterisation (octo_param.cu)
#define OCTO_PARAM_MAX 64         // max D and max n_in of a parameterised model
#define OCTO_PARAM_TPERI_MAX 8    // more θ_at_epoch_to_tperi definitions than this: stand-alone K0 kernels instead of the fused stage
struct DevParam {
    int32_t D, n_in;
    OctoPrior priors[OCTO_PARAM_MAX];
    double pc[OCTO_PARAM_MAX][5];        // per-prior constants: lo, hi, constant part of the log density, 1/(hi-lo), 1/σ
    OctoInputDef defs[OCTO_PARAM_MAX];
    int32_t n_tperi, pad;                // θ_at_epoch_to_tperi definitions, ascending input index
    int32_t tperi_k[OCTO_PARAM_TPERI_MAX];
    // reverse map for the gradient: parameter j gathers entries gat[gat_start[j] .. gat_start[j+1]), last input first;
    // entry = input index | role << 8 (role 0: the parameter itself, 1: x of a UniformCircular pair, 2: y, 3: both)
    int16_t gat_start[OCTO_PARAM_MAX + 1];
    int16_t gat[2 * OCTO_PARAM_MAX];
    // evaluation orders of the fused stage, most expensive item first (warps take items round-robin, so the expensive
    // ones land on different warps in the first round): priors (invlink + log density), input definitions, gathers
    uint8_t order_prior[OCTO_PARAM_MAX], order_input[OCTO_PARAM_MAX], order_gather[OCTO_PARAM_MAX];
    // inputs whose sine and cosine the evaluation needs (inclination / ω / Ω of a planet, the four angles of a
    // θ_at_epoch_to_tperi definition): the fused stage of the lean kernels produces them together with the input itself —
    // for a UniformCircular pair with domain 2π as (y, x) / r, without atan2 and sincos (octo_kernels.cu, param_forward)
    uint8_t in_trig[OCTO_PARAM_MAX];
};

// Tiny batches (a single chain, as the reference's samplers call the model): the inputs travel inside the kernel
// parameters instead of through a host-to-device copy of their own (saves that copy's ~3 us of a ~26 us call).
#define OCTO_INLINE_MAX 64
#define OCTO_MODE_INLINE 256      // post_mode flag: read the inputs from the InlineIn parameter, column-major [n x cols], ld = n
struct InlineIn { double v[OCTO_INLINE_MAX]; };

// leapfrog update folded into the fused log-posterior launch (octo_hmc.cu): after the gradient of a chain is known,
// p += kick * g and, unless it was the last leapfrog of the trajectory, q += eps * p * inv_mass.  p == nullptr: off.
// beta != nullptr: tempering — the likelihood part of chain c is scaled by beta[c] (log posterior = prior terms +
// beta * ln_like, the path between Pigeons' prior-only reference and the target); ll_raw receives ln_like itself.
struct HmcLeap { double* p; double* q; const double* inv_mass; double eps, kick; int drift, pad; const double* beta; double* ll_raw; };
// lat: the latency-tuned instantiation.  ch: chains per CTA = 32 / sub-lanes (octo_kernels.cu, "SUB-LANES"); gx counts
// groups of ch chains; slice = epochs per (warp, sub-lane) unit
struct LaunchGeom { int gx, gy, block, slice; size_t smem; bool lat = false; int ch = 32; };

// trajectory-resident explorer (octo_kernels.cu, k_hmc_resident): arrays are column-major [n x D], chain fastest
struct ResidentArgs {
    double *q, *lp, *g;              // in/out: current states; out: their log posterior and gradient
    double* acc;                     // [n] accepted transitions, incremented
    const double* inv_mass;          // [D]
    double *out_theta, *out_lp;      // optional sample stores [.. x D x n], [.. x n], rows it0 .. it0 + n_iter - 1
    const double* beta;              // [n] tempering weights or nullptr
    double* ll;                      // [n] raw ln_like of the current states (tempering) or nullptr
    int64_t n, chain_offset;         // chains in this launch; global index of chain 0 (keys the random streams)
    int32_t D, n_iter, n_leapfrog, it0;
    double eps;
    uint64_t seed;
    double* pair_out;                // [n x 2] (l_ref, l_target) of the final states for a sharded swap round, or nullptr
};
// sharded parallel tempering (octo_hmc.cu): replicated rung assignment + the all-gathered pairs, all on the device
struct PtDist { double *pairs_all, *pairs_local, *ladder, *swap_acc; int32_t *chain_of_rung, *rung_of_chain; };
struct PtDistRun {
    PtDist d; int R; int64_t chain0;
    int (*allgather)(void* user, const double* d_send, double* d_recv, size_t count);   // enqueues on the run's stream
    void* user;
};
size_t octo_resident_smem_bytes(const DevModel& m, int D, int n_tperi);
cudaError_t octo_resident_init(const DevModel& m, size_t smem_optin);
cudaError_t octo_resident_launch(const DevModel& m, const DevParam* d_param, int n_tperi, const ResidentArgs& R, int ch, cudaStream_t st);

// kernels (octo_kernels.cu).  The file is compiled once per planet-count instantiation (1, 2, 4 planets per table loop);
// each object exports its entry points through one of these
struct OctoNptEntry {
    cudaError_t (*attr)(size_t smem_optin);
    cudaError_t (*occupancy)(const DevModel& m, int warps, size_t smem_bytes, int* ctas_per_sm);
    cudaError_t (*launch)(const DevModel& m, const LaunchGeom& g, bool grad, const double* d_in, int64_t n_chains, int64_t ld,
                          double* d_ll, double* d_g, int64_t ldg, double* d_partial, unsigned int* d_tickets,
                          const DevParam* d_param, int post_mode, const double* d_pw_const, const HmcLeap& leap,
                          cudaStream_t stream, const InlineIn* inl);
    cudaError_t (*resident)(const DevModel& m, const cudaLaunchConfig_t* cfg, const DevParam* d_param, const ResidentArgs& R,
                            int ch, int eval_doubles);
};
cudaError_t octo_launch(const DevModel& m, const LaunchGeom& g, bool grad, const double* d_in, int64_t n_chains,
                        int64_t ld, double* d_ll, double* d_g, int64_t ldg, double* d_partial,
                        unsigned int* d_tickets, const DevParam* d_param, int post_mode, const double* d_pw_const,
                        const HmcLeap& leap, cudaStream_t stream, const InlineIn* inl = nullptr);
size_t octo_smem_bytes(const DevModel& m, int warps, int D = 0, int n_tperi = 0);
cudaError_t octo_kernels_init(const DevModel& m, size_t smem_bytes, size_t smem_optin, int warps, int* ctas_per_sm);
cudaError_t octo_selftest_kepler_launch(const double* d_MA, const double* d_e, int64_t n, double* d_s, double* d_c);
cudaError_t octo_param_init(int D, int n_in);
cudaError_t octo_param_forward(const DevParam* d_param, int D, const DevModel& m, const double* d_theta, int64_t n,
                               int64_t ld, double* d_in, double* d_save, cudaStream_t st);
cudaError_t octo_param_backward(const DevParam* d_param, int D, const DevModel& m, int64_t n, const double* d_in,
                                const double* d_save, const double* d_ll, const double* d_g_in, double* d_lp,
                                double* d_g_t, int64_t ldg, int post_mode, cudaStream_t st);
cudaError_t octo_param_invlink(const DevParam* d_param, const double* d_theta, int64_t n, int64_t ld, double* d_out,
                               cudaStream_t st);

// device-resident HMC explorer (octo_hmc.cu)
size_t octo_hmc_state_doubles(int64_t n, int D);
cudaError_t octo_hmc_enqueue(double* d_state, int64_t n, int D, int n_iter, int n_leapfrog, double eps, uint64_t seed,
                             double* d_out_theta, double* d_out_lp, cudaStream_t st,
                             int (*logpost)(void*, const double*, double*, double*, const HmcLeap*), void* user,
                             bool fused_leap, int* rc_out, const double* h_ladder = nullptr, int n_rounds = 0,
                             double* d_cold = nullptr, int (*resident)(void*, const ResidentArgs*) = nullptr,
                             const PtDistRun* dist = nullptr);
cudaError_t octo_pt_swap_dist_launch(const PtDist& d, double* d_beta_local, int64_t chain0, int64_t n_local, int R, int64_t round,
                                     uint64_t seed, cudaStream_t st);
// after a tempered run: per-chain beta, rung of each chain, swap acceptance counts per adjacent pair (device pointers into the state)
void octo_hmc_pt_views(double* d_state, int64_t n, int D, double** beta, double** ll, int32_t** rung_of_chain, double** swap_acc);

