// octo_hmc.cu — a device-resident, chain-batched static-trajectory HMC explorer over the fused log-posterior launch
// (SURVEY.md §8f N2).  The reference advances ONE chain at a time with AdvancedHMC (src/sampling.jl:412-423); here
// all chains of a batch move in lockstep and nothing returns to the host between the first and the last launch of a
// run: per leapfrog one small update kernel + one log-posterior launch, chained on one stream with programmatic
// dependent launch.  Randomness is counter-based (splitmix64 of seed, iteration, chain, coordinate): a run is a pure
// function of its arguments.  Arrays are column-major [n_chains x D] (chain fastest), one thread per chain.
#include <math_constants.h>

#include <vector>

#include "octo_internal.h"
#include "octo_hmc_dev.cuh"

namespace {

using namespace octo_hmc_dev;

__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct HmcBuf {
    double *q, *lp, *g;              // current state, its log posterior and gradient
    double *qp, *lpp, *gp;           // proposal (what the log-posterior launch reads and writes)
    double *p, *h0;                  // momentum, initial Hamiltonian
    double *inv_mass;                // [D]
    double *acc;                     // [n] accepted transitions
    double *out_theta, *out_lp;      // optional sample store [n_iter][n x D], [n_iter][n]
    // parallel tempering (one chain per rung of the ladder): tempering weight and raw ln_like of every chain (current
    // state and proposal), the ladder itself, who sits where, swap acceptance counts per adjacent pair
    double *beta, *ll, *llp, *ladder, *swap_acc;
    int32_t *rung_of_chain, *chain_of_rung;
};

// accept/reject transition `it - 1` (it > 0), record the sample, then start transition `it` (it < n_iter):
// fresh momentum, H0, first half kick and drift.
__global__ void k_hmc_turn(HmcBuf b, int64_t n, int D, int it, int n_iter, double eps, uint64_t seed) {
    pdl_sync();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (it > 0) {
        const double lpp = b.lpp[c];
        double kin = 0.0;
        for (int j = 0; j < D; ++j) kin = __dadd_rn(kin, hmc_kin(b.p[c + (int64_t)j * n], b.inv_mass[j]));
        if (hmc_accept(seed, it - 1, (uint64_t)c, D, b.h0[c], lpp, kin)) {
            for (int j = 0; j < D; ++j) { b.q[c + (int64_t)j * n] = b.qp[c + (int64_t)j * n]; b.g[c + (int64_t)j * n] = b.gp[c + (int64_t)j * n]; }
            b.lp[c] = lpp; b.acc[c] += 1.0;
            if (b.ll) b.ll[c] = b.llp[c];
        }
        if (b.out_lp) b.out_lp[(int64_t)(it - 1) * n + c] = b.lp[c];
        if (b.out_theta) for (int j = 0; j < D; ++j) b.out_theta[((int64_t)(it - 1) * D + j) * n + c] = b.q[c + (int64_t)j * n];
    }
    if (it >= n_iter) return;
    const uint64_t key = hmc_key(seed, it);
    double kin = 0.0;
    for (int j = 0; j < D; ++j) {
        const HmcStart r = hmc_start(key, (uint64_t)c, j, b.inv_mass[j], eps, b.g[c + (int64_t)j * n], b.q[c + (int64_t)j * n]);
        kin = __dadd_rn(kin, r.kin);
        b.p[c + (int64_t)j * n] = r.p;
        b.qp[c + (int64_t)j * n] = r.q;
    }
    b.h0[c] = hmc_h0(b.lp[c], kin);
}

// after a log-posterior launch on the proposal: (half) kick, and drift unless it was the last leapfrog
__global__ void k_hmc_leap(HmcBuf b, int64_t n, int D, int last, double eps) {
    pdl_sync();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const bool ok = isfinite(b.lpp[c]);
    const double k = last ? 0.5 * eps : eps;
    for (int j = 0; j < D; ++j) {
        const double pj = fma(k, ok ? b.gp[c + (int64_t)j * n] : 0.0, b.p[c + (int64_t)j * n]);    // same roundings as the fused form
        b.p[c + (int64_t)j * n] = pj;
        if (!last) b.qp[c + (int64_t)j * n] = fma(eps * pj, b.inv_mass[j], b.qp[c + (int64_t)j * n]);
    }
}

// One deterministic even-odd swap round (Pigeons' non-reversible scheme, as octo_pt_decide takes it on the host): pair i
// = rungs (i, i+1), i of the round's parity.  With l_ref = lp - beta ll (prior terms) and l_target = l_ref + ll the
// tempered density of a chain on rung k is (1 - beta_k) l_ref + beta_k l_target.  Swaps exchange rungs, not states;
// afterwards the tempered log posterior of a swapped chain is stale and its gradient too: the caller re-evaluates.
__global__ void k_pt_swap(HmcBuf b, int64_t n, int D, int64_t round, uint64_t seed, double* cold_out) {
    pdl_sync();
    const int i = (int)(round & 1) + 2 * (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1 < n) {
        const int ca = b.chain_of_rung[i], cb = b.chain_of_rung[i + 1];
        const double bi = b.ladder[i], bj = b.ladder[i + 1];
        double ra, ta, rb, tb;                                                 // l_ref, l_target of the two chains
        pt_pair(b.lp[ca], b.beta[ca], b.ll[ca], ra, ta);
        pt_pair(b.lp[cb], b.beta[cb], b.ll[cb], rb, tb);
        const double log_ratio = pt_log_ratio(ra, ta, rb, tb, bi, bj);
        const double u = pt_uniform_dev(seed, (uint64_t)round, (uint64_t)i);
        const bool acc = isfinite(log_ratio) ? (log(u) < log_ratio) : (log_ratio > 0);
        if (acc) {
            b.chain_of_rung[i] = cb; b.chain_of_rung[i + 1] = ca;
            b.rung_of_chain[ca] = i + 1; b.rung_of_chain[cb] = i;
            b.beta[ca] = bj; b.beta[cb] = bi;
            b.swap_acc[i] += 1.0;
        }
    }
}
// the state of the chain on the last rung (beta = 1 by convention) after a round
__global__ void k_pt_record(HmcBuf b, int64_t n, int D, int64_t round, double* cold_out) {
    pdl_sync();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const int c = b.chain_of_rung[n - 1];
    cold_out[round * D + j] = b.q[c + (int64_t)j * n];
}

// ---- the ladder sharded over ranks (one process per GPU; octo_pt_hmc_run_dist): every rank holds n_local chains
// [chain0, chain0 + n_local) of the R, the all-gathered (l_ref, l_target) pairs of all R, and a replica of the rung
// assignment.  Every rank takes all the decisions (same inputs, same arithmetic, same counter-based uniforms => same
// outcome everywhere, no second exchange) and updates the weights of its own chains.
__global__ void k_pt_swap_dist(PtDist d, double* beta_local, int64_t chain0, int64_t n_local, int R, int64_t round, uint64_t seed) {
    pdl_sync();
    const int i = (int)(round & 1) + 2 * (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1 < R) {
        const int ca = d.chain_of_rung[i], cb = d.chain_of_rung[i + 1];
        const double bi = d.ladder[i], bj = d.ladder[i + 1];
        const double log_ratio = pt_log_ratio(d.pairs_all[2 * ca], d.pairs_all[2 * ca + 1], d.pairs_all[2 * cb], d.pairs_all[2 * cb + 1], bi, bj);
        const double u = pt_uniform_dev(seed, (uint64_t)round, (uint64_t)i);
        const bool acc = isfinite(log_ratio) ? (log(u) < log_ratio) : (log_ratio > 0);
        if (acc) {
            d.chain_of_rung[i] = cb; d.chain_of_rung[i + 1] = ca;
            d.rung_of_chain[ca] = i + 1; d.rung_of_chain[cb] = i;
            if (ca >= chain0 && ca < chain0 + n_local) beta_local[ca - chain0] = bj;
            if (cb >= chain0 && cb < chain0 + n_local) beta_local[cb - chain0] = bi;
            d.swap_acc[i] += 1.0;
        }
    }
}
// the state of the chain on the last rung after a round: written by the rank that holds it, zero elsewhere (summed over
// ranks at the end of the run)
__global__ void k_pt_record_dist(PtDist d, const double* q_local, int64_t chain0, int64_t n_local, int R, int D, int64_t round,
                                 double* cold_out) {
    pdl_sync();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const int c = d.chain_of_rung[R - 1];
    const bool mine = c >= chain0 && c < chain0 + n_local;
    cold_out[round * D + j] = mine ? q_local[(c - chain0) + (int64_t)j * n_local] : 0.0;
}

template <class... Args>
cudaError_t launch_pdl(void (*kern)(Args...), int64_t n, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((n + 127) / 128)); cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace

// d_state: 2 * (2 n D + n) + n D + 3 n + D doubles laid out as HmcBuf expects (see octo_hmc_state_doubles)
size_t octo_hmc_state_doubles(int64_t n, int D) { return (size_t)(2 * (2 * n * D + n) + n * D + 2 * n + D) + 6 * (size_t)n + 2; }

static HmcBuf hmc_views(double* d_state, int64_t n, int D) {
    HmcBuf b;
    double* p = d_state;
    const size_t nD = (size_t)n * D;
    b.q = p; p += nD; b.lp = p; p += n; b.g = p; p += nD;
    b.qp = p; p += nD; b.lpp = p; p += n; b.gp = p; p += nD;
    b.p = p; p += nD; b.h0 = p; p += n; b.acc = p; p += n; b.inv_mass = p; p += D;
    b.beta = p; p += n; b.ll = p; p += n; b.llp = p; p += n; b.ladder = p; p += n; b.swap_acc = p; p += n;
    b.rung_of_chain = reinterpret_cast<int32_t*>(p); b.chain_of_rung = b.rung_of_chain + n;
    b.out_theta = nullptr; b.out_lp = nullptr;
    return b;
}
void octo_hmc_pt_views(double* d_state, int64_t n, int D, double** beta, double** ll, int32_t** rung_of_chain, double** swap_acc) {
    const HmcBuf b = hmc_views(d_state, n, D);
    *beta = b.beta; *ll = b.ll; *rung_of_chain = b.rung_of_chain; *swap_acc = b.swap_acc;
}

// logpost(d_theta [n x D], d_lp, d_g, leap) enqueues one log-posterior + gradient evaluation on `st`.
// h_ladder != nullptr: parallel tempering — chain c starts on rung c with weight h_ladder[c]; the run is n_rounds
// rounds of n_iter tempered HMC transitions followed by one swap round and a re-evaluation at the new weights; d_cold
// [n_rounds x D] receives the state of the chain on the last rung after every round.  Needs the fused log posterior.
cudaError_t octo_hmc_enqueue(double* d_state, int64_t n, int D, int n_iter, int n_leapfrog, double eps, uint64_t seed,
                             double* d_out_theta, double* d_out_lp, cudaStream_t st,
                             int (*logpost)(void*, const double*, double*, double*, const HmcLeap*), void* user,
                             bool fused_leap, int* rc_out, const double* h_ladder, int n_rounds, double* d_cold,
                             int (*resident)(void*, const ResidentArgs*), const PtDistRun* dist) {
    HmcBuf b = hmc_views(d_state, n, D);
    b.out_theta = d_out_theta; b.out_lp = d_out_lp;
    const bool pt = h_ladder != nullptr;
    if (dist && (!pt || !resident)) { *rc_out = -1; return cudaSuccess; }
    if (!pt) { b.beta = nullptr; b.ll = nullptr; b.llp = nullptr; n_rounds = 1; }
    *rc_out = 0;
    cudaError_t e;
    if (pt) {
        std::vector<double> hb(n);
        std::vector<int32_t> id(2 * n);
        // sharded: h_ladder is the whole ladder [R]; this rank's chains start on rungs chain0 .. chain0 + n - 1
        const int64_t c0 = dist ? dist->chain0 : 0;
        for (int64_t c = 0; c < n; ++c) { hb[c] = h_ladder[c0 + c]; id[c] = (int32_t)c; id[n + c] = (int32_t)c; }
        if (dist) {
            std::vector<int32_t> idR(2 * (size_t)dist->R);
            for (int r = 0; r < dist->R; ++r) { idR[r] = r; idR[dist->R + r] = r; }
            e = cudaMemcpyAsync(dist->d.ladder, h_ladder, dist->R * sizeof(double), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dist->d.chain_of_rung, idR.data(), dist->R * sizeof(int32_t), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dist->d.rung_of_chain, idR.data(), dist->R * sizeof(int32_t), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(dist->d.swap_acc, 0, dist->R * sizeof(double), st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
        }
        e = cudaMemcpyAsync(b.beta, hb.data(), n * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b.ladder, hb.data(), n * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b.rung_of_chain, id.data(), 2 * n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(b.swap_acc, 0, n * sizeof(double), st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);                         // the staging vectors are locals
        if (e != cudaSuccess) return e;
    }
    const HmcLeap at_state{nullptr, nullptr, nullptr, 0.0, 0.0, 0, 0, b.beta, b.ll};
    for (int round = 0; round < n_rounds; ++round) {
        const uint64_t round_seed = (uint64_t)(seed + 0x51ED27ULL * (uint64_t)round);
        if (resident) {
            // trajectory-resident kernel: (re-)evaluation at the current weights and all n_iter transitions in ONE launch
            ResidentArgs R;
            R.q = b.q; R.lp = b.lp; R.g = b.g; R.acc = b.acc; R.inv_mass = b.inv_mass; R.out_theta = b.out_theta; R.out_lp = b.out_lp;
            R.beta = pt ? b.beta : nullptr; R.ll = pt ? b.ll : nullptr; R.n = n; R.chain_offset = dist ? dist->chain0 : 0; R.D = D;
            R.n_iter = n_iter; R.n_leapfrog = n_leapfrog; R.it0 = 0; R.eps = eps; R.seed = round_seed;
            R.pair_out = dist ? dist->d.pairs_local : nullptr;
            if ((*rc_out = resident(user, &R))) return cudaSuccess;
        } else {
        // (re-)evaluate the current states at the current weights
        if ((*rc_out = logpost(user, b.q, b.lp, b.g, pt ? &at_state : nullptr))) return cudaSuccess;
        for (int it = 0; it <= n_iter; ++it) {
            // transition counter `round * n_iter + it` keys the random stream; the sample store is per run, not per round
            e = launch_pdl(k_hmc_turn, n, st, b, n, D, it, n_iter, eps, round_seed);
            if (e != cudaSuccess) return e;
            if (it == n_iter) break;
            for (int l = 0; l < n_leapfrog; ++l) {
                const int last = (int)(l == n_leapfrog - 1);
                if (fused_leap) {    // the log-posterior launch applies the kick (and drift) itself: one launch per leapfrog
                    const HmcLeap leap{b.p, b.qp, b.inv_mass, eps, last ? 0.5 * eps : eps, !last, 0, b.beta, b.llp};
                    if ((*rc_out = logpost(user, b.qp, b.lpp, b.gp, &leap))) return cudaSuccess;
                    continue;
                }
                const HmcLeap only_beta{nullptr, nullptr, nullptr, 0.0, 0.0, 0, 0, b.beta, b.llp};
                if ((*rc_out = logpost(user, b.qp, b.lpp, b.gp, pt ? &only_beta : nullptr))) return cudaSuccess;
                e = launch_pdl(k_hmc_leap, n, st, b, n, D, last, eps);
                if (e != cudaSuccess) return e;
            }
        }
        }
        if (dist) {
            // one all-gather of (l_ref, l_target) per round on this stream, then every rank decides every pair
            if ((*rc_out = dist->allgather(dist->user, dist->d.pairs_local, dist->d.pairs_all, (size_t)n * 2))) return cudaSuccess;
            e = launch_pdl(k_pt_swap_dist, (int64_t)(dist->R + 1) / 2, st, dist->d, b.beta, dist->chain0, n, dist->R, (int64_t)round, seed);
            if (e != cudaSuccess) return e;
            if (d_cold) {
                e = launch_pdl(k_pt_record_dist, (int64_t)D, st, dist->d, (const double*)b.q, dist->chain0, n, dist->R, D, (int64_t)round, d_cold);
                if (e != cudaSuccess) return e;
            }
        } else if (pt) {
            e = launch_pdl(k_pt_swap, (n + 1) / 2, st, b, n, D, (int64_t)round, seed, d_cold);
            if (e != cudaSuccess) return e;
            if (d_cold) {
                e = launch_pdl(k_pt_record, (int64_t)D, st, b, n, D, (int64_t)round, d_cold);
                if (e != cudaSuccess) return e;
            }
        }
    }
    if (pt && (*rc_out = logpost(user, b.q, b.lp, b.g, &at_state))) return cudaSuccess;    // lp at the final weights
    return cudaSuccess;
}

// one sharded swap decision step on its own (octo_pt_swap_round_device): d.pairs_all already gathered on `st`
cudaError_t octo_pt_swap_dist_launch(const PtDist& d, double* d_beta_local, int64_t chain0, int64_t n_local, int R, int64_t round,
                                     uint64_t seed, cudaStream_t st) {
    return launch_pdl(k_pt_swap_dist, (int64_t)(R + 1) / 2, st, d, d_beta_local, chain0, n_local, R, round, seed);
}

// host twin of the random stream (tests reproduce a transition with it)
extern "C" void octo_hmc_random(uint64_t seed, int64_t it, int64_t chain, int32_t D, double* z, double* u) {
    const uint64_t key = hmc_key(seed, (uint64_t)it);
    for (int j = 0; j < D; ++j) {
        const uint64_t s = hmc_draw(key, (uint64_t)chain, (uint64_t)j);
        const double u1 = u01(s), u2 = u01(splitmix64(s));
        z[j] = sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2);
    }
    *u = u01(hmc_draw(key, (uint64_t)chain, (uint64_t)D));
}
