// octo_shim.cu — the C ABI of libocto_b200.so (include/octo_b200.h): context creation
// (table upload, per-epoch weight precomputation, accumulator slot map), the workspace/stream pool that
// makes the entry points re-entrant, host-buffer and device-buffer entry points, and the NCCL
// parallel-tempering swap round.  No CPU evaluation path exists here: without a CUDA device
// octo_create fails.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "octo_internal.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return OCTO_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_cuda(e_, #call); } while (0)

// One lease = everything a call needs so that concurrent callers never share mutable state.
struct Workspace {
    cudaStream_t stream = nullptr;
    double *d_in = nullptr, *d_ll = nullptr, *d_g = nullptr, *d_partial = nullptr;
    unsigned int* d_tickets = nullptr;
    double *h_in = nullptr, *h_out = nullptr;          // pinned staging
    double *d_theta = nullptr, *d_post = nullptr;      // parameterised entry points: θ_t and [lp | g_t]
    size_t cap_theta = 0, cap_post = 0;
    size_t cap_in = 0, cap_ll = 0, cap_g = 0, cap_partial = 0, cap_tickets = 0, cap_hin = 0, cap_hout = 0;
    bool busy = false;
    std::mutex mu;        // device-buffer entry points: two host threads may enqueue on the same caller stream
};

// minimal NCCL surface, resolved with dlopen so the library has no link-time NCCL dependency
struct Id128 { char b[128]; };   // ncclUniqueId, passed by value
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

}  // namespace

struct OctoCtx {
    DevModel m;
    int device = 0;
    int n_sm = 148;
    size_t smem = 0;
    double* d_tables = nullptr;
    double* d_pw_const = nullptr;   // per-epoch normalisation constants (pointwise mode)
    std::mutex mu;
    std::vector<Workspace*> pool;                                  // host-buffer calls: leased per call
    std::vector<std::pair<cudaStream_t, Workspace*>> stream_ws;    // device-buffer calls: one per caller stream
    std::atomic<int64_t> launches{0};
    struct GeomEntry { int64_t n; bool fused; int sub; int param_gen; LaunchGeom g; };
    std::mutex geom_mu;
    std::vector<GeomEntry> geom_cache;                             // geometry(): results by batch size
    int param_gen = 0;                                             // bumped by octo_set_parameterization (fused shared-memory sizes change)
    int ctas_per_sm = 2;
    int warps = OCTO_WARPS;        // warps per CTA; halved until the model's accumulator slots fit in shared memory
    int slice_override = 0;
    int latency_mode = 1;          // OCTO_B200_LATENCY: 0 never, 1 automatic, 2 whenever the chain groups fit one per SM
    int resident_mode = 1;         // OCTO_B200_RESIDENT: 0 = explorers never use the trajectory-resident kernel
    int force[3] = {0, 0, 0};      // OCTO_B200_FORCE (experiments): sub-lanes, latency instantiation, epoch splits
    // page-locked host inputs up to this many bytes are read by the kernel in place (UVA, over PCIe) instead of being copied
    // to the device first: one driver call less per evaluation.  C2 (90 KB): 12.9 -> 9.2 us per step with three evaluations in
    // flight, 28.8 -> 25.5 us blocking; 4096 x 100 (295 KB): 16.8 -> 11.7 / 39.5 -> 29.6; 16384 x 100 (1.2 MB): 57 -> 116 (the
    // copy engine is the faster way for large inputs).  OCTO_B200_ZEROCOPY_MAX (bytes; 0 = always copy)
    size_t zerocopy_max = 512 << 10;
    double lat_cap = 300.0;        // OCTO_B200_LAT_CAP: most dependent pairs per lane a single wave of the latency instantiation may get
    int sublane_mode = 0;          // OCTO_B200_SUBLANES: 0 automatic, 1 never (lane = chain), 2..32 that many sub-lanes per chain
    // device-side parameterisation (N1)
    DevParam* d_param = nullptr;
    int param_D = 0, param_T = 0;
    bool param_fused = false;      // the parameterisation runs inside K1 (one launch) instead of K0f + K1 + K0b
    size_t smem_fused = 0;
    int ctas_per_sm_fused = 1;
    int warps_fused = OCTO_WARPS;  // the fused stage's arrays may need a smaller CTA than the plain kernel
    size_t smem_optin = 0;
    // parallel tempering
    void* nccl_comm = nullptr;
    int pt_rank = 0, pt_world = 1, pt_local = 0;
    uint64_t pt_seed = 0;
    double* d_gather = nullptr;
    double* h_gather = nullptr;
    cudaStream_t pt_stream = nullptr;
};

namespace {

NcclApi g_nccl;
std::mutex g_nccl_mu;

// host buffers handed out by octo_alloc_pinned: copies to/from them need no staging
std::mutex g_pin_mu;
std::vector<std::pair<const char*, size_t>> g_pinned;
bool is_pinned(const void* p, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (auto& r : g_pinned) if ((const char*)p >= r.first && (const char*)p + bytes <= r.first + r.second) return true;
    return false;
}

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.h) return OCTO_OK;
    const char* names[] = {getenv("OCTO_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!h) return fail(OCTO_ERR_NCCL, "cannot dlopen libnccl.so.2 (set OCTO_B200_NCCL)");
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(OCTO_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.h = h;
    return OCTO_OK;
}
int fail_nccl(int rc, const char* what) {
    g_err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
    return OCTO_ERR_NCCL;
}

template <class T>
int ensure(T** p, size_t* cap, size_t need, bool pinned = false) {
    if (need <= *cap) return OCTO_OK;
    size_t n = need + need / 2;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    if (pinned) CU(cudaMallocHost((void**)p, n * sizeof(T)));
    else CU(cudaMalloc((void**)p, n * sizeof(T)));
    *cap = n;
    return OCTO_OK;
}

Workspace* lease(OctoCtx* ctx) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (Workspace* w : ctx->pool) if (!w->busy) { w->busy = true; return w; }
    Workspace* w = new Workspace();
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) { delete w; return nullptr; }
    w->busy = true;
    ctx->pool.push_back(w);
    return w;
}
void release(OctoCtx* ctx, Workspace* w) { std::lock_guard<std::mutex> lk(ctx->mu); w->busy = false; }

// device-buffer entry point: the partial buffer/tickets are tied to the caller's stream, on which the
// launches that use them are ordered
Workspace* stream_workspace(OctoCtx* ctx, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& kv : ctx->stream_ws) if (kv.first == st) return kv.second;
    Workspace* w = new Workspace();
    ctx->stream_ws.emplace_back(st, w);
    return w;
}

void free_ws(Workspace* w) {
    if (w->d_in) cudaFree(w->d_in);
    if (w->d_ll) cudaFree(w->d_ll);
    if (w->d_g) cudaFree(w->d_g);
    if (w->d_partial) cudaFree(w->d_partial);
    if (w->d_tickets) cudaFree(w->d_tickets);
    if (w->d_theta) cudaFree(w->d_theta);
    if (w->d_post) cudaFree(w->d_post);
    if (w->h_in) cudaFreeHost(w->h_in);
    if (w->h_out) cudaFreeHost(w->h_out);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
}

// grid: chain groups x epoch splits.  Small problems: as many splits as fill the resident CTA slots exactly once
// (an integral number of waves, never below min_slice epochs per warp).  Large problems (>= 4 waves): ~32 epochs
// per warp so the per-CTA prologue/epilogue is amortised and the hardware CTA scheduler balances the tail.
// This is the lane = chain mapping on its own (ch = 32); geometry() below adds the sub-lane mapping for small batches.
LaunchGeom geometry_classic(const OctoCtx* ctx, int64_t n_chains, bool fused) {
    LaunchGeom g;
    const int W = fused ? ctx->warps_fused : ctx->warps;
    g.block = W * 32;
    g.smem = fused ? ctx->smem_fused : ctx->smem;
    const int64_t E = ctx->m.n_epochs;
    g.gx = (int)((n_chains + 31) / 32);
    const int min_slice = ctx->slice_override > 0 ? ctx->slice_override : OCTO_MIN_SLICE;
    int64_t max_gy = E / ((int64_t)min_slice * W);
    if (max_gy < 1) max_gy = 1;
    if (max_gy > 65535) max_gy = 65535;
    const int64_t resident = (int64_t)ctx->n_sm * (fused ? ctx->ctas_per_sm_fused : ctx->ctas_per_sm);
    int64_t gy_fill = (resident + g.gx - 1) / g.gx;
    if (gy_fill > max_gy) gy_fill = max_gy;
    int64_t gy = gy_fill;
    if ((int64_t)g.gx * gy_fill >= 4 * resident) {
        const int64_t gy_big = E / (32 * W);
        gy = gy_big > gy_fill ? (gy_big < max_gy ? gy_big : max_gy) : gy_fill;
    } else if ((int64_t)g.gx * max_gy > resident) {
        // pick the split count whose CTA total is closest below an integral number of waves
        double best_eff = -1.0;
        const int64_t hi = max_gy < gy_fill * 4 + 8 ? max_gy : gy_fill * 4 + 8;
        for (int64_t c = gy_fill > 2 ? gy_fill - 2 : 1; c <= hi; ++c) {
            const double waves = (double)(g.gx * c) / (double)resident;
            if (waves < 0.9) continue;
            const double eff = waves / std::ceil(waves);
            if (eff > best_eff + 1e-9) { best_eff = eff; gy = c; if (eff >= 0.95) break; }
        }
    }
    if (gy < 1) gy = 1;
    // very long slices (> 1024 epochs per warp): more, shorter CTAs balance the tail of the last wave better (4096 x 1e5:
    // 37 splits of 338 epochs per warp instead of 9 of 1389: +2 %); keep >= 300 epochs per warp and whole waves
    if ((E + gy * W - 1) / (gy * W) > 1024) {
        for (int64_t c = std::min<int64_t>(max_gy, E / (300 * (int64_t)W)); c > gy; --c) {
            const double waves = (double)(g.gx * c) / (double)resident;
            if (waves / std::ceil(waves) >= 0.97) { gy = c; break; }
        }
    }
    // tiny problems: the cross-CTA combine (write + fence + ticket + reads, ~3.5 us) costs more than the epochs it
    // takes off each warp (~0.6 us per lean-astrometry-equivalent epoch, measured on C2's timeline)
    if (gy > 1) {
        const double t_iter = 0.6, t_k2 = 3.5, ew = ctx->m.wtot;
        const double split = std::ceil(ew / (double)(W * gy)) * t_iter + t_k2, single = std::ceil(ew / (double)W) * t_iter;
        if (single <= split) gy = 1;
    }
    // Latency-tuned instantiation (no register cap, one CTA per SM): when the grid above is a single wave anyway and
    // the chain groups fit one per SM, run fewer, fatter CTAs of it instead — on C2 128 CTAs x 7 epochs per warp beat
    // 288 x 3 (14.6 vs 15.4 us per step): no spills, a 4- instead of 9-way combine, the SM to itself
    int Wg = W;
    if (ctx->latency_mode != 0 && g.gx <= ctx->n_sm && ((int64_t)g.gx * gy <= resident || ctx->latency_mode == 2)) {
        // its CTA may be wider than the throughput one when the model's shared memory allows
        int Wl = (W == OCTO_WARPS) ? OCTO_LAT_WARPS : W;
        const int D = fused ? ctx->param_D : 0, T = fused ? ctx->param_T : 0;
        if (octo_smem_bytes(ctx->m, Wl, D, T) > ctx->smem_optin) Wl = W;
        int64_t max_gl = E / ((int64_t)min_slice * Wl);
        if (max_gl < 1) max_gl = 1;
        int64_t gl = ctx->n_sm / g.gx;
        if (gl > max_gl) gl = max_gl;
        if (gl >= 1) {
            if (gy > 1 || ctx->latency_mode == 2) gy = gl;
            g.lat = true; Wg = Wl; g.block = Wl * 32; g.smem = octo_smem_bytes(ctx->m, Wl, D, T);
        }
    }
    g.gy = (int)gy;
    g.slice = (int)((E + gy * Wg - 1) / (gy * Wg));
    return g;
}

// Small batches (the chain groups of the classic mapping would leave SMs idle, or need the cross-CTA combine to fill
// them): candidates (sub-lanes S, instantiation, epoch splits) whose grid is a single wave, ranked by a two-term cost
// model measured on C2's timeline — dependent pair evaluations per (warp, sub-lane) unit x ~0.55 us, + ~3.5 us when
// the splits of a chain group have to be combined through L2.  Larger batches keep the classic geometry.
// sub_override: 0 = automatic, otherwise the sub-lane count to use (1 = classic mapping).
LaunchGeom geometry_search(const OctoCtx* ctx, int64_t n_chains, bool fused, int sub_override) {
    if (sub_override == 0) sub_override = ctx->sublane_mode;
    LaunchGeom best = geometry_classic(ctx, n_chains, fused);
    const int64_t E = ctx->m.n_epochs;
    const int W = fused ? ctx->warps_fused : ctx->warps;
    const int64_t resident = (int64_t)ctx->n_sm * (fused ? ctx->ctas_per_sm_fused : ctx->ctas_per_sm);
    const int D = fused ? ctx->param_D : 0, T = fused ? ctx->param_T : 0;
    int Wl = (W == OCTO_WARPS) ? OCTO_LAT_WARPS : W;
    if (octo_smem_bytes(ctx->m, Wl, D, T) > ctx->smem_optin) Wl = W;
    if (ctx->force[0] > 0 && sub_override != 1 && E > 0) {       // experiments: OCTO_B200_FORCE="sub-lanes,latency,splits"
        const int S = ctx->force[0], Wc = ctx->force[1] ? Wl : W;
        best.ch = 32 / S; best.gx = (int)((n_chains + best.ch - 1) / best.ch); best.gy = ctx->force[2] > 0 ? ctx->force[2] : 1;
        best.lat = ctx->force[1] != 0; best.block = Wc * 32;
        best.smem = best.lat ? octo_smem_bytes(ctx->m, Wc, D, T) : (fused ? ctx->smem_fused : ctx->smem);
        best.slice = (int)((E + (int64_t)best.gy * Wc * S - 1) / ((int64_t)best.gy * Wc * S));
        return best;
    }
    if (sub_override == 1 || E < 1 || (sub_override == 0 && (n_chains + 31) / 32 >= resident)) return best;
    // measured on C1 - C4 / 4096 x {10 ... 3162} / RV tables (profiles/r02_geometry_sweep.txt): a lean-astrometry-equivalent
    // pair costs a lane of the latency-tuned instantiation ~0.42 us (12 warps, two pairs in flight), ~0.63 us in the throughput
    // instantiation (two CTAs share the SM), which also has ~0.5 us more fixed cost per launch; the combine through L2 costs
    // ~3.5 us.  The lane = chain choice above is priced by the same model, so a
    // single wave of the latency instantiation is taken when the model says it is the faster one, up to 128 dependent
    // pairs per lane (see below)
    const double t_it[2] = {0.63, 0.42}, t_k2 = 3.5, t_thr = 0.5;
    const double ew[2] = {ctx->m.wtot > 0 ? ctx->m.wtot : 1.0, ctx->m.wtot_lat > 0 ? ctx->m.wtot_lat : 1.0};
    auto cost = [&](int S, bool lat, int64_t gy) {
        const int Wc = lat ? Wl : W;
        return std::ceil(ew[lat ? 1 : 0] / (double)(gy * Wc * S)) * t_it[lat ? 1 : 0] + (gy > 1 ? t_k2 + 0.02 * (double)gy : 0.0) + (lat ? 0.0 : t_thr);
    };
    // the classic choice, priced by the same model (multi-wave grids pay per wave)
    double best_cost;
    {
        const int64_t slots = best.lat ? ctx->n_sm : resident;
        // (whole waves up to four, then the hardware's CTA scheduler evens the tail out: fractional)
        double waves = (double)best.gx * best.gy / (double)slots;
        if (waves < 4.0) waves = std::ceil(waves);
        best_cost = waves * cost(1, best.lat, best.gy);
        if (sub_override > 1) best_cost = 1e30;
    }
    for (int S = 1; S <= 32; S *= 2) {
        if (sub_override > 1 && S != sub_override) continue;
        const int ch = 32 / S;
        const int64_t gx = (n_chains + ch - 1) / ch;
        for (int lat = 1; lat >= 0; --lat) {
            if (lat && ctx->latency_mode == 0) continue;
            const int64_t slots = lat ? ctx->n_sm : resident;
            if (gx > slots) continue;
            const int Wc = lat ? Wl : W;
            int64_t max_gy = slots / gx;
            // no more units than twice the epochs (a unit may end up with nothing to do, never all but a few)
            const int64_t cap = (2 * E) / ((int64_t)Wc * S);
            if (max_gy > cap) max_gy = cap;
            if (max_gy > 65535) max_gy = 65535;
            if (max_gy < 1) { if (S > 1 && sub_override != S) continue; max_gy = 1; }
            int64_t gy_c = 1; double c_c = cost(S, lat, 1);
            for (int64_t gy = 2; gy <= max_gy; ++gy) {
                const double c = cost(S, lat, gy);
                if (c < c_c - 1e-9) { c_c = c; gy_c = gy; }
            }
            // a single wave is a latency play: for long loops the throughput instantiation (16 warps per SM) catches up —
            // 4096 x 3162 lean astrometry (264 pairs per lane): 115 vs 124 us, x 10000: 352 vs 347; RV + jitter (weight
            // 1.5): x 3162 171 vs 166, x 10000 521 vs 479 — so stop considering it beyond 300 weighted pairs per lane
            if (sub_override <= 1 && std::ceil(ew[lat ? 1 : 0] / (double)(gy_c * Wc * S)) > ctx->lat_cap) continue;
            if (c_c < best_cost - 1e-9) {
                best_cost = c_c;
                best.gx = (int)gx; best.gy = (int)gy_c; best.ch = ch; best.lat = lat != 0; best.block = Wc * 32;
                best.smem = lat ? octo_smem_bytes(ctx->m, Wc, D, T) : (fused ? ctx->smem_fused : ctx->smem);
                best.slice = (int)((E + gy_c * Wc * S - 1) / (gy_c * Wc * S));
            }
        }
    }
    return best;
}

// the search runs once per (batch size, fused, override): samplers call with the same batch size millions of times
LaunchGeom geometry(const OctoCtx* ctx, int64_t n_chains, bool fused = false, int sub_override = 0) {
    OctoCtx* c = const_cast<OctoCtx*>(ctx);
    {
        std::lock_guard<std::mutex> lk(c->geom_mu);
        for (const auto& e : c->geom_cache)
            if (e.n == n_chains && e.fused == fused && e.sub == sub_override && e.param_gen == c->param_gen) return e.g;
    }
    const LaunchGeom g = geometry_search(ctx, n_chains, fused, sub_override);
    std::lock_guard<std::mutex> lk(c->geom_mu);
    if (c->geom_cache.size() >= 16) c->geom_cache.erase(c->geom_cache.begin());
    c->geom_cache.push_back({n_chains, fused, sub_override, c->param_gen, g});
    return g;
}

// d_param != nullptr: fused parameterisation — d_in is θ_t, d_ll / d_g receive the log posterior and its gradient
// post_mode 1 (with d_param): likelihood part only.  pointwise: value-only, one CTA row per epoch, d_ll is [n x E] (ld ldg).
int enqueue(OctoCtx* ctx, Workspace* w, bool grad, const double* d_in, int64_t n, int64_t ld, double* d_ll, double* d_g,
            int64_t ldg, cudaStream_t st, const DevParam* d_param = nullptr, int post_mode = 0, bool pointwise = false,
            const HmcLeap* leap = nullptr, int64_t pw_e0 = 0, int64_t pw_n = 0, const InlineIn* inl = nullptr) {
    LaunchGeom g = geometry(ctx, n, d_param != nullptr, pointwise ? 1 : 0);
    if (pointwise) {      // epochs [pw_e0, pw_e0 + pw_n) of the concatenated list, one CTA row each
        if (pw_n < 1 || pw_n > 65535 || pw_e0 < 0 || pw_e0 + pw_n > ctx->m.n_epochs || pw_e0 >= (1 << 22)) return fail(OCTO_ERR_ARG, "bad pointwise chunk");
        const int W = ctx->warps > 4 ? 4 : ctx->warps;        // one warp does the epoch; the others only help the prologue
        g.block = W * 32; g.gy = (int)pw_n; g.smem = octo_smem_bytes(ctx->m, W); g.lat = false;
        post_mode |= (int)(pw_e0 << 9);
    }
    if (g.gy > 1 && !pointwise) {
        size_t need = (size_t)g.gx * g.gy * ctx->m.n_acc * 32;
        if (int rc = ensure(&w->d_partial, &w->cap_partial, need)) return rc;
        if ((size_t)g.gx > w->cap_tickets) {
            if (int rc = ensure(&w->d_tickets, &w->cap_tickets, (size_t)g.gx)) return rc;
            CU(cudaMemsetAsync(w->d_tickets, 0, w->cap_tickets * sizeof(unsigned int), st));
        }
    }
    cudaError_t e = octo_launch(ctx->m, g, grad, d_in, n, ld, d_ll, d_g, ldg, w->d_partial, w->d_tickets, d_param, post_mode,
                                pointwise ? ctx->d_pw_const + pw_e0 : nullptr, leap ? *leap : HmcLeap{}, st, inl);
    if (e != cudaSuccess) return fail_cuda(e, "kernel launch");
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return OCTO_OK;
}

int logpost_enqueue(OctoCtx* ctx, Workspace* w, const double* d_theta, int64_t n, int64_t ld, double* d_lp,
                    double* d_g_t, int64_t ldg, double* d_work, cudaStream_t st, int post_mode = 0,
                    const HmcLeap* leap = nullptr, const InlineIn* inl = nullptr) {
    if (ctx->param_fused)     // one launch: θ_t -> inputs in K1's prologue, ∂/∂θ_t in its epilogue; no workspace
        return enqueue(ctx, w, d_g_t != nullptr, d_theta, n, ld, d_lp, d_g_t, ldg, st, ctx->d_param, post_mode, false, leap, 0, 0, inl);
    const int n_in = ctx->m.n_in;
    double* d_in = d_work;                      // [n x n_in]
    double* d_ll = d_work + (size_t)n * n_in;   // [n]
    double* d_gin = d_ll + n;                   // [n x n_in]
    double* d_save = d_gin + (size_t)n * n_in;  // [n x (3D + 3)]
    cudaError_t e = octo_param_forward(ctx->d_param, ctx->param_D, ctx->m, d_theta, n, ld, d_in, d_save, st);
    if (e != cudaSuccess) return fail_cuda(e, "k_param_forward");
    if (int rc = enqueue(ctx, w, d_g_t != nullptr, d_in, n, n, d_ll, d_g_t ? d_gin : nullptr, n, st)) return rc;
    e = octo_param_backward(ctx->d_param, ctx->param_D, ctx->m, n, d_in, d_save, d_ll, d_gin, d_lp, d_g_t, ldg, post_mode, st);
    if (e != cudaSuccess) return fail_cuda(e, "k_param_backward");
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    return OCTO_OK;
}

// One evaluation with host buffers, in two halves so that callers can overlap their own work (or further evaluations)
// with it: begin_host leases a workspace (a stream of its own) and enqueues copy-in, kernel(s) and copy-out;
// finish_host waits for that stream, hands staged results to the caller's buffers and returns the workspace to the pool.
// post = false: `in` are the kernel inputs [n x n_in] (octo_logp[_grad]); post = true: θ_t [n x D] -> log posterior
// (octo_logpost_grad).  Gradient columns = input columns in both cases.  post_mode 1 (post only, value-only): the
// likelihood part, ln_like(system, arr2nt(invlink(θ_t))).
struct Pending {
    OctoCtx* ctx = nullptr; Workspace* w = nullptr;
    bool staged_out = false, grad = false;
    double *ll = nullptr, *g = nullptr;
    int64_t n = 0, ld = 0; int nc = 0;
};

int begin_host(OctoCtx* ctx, bool post, bool grad, const double* in, int64_t n, int64_t ld, double* ll, double* g,
               int post_mode, Pending* P) {
    if (!ctx) return fail(OCTO_ERR_ARG, "null context");
    if (post && !ctx->d_param) return fail(OCTO_ERR_STATE, "octo_set_parameterization has not been called");
    P->ctx = ctx; P->w = nullptr; P->n = n;
    if (n == 0) return OCTO_OK;
    if (!in || !ll || (grad && !g) || n < 0 || ld < n) return fail(OCTO_ERR_ARG, "bad buffers / leading dimension");
    CU(cudaSetDevice(ctx->device));
    Workspace* w = lease(ctx);
    if (!w) return fail(OCTO_ERR_CUDA, "cannot create stream");
    const int n_in = ctx->m.n_in, nc = post ? ctx->param_D : n_in;
    const size_t col = (size_t)n * sizeof(double), pitch = (size_t)ld * sizeof(double);
    // buffers from octo_alloc_pinned are copied directly; anything else is staged through pinned memory
    const bool pin_in = is_pinned(in, pitch * (nc - 1) + col);
    const bool pin_out = is_pinned(ll, col) && (!grad || is_pinned(g, pitch * (nc - 1) + col));
    int rc = OCTO_OK;
    do {
        double*& d_x = post ? w->d_theta : w->d_in;
        size_t& cap_x = post ? w->cap_theta : w->cap_in;
        double*& d_y = post ? w->d_post : w->d_ll;
        size_t& cap_y = post ? w->cap_post : w->cap_ll;
        if ((rc = ensure(&d_x, &cap_x, (size_t)n * nc))) break;
        if ((rc = ensure(&d_y, &cap_y, (size_t)n * (nc + 1)))) break;                // [ll | g] contiguous
        if (post && !ctx->param_fused && (rc = ensure(&w->d_in, &w->cap_in, (size_t)n * (2 * n_in + 1 + 3 * nc + 3)))) break;
        // pinned outputs: the last kernel stores ll / g rows straight into the caller's page-locked
        // buffers over PCIe (UVA: host pointer == device pointer) — no D2H copy launches at all
        const bool direct_out = pin_out;
        double* d_ll = direct_out ? ll : d_y;
        double* d_g = direct_out ? g : d_y + n;
        const int64_t ldg = direct_out ? ld : n;
        cudaError_t e = cudaSuccess;
        // tiny batch: the inputs ride in the kernel parameters, no copy launch at all
        InlineIn inl;
        const bool inline_in = (size_t)n * nc <= OCTO_INLINE_MAX && (!post || ctx->param_fused);
        const bool zc = pin_in && !inline_in && (size_t)n * nc * sizeof(double) <= ctx->zerocopy_max;
        const double* k_in = zc ? in : d_x;                 // UVA: the page-locked host pointer is a device pointer
        const int64_t k_ld = zc ? ld : n;
        if (inline_in) {
            for (int k = 0; k < nc; ++k) memcpy(inl.v + (size_t)k * n, in + (size_t)k * ld, col);
        } else if (zc) {
        } else if (pin_in) {
            e = (ld == n) ? cudaMemcpyAsync(d_x, in, col * nc, cudaMemcpyHostToDevice, w->stream)
                          : cudaMemcpy2DAsync(d_x, col, in, pitch, col, nc, cudaMemcpyHostToDevice, w->stream);
        } else {
            if ((rc = ensure(&w->h_in, &w->cap_hin, (size_t)n * nc, true))) break;
            for (int k = 0; k < nc; ++k) memcpy(w->h_in + (size_t)k * n, in + (size_t)k * ld, col);
            e = cudaMemcpyAsync(d_x, w->h_in, col * nc, cudaMemcpyHostToDevice, w->stream);
        }
        if (e != cudaSuccess) { rc = fail_cuda(e, "H2D"); break; }
        const InlineIn* pinl = inline_in ? &inl : nullptr;
        if (post) rc = logpost_enqueue(ctx, w, k_in, n, k_ld, d_ll, grad ? d_g : nullptr, ldg, w->d_in, w->stream, post_mode, nullptr, pinl);
        else rc = enqueue(ctx, w, grad, k_in, n, k_ld, d_ll, grad ? d_g : nullptr, ldg, w->stream, nullptr, 0, false, nullptr, 0, 0, pinl);
        if (rc) break;
        if (!direct_out) {
            if ((rc = ensure(&w->h_out, &w->cap_hout, (size_t)n * (nc + 1), true))) break;
            e = cudaMemcpyAsync(w->h_out, d_ll, col * (grad ? nc + 1 : 1), cudaMemcpyDeviceToHost, w->stream);
            if (e != cudaSuccess) { rc = fail_cuda(e, "D2H"); break; }
        }
        P->w = w; P->staged_out = !direct_out; P->grad = grad; P->ll = ll; P->g = g; P->ld = ld; P->nc = nc;
    } while (0);
    if (rc) { cudaStreamSynchronize(w->stream); release(ctx, w); }
    return rc;
}

int finish_host(Pending* P) {
    if (!P->w) return OCTO_OK;                         // empty batch
    Workspace* w = P->w;
    int rc = OCTO_OK;
    cudaError_t e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) rc = fail_cuda(e, "kernel execution");
    else if (P->staged_out) {
        const size_t col = (size_t)P->n * sizeof(double);
        memcpy(P->ll, w->h_out, col);
        if (P->grad) for (int k = 0; k < P->nc; ++k) memcpy(P->g + (size_t)k * P->ld, w->h_out + P->n + (size_t)k * P->n, col);
    }
    release(P->ctx, w);
    P->w = nullptr;
    return rc;
}

int run_host(OctoCtx* ctx, bool post, bool grad, const double* in, int64_t n, int64_t ld, double* ll, double* g,
             int post_mode = 0) {
    Pending P;
    if (int rc = begin_host(ctx, post, grad, in, n, ld, ll, g, post_mode, &P)) return rc;
    return finish_host(&P);
}

}  // namespace

extern "C" {

const char* octo_last_error(void) { return g_err.c_str(); }
int octo_abi_version(void) { return OCTO_ABI_VERSION; }

void octo_default_constants(OctoConstants* c) {
    if (!c) return;
    c->kepler_year_days = 365.2568983840419; c->year2day = 365.25; c->rad2as = 206265.0; c->pc2au = 206265.0;
    c->au2m = 1.495978707e11; c->sec2year = 1.0 / 31557600.0; c->mjup2msol = 0.0009545942339693249;
}

int octo_create(const OctoConstants* consts, const OctoLayout* L, const OctoObsBlock* blocks, int32_t n_blocks,
                int32_t device, OctoCtx** out) {
    if (!out) return fail(OCTO_ERR_ARG, "out is null");
    *out = nullptr;
    if (!consts || !L || (n_blocks > 0 && !blocks)) return fail(OCTO_ERR_ARG, "null argument");
    if (L->n_planets < 1 || L->n_planets > OCTO_MAX_PLANETS) return fail(OCTO_ERR_ARG, "n_planets must be 1..4");
    if (L->n_in < 1 || L->n_in > 4096) return fail(OCTO_ERR_ARG, "n_in out of range");
    if (n_blocks < 0 || n_blocks > OCTO_MAX_BLOCKS) return fail(OCTO_ERR_ARG, "too many observation tables");
    auto col_ok = [&](int k, bool optional) { return (optional && k == -1) || (k >= 0 && k < L->n_in); };
    bool any_ti = false;
    for (int p = 0; p < L->n_planets; ++p) {
        if (L->basis[p] != OCTO_BASIS_CAMPBELL && L->basis[p] != OCTO_BASIS_THIELE_INNES) return fail(OCTO_ERR_ARG, "unknown orbit basis");
        const bool ti = L->basis[p] == OCTO_BASIS_THIELE_INNES;
        any_ti = any_ti || ti;
        if (!col_ok(L->idx_plx[p], false) || !col_ok(L->idx_e[p], false) || !col_ok(L->idx_tp[p], false) ||
            !col_ok(L->idx_M[p], false) || !col_ok(L->idx_mass[p], true))
            return fail(OCTO_ERR_ARG, "layout column index out of range");
        if (ti ? (!col_ok(L->idx_A[p], false) || !col_ok(L->idx_B[p], false) || !col_ok(L->idx_F[p], false) || !col_ok(L->idx_G[p], false))
               : (!col_ok(L->idx_a[p], false) || !col_ok(L->idx_i[p], false) || !col_ok(L->idx_w[p], false) || !col_ok(L->idx_W[p], false)))
            return fail(OCTO_ERR_ARG, "layout column index out of range");
    }
    if (any_ti)
        for (int b = 0; b < n_blocks; ++b)
            if (blocks[b].kind == OCTO_KIND_RV_STAR_ABS || blocks[b].kind == OCTO_KIND_RV_STAR_MARGIN || blocks[b].kind == OCTO_KIND_RV_PLANET_REL)
                return fail(OCTO_ERR_ARG, "radial-velocity tables cannot be combined with a Thiele-Innes planet (not offloaded)");
    // HGCAInstantaneousObs tables (kind 5) are not part of the epoch list: set them aside
    std::vector<OctoObsBlock> regular;
    std::vector<OctoObsBlock> hgs;
    for (int b = 0; b < n_blocks; ++b) (blocks[b].kind == OCTO_KIND_HGCA_INSTANT ? hgs : regular).push_back(blocks[b]);
    if ((int)hgs.size() > OCTO_MAX_HGCA) return fail(OCTO_ERR_ARG, "too many HGCA tables");
    blocks = regular.data(); n_blocks = (int32_t)regular.size();
    int dev_count = 0;
    cudaError_t ce = cudaGetDeviceCount(&dev_count);
    if (ce != cudaSuccess || dev_count == 0)
        return fail(OCTO_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (this library has no CPU path)");
    if (device < 0 || device >= dev_count) return fail(OCTO_ERR_ARG, "device ordinal out of range");
    CU(cudaSetDevice(device));

    OctoCtx* ctx = new OctoCtx();
    DevModel& m = ctx->m;
    memset(&m, 0, sizeof(m));
    m.c = *consts;
    m.kappa = 6.283185307179586 * consts->year2day / consts->kepler_year_days * consts->au2m * consts->sec2year;
    m.c2a_per_plx = consts->rad2as * 1e3 / (1000.0 * consts->pc2au);
    m.n_planets = L->n_planets; m.n_in = L->n_in; m.n_blocks = n_blocks;
    for (int p = 0; p < OCTO_MAX_PLANETS; ++p) {
        m.idx_plx[p] = L->idx_plx[p]; m.idx_a[p] = L->idx_a[p]; m.idx_e[p] = L->idx_e[p]; m.idx_i[p] = L->idx_i[p];
        m.idx_w[p] = L->idx_w[p]; m.idx_W[p] = L->idx_W[p]; m.idx_tp[p] = L->idx_tp[p]; m.idx_M[p] = L->idx_M[p];
        m.idx_mass[p] = p < L->n_planets ? L->idx_mass[p] : -1;
        m.basis[p] = p < L->n_planets ? L->basis[p] : 0;
        m.idx_A[p] = L->idx_A[p]; m.idx_B[p] = L->idx_B[p]; m.idx_F[p] = L->idx_F[p]; m.idx_G[p] = L->idx_G[p];
        if (p < L->n_planets && m.basis[p] == OCTO_BASIS_THIELE_INNES) {
            m.any_ti = 1;
            m.idx_a[p] = L->n_in + p;            // virtual gradient column: a is an intermediate of this basis
            m.idx_i[p] = m.idx_w[p] = m.idx_W[p] = -1;
        }
    }
    // validate tables, assign accumulator slots
    int64_t E = 0;
    int n_acc = 1 + PA_COUNT * L->n_planets;
    for (int b = 0; b < n_blocks; ++b) {
        const OctoObsBlock& B = blocks[b];
        DevBlock& D = m.blocks[b];
        const bool astrom = (B.kind == OCTO_KIND_ASTROM_RADEC || B.kind == OCTO_KIND_ASTROM_PASEP);
        const bool sys = (B.kind == OCTO_KIND_RV_STAR_ABS || B.kind == OCTO_KIND_RV_STAR_MARGIN);
        auto bad = [&](const char* msg) { delete ctx; return fail(OCTO_ERR_ARG, std::string("table ") + std::to_string(b) + ": " + msg); };
        if (B.kind < 0 || B.kind > OCTO_KIND_RV_PLANET_REL) return bad("unknown kind");
        if (B.n_epochs < 0 || (B.n_epochs > 0 && (!B.epoch || !B.y1 || !B.s1))) return bad("missing columns");
        if (astrom && B.n_epochs > 0 && (!B.y2 || !B.s2)) return bad("astrometry needs y2 and s2");
        if (astrom && B.has_cor && B.n_epochs > 0 && !B.cor) return bad("has_cor set but cor is null");
        if (!sys && (B.planet < 0 || B.planet >= L->n_planets)) return bad("planet index out of range");
        if (sys) for (int p = 0; p < L->n_planets; ++p)
            if (L->idx_mass[p] < 0) return bad("star RV needs a mass variable on every planet (rv-absolute.jl:147)");
        if (!col_ok(B.idx_jitter, true) || !col_ok(B.idx_platescale, true) || !col_ok(B.idx_northangle, true) ||
            !col_ok(B.idx_offset, true)) return bad("observation variable column out of range");
        if (B.kind == OCTO_KIND_RV_STAR_MARGIN && B.idx_jitter < 0) return bad("marginalised RV requires jitter (rv-absolute-margin.jl:149)");
        D.kind = B.kind; D.planet = sys ? -1 : B.planet; D.start = (int32_t)E; D.n = B.n_epochs;
        D.has_cor = (astrom && B.has_cor) ? 1 : 0;
        D.jit = B.idx_jitter >= 0 ? 1 : 0;
        D.idx_jitter = B.idx_jitter; D.idx_offset = (B.kind == OCTO_KIND_RV_STAR_MARGIN || astrom) ? -1 : B.idx_offset;
        D.idx_platescale = astrom ? B.idx_platescale : -1; D.idx_northangle = astrom ? B.idx_northangle : -1;
        D.slot_jitter = D.slot_platescale = D.slot_northangle = D.slot_offset = D.slot_margin = D.slot_obsprior = -1;
        D.n_trend = 0;
        for (int v = 0; v < 3; ++v) { D.idx_trend[v] = -1; D.slot_trend[v] = -1; }
        if (B.n_trend != 0 || B.trend_const) {
            if (astrom) return bad("trend_function belongs to the radial-velocity kinds");
            if (B.n_trend < 0 || B.n_trend > 3) return bad("at most 3 trend variables");
            if (B.n_trend > 0 && B.n_epochs > 0 && !B.trend_basis) return bad("n_trend set but trend_basis is null");
            for (int v = 0; v < B.n_trend; ++v) {
                if (!col_ok(B.idx_trend[v], false)) return bad("trend variable column out of range");
                for (int k = 0; k < B.n_epochs; ++k) if (!std::isfinite(B.trend_basis[(size_t)v * B.n_epochs + k])) return bad("non-finite trend basis value");
                D.idx_trend[v] = B.idx_trend[v];
            }
            if (B.trend_const) for (int k = 0; k < B.n_epochs; ++k) if (!std::isfinite(B.trend_const[k])) return bad("non-finite trend constant");
            D.n_trend = B.n_trend;
        }
        if (B.obs_prior) {
            if (!astrom) return bad("the observable-based prior is offloaded for relative astrometry only (prior-observable.jl:78-137)");
            D.slot_obsprior = n_acc; n_acc += OP_COUNT;
        }
        if (B.kind == OCTO_KIND_RV_STAR_MARGIN) {
            D.slot_margin = n_acc; n_acc += MA_COUNT + MV_COUNT * L->n_planets;
            for (int v = 0; v < D.n_trend; ++v) { D.slot_trend[v] = n_acc; n_acc += 2; }
        } else {
            for (int v = 0; v < D.n_trend; ++v) D.slot_trend[v] = n_acc++;
            if (D.idx_jitter >= 0) D.slot_jitter = n_acc++;
            if (D.idx_offset >= 0) D.slot_offset = n_acc++;
        }
        if (D.idx_platescale >= 0) D.slot_platescale = n_acc++;
        if (D.idx_northangle >= 0) D.slot_northangle = n_acc++;
        E += B.n_epochs;
        if (E > 0x7fffffff) return bad("too many epochs");
    }
    // HGCA tables: validation, averaging constants, precision matrices; two accumulator slots each (pmra, pmdec)
    long double cll_hg = 0.0L;
    size_t hg_rows_total = 0;
    m.n_hg = (int32_t)hgs.size();
    for (size_t h = 0; h < hgs.size(); ++h) {
        const OctoObsBlock& B = hgs[h];
        DevHg& H = m.hg[h];
        auto bad = [&](const char* msg) { delete ctx; return fail(OCTO_ERR_ARG, std::string("HGCA table: ") + msg); };
        if (B.n_epochs < 4 || !B.epoch || !B.y1 || !B.aux) return bad("needs rows (epoch, code) and the 15 catalogue numbers");
        if (!col_ok(B.idx_pmra, false) || !col_ok(B.idx_pmdec, false)) return bad("pmra / pmdec column out of range");
        for (int p = 0; p < L->n_planets; ++p)
            if (L->idx_mass[p] < 0) return bad("needs a mass variable on every planet (hgca.jl:271)");
        int cnt[4] = {0, 0, 0, 0};
        double ep[4] = {0, 0, 0, 0};
        for (int k = 0; k < B.n_epochs; ++k) {
            const int code = (int)B.y1[k];
            if (code < 0 || code > 3 || (double)code != B.y1[k]) return bad("row code must be 0..3");
            cnt[code]++; ep[code] += B.epoch[k];
        }
        for (int q = 0; q < 4; ++q) {
            if (cnt[q] == 0) return bad("needs at least one row of every kind (Hipparcos/Gaia x RA/Dec)");
            ep[q] /= cnt[q];
            H.inv_N[q] = 1.0 / ((double)cnt[q] * L->n_planets);
        }
        if (ep[2] == ep[0] || ep[3] == ep[1]) return bad("Hipparcos and Gaia epochs coincide");
        H.k_ra = 365.25 / (ep[2] - ep[0]); H.k_dec = 365.25 / (ep[3] - ep[1]);
        for (int d = 0; d < 3; ++d) {
            const double* q = B.aux + 5 * d;
            const double s1 = q[2], s2 = q[3], cor = q[4];
            if (!(s1 > 0) || !(s2 > 0) || !(std::fabs(cor) < 1.0)) return bad("catalogue errors must be > 0 and |correlation| < 1");
            const double om = 1.0 - cor * cor;
            H.cat[d][0] = q[0]; H.cat[d][1] = q[1];
            H.w[d][0] = 1.0 / (s1 * s1 * om); H.w[d][1] = -cor / (s1 * s2 * om); H.w[d][2] = 1.0 / (s2 * s2 * om);
            cll_hg += -1.8378770664093454835606594728112353L - 0.5L * std::log((long double)s1 * s1 * s2 * s2 * om);
        }
        H.n_rows = B.n_epochs; H.idx_pmra = B.idx_pmra; H.idx_pmdec = B.idx_pmdec;
        H.slot_pmra = n_acc++; H.slot_pmdec = n_acc++;
        hg_rows_total += (size_t)B.n_epochs;
    }
    m.n_epochs = E; m.n_acc = n_acc;
    for (int b = 0; b < n_blocks; ++b) if (m.blocks[b].kind == OCTO_KIND_RV_STAR_MARGIN || m.blocks[b].slot_obsprior >= 0) m.has_margin = 1;
    m.lean = (m.n_hg == 0 && !m.any_ti && !getenv("OCTO_B200_NO_LEAN")) ? 1 : 0;
    for (int b = 0; b < n_blocks; ++b) {
        const DevBlock& D = m.blocks[b];
        const bool astrom = D.kind <= OCTO_KIND_ASTROM_PASEP;
        if (astrom ? !(D.kind == OCTO_KIND_ASTROM_RADEC && !D.jit && D.idx_platescale < 0 && D.idx_northangle < 0 && D.slot_obsprior < 0)
                   : (D.kind == OCTO_KIND_RV_STAR_MARGIN || D.n_trend > 0)) m.lean = 0;
    }
    // cost model for the epoch split (instructions per epoch of each specialised loop, relative to lean astrometry)
    double cum = 0.0, cum_lat = 0.0;
    for (int b = 0; b < n_blocks; ++b) {
        DevBlock& D = m.blocks[b];
        const bool astrom = D.kind <= OCTO_KIND_ASTROM_PASEP;
        const bool lean = D.kind == OCTO_KIND_ASTROM_RADEC && !D.jit && D.idx_platescale < 0 && D.idx_northangle < 0 && D.slot_obsprior < 0;
        const bool plain = D.kind == OCTO_KIND_ASTROM_RADEC && D.idx_platescale < 0 && D.idx_northangle < 0 && D.slot_obsprior < 0;
        double w = astrom ? (lean ? 1.0 : (plain ? 1.6 : 2.5)) : (D.kind == OCTO_KIND_RV_STAR_MARGIN ? 1.9 : (D.jit ? 1.8 : 1.1));
        int solves = 1;
        if (D.kind == OCTO_KIND_RV_STAR_ABS || D.kind == OCTO_KIND_RV_STAR_MARGIN) solves = L->n_planets;
        else for (int p = 0; p < L->n_planets; ++p) if (p != D.planet && L->idx_mass[p] >= 0) ++solves;   // upper bound
        D.wgt = w + 0.8 * (solves - 1);
        D.cum = cum;
        cum += D.wgt * D.n;
        // latency-bound launches (one CTA per SM, a few pairs per lane): what counts is the dependent chain of one pair,
        // measured on the resident kernel's timeline: lean astrometry 1340 cycles, RV + jitter 1870
        const double wl = astrom ? (lean ? 1.0 : (plain ? 1.3 : 1.8)) : (D.kind == OCTO_KIND_RV_STAR_MARGIN ? 1.45 : (D.jit ? 1.4 : 1.05));
        D.wgt_lat = wl + 0.8 * (solves - 1);
        D.cum_lat = cum_lat;
        cum_lat += D.wgt_lat * D.n;
    }
    m.wtot = cum; m.wtot_lat = cum_lat;

    // host tables: t, y1, y2, c1, c2, c3 (see DevModel); chain-independent normalisation summed in long double
    // + padding: the kernels prefetch one lane-stride (<= 32*8 records) past the record they read
    const size_t hg_off0 = (size_t)6 * (E > 0 ? E : 1) + 6 * 32;        // HGCA rows live behind the (padded) epoch records
    std::vector<double> T(hg_off0 + 2 * hg_rows_total + 2, 0.0);
    {
        size_t off = hg_off0;
        for (size_t h = 0; h < hgs.size(); ++h) {
            m.hg[h].row_off = (int32_t)off;
            for (int k = 0; k < hgs[h].n_epochs; ++k) { T[off++] = hgs[h].epoch[k]; T[off++] = hgs[h].y1[k]; }
        }
    }
    struct Col { double* b; double& operator[](size_t o) const { return b[6 * o]; } };   // AoS record field view
    const Col t{T.data()}, y1{T.data() + 1}, c1{T.data() + 2}, y2{T.data() + 3}, c2{T.data() + 4}, c3{T.data() + 5};
    long double cll = 0.0L;
    std::vector<double> pwc((size_t)(E > 0 ? E : 1), 0.0);        // the same normalisation terms, per epoch (pointwise mode)
    const long double log2pi = 1.8378770664093454835606594728112353L;
    for (int b = 0; b < n_blocks; ++b) {
        const OctoObsBlock& B = blocks[b];
        const DevBlock& D = m.blocks[b];
        const bool astrom = (B.kind <= OCTO_KIND_ASTROM_PASEP);
        for (int k = 0; k < B.n_epochs; ++k) {
            const size_t o = (size_t)D.start + k;
            // a zero / negative / non-finite uncertainty or a non-finite datum would turn EVERY chain into NaN without an
            // error (inf weights, NaN normalisation): refuse the table instead
            const bool fin = std::isfinite(B.epoch[k]) && std::isfinite(B.y1[k]) && std::isfinite(B.s1[k]) && B.s1[k] > 0.0 &&
                             (!astrom || (std::isfinite(B.y2[k]) && std::isfinite(B.s2[k]) && B.s2[k] > 0.0));
            if (!fin) {
                delete ctx;
                return fail(OCTO_ERR_ARG, "table " + std::to_string(b) + ", row " + std::to_string(k) + ": epochs and data must be finite, uncertainties finite and > 0");
            }
            t[o] = B.epoch[k]; y1[o] = B.y1[k];
            if (astrom) {
                y2[o] = B.y2[k];
                const double s1 = B.s1[k], s2 = B.s2[k], cor = D.has_cor ? B.cor[k] : 0.0;
                // the reference ctor rejects |cor| > 1 - 1e-5 ("may not be well-specified", relative-astrometry.jl:69-71)
                if (!(std::fabs(cor) <= 1.0 - 1e-5)) { delete ctx; return fail(OCTO_ERR_ARG, "|cor| > 1 - 1e-5 (relative-astrometry.jl:69-71)"); }
                if (B.kind == OCTO_KIND_ASTROM_RADEC && D.idx_platescale < 0 && D.idx_northangle < 0) {
                    // the reference pushes the data through atan/hypot/cos/sin even with platescale = 1,
                    // northangle = 0 (relative-astrometry.jl:209-213): reproduce that <= 2 ulp perturbation here
                    const double pa = std::atan2(B.y2[k], B.y1[k]) - 0.0, sep = std::hypot(B.y2[k], B.y1[k]) * 1.0;
                    y1[o] = sep * std::cos(pa); y2[o] = sep * std::sin(pa);
                }
                if (!D.jit) {
                    const double om = 1.0 - cor * cor;
                    c1[o] = 1.0 / (s1 * s1 * om); c3[o] = 1.0 / (s2 * s2 * om); c2[o] = -cor / (s1 * s2 * om);
                    const long double term = -log2pi - 0.5L * std::log((long double)s1 * s1 * s2 * s2 * om);
                    cll += term; pwc[o] = (double)term;
                } else { c1[o] = s1 * s1; c2[o] = s2 * s2; c3[o] = cor; }
            } else {
                // trend: the constant part leaves the data, the basis values of the (<= 3) coefficients ride in the record
                if (B.trend_const) y1[o] = B.y1[k] - B.trend_const[k];
                if (D.n_trend > 0) y2[o] = B.trend_basis[k];
                if (D.n_trend > 1) c2[o] = B.trend_basis[(size_t)B.n_epochs + k];
                if (D.n_trend > 2) c3[o] = B.trend_basis[2 * (size_t)B.n_epochs + k];
                const double s = B.s1[k];
                if (!D.jit) {
                    const long double term = -0.5L * (log2pi + std::log((long double)s * s));
                    c1[o] = 1.0 / (s * s); cll += term; pwc[o] = (double)term;
                }
                else c1[o] = s * s;
            }
        }
    }
    m.const_ll = (double)(cll + cll_hg);

    cudaDeviceProp prop;
    ce = cudaGetDeviceProperties(&prop, device);
    if (ce != cudaSuccess) { delete ctx; return fail_cuda(ce, "cudaGetDeviceProperties"); }
    if (prop.major < 10) { delete ctx; return fail(OCTO_ERR_CUDA, "libocto_b200 is built for sm_100a only"); }
    ctx->device = device; ctx->n_sm = prop.multiProcessorCount;
    ctx->warps = OCTO_WARPS;
    ctx->smem_optin = (size_t)prop.sharedMemPerBlockOptin - 1024;      // room for the kernels' static shared memory
    while (ctx->warps > 1 && octo_smem_bytes(m, ctx->warps) > ctx->smem_optin) ctx->warps /= 2;
    ctx->smem = octo_smem_bytes(m, ctx->warps);
    if (ctx->smem > (size_t)prop.sharedMemPerBlockOptin) {
        delete ctx; return fail(OCTO_ERR_ARG, "model too large: accumulator slots exceed shared memory");
    }
    if (const char* s = getenv("OCTO_B200_SLICE")) ctx->slice_override = std::max(1, atoi(s));
    if (const char* s = getenv("OCTO_B200_LATENCY")) ctx->latency_mode = atoi(s);
    if (const char* s = getenv("OCTO_B200_RESIDENT")) ctx->resident_mode = atoi(s);
    if (const char* s = getenv("OCTO_B200_FORCE")) {
        int a = 0, b = 0, c = 0;
        if (sscanf(s, "%d,%d,%d", &a, &b, &c) == 3 && (a == 1 || a == 2 || a == 4 || a == 8 || a == 16 || a == 32) && c >= 1 && c <= 65535) {
            ctx->force[0] = a; ctx->force[1] = b; ctx->force[2] = c;
        }
    }
    if (const char* s = getenv("OCTO_B200_ZEROCOPY_MAX")) ctx->zerocopy_max = (size_t)atoll(s);
    if (const char* s = getenv("OCTO_B200_LAT_CAP")) { const double v = atof(s); if (v >= 1.0) ctx->lat_cap = v; }
    if (const char* s = getenv("OCTO_B200_SUBLANES")) {
        const int v = atoi(s);
        if (v == 0 || v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ctx->sublane_mode = v;
    }
    ce = cudaMalloc((void**)&ctx->d_tables, T.size() * sizeof(double));
    if (ce != cudaSuccess) { delete ctx; return fail_cuda(ce, "cudaMalloc tables"); }
    ce = cudaMemcpy(ctx->d_tables, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(ctx->d_tables); delete ctx; return fail_cuda(ce, "upload tables"); }
    m.tab = ctx->d_tables;
    ce = cudaMalloc((void**)&ctx->d_pw_const, pwc.size() * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemcpy(ctx->d_pw_const, pwc.data(), pwc.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(ctx->d_tables); if (ctx->d_pw_const) cudaFree(ctx->d_pw_const); delete ctx; return fail_cuda(ce, "upload per-epoch constants"); }
    int occ = 0;
    ce = octo_kernels_init(m, ctx->smem, ctx->smem_optin, ctx->warps, &occ);
    if (ce != cudaSuccess) { cudaFree(ctx->d_tables); delete ctx; return fail_cuda(ce, "cudaFuncSetAttribute"); }
    ctx->ctas_per_sm = occ > 0 ? occ : 1;
    if (const char* s = getenv("OCTO_B200_CTAS_PER_SM")) ctx->ctas_per_sm = std::max(1, atoi(s));
    *out = ctx;
    return OCTO_OK;
}

void octo_destroy(OctoCtx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    octo_pt_finalize(ctx);
    cudaDeviceSynchronize();
    for (Workspace* w : ctx->pool) free_ws(w);
    for (auto& kv : ctx->stream_ws) free_ws(kv.second);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->d_pw_const) cudaFree(ctx->d_pw_const);
    if (ctx->d_param) cudaFree(ctx->d_param);
    delete ctx;
}

int octo_logp(OctoCtx* ctx, const double* in, int64_t n, int64_t ld, double* ll) { return run_host(ctx, false, false, in, n, ld, ll, nullptr); }
int octo_logp_grad(OctoCtx* ctx, const double* in, int64_t n, int64_t ld, double* ll, double* g) { return run_host(ctx, false, true, in, n, ld, ll, g); }

// ---- asynchronous halves of the host-buffer entry points
struct OctoTicket { Pending p; };

static int begin_ticket(OctoCtx* ctx, bool post, const double* in, int64_t n, int64_t ld, double* out, double* g, OctoTicket** ticket) {
    if (!ticket) return fail(OCTO_ERR_ARG, "ticket is null");
    *ticket = nullptr;
    OctoTicket* t = new OctoTicket();
    if (int rc = begin_host(ctx, post, g != nullptr, in, n, ld, out, g, 0, &t->p)) { delete t; return rc; }
    *ticket = t;
    return OCTO_OK;
}
int octo_logp_grad_begin(OctoCtx* ctx, const double* in, int64_t n, int64_t ld, double* ll, double* g_in, OctoTicket** ticket) {
    return begin_ticket(ctx, false, in, n, ld, ll, g_in, ticket);
}
int octo_logpost_grad_begin(OctoCtx* ctx, const double* theta_t, int64_t n, int64_t ld, double* lp, double* g_t, OctoTicket** ticket) {
    return begin_ticket(ctx, true, theta_t, n, ld, lp, g_t, ticket);
}
int octo_ready(OctoTicket* t) {
    if (!t) return -OCTO_ERR_ARG;
    if (!t->p.w) return 1;
    cudaError_t e = cudaStreamQuery(t->p.w->stream);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    fail_cuda(e, "cudaStreamQuery");
    return -OCTO_ERR_CUDA;
}
int octo_wait(OctoTicket* t) {
    if (!t) return fail(OCTO_ERR_ARG, "null ticket");
    if (t->p.ctx) cudaSetDevice(t->p.ctx->device);
    const int rc = finish_host(&t->p);
    delete t;
    return rc;
}

int octo_logp_grad_device(OctoCtx* ctx, const double* d_in, int64_t n, int64_t ld, double* d_ll, double* d_g, void* stream) {
    if (!ctx) return fail(OCTO_ERR_ARG, "null context");
    if (n == 0) return OCTO_OK;
    if (!d_in || !d_ll || n < 0 || ld < n) return fail(OCTO_ERR_ARG, "bad buffers / leading dimension");
    CU(cudaSetDevice(ctx->device));
    Workspace* w = stream_workspace(ctx, (cudaStream_t)stream);   // partial buffer + tickets of this stream
    std::lock_guard<std::mutex> lk(w->mu);                        // growing them and launching is one critical section
    return enqueue(ctx, w, d_g != nullptr, d_in, n, ld, d_ll, d_g, ld, (cudaStream_t)stream);
}

// drop the workspace the device-buffer entry points keep for a caller stream (call it before destroying the stream: a
// recycled stream handle would otherwise inherit it).  Waits for the stream's pending work first.
int octo_release_stream(OctoCtx* ctx, void* stream) {
    if (!ctx) return fail(OCTO_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    Workspace* w = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        for (size_t i = 0; i < ctx->stream_ws.size(); ++i)
            if (ctx->stream_ws[i].first == (cudaStream_t)stream) { w = ctx->stream_ws[i].second; ctx->stream_ws.erase(ctx->stream_ws.begin() + i); break; }
    }
    if (!w) return OCTO_OK;
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    { std::lock_guard<std::mutex> lk(w->mu); }
    free_ws(w);
    return OCTO_OK;
}

// ------------------------------------------------------------------------------------------------
// Standard parameterisation on the device
// ------------------------------------------------------------------------------------------------
int octo_set_parameterization(OctoCtx* ctx, const OctoPrior* priors, int32_t D, const OctoInputDef* defs) {
    if (!ctx || !priors || !defs) return fail(OCTO_ERR_ARG, "null argument");
    const int n_in = ctx->m.n_in;
    if (D < 1 || D > OCTO_PARAM_MAX || n_in > OCTO_PARAM_MAX)
        return fail(OCTO_ERR_ARG, "parameterised models support D, n_in <= " + std::to_string(OCTO_PARAM_MAX));
    DevParam P;
    memset(&P, 0, sizeof(P));
    P.D = D; P.n_in = n_in;
    for (int j = 0; j < D; ++j) {
        const OctoPrior& pr = priors[j];
        P.priors[j] = pr;
        double* pc = P.pc[j];                      // lo, hi, constant part of the log density, 1/(hi-lo), 1/σ
        const double kHalfLog2Pi = 0.91893853320467274178, kPi = 3.14159265358979323846;
        pc[0] = -INFINITY; pc[1] = INFINITY; pc[2] = 0.0; pc[3] = 0.0; pc[4] = 0.0;
        switch (pr.family) {
            case OCTO_PRIOR_NORMAL:
                if (!(pr.p[1] > 0)) return fail(OCTO_ERR_ARG, "Normal: sigma must be > 0");
                pc[2] = -std::log(pr.p[1]) - kHalfLog2Pi; pc[4] = 1.0 / pr.p[1];
                break;
            case OCTO_PRIOR_UNIFORM:
                if (!(pr.p[0] < pr.p[1])) return fail(OCTO_ERR_ARG, "Uniform: a < b required");
                pc[0] = pr.p[0]; pc[1] = pr.p[1]; pc[2] = -std::log(pr.p[1] - pr.p[0]);
                break;
            case OCTO_PRIOR_LOGUNIFORM:
                if (!(0 < pr.p[0] && pr.p[0] < pr.p[1])) return fail(OCTO_ERR_ARG, "LogUniform: 0 < a < b required");
                pc[0] = pr.p[0]; pc[1] = pr.p[1]; pc[2] = -std::log(std::log(pr.p[1] / pr.p[0]));
                break;
            case OCTO_PRIOR_SINE: pc[0] = 2.220446049250313e-16; pc[1] = kPi - 2.220446049250313e-16; break;
            case OCTO_PRIOR_TRUNCNORMAL: {
                if (!(pr.p[1] > 0) || !(pr.p[2] < pr.p[3])) return fail(OCTO_ERR_ARG, "truncated Normal: sigma > 0, lower < upper required");
                // log(Φ(β) - Φ(α)), on the side of the distribution that avoids cancellation
                const double is2 = 0.7071067811865476, mu = pr.p[0], sg = pr.p[1];
                const double a = std::isfinite(pr.p[2]) ? (pr.p[2] - mu) / sg : -INFINITY;
                const double b = std::isfinite(pr.p[3]) ? (pr.p[3] - mu) / sg : INFINITY;
                const double tp = a > 0 ? 0.5 * (std::erfc(a * is2) - std::erfc(b * is2))
                                        : 0.5 * (std::erfc(-b * is2) - std::erfc(-a * is2));
                pc[0] = pr.p[2]; pc[1] = pr.p[3];
                pc[2] = -std::log(sg) - kHalfLog2Pi - std::log(tp); pc[4] = 1.0 / sg;
                break;
            }
            default: return fail(OCTO_ERR_ARG, "unknown prior family");
        }
        if (std::isfinite(pc[0]) && std::isfinite(pc[1])) pc[3] = 1.0 / (pc[1] - pc[0]);
    }
    for (int k = 0; k < n_in; ++k) {
        const OctoInputDef& d = defs[k];
        P.defs[k] = d;
        auto th_ok = [&](int j) { return j >= 0 && j < D; };
        switch (d.op) {
            case OCTO_IN_PARAM: if (!th_ok(d.a[0])) return fail(OCTO_ERR_ARG, "input definition: parameter index out of range"); break;
            case OCTO_IN_CONST: break;
            case OCTO_IN_CIRC: if (!th_ok(d.a[0]) || !th_ok(d.a[1])) return fail(OCTO_ERR_ARG, "UniformCircular: parameter index out of range"); break;
            case OCTO_IN_TPERI: case OCTO_IN_TPERI_TI:
                for (int q = 0; q < (d.op == OCTO_IN_TPERI_TI ? 8 : 7); ++q)
                    if (d.a[q] < 0 || d.a[q] >= k) return fail(OCTO_ERR_ARG, "θ_at_epoch_to_tperi: arguments must be earlier kernel inputs");
                break;
            default: return fail(OCTO_ERR_ARG, "unknown input definition");
        }
    }
    // reverse map parameter -> inputs that read it, last input first (the order of the reference's reverse pass)
    {
        int n = 0;
        for (int j = 0; j < D; ++j) {
            P.gat_start[j] = (int16_t)n;
            for (int k = n_in - 1; k >= 0; --k) {
                const OctoInputDef& d = P.defs[k];
                int role = -1;
                if (d.op == OCTO_IN_PARAM && d.a[0] == j) role = 0;
                else if (d.op == OCTO_IN_CIRC && (d.a[0] == j || d.a[1] == j)) role = (d.a[0] == j ? 1 : 0) | (d.a[1] == j ? 2 : 0);
                if (role >= 0) P.gat[n++] = (int16_t)(k | (role << 8));
            }
        }
        P.gat_start[D] = (int16_t)n;
    }
    // lean kernels (Campbell planets, lean tables): which inputs need their sine / cosine, and what their fused stage
    // assumes about the other orbital elements (octo_kernels.cu, param_forward)
    memset(P.in_trig, 0, sizeof(P.in_trig));
    bool lean_ok = true;
    {
        const DevModel& m = ctx->m;
        auto plain = [&](int k) { return k >= 0 && (P.defs[k].op == OCTO_IN_PARAM || P.defs[k].op == OCTO_IN_CONST); };
        auto angle = [&](int k) { return plain(k) || (k >= 0 && P.defs[k].op == OCTO_IN_CIRC); };
        for (int p = 0; p < m.n_planets && m.lean; ++p) {
            for (int k : {m.idx_i[p], m.idx_w[p], m.idx_W[p]}) { if (angle(k)) P.in_trig[k] = 1; else lean_ok = false; }
            for (int k : {m.idx_e[p], m.idx_a[p], m.idx_M[p], m.idx_plx[p]}) if (!plain(k)) lean_ok = false;
            if (m.idx_mass[p] >= 0 && !plain(m.idx_mass[p])) lean_ok = false;
            const int ktp = m.idx_tp[p];
            if (!(plain(ktp) || (ktp >= 0 && P.defs[ktp].op == OCTO_IN_TPERI))) lean_ok = false;
        }
        for (int k = 0; k < n_in && m.lean; ++k) {
            if (P.defs[k].op == OCTO_IN_TPERI_TI) lean_ok = false;
            if (P.defs[k].op != OCTO_IN_TPERI) continue;
            for (int q : {0, 4, 5, 6}) { const int a = P.defs[k].a[q]; if (angle(a)) P.in_trig[a] = 1; else lean_ok = false; }
        }
    }
    // evaluation orders, most expensive first (stable): see DevParam
    {
        auto order_by = [](uint8_t* out, int n, const std::vector<int>& cost) {
            std::vector<int> idx(n);
            for (int i = 0; i < n; ++i) idx[i] = i;
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return cost[a] > cost[b]; });
            for (int i = 0; i < n; ++i) out[i] = (uint8_t)idx[i];
        };
        std::vector<int> cp(D), ci(n_in), cg(D, 0);
        for (int j = 0; j < D; ++j) {
            const bool lb = std::isfinite(P.pc[j][0]), ub = std::isfinite(P.pc[j][1]);
            cp[j] = (lb || ub ? 3 : 0) + (P.priors[j].family == OCTO_PRIOR_LOGUNIFORM ? 1 : 0) + (P.priors[j].family == OCTO_PRIOR_SINE ? 2 : 0);
        }
        for (int k = 0; k < n_in; ++k) ci[k] = P.defs[k].op == OCTO_IN_CIRC ? 3 : (P.in_trig[k] ? 2 : 0);
        for (int j = 0; j < D; ++j)
            for (int it = P.gat_start[j]; it < P.gat_start[j + 1]; ++it) cg[j] += (P.gat[it] >> 8) ? 3 : 1;
        order_by(P.order_prior, D, cp); order_by(P.order_input, n_in, ci); order_by(P.order_gather, D, cg);
    }
    // fused stage inside K1: needs every θ_at_epoch_to_tperi to depend on non-tperi inputs only (they are evaluated
    // together), at most OCTO_PARAM_TPERI_MAX of them, and the extra shared memory; otherwise K0f + K1 + K0b
    bool fusable = true;
    P.n_tperi = 0;
    for (int k = 0; k < n_in; ++k) {
        if (P.defs[k].op != OCTO_IN_TPERI && P.defs[k].op != OCTO_IN_TPERI_TI) continue;
        for (int q = 0; q < (P.defs[k].op == OCTO_IN_TPERI_TI ? 8 : 7); ++q) {
            const int op = P.defs[P.defs[k].a[q]].op;
            if (op == OCTO_IN_TPERI || op == OCTO_IN_TPERI_TI) fusable = false;
        }
        if (P.n_tperi < OCTO_PARAM_TPERI_MAX) P.tperi_k[P.n_tperi] = k;
        if (++P.n_tperi > OCTO_PARAM_TPERI_MAX) { fusable = false; P.n_tperi = OCTO_PARAM_TPERI_MAX; }
    }
    if (ctx->m.lean && !lean_ok) fusable = false;          // (an element derived in an unusual way: the stand-alone stage handles it)
    if (const char* e = getenv("OCTO_B200_FUSE_PARAM")) if (atoi(e) == 0) fusable = false;
    CU(cudaSetDevice(ctx->device));
    int wf = ctx->warps;
    while (wf > 1 && octo_smem_bytes(ctx->m, wf, D, P.n_tperi) > ctx->smem_optin) wf /= 2;
    const size_t smem_f = octo_smem_bytes(ctx->m, wf, D, P.n_tperi);
    if (smem_f > ctx->smem_optin) fusable = false;
    if (fusable) {
        int occ = 0;
        ctx->warps_fused = wf;
        CU(octo_kernels_init(ctx->m, smem_f, ctx->smem_optin, wf, &occ));
        if (occ < 1) fusable = false;
        ctx->smem_fused = smem_f; ctx->ctas_per_sm_fused = occ > 0 ? occ : 1;
        if (const char* e = getenv("OCTO_B200_CTAS_PER_SM")) ctx->ctas_per_sm_fused = std::max(1, atoi(e));
    }
    if (fusable) CU(octo_resident_init(ctx->m, ctx->smem_optin));
    CU(octo_param_init(D, n_in));
    if (!ctx->d_param) CU(cudaMalloc((void**)&ctx->d_param, sizeof(DevParam)));
    CU(cudaMemcpy(ctx->d_param, &P, sizeof(DevParam), cudaMemcpyHostToDevice));
    ctx->param_D = D; ctx->param_T = P.n_tperi;
    ctx->param_fused = fusable;
    { std::lock_guard<std::mutex> lk(ctx->geom_mu); ++ctx->param_gen; ctx->geom_cache.clear(); }
    return OCTO_OK;
}

int64_t octo_logpost_workspace(const OctoCtx* ctx, int64_t n) {
    if (!ctx || n < 0) return -1;
    if (ctx->param_fused) return 0;            // one fused launch: nothing to hand from kernel to kernel
    return (int64_t)sizeof(double) * n * (2 * (int64_t)ctx->m.n_in + 1 + 3 * (int64_t)ctx->param_D + 3);
}


int octo_logpost_grad_device(OctoCtx* ctx, const double* d_theta, int64_t n, int64_t ld, double* d_lp, double* d_g_t,
                             void* d_work, void* stream) {
    if (!ctx) return fail(OCTO_ERR_ARG, "null context");
    if (!ctx->d_param) return fail(OCTO_ERR_STATE, "octo_set_parameterization has not been called");
    if (n == 0) return OCTO_OK;
    if (!d_theta || !d_lp || (!d_work && !ctx->param_fused) || n < 0 || ld < n)
        return fail(OCTO_ERR_ARG, "bad buffers / leading dimension");
    CU(cudaSetDevice(ctx->device));
    Workspace* w = stream_workspace(ctx, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lk(w->mu);
    return logpost_enqueue(ctx, w, d_theta, n, ld, d_lp, d_g_t, ld, (double*)d_work, (cudaStream_t)stream);
}

int octo_logpost_grad(OctoCtx* ctx, const double* theta_t, int64_t n, int64_t ld, double* lp, double* g_t) {
    return run_host(ctx, true, g_t != nullptr, theta_t, n, ld, lp, g_t);
}

int octo_loglike_theta(OctoCtx* ctx, const double* theta_t, int64_t n, int64_t ld, double* ll) {
    return run_host(ctx, true, false, theta_t, n, ld, ll, nullptr, 1);
}

// per-epoch log-likelihoods: out[chain + e * ldo] = ln_like of the model reduced to epoch e alone (e in the order of
// the concatenated tables).  What `pointwise_like` (src/cross-validation.jl:6-49) evaluates per posterior sample.
int octo_logp_pointwise(OctoCtx* ctx, const double* in, int64_t n, int64_t ld, double* out, int64_t ldo) {
    if (!ctx) return fail(OCTO_ERR_ARG, "null context");
    const int64_t E = ctx->m.n_epochs;
    if (n == 0 || E == 0) return OCTO_OK;
    if (!in || !out || n < 0 || ld < n || ldo < n) return fail(OCTO_ERR_ARG, "bad buffers / leading dimension");
    if ((double)n * (double)E > 2.0e9) return fail(OCTO_ERR_ARG, "pointwise output too large: split the batch of chains");
    CU(cudaSetDevice(ctx->device));
    Workspace* w = lease(ctx);
    if (!w) return fail(OCTO_ERR_CUDA, "cannot create stream");
    const int n_in = ctx->m.n_in;
    const size_t col = (size_t)n * sizeof(double);
    int rc = OCTO_OK;
    do {
        if ((rc = ensure(&w->d_in, &w->cap_in, (size_t)n * n_in))) break;
        if ((rc = ensure(&w->d_post, &w->cap_post, (size_t)n * E))) break;
        cudaError_t e = cudaMemcpy2DAsync(w->d_in, col, in, (size_t)ld * sizeof(double), col, n_in, cudaMemcpyHostToDevice, w->stream);
        if (e != cudaSuccess) { rc = fail_cuda(e, "H2D"); break; }
        for (int64_t e0 = 0; e0 < E && !rc; e0 += 65535)        // grid.y is limited to 65535 rows: chunks of epochs
            rc = enqueue(ctx, w, false, w->d_in, n, n, w->d_post + e0 * n, nullptr, n, w->stream, nullptr, 0, true, nullptr, e0,
                         std::min<int64_t>(65535, E - e0));
        if (rc) break;
        e = cudaMemcpy2DAsync(out, (size_t)ldo * sizeof(double), w->d_post, col, col, (size_t)E, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
        if (e != cudaSuccess) { rc = fail_cuda(e, "pointwise evaluation"); break; }
    } while (0);
    release(ctx, w);
    return rc;
}

// ---- device-resident HMC explorer (octo_hmc.cu): the whole run is enqueued on one stream, one sync at the end
namespace {
struct HmcUser { OctoCtx* ctx; Workspace* w; int64_t n; int ch; };
// one all-gather of `count` doubles per rank on the run's stream (sharded parallel tempering); a single rank copies
int hmc_allgather(void* user, const double* d_send, double* d_recv, size_t count) {
    HmcUser* u = (HmcUser*)user;
    if (u->ctx->pt_world > 1) {
        int r = g_nccl.AllGather(d_send, d_recv, count, /*ncclFloat64*/ 8, u->ctx->nccl_comm, u->w->stream);
        return r ? fail_nccl(r, "ncclAllGather") : OCTO_OK;
    }
    cudaError_t e = cudaMemcpyAsync(d_recv, d_send, count * sizeof(double), cudaMemcpyDeviceToDevice, u->w->stream);
    return e == cudaSuccess ? OCTO_OK : fail_cuda(e, "pair copy");
}
int hmc_logpost(void* user, const double* d_theta, double* d_lp, double* d_g, const HmcLeap* leap) {
    HmcUser* u = (HmcUser*)user;
    return logpost_enqueue(u->ctx, u->w, d_theta, u->n, u->n, d_lp, d_g, u->n, u->w->d_in, u->w->stream, 0, leap);
}
int hmc_resident(void* user, const ResidentArgs* R) {
    HmcUser* u = (HmcUser*)user;
    cudaError_t e = octo_resident_launch(u->ctx->m, u->ctx->d_param, u->ctx->param_T, *R, u->ch, u->w->stream);
    if (e != cudaSuccess) return fail_cuda(e, "k_hmc_resident launch");
    u->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return OCTO_OK;
}
// chains per CTA of the trajectory-resident explorer: as many sub-lanes per chain as keep one CTA per SM busy (its chain
// groups never split epochs across CTAs), never more (warp, sub-lane) units than twice the epochs; 0 = not available
int resident_ch(const OctoCtx* ctx, int64_t n) {
    if (!ctx->param_fused || ctx->resident_mode == 0) return 0;
    if (octo_resident_smem_bytes(ctx->m, ctx->param_D, ctx->param_T) > ctx->smem_optin) return 0;
    if (ctx->force[0] > 0) return 32 / ctx->force[0];
    if (ctx->sublane_mode > 0) return 32 / ctx->sublane_mode;
    int S = 1;
    while (S < 32 && (n + (32 / (2 * S)) - 1) / (32 / (2 * S)) <= ctx->n_sm && (int64_t)OCTO_LAT_WARPS * 2 * S <= 2 * std::max<int64_t>(ctx->m.n_epochs, 1)) S *= 2;
    return 32 / S;
}
}  // namespace

namespace {
// shared by octo_hmc_run (ladder == nullptr) and octo_pt_hmc_run
int hmc_run_impl(OctoCtx* ctx, const double* theta0, int64_t n, int64_t ld, int32_t n_iter, int32_t n_leapfrog,
                 double step_size, const double* inv_mass, uint64_t seed, double* theta_samples, double* lp_samples,
                 double* theta_final, double* lp_final, double* accept_rate, const double* ladder, int32_t n_rounds,
                 double* beta_final, int32_t* rung_final, double* swap_accept, double* cold_samples, double* ll_final,
                 bool sharded = false) {
    if (!ctx) return fail(OCTO_ERR_ARG, "null context");
    // sharded: this rank's n chains are chains [rank n, (rank + 1) n) of R = world n; `ladder` has R entries
    const int64_t R = sharded ? n * ctx->pt_world : n, chain0 = sharded ? n * ctx->pt_rank : 0;
    if (sharded && (!ctx->pt_stream || ctx->pt_local != n)) return fail(OCTO_ERR_STATE, "octo_pt_init has not been called with this number of local replicas");
    if (!ctx->d_param) return fail(OCTO_ERR_STATE, "octo_set_parameterization has not been called");
    if (!theta0 || n < 1 || ld < n || n_iter < 1 || n_leapfrog < 1 || !(step_size > 0)) return fail(OCTO_ERR_ARG, "bad arguments");
    const bool pt = ladder != nullptr;
    if (pt && !ctx->param_fused) return fail(OCTO_ERR_STATE, "tempering needs the fused log-posterior launch (not available for this model)");
    if (pt && (n_rounds < 1 || R < 2)) return fail(OCTO_ERR_ARG, "parallel tempering needs >= 2 chains and >= 1 round");
    if (pt) for (int64_t c = 0; c < R; ++c) if (!(ladder[c] >= 0.0 && ladder[c] <= 1.0)) return fail(OCTO_ERR_ARG, "ladder weights must lie in [0, 1]");
    const int D = ctx->param_D, n_in = ctx->m.n_in;
    if (inv_mass) for (int j = 0; j < D; ++j) if (!(inv_mass[j] > 0) || !std::isfinite(inv_mass[j])) return fail(OCTO_ERR_ARG, "inverse mass must be positive");
    CU(cudaSetDevice(ctx->device));
    Workspace* w = lease(ctx);
    if (!w) return fail(OCTO_ERR_CUDA, "cannot create stream");
    const size_t col = (size_t)n * sizeof(double), nD = (size_t)n * D;
    double *d_state = nullptr, *d_ot = nullptr, *d_ol = nullptr, *d_cold = nullptr, *d_dist = nullptr;
    int rc = OCTO_OK;
    do {
        if (!ctx->param_fused && (rc = ensure(&w->d_in, &w->cap_in, (size_t)n * (2 * n_in + 1 + 3 * D + 3)))) break;
        cudaError_t e = cudaMalloc((void**)&d_state, octo_hmc_state_doubles(n, D) * sizeof(double));
        if (e == cudaSuccess && theta_samples) e = cudaMalloc((void**)&d_ot, (size_t)n_iter * nD * sizeof(double));
        if (e == cudaSuccess && lp_samples) e = cudaMalloc((void**)&d_ol, (size_t)n_iter * col);
        if (e == cudaSuccess && pt && cold_samples) e = cudaMalloc((void**)&d_cold, (size_t)n_rounds * D * sizeof(double));
        if (e != cudaSuccess) { rc = fail_cuda(e, "cudaMalloc (HMC state)"); break; }
        double* d_acc = d_state + 5 * nD + 3 * (size_t)n;
        double* d_im = d_acc + n;
        std::vector<double> im(D, 1.0);
        if (inv_mass) im.assign(inv_mass, inv_mass + D);
        e = cudaMemcpy2DAsync(d_state, col, theta0, (size_t)ld * sizeof(double), col, D, cudaMemcpyHostToDevice, w->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_im, im.data(), D * sizeof(double), cudaMemcpyHostToDevice, w->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_acc, 0, col, w->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);          // `im` is a local
        if (e != cudaSuccess) { rc = fail_cuda(e, "HMC setup"); break; }
        const bool fused_leap = ctx->param_fused && !getenv("OCTO_B200_HMC_SEPARATE_LEAP");
        // the trajectory-resident kernel when the model allows it (fused parameterisation, shared memory); a sharded
        // ladder sizes its chain groups by the total replica count, so that every chain sees the arithmetic it would
        // see in a single-GPU run of all R replicas
        const int rch = (fused_leap && (sharded || !getenv("OCTO_B200_HMC_LAUNCH_PER_LEAPFROG"))) ? resident_ch(ctx, R) : 0;
        HmcUser user{ctx, w, n, rch};
        PtDistRun dist{};
        if (sharded) {
            if (rch <= 0) { rc = fail(OCTO_ERR_STATE, "the sharded ladder needs the trajectory-resident explorer (not available for this model)"); break; }
            // [pairs_all 2R | pairs_local 2n | ladder R | swap_acc R | chain_of_rung R (int32) | rung_of_chain R (int32)]
            e = cudaMalloc((void**)&d_dist, (size_t)(2 * R + 2 * n + 2 * R + R + 2) * sizeof(double));
            if (e != cudaSuccess) { rc = fail_cuda(e, "cudaMalloc (sharded tempering)"); break; }
            dist.d.pairs_all = d_dist; dist.d.pairs_local = d_dist + 2 * R; dist.d.ladder = dist.d.pairs_local + 2 * n;
            dist.d.swap_acc = dist.d.ladder + R;
            dist.d.chain_of_rung = reinterpret_cast<int32_t*>(dist.d.swap_acc + R); dist.d.rung_of_chain = dist.d.chain_of_rung + R;
            dist.R = (int)R; dist.chain0 = chain0; dist.allgather = hmc_allgather; dist.user = &user;
        }
        int cb_rc = 0;
        e = octo_hmc_enqueue(d_state, n, D, n_iter, n_leapfrog, step_size, seed, d_ot, d_ol, w->stream, hmc_logpost, &user,
                             fused_leap, &cb_rc, ladder, pt ? n_rounds : 0, d_cold, rch > 0 ? hmc_resident : nullptr,
                             sharded ? &dist : nullptr);
        if (cb_rc) { rc = cb_rc; cudaStreamSynchronize(w->stream); break; }
        if (e != cudaSuccess) { rc = fail_cuda(e, "HMC launch"); cudaStreamSynchronize(w->stream); break; }
        const int64_t rounds = pt ? n_rounds : 1;
        // helper kernels (the log-posterior launches and the resident launches count themselves)
        ctx->launches.fetch_add(rounds * ((rch > 0 ? 0 : (int64_t)n_iter * ((fused_leap ? 0 : n_leapfrog) + 1) + 1) + (pt ? 1 + (d_cold ? 1 : 0) : 0)),
                                std::memory_order_relaxed);
        if (theta_final) e = cudaMemcpy2DAsync(theta_final, (size_t)ld * sizeof(double), d_state, col, col, D, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess && lp_final) e = cudaMemcpyAsync(lp_final, d_state + nD, col, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess && accept_rate) e = cudaMemcpyAsync(accept_rate, d_acc, col, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess && theta_samples) e = cudaMemcpyAsync(theta_samples, d_ot, (size_t)n_iter * nD * sizeof(double), cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess && lp_samples) e = cudaMemcpyAsync(lp_samples, d_ol, (size_t)n_iter * col, cudaMemcpyDeviceToHost, w->stream);
        if (pt) {
            double *d_beta, *d_ll, *d_swap; int32_t* d_rung;
            octo_hmc_pt_views(d_state, n, D, &d_beta, &d_ll, &d_rung, &d_swap);
            if (sharded) {
                d_swap = dist.d.swap_acc; d_rung = dist.d.rung_of_chain + chain0;
                if (cold_samples && ctx->pt_world > 1) {      // every round's row was written by the rank that held the chain
                    int r = g_nccl.AllReduce(d_cold, d_cold, (size_t)n_rounds * D, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->nccl_comm, w->stream);
                    if (r) { rc = fail_nccl(r, "ncclAllReduce"); cudaStreamSynchronize(w->stream); break; }
                }
            }
            if (e == cudaSuccess && beta_final) e = cudaMemcpyAsync(beta_final, d_beta, col, cudaMemcpyDeviceToHost, w->stream);
            if (e == cudaSuccess && ll_final) e = cudaMemcpyAsync(ll_final, d_ll, col, cudaMemcpyDeviceToHost, w->stream);
            if (e == cudaSuccess && rung_final) e = cudaMemcpyAsync(rung_final, d_rung, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, w->stream);
            if (e == cudaSuccess && swap_accept) e = cudaMemcpyAsync(swap_accept, d_swap, (size_t)(R - 1) * sizeof(double), cudaMemcpyDeviceToHost, w->stream);
            if (e == cudaSuccess && cold_samples) e = cudaMemcpyAsync(cold_samples, d_cold, (size_t)n_rounds * D * sizeof(double), cudaMemcpyDeviceToHost, w->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
        if (e != cudaSuccess) { rc = fail_cuda(e, "HMC run"); break; }
        if (accept_rate) for (int64_t c = 0; c < n; ++c) accept_rate[c] /= (double)(n_iter * rounds);
    } while (0);
    if (d_state) cudaFree(d_state);
    if (d_ot) cudaFree(d_ot);
    if (d_ol) cudaFree(d_ol);
    if (d_cold) cudaFree(d_cold);
    if (d_dist) cudaFree(d_dist);
    release(ctx, w);
    return rc;
}
}  // namespace

int octo_hmc_run(OctoCtx* ctx, const double* theta0, int64_t n, int64_t ld, int32_t n_iter, int32_t n_leapfrog,
                 double step_size, const double* inv_mass, uint64_t seed, double* theta_samples, double* lp_samples,
                 double* theta_final, double* lp_final, double* accept_rate) {
    return hmc_run_impl(ctx, theta0, n, ld, n_iter, n_leapfrog, step_size, inv_mass, seed, theta_samples, lp_samples, theta_final,
                        lp_final, accept_rate, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int octo_pt_hmc_run(OctoCtx* ctx, const double* theta0, int64_t n, int64_t ld, const double* ladder, int32_t n_rounds,
                    int32_t n_iter, int32_t n_leapfrog, double step_size, const double* inv_mass, uint64_t seed,
                    double* theta_final, double* lp_final, double* ll_final, double* beta_final, int32_t* rung_final,
                    double* swap_accept, double* cold_samples, double* accept_rate) {
    if (!ladder) return fail(OCTO_ERR_ARG, "ladder is null");
    return hmc_run_impl(ctx, theta0, n, ld, n_iter, n_leapfrog, step_size, inv_mass, seed, nullptr, nullptr, theta_final, lp_final,
                        accept_rate, ladder, n_rounds, beta_final, rung_final, swap_accept, cold_samples, ll_final);
}

int octo_pt_hmc_run_dist(OctoCtx* ctx, const double* theta0_local, int64_t n_local, int64_t ld, const double* ladder_all,
                         int32_t n_rounds, int32_t n_iter, int32_t n_leapfrog, double step_size, const double* inv_mass,
                         uint64_t seed, double* theta_final, double* lp_final, double* ll_final, double* beta_final,
                         int32_t* rung_final, double* swap_accept, double* cold_samples, double* accept_rate) {
    if (!ladder_all) return fail(OCTO_ERR_ARG, "ladder is null");
    return hmc_run_impl(ctx, theta0_local, n_local, ld, n_iter, n_leapfrog, step_size, inv_mass, seed, nullptr, nullptr, theta_final,
                        lp_final, accept_rate, ladder_all, n_rounds, beta_final, rung_final, swap_accept, cold_samples, ll_final, true);
}

int octo_invlink(OctoCtx* ctx, const double* theta_t, int64_t n, int64_t ld, double* theta_nat) {
    if (!ctx) return fail(OCTO_ERR_ARG, "null context");
    if (!ctx->d_param) return fail(OCTO_ERR_STATE, "octo_set_parameterization has not been called");
    if (n == 0) return OCTO_OK;
    if (!theta_t || !theta_nat || n < 0 || ld < n) return fail(OCTO_ERR_ARG, "bad buffers / leading dimension");
    CU(cudaSetDevice(ctx->device));
    Workspace* w = lease(ctx);
    if (!w) return fail(OCTO_ERR_CUDA, "cannot create stream");
    const int D = ctx->param_D;
    int rc = OCTO_OK;
    do {
        const size_t col = (size_t)n * sizeof(double), pitch = (size_t)ld * sizeof(double);
        if ((rc = ensure(&w->d_theta, &w->cap_theta, (size_t)n * D))) break;
        if ((rc = ensure(&w->d_post, &w->cap_post, (size_t)n * (D + 1)))) break;
        cudaError_t e = cudaMemcpy2DAsync(w->d_theta, col, theta_t, pitch, col, D, cudaMemcpyHostToDevice, w->stream);
        if (e == cudaSuccess) e = octo_param_invlink(ctx->d_param, w->d_theta, n, n, w->d_post, w->stream);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(theta_nat, pitch, w->d_post, col, col, D, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
        if (e != cudaSuccess) rc = fail_cuda(e, "invlink");
    } while (0);
    release(ctx, w);
    return rc;
}

// page-locked host memory for `in` / `ll` / `g_in`: octo_logp[_grad] then copies without staging
void* octo_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 8;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) { fail_cuda(e, "cudaMallocHost"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pinned.emplace_back((const char*)p, bytes);
    return p;
}
void octo_free_pinned(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        for (size_t i = 0; i < g_pinned.size(); ++i)
            if (g_pinned[i].first == (const char*)p) { g_pinned.erase(g_pinned.begin() + i); break; }
    }
    cudaFreeHost(p);
}

// diagnostic: the device Kepler solve on its own (mean anomaly MA, eccentricity e) -> sin E, cos E
int octo_selftest_kepler(int32_t device, const double* MA, const double* e, int64_t n, double* sinE, double* cosE) {
    if (!MA || !e || !sinE || !cosE || n < 0) return fail(OCTO_ERR_ARG, "null argument");
    if (n == 0) return OCTO_OK;
    CU(cudaSetDevice(device));
    double* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)4 * n * sizeof(double)));
    int rc = OCTO_OK;
    cudaError_t ce = cudaMemcpy(d, MA, n * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(d + n, e, n * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = octo_selftest_kepler_launch(d, d + n, n, d + 2 * n, d + 3 * n);
    if (ce == cudaSuccess) ce = cudaMemcpy(sinE, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(cosE, d + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) rc = fail_cuda(ce, "selftest");
    cudaFree(d);
    return rc;
}

int32_t octo_n_in(const OctoCtx* ctx) { return ctx ? ctx->m.n_in : -1; }
int32_t octo_n_planets(const OctoCtx* ctx) { return ctx ? ctx->m.n_planets : -1; }
int64_t octo_total_epochs(const OctoCtx* ctx) { return ctx ? ctx->m.n_epochs : -1; }
int32_t octo_device(const OctoCtx* ctx) { return ctx ? ctx->device : -1; }
int64_t octo_kernel_launches(const OctoCtx* ctx) { return ctx ? ctx->launches.load() : -1; }
int octo_launch_geometry(const OctoCtx* ctx, int64_t n_chains, int32_t out[6]) {
    if (!ctx || !out || n_chains < 1) return fail(OCTO_ERR_ARG, "bad argument");
    LaunchGeom g = geometry(ctx, n_chains);
    out[0] = g.gx; out[1] = g.gy; out[2] = g.block; out[3] = g.slice; out[4] = 32 / g.ch; out[5] = g.lat ? 1 : 0;
    return OCTO_OK;
}

// ------------------------------------------------------------------------------------------------
// Parallel tempering swap round
// ------------------------------------------------------------------------------------------------
int octo_pt_unique_id(void* out128) {
    if (!out128) return fail(OCTO_ERR_ARG, "null id buffer");
    if (int rc = load_nccl()) return rc;
    int r = g_nccl.GetUniqueId(out128);
    return r ? fail_nccl(r, "ncclGetUniqueId") : OCTO_OK;
}

int octo_pt_init(OctoCtx* ctx, const void* id, int32_t rank, int32_t world, int32_t n_local, uint64_t seed) {
    if (!ctx || world < 1 || rank < 0 || rank >= world || n_local < 1) return fail(OCTO_ERR_ARG, "bad pt arguments");
    CU(cudaSetDevice(ctx->device));
    octo_pt_finalize(ctx);
    ctx->pt_rank = rank; ctx->pt_world = world; ctx->pt_local = n_local; ctx->pt_seed = seed;
    CU(cudaStreamCreateWithFlags(&ctx->pt_stream, cudaStreamNonBlocking));
    CU(cudaMalloc((void**)&ctx->d_gather, (size_t)(world + 1) * n_local * 2 * sizeof(double)));
    CU(cudaMallocHost((void**)&ctx->h_gather, (size_t)(world + 1) * n_local * 2 * sizeof(double)));
    if (world > 1) {
        if (!id) return fail(OCTO_ERR_ARG, "nccl unique id required for world > 1");
        if (int rc = load_nccl()) return rc;
        Id128 uid; memcpy(uid.b, id, 128);
        int r = g_nccl.CommInitRank(&ctx->nccl_comm, world, uid, rank);
        if (r) return fail_nccl(r, "ncclCommInitRank");
    }
    return OCTO_OK;
}

// counter-based uniform in (0,1): splitmix64 of (seed, round, pair) — identical on every rank
static double pt_uniform(uint64_t seed, uint64_t round, uint64_t pair) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (round * 0x100000001B3ULL + pair + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// Pure host part of a swap round (no CUDA, no NCCL): given every replica's (l_ref, l_target), decide the
// deterministic even-odd swaps.  Identical inputs give identical decisions on every rank.
int octo_pt_decide(const double* ll_pair_all, const double* beta, int32_t* chain_of_replica, int32_t R,
                   int64_t round, uint64_t seed, int32_t* accepted) {
    if (!ll_pair_all || !beta || !chain_of_replica || R < 1) return fail(OCTO_ERR_ARG, "null argument");
    std::vector<int> rep_of(R, -1);                 // inverse permutation: which replica sits on ladder rung i
    for (int r = 0; r < R; ++r) {
        const int ch = chain_of_replica[r];
        if (ch < 0 || ch >= R || rep_of[ch] != -1) return fail(OCTO_ERR_ARG, "chain_of_replica is not a permutation");
        rep_of[ch] = r;
    }
    if (accepted) for (int i = 0; i < R - 1; ++i) accepted[i] = 0;
    // even rounds pair rungs (0,1),(2,3)...; odd rounds (1,2),(3,4)...
    for (int i = (int)(round & 1); i + 1 < R; i += 2) {
        const int ra = rep_of[i], rb = rep_of[i + 1];
        const double* A = ll_pair_all + 2 * (size_t)ra;   // (l_ref, l_target) of the replica on rung i
        const double* B = ll_pair_all + 2 * (size_t)rb;
        const double bi = beta[i], bj = beta[i + 1];
        auto V = [](const double* x, double b) { return (1.0 - b) * x[0] + b * x[1]; };   // tempered log-density
        const double log_ratio = (V(A, bj) + V(B, bi)) - (V(A, bi) + V(B, bj));
        const double u = pt_uniform(seed, (uint64_t)round, (uint64_t)i);
        const bool acc = std::isfinite(log_ratio) ? (std::log(u) < log_ratio) : (log_ratio > 0);
        if (acc) {
            chain_of_replica[ra] = i + 1; chain_of_replica[rb] = i;
            rep_of[i] = rb; rep_of[i + 1] = ra;
            if (accepted) accepted[i] = 1;
        }
    }
    return OCTO_OK;
}

int octo_pt_swap_round(OctoCtx* ctx, const double* ll_pair, const double* beta, int32_t* chain_of_replica,
                       int64_t round, int32_t* accepted) {
    if (!ctx || !ctx->pt_stream) return fail(OCTO_ERR_STATE, "octo_pt_init has not been called");
    if (!ll_pair || !beta || !chain_of_replica) return fail(OCTO_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    const int nl = ctx->pt_local, R = nl * ctx->pt_world;
    if (ctx->pt_world > 1) {
        // local pairs -> device slot of this rank; one all-gather over NVLink; back to pinned host memory
        double* mine = ctx->d_gather + (size_t)R * 2;       // send buffer lives after the receive buffer
        memcpy(ctx->h_gather + (size_t)R * 2, ll_pair, (size_t)nl * 2 * sizeof(double));
        CU(cudaMemcpyAsync(mine, ctx->h_gather + (size_t)R * 2, (size_t)nl * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->pt_stream));
        int r = g_nccl.AllGather(mine, ctx->d_gather, (size_t)nl * 2, /*ncclFloat64*/ 8, ctx->nccl_comm, ctx->pt_stream);
        if (r) return fail_nccl(r, "ncclAllGather");
        CU(cudaMemcpyAsync(ctx->h_gather, ctx->d_gather, (size_t)R * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->pt_stream));
        CU(cudaStreamSynchronize(ctx->pt_stream));
        return octo_pt_decide(ctx->h_gather, beta, chain_of_replica, R, round, ctx->pt_seed, accepted);
    }
    return octo_pt_decide(ll_pair, beta, chain_of_replica, R, round, ctx->pt_seed, accepted);
}

// Device-ordered swap round: the local (l_ref, l_target) pairs are already on the device; one
// ncclAllGather on `stream`, then the decision kernel on the gathered buffer, no host synchronisation.  The rung
// assignment lives on the device, replicated on every rank (every rank takes every decision).
int octo_pt_swap_round_device(OctoCtx* ctx, const double* d_ll_pair_local, const double* d_ladder, int32_t* d_chain_of_rung,
                              int32_t* d_rung_of_chain, double* d_swap_count, double* d_beta_local, int64_t round, void* stream) {
    if (!ctx || !ctx->pt_stream) return fail(OCTO_ERR_STATE, "octo_pt_init has not been called");
    if (!d_ll_pair_local || !d_ladder || !d_chain_of_rung || !d_rung_of_chain || !d_swap_count) return fail(OCTO_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    const int nl = ctx->pt_local, R = nl * ctx->pt_world;
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->pt_world > 1) {
        int r = g_nccl.AllGather(d_ll_pair_local, ctx->d_gather, (size_t)nl * 2, /*ncclFloat64*/ 8, ctx->nccl_comm, st);
        if (r) return fail_nccl(r, "ncclAllGather");
    } else {
        CU(cudaMemcpyAsync(ctx->d_gather, d_ll_pair_local, (size_t)nl * 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    PtDist d{ctx->d_gather, nullptr, const_cast<double*>(d_ladder), d_swap_count, d_chain_of_rung, d_rung_of_chain};
    cudaError_t e = octo_pt_swap_dist_launch(d, d_beta_local, (int64_t)ctx->pt_rank * nl, d_beta_local ? nl : 0, R, round, ctx->pt_seed, st);
    if (e != cudaSuccess) return fail_cuda(e, "k_pt_swap_dist launch");
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return OCTO_OK;
}

void octo_pt_finalize(OctoCtx* ctx) {
    if (!ctx) return;
    if (ctx->nccl_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    if (ctx->d_gather) { cudaFree(ctx->d_gather); ctx->d_gather = nullptr; }
    if (ctx->h_gather) { cudaFreeHost(ctx->h_gather); ctx->h_gather = nullptr; }
    if (ctx->pt_stream) { cudaStreamDestroy(ctx->pt_stream); ctx->pt_stream = nullptr; }
}

}  // extern "C"
