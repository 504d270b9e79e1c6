// octo_oracle.cpp — C entry points of the CPU oracle (see octo_oracle.hpp for scope and
// parity status).  TEST INFRASTRUCTURE ONLY: never linked into libocto_b200.so.
#include "octo_oracle.hpp"
#include "octo_oracle_param.hpp"
#include <cstring>
#include <string>
#include <algorithm>
#include <thread>

// chains are independent: static block partition over std::threads
template <class F>
static void parallel_chains(int64_t n, int n_threads, F&& body) {
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : 1, n));
    if (nt == 1) { for (int64_t c = 0; c < n; ++c) body(c); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        th.emplace_back([lo, hi, &body]() { for (int64_t c = lo; c < hi; ++c) body(c); });
    }
    for (auto& x : th) x.join();
}

using namespace octo_oracle;

static thread_local std::string g_err;

static int validate(const OctoLayout* L, const OctoObsBlock* blocks, int n_blocks) {
    if (!L || L->n_planets < 1 || L->n_planets > OCTO_MAX_PLANETS || L->n_in < 1) { g_err = "bad layout"; return OCTO_ERR_ARG; }
    for (int b = 0; b < n_blocks; ++b) {
        const OctoObsBlock& B = blocks[b];
        if (B.kind < 0 || B.kind > 5) { g_err = "bad kind"; return OCTO_ERR_ARG; }
        const bool sys = (B.kind == OCTO_KIND_RV_STAR_ABS || B.kind == OCTO_KIND_RV_STAR_MARGIN || B.kind == OCTO_KIND_HGCA_INSTANT);
        if (B.kind == OCTO_KIND_HGCA_INSTANT && (!B.aux || B.idx_pmra < 0 || B.idx_pmdec < 0)) { g_err = "HGCA needs aux, pmra, pmdec"; return OCTO_ERR_ARG; }
        if (!sys && (B.planet < 0 || B.planet >= L->n_planets)) { g_err = "bad planet index"; return OCTO_ERR_ARG; }
        if (sys) for (int p = 0; p < L->n_planets; ++p)
            if (L->idx_mass[p] < 0) { g_err = "star RV needs a mass variable on every planet"; return OCTO_ERR_ARG; }
        if (B.kind == OCTO_KIND_RV_STAR_MARGIN && B.idx_jitter < 0) { g_err = "marginalised RV needs jitter"; return OCTO_ERR_ARG; }
        const bool rv = (B.kind == OCTO_KIND_RV_STAR_ABS || B.kind == OCTO_KIND_RV_STAR_MARGIN || B.kind == OCTO_KIND_RV_PLANET_REL);
        if (rv) for (int p = 0; p < L->n_planets; ++p)
            if (L->basis[p] == OCTO_BASIS_THIELE_INNES) { g_err = "radial velocities with a Thiele-Innes planet are not offloaded"; return OCTO_ERR_ARG; }
    }
    return OCTO_OK;
}

template <int N>
static void eval_grad_chain(const OctoConstants& c, const OctoLayout& L, const OctoObsBlock* blocks, int n_blocks,
                            const double* in, int64_t ld, int64_t ch, double* ll, double* g, int64_t ldg) {
    Dual<N> X[N];
    for (int k = 0; k < L.n_in; ++k) X[k] = seed<N>(in[ch + k * ld], k);
    Dual<N> r = ln_like_chain<Dual<N>>(c, L, blocks, n_blocks, X);
    ll[ch] = r.v;
    for (int k = 0; k < L.n_in; ++k) g[ch + k * ldg] = r.d[k];
}

extern "C" {

const char* octo_oracle_last_error(void) { return g_err.c_str(); }

double octo_oracle_kepler(double MA, double e) { return kepler_markley(MA, e); }
double octo_oracle_rem2pi(double x) { return rem2pi_nearest(x); }

// ra/dec [mas] and rv [m/s] of one orbit at n epochs (a3, a4, a6, a7) — for fixture pins
int octo_oracle_orbit_radecrv(const OctoConstants* c, double a, double e, double i, double w, double W, double tp,
                              double M, double plx, const double* t, int n, double* ra, double* dec, double* rv) {
    Orbit<double> o = make_orbit<double>(*c, a, e, i, w, W, tp, M, plx);
    for (int k = 0; k < n; ++k) {
        Solution<double> s = orbitsolve(*c, o, t[k]);
        if (ra) ra[k] = raoff(o, s);
        if (dec) dec[k] = decoff(o, s);
        if (rv) rv[k] = radvel(o, s);
    }
    return 0;
}

int octo_oracle_logp(const OctoConstants* c, const OctoLayout* L, const OctoObsBlock* blocks, int n_blocks,
                     const double* in, int64_t n_chains, int64_t ld, double* ll, int n_threads) {
    if (int rc = validate(L, blocks, n_blocks)) return rc;
    const double ninf = -std::numeric_limits<double>::infinity();
    if (L->n_in > 256) { g_err = "oracle supports n_in <= 256"; return OCTO_ERR_ARG; }
    parallel_chains(n_chains, n_threads, [&](int64_t ch) {
        if (!chain_valid(*L, in, ld, ch)) { ll[ch] = ninf; return; }
        double X[256];
        for (int k = 0; k < L->n_in; ++k) X[k] = in[ch + k * ld];
        ll[ch] = ln_like_chain<double>(*c, *L, blocks, n_blocks, X);
    });
    return OCTO_OK;
}

int octo_oracle_logp_grad(const OctoConstants* c, const OctoLayout* L, const OctoObsBlock* blocks, int n_blocks,
                          const double* in, int64_t n_chains, int64_t ld, double* ll, double* g, int n_threads) {
    if (int rc = validate(L, blocks, n_blocks)) return rc;
    if (L->n_in > 48) { g_err = "oracle gradient supports n_in <= 48"; return OCTO_ERR_ARG; }
    const double ninf = -std::numeric_limits<double>::infinity();
    parallel_chains(n_chains, n_threads, [&](int64_t ch) {
        if (!chain_valid(*L, in, ld, ch)) {
            ll[ch] = ninf;
            for (int k = 0; k < L->n_in; ++k) g[ch + k * ld] = 0.0;
            return;
        }
        const int n = L->n_in;   // chunk = n_in, rounded up to the next instantiated width
        if (n <= 8)        eval_grad_chain<8>(*c, *L, blocks, n_blocks, in, ld, ch, ll, g, ld);
        else if (n <= 12)  eval_grad_chain<12>(*c, *L, blocks, n_blocks, in, ld, ch, ll, g, ld);
        else if (n <= 16)  eval_grad_chain<16>(*c, *L, blocks, n_blocks, in, ld, ch, ll, g, ld);
        else if (n <= 24)  eval_grad_chain<24>(*c, *L, blocks, n_blocks, in, ld, ch, ll, g, ld);
        else if (n <= 32)  eval_grad_chain<32>(*c, *L, blocks, n_blocks, in, ld, ch, ll, g, ld);
        else               eval_grad_chain<48>(*c, *L, blocks, n_blocks, in, ld, ch, ll, g, ld);
    });
    return OCTO_OK;
}

}  // extern "C"

template <int N>
static void eval_post_chain(const OctoConstants& c, const OctoLayout& L, const OctoObsBlock* blocks, int n_blocks,
                            const OctoPrior* priors, int D, const OctoInputDef* defs, const double* th, int64_t ld,
                            int64_t ch, double* lp, double* g) {
    Dual<N> X[N];
    for (int k = 0; k < D; ++k) X[k] = seed<N>(th[ch + k * ld], k);
    Dual<N> r = logpost_chain<Dual<N>>(c, L, blocks, n_blocks, priors, D, defs, X);
    lp[ch] = r.v;
    for (int k = 0; k < D; ++k) g[ch + k * ld] = std::isfinite(r.v) ? r.d[k] : 0.0;
}

extern "C" {

// ℓπcallback / ∇ℓπcallback of the standard parameterisation (theta_t: column-major [n x D])
int octo_oracle_logpost(const OctoConstants* c, const OctoLayout* L, const OctoObsBlock* blocks, int n_blocks,
                        const OctoPrior* priors, int D, const OctoInputDef* defs, const double* theta_t, int64_t n,
                        int64_t ld, double* lp, double* g, int n_threads) {
    if (int rc = validate(L, blocks, n_blocks)) return rc;
    if (D < 1 || D > 48) { g_err = "oracle logpost supports 1 <= D <= 48"; return OCTO_ERR_ARG; }
    parallel_chains(n, n_threads, [&](int64_t ch) {
        if (!g) {
            double X[48];
            for (int k = 0; k < D; ++k) X[k] = theta_t[ch + k * ld];
            lp[ch] = logpost_chain<double>(*c, *L, blocks, n_blocks, priors, D, defs, X);
            return;
        }
        if (D <= 8)       eval_post_chain<8>(*c, *L, blocks, n_blocks, priors, D, defs, theta_t, ld, ch, lp, g);
        else if (D <= 12) eval_post_chain<12>(*c, *L, blocks, n_blocks, priors, D, defs, theta_t, ld, ch, lp, g);
        else if (D <= 16) eval_post_chain<16>(*c, *L, blocks, n_blocks, priors, D, defs, theta_t, ld, ch, lp, g);
        else if (D <= 24) eval_post_chain<24>(*c, *L, blocks, n_blocks, priors, D, defs, theta_t, ld, ch, lp, g);
        else if (D <= 32) eval_post_chain<32>(*c, *L, blocks, n_blocks, priors, D, defs, theta_t, ld, ch, lp, g);
        else              eval_post_chain<48>(*c, *L, blocks, n_blocks, priors, D, defs, theta_t, ld, ch, lp, g);
    });
    return OCTO_OK;
}

// ln_like(system, arr2nt(invlink(θ_t))) alone, -Inf where not finite (octofit_rejection, src/sampling.jl:261-270)
int octo_oracle_loglike_theta(const OctoConstants* c, const OctoLayout* L, const OctoObsBlock* blocks, int n_blocks,
                              const OctoPrior* priors, int D, const OctoInputDef* defs, const double* theta_t, int64_t n,
                              int64_t ld, double* ll, int n_threads) {
    if (int rc = validate(L, blocks, n_blocks)) return rc;
    if (D < 1 || D > 48) { g_err = "oracle logpost supports 1 <= D <= 48"; return OCTO_ERR_ARG; }
    parallel_chains(n, n_threads, [&](int64_t ch) {
        double X[48];
        for (int k = 0; k < D; ++k) X[k] = theta_t[ch + k * ld];
        const double v = logpost_chain<double>(*c, *L, blocks, n_blocks, priors, D, defs, X, true);
        ll[ch] = std::isfinite(v) ? v : -std::numeric_limits<double>::infinity();
    });
    return OCTO_OK;
}

int octo_oracle_invlink(const OctoPrior* priors, int D, const double* theta_t, int64_t n, int64_t ld, double* theta_nat) {
    for (int64_t ch = 0; ch < n; ++ch)
        for (int k = 0; k < D; ++k) theta_nat[ch + k * ld] = prior_invlink<double>(priors[k], theta_t[ch + k * ld]);
    return OCTO_OK;
}

double octo_oracle_tperi(const OctoConstants* c, double theta, double t_ref, double M, double e, double a, double i,
                         double w, double W) {
    return theta_at_epoch_to_tperi<double>(*c, theta, t_ref, M, e, a, i, w, W);
}

int octo_oracle_max_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

}  // extern "C"
