// octo_hmc_dev.cuh — the per-coordinate arithmetic of the chain-batched HMC explorer, shared by the launch-per-leapfrog
// path (octo_hmc.cu: k_hmc_turn / k_hmc_leap) and the trajectory-resident kernel (octo_kernels.cu: k_hmc_resident) so
// that both produce the same bits.  Every rounding is spelled out (no implicit contraction): __dmul_rn / __dadd_rn / fma.
// Randomness is counter-based: splitmix64 of (seed, transition, chain, coordinate); octo_hmc_random is its host twin.
#pragma once
#include <stdint.h>

namespace octo_hmc_dev {

__host__ __device__ inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ inline double u01(uint64_t s) { return ((double)(s >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
// stream of (seed, iteration): coordinate j of chain c draws from splitmix64(key ^ (c * K1 + j * K2))
__host__ __device__ inline uint64_t hmc_key(uint64_t seed, uint64_t it) { return splitmix64(seed ^ splitmix64(it + 1)); }
__host__ __device__ inline uint64_t hmc_draw(uint64_t key, uint64_t chain, uint64_t j) {
    return splitmix64(key ^ (chain * 0x9E3779B97F4A7C15ULL + j * 0xD1B54A32D192ED03ULL));
}
// the uniform of octo_pt_decide (octo_shim.cu): the device swap rounds take the same decisions
__host__ __device__ inline double pt_uniform_dev(uint64_t seed, uint64_t round, uint64_t pair) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (round * 0x100000001B3ULL + pair + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

#ifdef __CUDACC__
// Parallel tempering: what a chain contributes to a swap decision — l_ref = lp - beta ll (the prior terms) and
// l_target = l_ref + ll — and the log acceptance ratio of exchanging the rungs (weights bi, bj) of chains a and b.
// Shared by the single-GPU swap kernel, the resident kernel's epilogue (which packs the pair for the all-gather) and the
// sharded swap kernel: identical roundings everywhere, hence identical decisions.
__device__ __forceinline__ void pt_pair(double lp, double beta, double ll, double& l_ref, double& l_target) {
    l_ref = fma(-beta, ll, lp);
    l_target = __dadd_rn(l_ref, ll);
}
__device__ __forceinline__ double pt_log_ratio(double ra, double ta, double rb, double tb, double bi, double bj) {
    const double Vaj = fma(bj, ta, __dmul_rn(__dadd_rn(1.0, -bj), ra)), Vbi = fma(bi, tb, __dmul_rn(__dadd_rn(1.0, -bi), rb));
    const double Vai = fma(bi, ta, __dmul_rn(__dadd_rn(1.0, -bi), ra)), Vbj = fma(bj, tb, __dmul_rn(__dadd_rn(1.0, -bj), rb));
    return __dadd_rn(__dadd_rn(Vaj, Vbi), -__dadd_rn(Vai, Vbj));
}
// Start of a transition, coordinate j of chain c: fresh momentum p ~ N(0, 1/inv_mass) (Box-Muller), its kinetic term
// p² inv_mass (BEFORE the kick), then the first half kick with the current gradient g and the first drift from q.
struct HmcStart { double kin, p, q; };
__device__ __forceinline__ HmcStart hmc_start(uint64_t key, uint64_t chain, int j, double im, double eps, double g, double q) {
    const uint64_t s = hmc_draw(key, chain, (uint64_t)j);
    const double z = __dmul_rn(sqrt(__dmul_rn(-2.0, log(u01(s)))), cospi(__dmul_rn(2.0, u01(splitmix64(s)))));
    double pj = __dmul_rn(z, rsqrt(im));
    HmcStart r;
    r.kin = __dmul_rn(__dmul_rn(pj, pj), im);
    pj = fma(__dmul_rn(0.5, eps), g, pj);
    r.p = pj;
    r.q = fma(__dmul_rn(eps, pj), im, q);
    return r;
}
// kinetic term of coordinate j at the end of a trajectory
__device__ __forceinline__ double hmc_kin(double pj, double im) { return __dmul_rn(__dmul_rn(pj, pj), im); }
// Metropolis decision of transition `it` for chain c: h0 = -lp + kin0/2 at the start, lpp / kin1 at the end
__device__ __forceinline__ bool hmc_accept(uint64_t seed, int it, uint64_t chain, int D, double h0, double lpp, double kin1) {
    const double h1 = fma(0.5, kin1, -lpp);
    const double u = u01(hmc_draw(hmc_key(seed, (uint64_t)it), chain, (uint64_t)D));
    return isfinite(lpp) && (log(u) < __dadd_rn(h0, -h1));
}
__device__ __forceinline__ double hmc_h0(double lp, double kin0) { return fma(0.5, kin0, -lp); }
#endif

}  // namespace octo_hmc_dev
