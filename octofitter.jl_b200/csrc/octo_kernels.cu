// octo_kernels.cu — the fused sm_100a kernel of the hot path.
//
// One launch evaluates, for every (chain, epoch) pair of a batch:
//   Kepler solve (Markley 1995, branch-uniform, non-iterative)      ref a4/a5: PlanetOrbits orbitsolve + kepler_solver
//   sky-plane ΔRA/ΔDec [mas] and radial velocity [m/s]              ref a6/a7: raoff/decoff/radvel
//   Gaussian log-likelihood terms of the five observation kinds     ref a9-a12: relative-astrometry.jl:166-253, rv-*.jl
//   the analytic adjoint w.r.t. the orbital elements / obs variables ref a13: replaces the ForwardDiff pass
// and reduces over epochs to one ll and one gradient row per chain.
//
// Mapping (all FP64, no tensor cores — the work is transcendental/irregular, not a contraction):
//   lane      -> chain of a 32-chain group (chain is the fastest index of `in`, so loads/stores coalesce)
//   warp      -> one contiguous range of the concatenated epoch list; epoch data are warp-uniform
//                (one broadcast load serves 32 pairs) and the observation kind never diverges in a warp
//   CTA       -> 8 warps = 8 epoch ranges of the same chain group; per-chain constants are computed once
//                per CTA by the prologue and staged in shared memory
//   grid      -> (chain groups) x (epoch splits); partial accumulators of the splits are combined in a fixed
//                order by the last CTA to finish (ticket per chain group) => run-to-run bit-reproducible
//   registers -> per-segment raw accumulators; folded into per-warp shared-memory slots at segment end
//   epilogue  -> one lane per chain maps the raw sums to ∂ll/∂(inputs) by the chain rule, writes coalesced rows
#include "octo_internal.h"
#include "octo_param_dev.cuh"
#include "octo_hmc_dev.cuh"
#include <math_constants.h>
#include <cstdio>

namespace {

constexpr int WMAX = OCTO_WARPS;   // warps per CTA (models whose accumulator slots would not fit launch with fewer)
#ifndef OCTO_MIN_CTAS
#define OCTO_MIN_CTAS 2          // resident CTAs/SM the register allocator targets
#endif
#ifndef OCTO_UNROLL
#define OCTO_UNROLL 1            // epochs in flight per warp in the lean loops
#endif
#ifndef OCTO_LAT_ILP
#define OCTO_LAT_ILP 2           // pairs in flight per lane in the lean loops of the latency-tuned instantiation / resident kernel
#endif
#ifndef OCTO_THR_ILP
#define OCTO_THR_ILP 2           // the same for the throughput instantiation (128 registers: +4 % on the C5 right end, measured)
#endif
#define OCTO_PRAGMA(x) _Pragma(#x)
#define OCTO_UNROLL_LOOP(n) OCTO_PRAGMA(unroll n)
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.283185307179586477;
constexpr double kLog2Pi = 1.8378770664093454836;
// Markley eq. 20: alpha = kA0 + ca1 * (pi - |M|),  ca1 = 1.6 pi / ((pi^2 - 6)(1 + e))
constexpr double kA0 = 3.0 * kPi * kPi / (kPi * kPi - 6.0);
constexpr double kA1 = 1.6 * kPi / (kPi * kPi - 6.0);

struct Orb { double nd, tp, e; float ef, omef, ca1f; };

// ---------------------------------------------------------------------------------------------
// Branch-free FP64 building blocks with known input ranges (no libm slow paths, no divergence).
// FP64 literals live in the constant bank: DFMA/DMUL/DADD read them as c[3][..] operands, so the hot loop
// spends no issue slots (MOV/UMOV pairs) or registers on materialising 64-bit immediates.
// ---------------------------------------------------------------------------------------------
struct KConst {
    double magic, inv_two_pi, two_pi1, two_pi2, two_pi3, two_over_pi, pio2_hi, pio2_lo;
    double s[6], c[6];
    double one, half, mhalf, sixth, msixth, r24, mr24, two, mtwo, five, two_pi, log2pi;
};
__constant__ KConst kc = {
    6755399441055744.0,            // 1.5 * 2^52: (x + magic) - magic == rint(x) for |x| < 2^51
    // 1/2pi, then 2pi split 33 + 33 + 53 bits for a 3-term Cody-Waite reduction (exact to < 1e-30 |k|)
    0.15915494309189533577, 0x1.921fb54400000p+2, 0x1.0b4611a600000p-32, 0x1.3198a2e037073p-67,
    0.63661977236758138243, 1.57079632679489655800e+00, 6.12323399573676603587e-17,
    // fdlibm __kernel_sin S1..S6 (highest degree first), __kernel_cos C1..C6 (highest first)
    {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
     -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01},
    {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
     2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02},
    1.0, 0.5, -0.5, 1.0 / 6.0, -1.0 / 6.0, 1.0 / 24.0, -1.0 / 24.0, 2.0, -2.0, 5.0,
    6.283185307179586477, 1.8378770664093454836};

__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 1/x: MUFU.RCP64H seed (~2^-20) + two Newton steps -> ~1 ulp.  x finite, normal, non-zero.
__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, kc.one);
    y = fma(y, e, y);
    e = fma(-x, y, kc.one);
    return fma(y, e, y);
}

// sin and cos for |x| <= ~4 (here x = E1 in [-pi, pi]): one quadrant reduction, fdlibm kernel polynomials on
// [-pi/4, pi/4] (< 1 ulp), quadrant fix-up by selects.
__device__ __forceinline__ void sincos_pi(double x, double& s, double& c) {
    const double tq = fma(x, kc.two_over_pi, kc.magic);
    const int q = __double2loint(tq);
    const double kq = tq - kc.magic;
    double r = fma(-kq, kc.pio2_hi, x);
    r = fma(-kq, kc.pio2_lo, r);
    const double z = r * r;
    double ps = fma(z, kc.s[0], kc.s[1]);
    ps = fma(z, ps, kc.s[2]);
    ps = fma(z, ps, kc.s[3]);
    ps = fma(z, ps, kc.s[4]);
    ps = fma(z, ps, kc.s[5]);
    const double sr = fma(r * z, ps, r);
    double pc = fma(z, kc.c[0], kc.c[1]);
    pc = fma(z, pc, kc.c[2]);
    pc = fma(z, pc, kc.c[3]);
    pc = fma(z, pc, kc.c[4]);
    pc = fma(z, pc, kc.c[5]);
    const double cr = fma(z * z, pc, fma(z, kc.mhalf, kc.one));
    const double s0 = (q & 1) ? cr : sr;
    const double c0 = (q & 1) ? sr : cr;
    s = (q & 2) ? -s0 : s0;
    c = ((q + 1) & 2) ? -c0 : c0;
}

__device__ __forceinline__ void sincos_any(double x, double& s, double& c) {   // any finite angle of sane size
    const double k = fma(x, kc.inv_two_pi, kc.magic) - kc.magic;
    double r = fma(-k, kc.two_pi1, x);
    r = fma(-k, kc.two_pi2, r);
    r = fma(-k, kc.two_pi3, r);
    sincos_pi(r, s, c);
}

// ---------------------------------------------------------------------------------------------
// Kepler solve: rem2pi (round to nearest) + Markley's starter + one fifth-order correction; returns sinE, cosE.
//  * The starter E1 (Markley eqs 5-15, |E1 - E| < 4.4e-4 rad everywhere) only seeds the correction, whose
//    result is accurate to O(|E1-E|^6); it is therefore evaluated in FP32 with MUFU rsqrt/lg2/ex2/rcp —
//    validated over e in [0, 1-1e-12], |M| down to 1e-30: identical max |d5| and residual <= 7e-16.
//  * f0 = E1 - e sinE1 - M and the fifth-order correction (eqs 21-28) are FP64, with ONE reciprocal (rcp_nr).
//  * sin/cos of E = E1 + d5 come from rotating sincos(E1) by d5 (three Taylor terms, exact to 1e-19):
//    one sincos per solve instead of two.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void kepler_sincos(const Orb& o, double t, double& dt, double& sE, double& cE, double* M_out = nullptr) {
    dt = t - o.tp;
    const double MA = o.nd * dt;
    const double k = fma(MA, kc.inv_two_pi, kc.magic) - kc.magic;     // rint(MA / 2pi)
    double M = fma(-k, kc.two_pi1, MA);
    M = fma(-k, kc.two_pi2, M);
    M = fma(-k, kc.two_pi3, M);
    if (M_out) *M_out = M;                                            // reduced mean anomaly = E - e sin E
    // ---- starter, FP32
    const float Mf = (float)M, ef = o.ef, omef = o.omef;
    const float alpha = fmaf(o.ca1f, (float)kPi - fabsf(Mf), (float)kA0);      // eq 20
    const float d = fmaf(alpha, ef, 3.0f * omef);                              // eq 5
    const float M2 = Mf * Mf;
    const float ad = alpha * d;
    const float q = fmaf(2.0f * ad, omef, -M2);                                // eq 9
    const float r = Mf * fmaf(3.0f * ad, d - omef, M2);                        // eq 10
    const float q2 = q * q;
    const float disc = fmaxf(fmaf(q2, q, r * r), 1e-30f);
    const float tt = fmaf(disc, mufu_rsqrt(disc), fabsf(r));                   // |r| + sqrt(q^3 + r^2)
    const float w = mufu_ex2(mufu_lg2(tt) * (2.0f / 3.0f));                    // eq 14: cbrt(tt^2)
    const float den = fmaf(w, w + q, q2);
    const double E1 = (double)(fmaf(Mf, den, 2.0f * r * w) * mufu_rcp(den * d));   // eq 15
    // ---- correction, FP64
    double s1, c1;
    sincos_pi(E1, s1, c1);
    const double f2 = o.e * s1, f3 = o.e * c1;                     // eqs 26, 27
    const double f0 = (E1 - M) - f2;                               // eq 21
    const double f1 = kc.one - f3;                                 // eq 25
    // eqs 22-24, 28: Markley's d5 is the root of f0 + f1 d + f2 d^2/2 + f3 d^3/6 - f2 d^4/24 = 0 reached by three
    // nested divisions.  The same root by series reversion in the Newton step u = -f0/f1 (|u| < 4.4e-4, and
    // |a2 u| < 2.5e-4 over the whole domain): d = u (1 - u (a2 - u (c3 - u c4))), one reciprocal, a third
    // of the dependent chain.  Agrees with the nested form to 1.1e-17 rad (validated e <= 1 - 1e-9, |M| >= 1e-30).
    const double h = rcp_nr(f1);
    const double u = -f0 * h;
    const double g2 = f2 * h, g3 = f3 * h;
    const double a2 = kc.half * g2, a3 = kc.sixth * g3, a4 = kc.mr24 * g2;
    const double c3 = fma(a2, g2, -a3);                            // 2 a2^2 - a3
    const double c4 = fma(kc.five * a2, fma(a2, a2, -a3), a4);     // 5 a2^3 - 5 a2 a3 + a4
    const double d5 = u * fma(-u, fma(-u, fma(-u, c4, c3), a2), kc.one);
    const double x2 = d5 * d5;
    const double sd = d5 * fma(x2, kc.msixth, kc.one);
    const double cd = fma(x2, fma(x2, kc.r24, kc.mhalf), kc.one);
    sE = fma(s1, cd, c1 * sd);
    cE = fma(c1, cd, -s1 * sd);
}

__device__ __forceinline__ Orb load_orb(const double* sc, int lane) {
    Orb o;
    o.nd = sc[PC_nd * 32 + lane]; o.tp = sc[PC_tp * 32 + lane]; o.e = sc[PC_e * 32 + lane];
    o.ef = (float)o.e; o.omef = (float)sc[PC_ome * 32 + lane]; o.ca1f = (float)sc[PC_ca1 * 32 + lane];
    return o;
}

__device__ __forceinline__ void acc_add(double* acc, int slot, int lane, double v) { acc[slot * 32 + lane] += v; }

// ---------------------------------------------------------------------------------------------
// Astrometry segment (kinds 0, 1): epochs [k0, k1) of table B for this warp's 32 chains.
// ---------------------------------------------------------------------------------------------
// MODE 0: RA/Dec table with fixed weights (no jitter / platescale / northangle) — the common case
// MODE 1: RA/Dec table with a sampled jitter (per-pair covariance), no platescale / northangle
// MODE 2: everything else (PA/sep tables, platescale, northangle), decided at run time
// Lean tables: the loop is branch-free — lanes past the end of their range evaluate neutral records (zero weights: every
// contribution is an exact zero) — and, for ILP > 1 (latency-tuned launches), unrolled so that the dependent chains of
// ILP consecutive pairs overlap in one thread.  Same sums, same bits as a loop that skips those lanes.
template <bool GRAD, int NPT, int MODE, int ILP>
__device__ __forceinline__ void seg_astrom(const DevModel& m, const DevBlock& B, int k0, int k1, const double* s_const,
                                        double* acc, double2* stage, const double* __restrict__ in, int64_t c, int64_t ld, int lane, int ch) {
    const int ip = B.planet;
    constexpr bool LEAN = (MODE == 0);
    constexpr bool PAD = LEAN;
    constexpr int UNR = (PAD && ILP > 1) ? ILP : OCTO_UNROLL;
    const bool pasep = (MODE == 2) && (B.kind == OCTO_KIND_ASTROM_PASEP);
    const bool jitm = (MODE == 1) || (MODE == 2 && B.jit);
    // involved planets: the observed one, then interior companions with a mass (relative-astrometry.jl:117-133)
    int pj[NPT]; double f[NPT]; Orb orb[NPT]; double Bh[NPT], Gs[NPT], Ah[NPT], Fs[NPT];
    int ni = 1;
    pj[0] = ip; f[0] = 1.0;
    const double a_ip = s_const[(ip * PC_COUNT + PC_a) * 32 + lane];
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
        if (NPT > 1 && j < m.n_planets && j != ip && m.idx_mass[j] >= 0) {
            const double* sc = s_const + j * PC_COUNT * 32;
            const bool inner = sc[PC_a * 32 + lane] < a_ip;
            if (__any_sync(0xffffffffu, inner)) {          // warp-uniform: skip the solve if no chain needs it
#pragma unroll
                for (int u = 1; u < NPT; ++u) if (u == ni) { pj[u] = j; f[u] = inner ? sc[PC_mu * 32 + lane] : 0.0; }
                ++ni;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NPT; ++u) if (u < ni) {
        const double* sc = s_const + pj[u] * PC_COUNT * 32;
        orb[u] = load_orb(sc, lane);
        Bh[u] = sc[PC_Bh * 32 + lane]; Gs[u] = sc[PC_Gs * 32 + lane];
        Ah[u] = sc[PC_Ah * 32 + lane]; Fs[u] = sc[PC_Fs * 32 + lane];
    }
    const double jit = (!LEAN && B.idx_jitter >= 0) ? in[c + (int64_t)B.idx_jitter * ld] : 0.0;
    const double ps = (MODE == 2 && B.idx_platescale >= 0) ? in[c + (int64_t)B.idx_platescale * ld] : 1.0;
    const double na = (MODE == 2 && B.idx_northangle >= 0) ? in[c + (int64_t)B.idx_northangle * ld] : 0.0;
    const bool rot = (MODE == 2) && ((B.idx_platescale >= 0) || (B.idx_northangle >= 0));
    double sna = 0.0, cna = 1.0;
    if (rot && !pasep) sincos_any(na, sna, cna);
    const double j2 = jit * jit;

    double ll = 0.0, g_jit = 0.0, g_ps = 0.0, g_na = 0.0;
    // ObsPriorAstromONeil2019 wrapper (generic path only): Σ|f| and Σ sign(f) ∂f/∂(MA, MA·dt, e) of the observed planet
    const bool obsprior = (MODE == 2) && B.slot_obsprior >= 0;
    double op_S = 0.0, op_Q0 = 0.0, op_Q1 = 0.0, op_Qe = 0.0;
    double L[NPT][7];
#pragma unroll
    for (int u = 0; u < NPT; ++u)
#pragma unroll
        for (int a = 0; a < 7; ++a) L[u][a] = 0.0;

    // Epoch records [t, y1 | c1, y2 | c2, c3] are staged through shared memory 32 at a time: one coalesced
    // read-only load per lane (three 16-byte words), then every iteration reads its record as a broadcast LDS —
    // no global-load latency inside the dependent chain.
    // With sub-lanes (ch < 32 chains per warp) every group of ch lanes walks its own range [k0, k1) through its own
    // window of ch records; the trip counts are made warp-uniform (lanes past their range idle).
    const double2* __restrict__ tab = reinterpret_cast<const double2*>(m.tab);
    const int col = lane & (ch - 1), sbase = lane - col;
    for (int kb = k0;; kb += ch) {
    const int nrec = max(0, min(ch, k1 - kb));
    const int nmax = __reduce_max_sync(0xffffffffu, nrec);
    if (nmax == 0) break;
    __syncwarp();
    if (col < nrec) {
        const double2* src = tab + 3 * (int64_t)(kb + col);
        const double2 a0 = __ldg(src), a1 = __ldg(src + 1), a2 = __ldg(src + 2);
        stage[3 * lane] = a0; stage[3 * lane + 1] = a1; stage[3 * lane + 2] = a2;
    } else if (PAD) {
        const double2 z = make_double2(0.0, 0.0);
        stage[3 * lane] = z; stage[3 * lane + 1] = z; stage[3 * lane + 2] = z;
    }
    __syncwarp();
#pragma unroll (UNR)
    for (int j = 0; j < nmax; ++j) {
        if (!PAD && j >= nrec) continue;
        const double2 ra0 = stage[3 * (sbase + j)], ra1 = stage[3 * (sbase + j) + 1], ra2 = stage[3 * (sbase + j) + 2];
        const double t = ra0.x, y1 = ra0.y, e1 = ra1.x, y2 = ra1.y, e2 = ra2.x, e3 = ra2.y;
        double sE[NPT], cE[NPT], dt[NPT];
        double ra = 0.0, dec = 0.0, Mred = 0.0;
#pragma unroll
        for (int u = 0; u < NPT; ++u) if (u < ni) {
            if (MODE == 2 && u == 0) kepler_sincos(orb[u], t, dt[u], sE[u], cE[u], &Mred);
            else kepler_sincos(orb[u], t, dt[u], sE[u], cE[u]);
            const double X = cE[u] - orb[u].e;
            const double ra_u = fma(X, Bh[u], sE[u] * Gs[u]);
            const double dec_u = fma(X, Ah[u], sE[u] * Fs[u]);
            if (u == 0) { ra = ra_u; dec = dec_u; } else { ra = fma(f[u], ra_u, ra); dec = fma(f[u], dec_u, dec); }
        }
        double r1, r2, rho = 0.0, irho = 0.0, ra_d = y1, dec_d = y2;
        if (pasep) {
            rho = sqrt(fma(ra, ra, dec * dec));
            irho = rcp_nr(rho);
            const double pa = atan2(ra, dec);
            double pd = fmod((y1 + na) - pa + kPi, kTwoPi) - kPi;      // Julia `%` == fmod
            if (pd < -kPi) pd += kTwoPi;
            r1 = pd;
            r2 = fma(y2, ps, -rho);
        } else {
            if (rot) {   // data rotated by -northangle and scaled (relative-astrometry.jl:209-213)
                ra_d = ps * fma(y1, cna, y2 * sna);
                dec_d = ps * fma(y2, cna, -y1 * sna);
            }
            r1 = ra_d - ra;
            r2 = dec_d - dec;
        }
        double w11, w12, w22, iv1 = 0.0, iv2 = 0.0;
        if (!jitm) { w11 = e1; w12 = e2; w22 = e3; }
        else {
            const double v1 = e1 + j2, v2 = e2 + j2;
            iv1 = rcp_nr(v1); iv2 = rcp_nr(v2);
            double iom = 1.0, lom = 0.0;
            w12 = 0.0;
            if (B.has_cor) {
                const double om = fma(-e3, e3, 1.0);
                iom = rcp_nr(om); lom = log(om);
                w12 = -e3 * sqrt(iv1 * iv2) * iom;
            }
            w11 = iv1 * iom; w22 = iv2 * iom;
            ll -= kLog2Pi + 0.5 * (log(v1 * v2) + lom);
        }
        const double q1 = fma(w11, r1, w12 * r2), q2 = fma(w12, r1, w22 * r2);
        ll = fma(kc.mhalf, fma(r1, q1, r2 * q2), ll);
        if (obsprior) {       // f = 3M(e + cos E) + 2(-2 + e² + e cos E) sin E,  M = meananom = E - e sin E
            const double e = orb[0].e, sn = sE[0], cs = cE[0];
            const double t2 = fma(e, cs, fma(e, e, -2.0));
            const double fo = fma(3.0 * Mred, e + cs, 2.0 * t2 * sn);
            op_S += fabs(fo);
            if (GRAD) {
                const double sg = fo < 0.0 ? -1.0 : 1.0;
                const double D1m = fma(-e, cs, kc.one), rD = rcp_nr(D1m);
                // ∂f/∂E at fixed e (dM/dE = 1 - e cos E) and ∂f/∂e at fixed E (∂M/∂e = -sin E)
                const double dfdE = 3.0 * D1m * (e + cs) - 3.0 * Mred * sn - 2.0 * e * sn * sn + 2.0 * t2 * cs;
                const double dfde = -3.0 * sn * (e + cs) + 3.0 * Mred + 2.0 * fma(2.0, e, cs) * sn;
                const double gM = sg * dfdE * rD;              // dE/dMA = 1 / (1 - e cos E)
                op_Q0 += gM;
                op_Q1 = fma(gM, dt[0], op_Q1);
                op_Qe += fma(gM, sn, sg * dfde);               // dE/de = sin E / (1 - e cos E)
            }
        }
        if (GRAD) {
            double gr, gd;
            if (pasep) {
                const double a1 = q1 * irho * irho, a2 = q2 * irho;
                gr = fma(a1, dec, a2 * ra);
                gd = fma(-a1, ra, a2 * dec);
                g_na -= q1;
                g_ps -= q2 * y2;
            } else {
                gr = q1; gd = q2;
                if (rot) {
                    g_ps -= fma(q1, ra_d, q2 * dec_d);
                    g_na += fma(q2, ra_d, -q1 * dec_d);
                }
            }
            if (jitm) g_jit += fma(fma(r1, q1, -1.0), iv1, fma(r2, q2, -1.0) * iv2);
#pragma unroll
            for (int u = 0; u < NPT; ++u) if (u < ni) {
                const double X = cE[u] - orb[u].e;
                const double rD = rcp_nr(fma(-orb[u].e, cE[u], kc.one));
                L[u][0] = fma(gr, X, L[u][0]);
                L[u][1] = fma(gr, sE[u], L[u][1]);
                L[u][2] = fma(gd, X, L[u][2]);
                L[u][3] = fma(gd, sE[u], L[u][3]);
                const double gX = fma(gr, Bh[u], gd * Ah[u]);
                const double gS = fma(gr, Gs[u], gd * Fs[u]);
                const double gM = fma(cE[u], gS, -sE[u] * gX) * rD;
                L[u][4] += fma(gM, sE[u], -gX);
                L[u][5] += gM;
                L[u][6] = fma(gM, dt[u], L[u][6]);
            }
        }
    }
    }
    acc_add(acc, 0, lane, ll);
    if (obsprior) {
        acc_add(acc, B.slot_obsprior + OP_S, lane, op_S);
        if (GRAD) {
            acc_add(acc, B.slot_obsprior + OP_Q0, lane, op_Q0); acc_add(acc, B.slot_obsprior + OP_Q1, lane, op_Q1);
            acc_add(acc, B.slot_obsprior + OP_Qe, lane, op_Qe);
        }
    }
    if (GRAD) {
        if (!LEAN) {
            if (B.slot_jitter >= 0) acc_add(acc, B.slot_jitter, lane, g_jit * jit);
            if (B.slot_platescale >= 0) acc_add(acc, B.slot_platescale, lane, pasep ? g_ps : g_ps / ps);
            if (B.slot_northangle >= 0) acc_add(acc, B.slot_northangle, lane, g_na);
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) if (u < ni) {
            const int p = pj[u];
            const double fu = f[u];
            acc_add(acc, slot_planet(p, PA_Bh), lane, fu * L[u][0]);
            acc_add(acc, slot_planet(p, PA_Gs), lane, fu * L[u][1]);
            acc_add(acc, slot_planet(p, PA_Ah), lane, fu * L[u][2]);
            acc_add(acc, slot_planet(p, PA_Fs), lane, fu * L[u][3]);
            acc_add(acc, slot_planet(p, PA_e), lane, fu * L[u][4]);
            acc_add(acc, slot_planet(p, PA_S0), lane, fu * L[u][5]);
            acc_add(acc, slot_planet(p, PA_S1), lane, fu * L[u][6]);
            if (u > 0) {   // d f / d mu = [a_j < a_i]
                const double ind = (s_const[(p * PC_COUNT + PC_a) * 32 + lane] < a_ip) ? 1.0 : 0.0;
                const double gf = fma(Bh[u], L[u][0], fma(Gs[u], L[u][1], fma(Ah[u], L[u][2], Fs[u] * L[u][3])));
                acc_add(acc, slot_planet(p, PA_mu), lane, ind * gf);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Radial-velocity segment (kinds 2, 3, 4).
// ---------------------------------------------------------------------------------------------
// TREND: the table has a linear trend_function (its own instantiations: the lean loops do not carry the terms)
template <bool GRAD, int NPT, bool MARGIN, bool JIT, bool TREND, int ILP>
__device__ __forceinline__ void seg_rv(const DevModel& m, const DevBlock& B, int k0, int k1, const double* s_const,
                                    double* acc, double2* stage, const double* __restrict__ in, int64_t c, int64_t ld, int lane, int ch) {
    const bool star = (B.kind != OCTO_KIND_RV_PLANET_REL);
    constexpr bool margin = MARGIN;
    constexpr bool PAD = !MARGIN;                     // see seg_astrom
    constexpr int UNR = (PAD && ILP > 1) ? ILP : OCTO_UNROLL;
    int pj[NPT]; double f[NPT], dmu[NPT]; Orb orb[NPT]; double Pc[NPT], Ps[NPT];
    int ni = 0;
    if (star) {   // every planet, reflex of the star: -mu * radvel (rv-absolute.jl:145-154)
#pragma unroll
        for (int j = 0; j < NPT; ++j) if (j < m.n_planets) {
            pj[j] = j; f[j] = -s_const[(j * PC_COUNT + PC_mu) * 32 + lane]; dmu[j] = -1.0; ni = j + 1;
        }
    } else {      // the planet itself, plus interior companions with a mass (rv-relative.jl:143-160)
        const int ip = B.planet;
        pj[0] = ip; f[0] = 1.0; dmu[0] = 0.0; ni = 1;
        const double a_ip = s_const[(ip * PC_COUNT + PC_a) * 32 + lane];
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            if (NPT > 1 && j < m.n_planets && j != ip && m.idx_mass[j] >= 0) {
                const double* sc = s_const + j * PC_COUNT * 32;
                const bool inner = sc[PC_a * 32 + lane] < a_ip;
                if (__any_sync(0xffffffffu, inner)) {
#pragma unroll
                    for (int u = 1; u < NPT; ++u) if (u == ni) {
                        pj[u] = j; f[u] = inner ? -sc[PC_mu * 32 + lane] : 0.0; dmu[u] = inner ? -1.0 : 0.0;
                    }
                    ++ni;
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NPT; ++u) if (u < ni) {
        const double* sc = s_const + pj[u] * PC_COUNT * 32;
        orb[u] = load_orb(sc, lane);
        Pc[u] = sc[PC_Pc * 32 + lane]; Ps[u] = sc[PC_Ps * 32 + lane];
    }
    const double jit = JIT ? in[c + (int64_t)B.idx_jitter * ld] : 0.0;
    const double off = (B.idx_offset >= 0 && !margin) ? in[c + (int64_t)B.idx_offset * ld] : 0.0;
    const double j2 = jit * jit;
    // trend_function linear in <= 3 observation variables: coefficients per chain, basis values in the record
    const int tn = TREND ? B.n_trend : 0;
    double tc[3] = {0.0, 0.0, 0.0}, gT[3] = {0.0, 0.0, 0.0}, vT[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int v = 0; v < 3; ++v) if (v < tn) tc[v] = in[c + (int64_t)B.idx_trend[v] * ld];

    double ll = 0.0, g_jit = 0.0, g_off = 0.0;
    double mA = 0.0, mS1 = 0.0, mC = 0.0, mLG = 0.0, mR2 = 0.0, mR1 = 0.0, mQ = 0.0;
    double L[NPT][5], V[MARGIN ? NPT : 1][5];
#pragma unroll
    for (int u = 0; u < NPT; ++u)
#pragma unroll
        for (int a = 0; a < 5; ++a) { L[u][a] = 0.0; if constexpr (MARGIN) V[u][a] = 0.0; }

    const double2* __restrict__ tab = reinterpret_cast<const double2*>(m.tab);
    const int col = lane & (ch - 1), sbase = lane - col;       // sub-lanes: see seg_astrom
    for (int kb = k0;; kb += ch) {
    const int nrec = max(0, min(ch, k1 - kb));
    const int nmax = __reduce_max_sync(0xffffffffu, nrec);
    if (nmax == 0) break;
    __syncwarp();
    if (col < nrec) {
        const double2* src = tab + 3 * (int64_t)(kb + col);
        const double2 a0 = __ldg(src), a1 = __ldg(src + 1);
        stage[3 * lane] = a0; stage[3 * lane + 1] = a1;
        if (TREND && tn > 1) stage[3 * lane + 2] = __ldg(src + 2);
    } else if (PAD) {
        const double2 z = make_double2(0.0, 0.0);
        stage[3 * lane] = z; stage[3 * lane + 1] = z;
        if (TREND && tn > 1) stage[3 * lane + 2] = z;
    }
    __syncwarp();
#pragma unroll (UNR)
    for (int j = 0; j < nmax; ++j) {
        if (!PAD && j >= nrec) continue;
        const bool valid = !PAD || j < nrec;
        const double2 ra0 = stage[3 * (sbase + j)], ra1 = stage[3 * (sbase + j) + 1];
        const double t = ra0.x, y = ra0.y, e1 = ra1.x;
        double sE[NPT], cE[NPT], dt[NPT], rD[NPT], rv[NPT];
        double model = off;
        double tb1 = 0.0, tb2 = 0.0;
        if (TREND && tn > 0) {
            model = fma(tc[0], ra1.y, model);
            if (tn > 1) {
                const double2 ra2 = stage[3 * (sbase + j) + 2];
                tb1 = ra2.x; tb2 = ra2.y;
                model = fma(tc[1], tb1, model); model = fma(tc[2], tb2, model);
            }
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) if (u < ni) {
            kepler_sincos(orb[u], t, dt[u], sE[u], cE[u]);
            rD[u] = rcp_nr(fma(-orb[u].e, cE[u], kc.one));
            rv[u] = fma(Pc[u], cE[u], -Ps[u] * sE[u]) * rD[u];
            model = fma(f[u], rv[u], model);
        }
        const double r = y - model;
        double iv;
        if constexpr (!JIT) iv = e1;             // 1/σ² precomputed; normalisation is in const_ll
        else {
            const double var = valid ? e1 + j2 : 1.0;
            iv = valid ? rcp_nr(var) : 0.0;
            if constexpr (MARGIN) mLG += log(kc.two_pi * var);
            else { const double lt = fma(kc.mhalf, kc.log2pi + log(var), ll); ll = valid ? lt : ll; }
        }
        const double riv = r * iv;
        double g;                                 // d ll / d model
        if constexpr (MARGIN) {
            mA += iv; mS1 += riv; mC = fma(r, riv, mC);
            mR2 = fma(riv, riv, mR2); mR1 = fma(riv, iv, mR1); mQ = fma(iv, iv, mQ);
            g = kc.two * riv;
        } else {
            ll = fma(kc.mhalf * r, riv, ll);
            g = riv;
            if (GRAD) { g_off += g; if constexpr (JIT) g_jit = fma(fma(r, riv, -kc.one), iv, g_jit); }
        }
        if (GRAD && TREND && tn > 0) {             // d model / d coefficient = basis value
            gT[0] = fma(g, ra1.y, gT[0]); gT[1] = fma(g, tb1, gT[1]); gT[2] = fma(g, tb2, gT[2]);
            if constexpr (MARGIN) { vT[0] = fma(iv, ra1.y, vT[0]); vT[1] = fma(iv, tb1, vT[1]); vT[2] = fma(iv, tb2, vT[2]); }
        }
        if (GRAD) {
#pragma unroll
            for (int u = 0; u < NPT; ++u) if (u < ni) {
                // d rv / d(Pc, Ps, e|E, E)
                const double dPc = cE[u] * rD[u], dPs = -sE[u] * rD[u];
                const double dE = -fma(Pc[u], sE[u], Ps[u] * cE[u]) * rD[u] - rv[u] * orb[u].e * sE[u] * rD[u];
                const double dM = dE * rD[u];
                const double de = fma(rv[u], dPc, dM * sE[u]);
                L[u][0] = fma(g, dPc, L[u][0]); L[u][1] = fma(g, dPs, L[u][1]);
                L[u][2] = fma(g, de, L[u][2]);  L[u][3] = fma(g, dM, L[u][3]);
                L[u][4] = fma(g * dM, dt[u], L[u][4]);
                if constexpr (MARGIN) {
                    V[u][0] = fma(iv, dPc, V[u][0]); V[u][1] = fma(iv, dPs, V[u][1]);
                    V[u][2] = fma(iv, de, V[u][2]);  V[u][3] = fma(iv, dM, V[u][3]);
                    V[u][4] = fma(iv * dM, dt[u], V[u][4]);
                }
            }
        }
    }
    }
    if constexpr (MARGIN) {
        const int s0 = B.slot_margin;
        acc_add(acc, s0 + MA_A, lane, mA);   acc_add(acc, s0 + MA_S1, lane, mS1); acc_add(acc, s0 + MA_C, lane, mC);
        acc_add(acc, s0 + MA_LG, lane, mLG); acc_add(acc, s0 + MA_R2, lane, mR2); acc_add(acc, s0 + MA_R1, lane, mR1);
        acc_add(acc, s0 + MA_Q, lane, mQ);
    } else {
        acc_add(acc, 0, lane, ll);
    }
    if (GRAD) {
        if (!margin) {
            if (B.slot_jitter >= 0) acc_add(acc, B.slot_jitter, lane, g_jit * jit);
            if (B.slot_offset >= 0) acc_add(acc, B.slot_offset, lane, g_off);
        }
#pragma unroll
        for (int v = 0; v < 3; ++v) if (TREND && v < tn) {
            acc_add(acc, B.slot_trend[v], lane, gT[v]);
            if constexpr (MARGIN) acc_add(acc, B.slot_trend[v] + 1, lane, vT[v]);
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) if (u < ni) {
            const int p = pj[u];
            const double fu = f[u];
            acc_add(acc, slot_planet(p, PA_Pc), lane, fu * L[u][0]);
            acc_add(acc, slot_planet(p, PA_Ps), lane, fu * L[u][1]);
            acc_add(acc, slot_planet(p, PA_e), lane, fu * L[u][2]);
            acc_add(acc, slot_planet(p, PA_S0), lane, fu * L[u][3]);
            acc_add(acc, slot_planet(p, PA_S1), lane, fu * L[u][4]);
            if (dmu[u] != 0.0) acc_add(acc, slot_planet(p, PA_mu), lane, dmu[u] * fma(Pc[u], L[u][0], Ps[u] * L[u][1]));
            if constexpr (MARGIN) {
                const int v0 = B.slot_margin + MA_COUNT + p * MV_COUNT;
                acc_add(acc, v0 + MV_Pc, lane, fu * V[u][0]); acc_add(acc, v0 + MV_Ps, lane, fu * V[u][1]);
                acc_add(acc, v0 + MV_e, lane, fu * V[u][2]);  acc_add(acc, v0 + MV_S0, lane, fu * V[u][3]);
                acc_add(acc, v0 + MV_S1, lane, fu * V[u][4]);
                acc_add(acc, v0 + MV_mu, lane, dmu[u] * fma(Pc[u], V[u][0], Ps[u] * V[u][1]));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Prologue: per chain*planet constants (ref a3: KepOrbit / Visual ctor caches) -> shared memory.
// Two phases so that the 8 warps of a CTA share the latency: phase 1 = five independent tasks per planet
// (sincos i | sincos ω | sincos Ω | eccentricity chain | size/mass/time chain: mean motion, mas/AU, mu), phase 2 =
// the Thiele-Innes / RV products.  All branch-free (rcp/rsqrt Newton, quadrant-reduced sincos).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_nr(double x) {      // 1/sqrt(x), x > 0 normal: seed + 3 Newton steps
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = kc.half * x;
#pragma unroll
    for (int it = 0; it < 3; ++it) y = fma(y, fma(-hx * y, y, kc.half), y);
    return y;
}

// ThieleInnesOrbit: a = sqrt(u + sqrt((u+v)(u-v))) / plx (src/parameterizations.jl:14-18); the constants and the pieces
// of that formula are kept for the projection and the chain rule (in slots this basis does not otherwise use).
// Out of line: models without such planets never fetch this code.
__device__ __noinline__ double ti_semimajor(double A, double B, double F, double G, double plx, double* sc, int lane) {
    const double u = 0.5 * (A * A + B * B + F * F + G * G), v = A * G - B * F;
    const double wq = sqrt((u + v) * (u - v)), alpha = sqrt(u + wq);
    sc[PC_A * 32 + lane] = A; sc[PC_B * 32 + lane] = B; sc[PC_F * 32 + lane] = F; sc[PC_G * 32 + lane] = G;
    sc[PC_K * 32 + lane] = u; sc[PC_Kb * 32 + lane] = v; sc[PC_Pc * 32 + lane] = wq; sc[PC_Ps * 32 + lane] = alpha;
    return (isfinite(plx) && plx > 0.0) ? alpha / plx : -1.0;
}

// returns validity of what the task looked at
// skip_tp: the fused parameterisation derives tp (θ_at_epoch_to_tperi) on another warp at the same time and stores it
// itself; the size / mass / time task then leaves tp alone
// rd(k): the value of kernel input k for this lane's chain (global memory, the staged inputs, or — fused stage, before the
// inputs are staged — straight from the parameter it is defined by)
template <bool LEAN, class RD>
__device__ __forceinline__ bool prologue_task(const DevModel& m, int p, int kind, RD rd, double* sc, int lane, bool skip_tp = false) {
    const bool ti = (!LEAN && m.any_ti) && m.basis[p] == OCTO_BASIS_THIELE_INNES;
    if (kind < 3) {
        if (ti) {            // no angles: neutral values for the slots the Campbell code reads
            const int ks = kind == 0 ? PC_sini : (kind == 1 ? PC_sinw : PC_sinW);
            sc[ks * 32 + lane] = 0.0; sc[(ks + 1) * 32 + lane] = 1.0;
            return true;
        }
        const int idx = kind == 0 ? m.idx_i[p] : (kind == 1 ? m.idx_w[p] : m.idx_W[p]);
        const double x = rd(idx);
        const bool ok = isfinite(x) && fabs(x) < 1e9;
        double sn, cs;
        sincos_any(ok ? x : 0.0, sn, cs);
        const int ks = kind == 0 ? PC_sini : (kind == 1 ? PC_sinw : PC_sinW);
        sc[ks * 32 + lane] = sn; sc[(ks + 1) * 32 + lane] = cs;
        return ok;
    }
    if (kind == 3) {          // eccentricity chain
        double e = rd(m.idx_e[p]);
        const bool ok = isfinite(e) && (e >= 0.0) && (e < 1.0);
        if (!ok) e = 0.1;
        const double s2 = fma(-e, e, 1.0);
        const double inv_s = rsqrt_nr(s2);
        sc[PC_e * 32 + lane] = e;          sc[PC_ome * 32 + lane] = 1.0 - e;  sc[PC_ca1 * 32 + lane] = kA1 * rcp_nr(1.0 + e);
        sc[PC_s * 32 + lane] = s2 * inv_s; sc[PC_inv_s * 32 + lane] = inv_s;
        return ok;
    }
    // kind 4: size / mass / time chain.  The mean motion follows the reference's operation order with IEEE
    // division and square root (KepOrbit ctor + orbitsolve: n = 2π / (√(a³/M)·kyd / y2d), MA = n / y2d · (t - tp)):
    // its last bit is multiplied by |MA| (thousands of radians for short periods), so anything else would cost
    // parity digits in exactly the regime where the problem is already ill-conditioned.
    double tp = skip_tp ? 0.0 : rd(m.idx_tp[p]);
    double M = rd(m.idx_M[p]), plx = rd(m.idx_plx[p]);
    double mass = m.idx_mass[p] >= 0 ? rd(m.idx_mass[p]) : 0.0;
    double a = ti ? ti_semimajor(rd(m.idx_A[p]), rd(m.idx_B[p]), rd(m.idx_F[p]), rd(m.idx_G[p]), plx, sc, lane) : rd(m.idx_a[p]);
    const bool fin = isfinite(a) && isfinite(tp) && isfinite(M) && isfinite(plx) && isfinite(mass);
    const bool ok = fin && (a > 0.0) && (M > 0.0) && (plx > 0.0);
    if (!ok) { a = 1.0; tp = 0.0; M = 1.0; plx = 1.0; mass = 0.0; }
    const double period_days = __dmul_rn(__dsqrt_rn(__ddiv_rn(__dmul_rn(__dmul_rn(a, a), a), M)), m.c.kepler_year_days);
    const double period_yrs = __ddiv_rn(period_days, m.c.year2day);
    const double n_yr = __ddiv_rn(kTwoPi, period_yrs);
    const double inv_a = rcp_nr(a), inv_M = rcp_nr(M);
    const double Moa = M * inv_a;
    const double c2a = plx * m.c2a_per_plx;                     // rad2as*1e3 / (1000/plx * pc2au)  [mas/AU]
    sc[PC_nd * 32 + lane] = __ddiv_rn(n_yr, m.c.year2day);      // [rad/day]
    if (!skip_tp) sc[PC_tp * 32 + lane] = tp;
    sc[PC_a * 32 + lane] = a;          sc[PC_inv_a * 32 + lane] = inv_a;
    sc[PC_M * 32 + lane] = M;          sc[PC_inv_M * 32 + lane] = inv_M;
    sc[PC_plx * 32 + lane] = plx;      sc[PC_c2a * 32 + lane] = c2a; sc[PC_sc * 32 + lane] = a * c2a;
    if (!ti) sc[PC_Kb * 32 + lane] = m.kappa * Moa * rsqrt_nr(Moa);      // K s / sin i  (finished in prologue_products)
    sc[PC_mu * 32 + lane] = mass * m.c.mjup2msol * inv_M;
    return ok;
}

__device__ __forceinline__ void prologue_products(double* sc, int lane, bool ti) {
    if (ti) {      // ra = X B + sinE s G, dec = X A + sinE s F [mas]; no radial velocity for this basis
        const double s = sc[PC_s * 32 + lane];
        sc[PC_Bh * 32 + lane] = sc[PC_B * 32 + lane]; sc[PC_Gs * 32 + lane] = s * sc[PC_G * 32 + lane];
        sc[PC_Ah * 32 + lane] = sc[PC_A * 32 + lane]; sc[PC_Fs * 32 + lane] = s * sc[PC_F * 32 + lane];
        return;
    }
    const double sW = sc[PC_sinW * 32 + lane], cW = sc[PC_cosW * 32 + lane], sw = sc[PC_sinw * 32 + lane];
    const double cw = sc[PC_cosw * 32 + lane], si = sc[PC_sini * 32 + lane], ci = sc[PC_cosi * 32 + lane];
    const double s = sc[PC_s * 32 + lane], scl = sc[PC_sc * 32 + lane];
    const double A = cW * cw - sW * sw * ci, Bc = sW * cw + cW * sw * ci;
    const double F = -cW * sw - sW * cw * ci, G = -sW * sw + cW * cw * ci;
    const double Kb = sc[PC_Kb * 32 + lane] * sc[PC_inv_s * 32 + lane];        // K / sin i
    const double K = Kb * si;
    sc[PC_Kb * 32 + lane] = Kb;
    sc[PC_A * 32 + lane] = A; sc[PC_B * 32 + lane] = Bc; sc[PC_F * 32 + lane] = F; sc[PC_G * 32 + lane] = G;
    sc[PC_Bh * 32 + lane] = scl * Bc; sc[PC_Gs * 32 + lane] = scl * s * G;
    sc[PC_Ah * 32 + lane] = scl * A;  sc[PC_Fs * 32 + lane] = scl * s * F;
    sc[PC_K * 32 + lane] = K;
    sc[PC_Pc * 32 + lane] = K * cw * s * s; sc[PC_Ps * 32 + lane] = K * sw * s;
}

// ---------------------------------------------------------------------------------------------
// Epilogue: raw epoch sums R[slot] -> ll and d ll / d inputs, one lane per chain, FOUR WARPS IN PARALLEL.
// This code runs once per chain group (in the last CTA to arrive) and is instruction-fetch bound when cold
// (measured 5.3k cycles cold vs 2.5k warm as a single function); splitting it into independent parts on
// different warps overlaps the fetches and the dependency chains.  Every part adds into its own gradient
// array gp[part][n_in][32]; the parts are summed in a fixed order afterwards (bit-reproducible).
//   part 0: table terms — ll, observation-variable slots
//   part 1: astrometry chain rule (scaled Thiele-Innes -> a, plx, e, i, ω, Ω)
//   part 2: radial-velocity chain rule (Pc, Ps -> K, ω, e, a, M, i)
//   part 3: mean motion / tp / reflex factor (S0, S1, e, mu -> tp, a, M, e, mass)
// Marginalised-RV tables fold their V-sums into the planet sums first (epilogue_margin, own barrier).
// ---------------------------------------------------------------------------------------------
constexpr int EPI_PARTS = 4;

template <bool GRAD>
__device__ __noinline__ void epilogue_margin(const DevModel& m, const double* s_const, double* R, double* gp0,
                                             const double* __restrict__ in, int64_t c, int64_t ld, int lane, bool skip_empty) {
    double ll = 0.0;
#pragma unroll 1
    for (int b = 0; b < m.n_blocks; ++b) {
        const DevBlock& B = m.blocks[b];
        if (B.slot_obsprior >= 0) {
            // ln_prior = 2 log(S cbrt(P) / sqrt(1 - e²)), P = period [days] / 365.25 (prior-observable.jl:96-134)
            const int p = B.planet;
            const double* sc = s_const + p * PC_COUNT * 32;
            const double S = R[(B.slot_obsprior + OP_S) * 32 + lane];
            if (skip_empty && S == 0.0) continue;       // pointwise mode: the epoch of this CTA is not in this table
            const double a = sc[PC_a * 32 + lane], Ms = sc[PC_M * 32 + lane], e = sc[PC_e * 32 + lane];
            const double inv_s = sc[PC_inv_s * 32 + lane], s = sc[PC_s * 32 + lane];
            const double P = sqrt(a * a * a / Ms) * m.c.kepler_year_days / 365.25;
            ll += 2.0 * log(S * cbrt(P) / s);
            if (GRAD) {
                const double fac = 2.0 / S;
                R[slot_planet(p, PA_S0) * 32 + lane] += fac * R[(B.slot_obsprior + OP_Q0) * 32 + lane];
                R[slot_planet(p, PA_S1) * 32 + lane] += fac * R[(B.slot_obsprior + OP_Q1) * 32 + lane];
                R[slot_planet(p, PA_e) * 32 + lane] += fac * R[(B.slot_obsprior + OP_Qe) * 32 + lane] + 2.0 * e * inv_s * inv_s;
                gp0[m.idx_a[p] * 32 + lane] += sc[PC_inv_a * 32 + lane];                  // (2/3) d log P / da
                gp0[m.idx_M[p] * 32 + lane] -= sc[PC_inv_M * 32 + lane] * (1.0 / 3.0);
            }
            continue;
        }
        if (B.kind != OCTO_KIND_RV_STAR_MARGIN) continue;
        const int s0 = B.slot_margin;
        const double A = R[(s0 + MA_A) * 32 + lane], S1 = R[(s0 + MA_S1) * 32 + lane];
        const double C = R[(s0 + MA_C) * 32 + lane], LG = R[(s0 + MA_LG) * 32 + lane];
        if (skip_empty && A == 0.0) continue;          // pointwise mode: the epoch of this CTA is not in this table
        const double rbar = S1 / A;
        ll += -LG - C + S1 * rbar - log(A);            // rv-absolute-margin.jl:171-181 with B = -2 S1
        if (GRAD) {
            const double R2 = R[(s0 + MA_R2) * 32 + lane], R1 = R[(s0 + MA_R1) * 32 + lane], Q = R[(s0 + MA_Q) * 32 + lane];
            const double jit = in[c + (int64_t)B.idx_jitter * ld];
            gp0[B.idx_jitter * 32 + lane] += 2.0 * jit * (-A + R2 - 2.0 * rbar * R1 + rbar * rbar * Q + Q / A);
            for (int v = 0; v < B.n_trend; ++v)
                gp0[B.idx_trend[v] * 32 + lane] += R[B.slot_trend[v] * 32 + lane] - 2.0 * rbar * R[(B.slot_trend[v] + 1) * 32 + lane];
#pragma unroll 1
            for (int p = 0; p < m.n_planets; ++p) {
                const int v0 = s0 + MA_COUNT + p * MV_COUNT;
                const double k2 = -2.0 * rbar;
                R[slot_planet(p, PA_Pc) * 32 + lane] += k2 * R[(v0 + MV_Pc) * 32 + lane];
                R[slot_planet(p, PA_Ps) * 32 + lane] += k2 * R[(v0 + MV_Ps) * 32 + lane];
                R[slot_planet(p, PA_e) * 32 + lane] += k2 * R[(v0 + MV_e) * 32 + lane];
                R[slot_planet(p, PA_S0) * 32 + lane] += k2 * R[(v0 + MV_S0) * 32 + lane];
                R[slot_planet(p, PA_S1) * 32 + lane] += k2 * R[(v0 + MV_S1) * 32 + lane];
                R[slot_planet(p, PA_mu) * 32 + lane] += k2 * R[(v0 + MV_mu) * 32 + lane];
            }
        }
    }
    R[0 * 32 + lane] += ll;
}

// Thiele-Innes planet: Bh = B, Gs = s G, Ah = A, Fs = s F (out of line, see ti_semimajor)
__device__ __noinline__ void epilogue_ti_astrom(const DevModel& m, int p, const double* sc, const double* R, double* gp, int lane) {
    const double e = sc[PC_e * 32 + lane], s = sc[PC_s * 32 + lane], inv_s = sc[PC_inv_s * 32 + lane];
    const double gGs = R[slot_planet(p, PA_Gs) * 32 + lane], gFs = R[slot_planet(p, PA_Fs) * 32 + lane];
    gp[m.idx_B[p] * 32 + lane] += R[slot_planet(p, PA_Bh) * 32 + lane]; gp[m.idx_G[p] * 32 + lane] += gGs * s;
    gp[m.idx_A[p] * 32 + lane] += R[slot_planet(p, PA_Ah) * 32 + lane]; gp[m.idx_F[p] * 32 + lane] += gFs * s;
    gp[m.idx_e[p] * 32 + lane] += -(e * inv_s) * (gGs * sc[PC_G * 32 + lane] + gFs * sc[PC_F * 32 + lane]);
}

// hands d ll / da of the Thiele-Innes planets (virtual column n_in + p, summed over the gradient parts) on to A, B, F, G
// and plx:  a = alpha / plx, alpha² = u + w, w² = (u+v)(u-v)  =>  d alpha/dX = (du/dX + (u du/dX - v dv/dX) / w) / (2 alpha).
// One warp, planets in order: they may share the plx column.
__device__ __noinline__ void epilogue_ti_distribute(const DevModel& m, const double* s_const, double* s_gp, int ng, int lane) {
#pragma unroll 1
    for (int p = 0; p < m.n_planets; ++p) {
        if (m.basis[p] != OCTO_BASIS_THIELE_INNES) continue;
        const double* sc = s_const + p * PC_COUNT * 32;
        const int va = (m.n_in + p) * 32 + lane;
        const double ga = ((s_gp[va] + s_gp[ng + va]) + s_gp[2 * ng + va]) + s_gp[3 * ng + va];
        const double A = sc[PC_A * 32 + lane], B = sc[PC_B * 32 + lane], F = sc[PC_F * 32 + lane], G = sc[PC_G * 32 + lane];
        const double u = sc[PC_K * 32 + lane], v = sc[PC_Kb * 32 + lane], wq = sc[PC_Pc * 32 + lane], alpha = sc[PC_Ps * 32 + lane];
        const double a = sc[PC_a * 32 + lane], plx = sc[PC_plx * 32 + lane];
        const double k = ga / (2.0 * alpha * plx), iw = 1.0 / wq;
        s_gp[m.idx_A[p] * 32 + lane] += k * (A + (u * A - v * G) * iw);
        s_gp[m.idx_B[p] * 32 + lane] += k * (B + (u * B + v * F) * iw);
        s_gp[m.idx_F[p] * 32 + lane] += k * (F + (u * F + v * B) * iw);
        s_gp[m.idx_G[p] * 32 + lane] += k * (G + (u * G - v * A) * iw);
        s_gp[m.idx_plx[p] * 32 + lane] -= ga * a / plx;
    }
}

template <bool LEAN>
__device__ __forceinline__ void epilogue_part(int part, const DevModel& m, const double* s_const, const double* R, double* gp,
                                           int lane) {
    if (part == 0) {
#pragma unroll 1
        for (int b = 0; b < m.n_blocks; ++b) {
            const DevBlock& B = m.blocks[b];
            if (B.kind == OCTO_KIND_RV_STAR_MARGIN) continue;
            if (B.slot_jitter >= 0) gp[B.idx_jitter * 32 + lane] += R[B.slot_jitter * 32 + lane];
            if (B.slot_platescale >= 0) gp[B.idx_platescale * 32 + lane] += R[B.slot_platescale * 32 + lane];
            if (B.slot_northangle >= 0) gp[B.idx_northangle * 32 + lane] += R[B.slot_northangle * 32 + lane];
            if (B.slot_offset >= 0) gp[B.idx_offset * 32 + lane] += R[B.slot_offset * 32 + lane];
            for (int v = 0; v < B.n_trend; ++v) gp[B.idx_trend[v] * 32 + lane] += R[B.slot_trend[v] * 32 + lane];
        }
#pragma unroll 1
        for (int h = 0; !LEAN && h < m.n_hg; ++h) {
            gp[m.hg[h].idx_pmra * 32 + lane] += R[m.hg[h].slot_pmra * 32 + lane];
            gp[m.hg[h].idx_pmdec * 32 + lane] += R[m.hg[h].slot_pmdec * 32 + lane];
        }
        return;
    }
#pragma unroll 1
    for (int p = 0; p < m.n_planets; ++p) {
        const double* sc = s_const + p * PC_COUNT * 32;
        auto C = [&](int k) { return sc[k * 32 + lane]; };
        auto Rp = [&](int a) { return R[slot_planet(p, a) * 32 + lane]; };
        const double e = C(PC_e), s = C(PC_s), inv_s = C(PC_inv_s), inv_a = C(PC_inv_a), inv_M = C(PC_inv_M);
        if ((!LEAN && m.any_ti) && part < 3 && m.basis[p] == OCTO_BASIS_THIELE_INNES) {
            if (part == 1) epilogue_ti_astrom(m, p, sc, R, gp, lane);
            continue;
        }
        if (part == 1) {
            // astrometry: Bh = sc*B, Gs = sc*s*G, Ah = sc*A, Fs = sc*s*F with sc = a*c2a
            const double sW = C(PC_sinW), cW = C(PC_cosW), sw = C(PC_sinw), cw = C(PC_cosw), si = C(PC_sini);
            const double A = C(PC_A), Bc = C(PC_B), F = C(PC_F), G = C(PC_G), scl = C(PC_sc);
            const double gBh = Rp(PA_Bh), gGs = Rp(PA_Gs), gAh = Rp(PA_Ah), gFs = Rp(PA_Fs);
            const double g_sc = gBh * Bc + gGs * s * G + gAh * A + gFs * s * F;
            const double gB = gBh * scl, gG = gGs * scl * s, gA = gAh * scl, gF = gFs * scl * s;
            gp[m.idx_a[p] * 32 + lane] += g_sc * C(PC_c2a);
            gp[m.idx_plx[p] * 32 + lane] += g_sc * C(PC_a) * m.c2a_per_plx;      // d(a*c2a)/d plx
            gp[m.idx_e[p] * 32 + lane] += -(e * inv_s) * scl * (gGs * G + gFs * F);
            gp[m.idx_w[p] * 32 + lane] += gA * F + gB * G - gF * A - gG * Bc;
            gp[m.idx_W[p] * 32 + lane] += -gA * Bc + gB * A - gF * G + gG * F;
            gp[m.idx_i[p] * 32 + lane] += si * (gA * sW * sw - gB * cW * sw + gF * sW * cw - gG * cW * cw);
        } else if (part == 2) {
            // radial velocity: Pc = K cosω s², Ps = K sinω s, K = kappa sqrt(M/a) sin i / s
            const double sw = C(PC_sinw), cw = C(PC_cosw), K = C(PC_K);
            const double gPc = Rp(PA_Pc), gPs = Rp(PA_Ps);
            const double s2 = s * s;
            const double gK = gPc * cw * s2 + gPs * sw * s;
            gp[m.idx_w[p] * 32 + lane] += -gPc * K * sw * s2 + gPs * K * cw * s;
            gp[m.idx_e[p] * 32 + lane] += gPc * K * cw * (-2.0 * e) + gPs * K * sw * (-e * inv_s) + gK * K * e * inv_s * inv_s;
            gp[m.idx_a[p] * 32 + lane] += gK * (-0.5 * K * inv_a);
            gp[m.idx_M[p] * 32 + lane] += gK * (0.5 * K * inv_M);
            gp[m.idx_i[p] * 32 + lane] += gK * C(PC_Kb) * C(PC_cosi);
        } else {
            // mean motion M_k = nd (t_k - tp); direct e terms; reflex factor mu = mass * mjup2msol / M
            const double nd = C(PC_nd), S0 = Rp(PA_S0), S1 = Rp(PA_S1), gmu = Rp(PA_mu);
            gp[m.idx_tp[p] * 32 + lane] += -nd * S0;
            gp[m.idx_a[p] * 32 + lane] += S1 * (-1.5 * nd * inv_a);
            gp[m.idx_M[p] * 32 + lane] += S1 * (0.5 * nd * inv_M) - gmu * C(PC_mu) * inv_M;
            gp[m.idx_e[p] * 32 + lane] += Rp(PA_e);
            if (m.idx_mass[p] >= 0) gp[m.idx_mass[p] * 32 + lane] += gmu * m.c.mjup2msol * inv_M;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// HGCAInstantaneousObs (kind 5; src/likelihoods/hgca.jl:155-417, Visual{KepOrbit}, absolute_orbits = false).
// A handful of rows whose likelihood is a non-linear function of eight sums (star position and proper motion at the
// Hipparcos / Gaia epochs, RA and Dec), so it is evaluated by the last CTA of a chain group once the epoch sums are
// complete: pass 1 (rows over warps) builds the eight sums, every lane then knows ll and its derivative w.r.t. each
// sum, pass 2 revisits the rows with those seeds and adds the adjoints into the per-planet sums R that the ordinary
// epilogue turns into gradients.  scratch = the per-warp accumulator area (free at this point).
//   position  = X B1 + sinE G1        (B1, G1) = (Bh, Gs) for RA rows, (Ah, Fs) for Dec rows      [mas]
//   velocity  = (-sinE B1 + cosE G1) nd year2day / (1 - e cosE)                                   [mas/yr]
//   star      = -mu x planet, summed over planets; the reference divides by planets x rows
// ---------------------------------------------------------------------------------------------
template <bool GRAD>
__device__ __noinline__ void hgca_tail(const DevModel& m, const DevHg& H, const double* s_const, double* R, double* scratch,
                                       const double* s_in, int n_acc, int w, int W, int lane) {
    const double* __restrict__ rows = m.tab + H.row_off;
    double* mine = scratch + w * n_acc * 32;
    double S[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) S[q] = 0.0;
#pragma unroll 1
    for (int r = w; r < H.n_rows; r += W) {
        const double t = rows[2 * r];
        const int code = (int)rows[2 * r + 1];
        double pos = 0.0, vel = 0.0;
#pragma unroll 1
        for (int p = 0; p < m.n_planets; ++p) {
            const double* sc = s_const + p * PC_COUNT * 32;
            const Orb o = load_orb(sc, lane);
            double dt, sE, cE;
            kepler_sincos(o, t, dt, sE, cE);
            const double B1 = sc[((code & 1) ? PC_Ah : PC_Bh) * 32 + lane], G1 = sc[((code & 1) ? PC_Fs : PC_Gs) * 32 + lane];
            const double mu = sc[PC_mu * 32 + lane];
            const double rD = rcp_nr(fma(-o.e, cE, kc.one));
            const double u = fma(cE, G1, -sE * B1);
            pos = fma(-mu, fma(cE - o.e, B1, sE * G1), pos);
            vel = fma(-mu, u * (o.nd * m.c.year2day * rD), vel);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) if (code == q) { S[2 * q] += pos; S[2 * q + 1] += vel; }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) mine[q * 32 + lane] = S[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        double v = scratch[q * 32 + lane];
        for (int ww = 1; ww < W; ++ww) v += scratch[(ww * n_acc + q) * 32 + lane];
        S[q] = v;
    }
    const double pmra = s_in[H.idx_pmra * 32 + lane], pmdec = s_in[H.idx_pmdec * 32 + lane];
    // model proper motions: Hipparcos, Hipparcos-Gaia (position difference), Gaia
    const double mod[3][2] = {
        {fma(S[1], H.inv_N[0], pmra), fma(S[3], H.inv_N[1], pmdec)},
        {fma(S[4] * H.inv_N[2] - S[0] * H.inv_N[0], H.k_ra, pmra), fma(S[6] * H.inv_N[3] - S[2] * H.inv_N[1], H.k_dec, pmdec)},
        {fma(S[5], H.inv_N[2], pmra), fma(S[7], H.inv_N[3], pmdec)}};
    double ll = 0.0, q1[3], q2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double r1 = mod[d][0] - H.cat[d][0], r2 = mod[d][1] - H.cat[d][1];
        q1[d] = fma(H.w[d][0], r1, H.w[d][1] * r2); q2[d] = fma(H.w[d][1], r1, H.w[d][2] * r2);
        ll = fma(kc.mhalf, fma(r1, q1[d], r2 * q2[d]), ll);
    }
    __syncthreads();                                   // everyone has read the partial sums
    if (w == 0) {
        R[0 * 32 + lane] += ll;
        if (GRAD) {
            R[H.slot_pmra * 32 + lane] = -(q1[0] + q1[1] + q1[2]);
            R[H.slot_pmdec * 32 + lane] = -(q2[0] + q2[1] + q2[2]);
        }
    }
    if (!GRAD) return;
    // seeds d ll / d S[q]
    double g[8];
    g[0] = q1[1] * H.k_ra * H.inv_N[0];  g[1] = -q1[0] * H.inv_N[0];
    g[2] = q2[1] * H.k_dec * H.inv_N[1]; g[3] = -q2[0] * H.inv_N[1];
    g[4] = -q1[1] * H.k_ra * H.inv_N[2]; g[5] = -q1[2] * H.inv_N[2];
    g[6] = -q2[1] * H.k_dec * H.inv_N[3]; g[7] = -q2[2] * H.inv_N[3];
    const int np = m.n_planets * PA_COUNT;
#pragma unroll 1
    for (int a = 0; a < np; ++a) mine[a * 32 + lane] = 0.0;
#pragma unroll 1
    for (int r = w; r < H.n_rows; r += W) {
        const double t = rows[2 * r];
        const int code = (int)rows[2 * r + 1];
        double gp = 0.0, gv = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (code == q) { gp = g[2 * q]; gv = g[2 * q + 1]; }
#pragma unroll 1
        for (int p = 0; p < m.n_planets; ++p) {
            const double* sc = s_const + p * PC_COUNT * 32;
            const Orb o = load_orb(sc, lane);
            double dt, sE, cE;
            kepler_sincos(o, t, dt, sE, cE);
            const double B1 = sc[((code & 1) ? PC_Ah : PC_Bh) * 32 + lane], G1 = sc[((code & 1) ? PC_Fs : PC_Gs) * 32 + lane];
            const double mu = sc[PC_mu * 32 + lane];
            const double rD = rcp_nr(fma(-o.e, cE, kc.one));
            const double kv = o.nd * m.c.year2day * rD;
            const double X = cE - o.e, u = fma(cE, G1, -sE * B1);
            const double pos = fma(X, B1, sE * G1), vel = u * kv;
            const double cp = -mu * gp, cv = -mu * gv;
            double* acc = mine + p * PA_COUNT * 32;
            // d/dB1, d/dG1
            acc[((code & 1) ? PA_Ah : PA_Bh) * 32 + lane] += fma(cp, X, -cv * sE * kv);
            acc[((code & 1) ? PA_Fs : PA_Gs) * 32 + lane] += fma(cp, sE, cv * cE * kv);
            // through E: d pos/dE = u, d vel/dE = (-(cosE B1 + sinE G1) - u e sinE rD) kv;  dE/dMA = rD, dE/de = sinE rD
            const double dvel = (-fma(cE, B1, sE * G1) - u * o.e * sE * rD) * kv;
            const double gM = fma(cp, u, cv * dvel) * rD;
            acc[PA_S0 * 32 + lane] += gM;
            acc[PA_S1 * 32 + lane] += fma(gM, dt, cv * u * (m.c.year2day * rD));      // + explicit mean-motion factor of the velocity
            acc[PA_e * 32 + lane] += fma(gM, sE, fma(cv * vel, cE * rD, -cp * B1)); // + explicit e in 1/(1 - e cosE) and in X
            acc[PA_mu * 32 + lane] -= fma(gp, pos, gv * vel);
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int idx = threadIdx.x; idx < np * 32; idx += W * 32) {
        double v = scratch[idx];
        for (int ww = 1; ww < W; ++ww) v += scratch[ww * n_acc * 32 + idx];
        R[32 + idx] += v;                              // planet slots start at slot 1 (slot_planet)
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Fused standard parameterisation: with a DevParam the kernel's input is θ_t.  Every CTA runs the
// forward stage for its 32 chains (lane = chain; WARPS stride over parameters / inputs / tperi items, so a prior
// family or input definition is warp-uniform), the last CTA of a chain group runs the reverse stage after the
// epilogue.  Same device functions and the same summation orders as the stand-alone K0 kernels (octo_param.cu):
// both paths agree to rounding (tests: 1e-12).  All arrays are [index][32 lanes].
// ---------------------------------------------------------------------------------------------
#ifdef OCTO_TIMING
#define PTICK(i) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ptk[i] = (long long)t_; } } while (0)
__device__ long long g_ptk[2][8];
__device__ long long g_tm_last[13];
__device__ long long g_wtk[16][12];
__device__ int g_quiet;      // set by the resident kernel: phase timings are printed once, at the end of the run
#else
#define PTICK(i) do {} while (0)
#endif
// trig: per θ_at_epoch_to_tperi definition 16 slots — sin, cos of (θ, i, ω, Ω), the mean anomaly, then the reciprocals and
// roots of its forward pass (tperi_mid `keep`); cir: [2][n_in] what circ_forward saves for the reverse pass
#ifndef OCTO_EARLY_TRIG
#define OCTO_EARLY_TRIG 1        // lean kernels: sines / cosines together with the inputs, prologue chains in the same phase (param_forward)
#endif
constexpr int TRIG_SLOTS = 16;
// DevParam staged in shared memory (the last kParamWords doubles of the evaluation's area): the parameterisation stages
// chase order -> prior / definition -> operand indices several times per phase, each a dependent L2 round trip (~300 cycles)
// when read from global memory — ~1.5 us of a 12 us leapfrog
constexpr int kParamWords = (int)((sizeof(DevParam) + 7) / 8);
struct ParamSmem { double *th, *dxdy, *gth, *L, *aux, *cir, *cs, *trig, *part, *lp, *extra, *beta; int* flags; };
__device__ __forceinline__ ParamSmem param_smem(double* base, int n_in, int D, int T) {
    ParamSmem S;
    S.th = base; S.dxdy = S.th + D * 32; S.gth = S.dxdy + D * 32; S.L = S.gth + D * 32; S.aux = S.L + D * 32;
    S.cir = S.aux + n_in * 32;
    S.cs = S.cir + 2 * n_in * 32;                              // [2][n_in]: sine, cosine of the inputs that are angles (DevParam::in_trig)
    S.trig = S.cs + 2 * n_in * 32; S.part = S.trig + T * TRIG_SLOTS * 32; S.lp = S.part + T * 8 * 32; S.extra = S.lp + 32;
    S.beta = S.extra + 32;
    S.flags = reinterpret_cast<int*>(S.beta + 32);
    return S;
}
__host__ __device__ inline size_t param_smem_doubles(int n_in, int D, int T) { return (size_t)(4 * D + 5 * n_in + (TRIG_SLOTS + 8) * T + 4) * 32; }

template <bool LEAN>
__device__ __forceinline__ void param_forward(const DevParam& P, const DevModel& m, const double* __restrict__ theta_t, int64_t c,
                                           int64_t ld, double* s_in, double* s_const, const ParamSmem& S, int* s_ok, int w, int W,
                                           int lane) {
    using namespace octo_param_dev;
    const int D = P.D, n_in = P.n_in, T = P.n_tperi;
#ifdef OCTO_TIMING
    long long* ptk = g_ptk[0];
#endif
    PTICK(0);
    // invlink + logpdf_with_trans
#pragma unroll 1
    for (int jj = w; jj < D; jj += W) {
        const int j = P.order_prior[jj];                              // expensive families first: they share the first round
        const double y = theta_t[c + (int64_t)j * ld];
        const bool fin = isfinite(y);
        const PriorEval r = prior_eval(P.priors[j].family, P.priors[j].p[0], P.pc[j], fin ? y : 0.0);
        S.th[j * 32 + lane] = r.x; S.dxdy[j * 32 + lane] = r.dxdy; S.gth[j * 32 + lane] = r.dLdx;
        S.L[j * 32 + lane] = fin ? r.L : CUDART_NAN;
        if (!fin) S.flags[lane] = 16;                                // non-finite θ_t entry (flags start at 0)
    }
    __syncthreads();
    PTICK(1);
    if constexpr (LEAN && OCTO_EARLY_TRIG) {
        // LEAN kernels (Campbell planets only).  ONE phase for: the derived inputs (arr2nt), the sines and cosines the
        // evaluation needs of them (DevParam::in_trig) — for a UniformCircular pair with the full circle as its domain
        // simply (y, x) / r: no atan2, no sincos; the angle itself is used nowhere in this kernel — and the two heavy
        // prologue chains of every planet, which read their inputs straight from the parameters that define them (the
        // host only fuses a lean model whose e, a, M, plx, mass are parameters or constants).  Round 1 - 2a had three
        // barrier-separated phases here (inputs; trigonometry of tperi + prologue tasks): 1.3k + 0.7k cycles -> ~1.0k.
        const int np = m.n_planets, n_item = 2 * np + n_in;
        auto rd_param = [&](int k) { const OctoInputDef& d = P.defs[k]; return d.op == OCTO_IN_PARAM ? S.th[d.a[0] * 32 + lane] : d.value; };
#pragma unroll 1
        for (int it = w; it < n_item; it += W) {
            if (it < 2 * np) {
                const int p = it >> 1, kind = 4 - (it & 1);
                bool derived_tp = false;
                for (int t = 0; t < T; ++t) derived_tp = derived_tp || P.tperi_k[t] == m.idx_tp[p];
                if (!prologue_task<LEAN>(m, p, kind, rd_param, s_const + p * PC_COUNT * 32, lane, derived_tp)) s_ok[lane] = 0;
                continue;
            }
            const int k = P.order_input[it - 2 * np];
            const OctoInputDef& d = P.defs[k];
            const bool trig = P.in_trig[k] != 0;
            double v = 0.0, ext = 0.0, sn = 0.0, cs = 1.0;
            bool ok = true;
            if (d.op == OCTO_IN_CIRC) {
                const double x = S.th[d.a[0] * 32 + lane], y = S.th[d.a[1] * 32 + lane];
                if (trig && d.value == octo_param_dev::kTwoPi) {
                    const double r2 = fma(x, x, y * y), l2 = p_log(r2), lr = 0.5 * l2;
                    ext = fma(-lr * lr, 50.0, -lr - (-2.302585092994045684 /* log 0.1 */) - kHalfLog2Pi);      // circ_forward's UnitLengthPrior term
                    const double ir2 = 1.0 / r2;
                    S.cir[k * 32 + lane] = ir2; S.cir[(n_in + k) * 32 + lane] = -1.0 - 0.5 * l2 * 100.0;
                    const bool pos = r2 > 0.0 && isfinite(r2);
                    const double ir = pos ? rsqrt_nr(r2) : 0.0;
                    sn = y * ir; cs = pos ? x * ir : 1.0;                  // atan2(0, 0) = 0
                    v = r2 - r2;                                           // stands for the angle in the validity checks: 0, or NaN
                } else {
                    circ_forward(x, y, d.value, v, ext, &S.cir[k * 32 + lane], &S.cir[(n_in + k) * 32 + lane]);
                    if (trig) sincos_any(isfinite(v) ? v : 0.0, sn, cs);
                }
            } else if (d.op == OCTO_IN_PARAM || d.op == OCTO_IN_CONST) {
                v = d.op == OCTO_IN_PARAM ? S.th[d.a[0] * 32 + lane] : d.value;
                if (trig) { ok = isfinite(v) && fabs(v) < 1e9; sincos_any(ok ? v : 0.0, sn, cs); }
            }
            s_in[k * 32 + lane] = v; S.aux[k * 32 + lane] = ext;
            if (trig) {
                S.cs[k * 32 + lane] = sn; S.cs[(n_in + k) * 32 + lane] = cs;
#pragma unroll 1
                for (int p = 0; p < np; ++p) {                           // the planets that use this angle
                    double* sc = s_const + p * PC_COUNT * 32;
                    if (k == m.idx_i[p]) { sc[PC_sini * 32 + lane] = sn; sc[PC_cosi * 32 + lane] = cs; if (!ok) s_ok[lane] = 0; }
                    if (k == m.idx_w[p]) { sc[PC_sinw * 32 + lane] = sn; sc[PC_cosw * 32 + lane] = cs; if (!ok) s_ok[lane] = 0; }
                    if (k == m.idx_W[p]) { sc[PC_sinW * 32 + lane] = sn; sc[PC_cosW * 32 + lane] = cs; if (!ok) s_ok[lane] = 0; }
                }
            }
        }
        __syncthreads();
        PTICK(2);
        PTICK(3);
    } else {
    // derived inputs that depend on parameters only (arr2nt)
#pragma unroll 1
        for (int kk = w; kk < n_in; kk += W) {
            const int k = P.order_input[kk];
            const OctoInputDef& d = P.defs[k];
            double v = 0.0, ext = 0.0;
            if (d.op == OCTO_IN_PARAM) v = S.th[d.a[0] * 32 + lane];
            else if (d.op == OCTO_IN_CONST) v = d.value;
            else if (d.op == OCTO_IN_CIRC) circ_forward(S.th[d.a[0] * 32 + lane], S.th[d.a[1] * 32 + lane], d.value, v, ext,
                                                        &S.cir[k * 32 + lane], &S.cir[(n_in + k) * 32 + lane]);
            s_in[k * 32 + lane] = v; S.aux[k * 32 + lane] = ext;
        }
        __syncthreads();
        PTICK(2);
        // One phase for two independent things: the trigonometry of θ_at_epoch_to_tperi — (definition, angle) items for
        // (θ, i, ω, Ω) — and phase 1 of K1's prologue (five tasks per planet; none of them needs tp).  Items over warps.
        {
            const int n_trig = 4 * T, n_item = n_trig + 5 * m.n_planets;
#pragma unroll 1
            for (int it = w; it < n_item; it += W) {
                if (it < n_trig) {
                    const int t = it >> 2, q = it & 3;
                    const OctoInputDef& d = P.defs[P.tperi_k[t]];
                    if (d.op == OCTO_IN_TPERI_TI && q > 0) continue;          // Thiele-Innes: only θ is an angle
                    double sn, cs;
                    p_sincos(s_in[d.a[q == 0 ? 0 : 3 + q] * 32 + lane], &sn, &cs);
                    S.trig[(t * TRIG_SLOTS + 2 * q) * 32 + lane] = sn; S.trig[(t * TRIG_SLOTS + 2 * q + 1) * 32 + lane] = cs;
                } else {
                    const int task = it - n_trig, p = task / 5, kind = task % 5;
                    bool derived_tp = false;                                 // this planet's tp is one of the tperi definitions
                    for (int t = 0; t < T; ++t) derived_tp = derived_tp || P.tperi_k[t] == m.idx_tp[p];
                    if (!prologue_task<LEAN>(m, p, kind, [&](int k) { return s_in[k * 32 + lane]; }, s_const + p * PC_COUNT * 32, lane, derived_tp)) s_ok[lane] = 0;
                }
            }
        }
        __syncthreads();
        PTICK(3);
    }
    // Second phase, again independent things on different warps: one warp per θ_at_epoch_to_tperi definition (it hands
    // tp to the planets that use it), the prologue's products (they do not involve tp) on the next warps, the ordered
    // prior sums on the last warp.
#pragma unroll 1
    for (int it = w; it < T + m.n_planets; it += W) {
        if (it < T) {
            const int t = it, k = P.tperi_k[t];
            const OctoInputDef& d = P.defs[k];
            const bool ti = d.op == OCTO_IN_TPERI_TI;
            double arg[8], trig[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) arg[q] = (q < 7 || ti) ? s_in[d.a[q] * 32 + lane] : 0.0;
            if constexpr (LEAN && OCTO_EARLY_TRIG) {                  // sines / cosines of (θ, i, ω, Ω) came with the inputs; kept for the reverse pass
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ak = d.a[q == 0 ? 0 : 3 + q];
                    trig[2 * q] = S.cs[ak * 32 + lane]; trig[2 * q + 1] = S.cs[(n_in + ak) * 32 + lane];
                    S.trig[(t * TRIG_SLOTS + 2 * q) * 32 + lane] = trig[2 * q]; S.trig[(t * TRIG_SLOTS + 2 * q + 1) * 32 + lane] = trig[2 * q + 1];
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) trig[q] = S.trig[(t * TRIG_SLOTS + q) * 32 + lane];
            }
            double MA;
            const double tp = tperi_value(m.c, d.value, arg, trig, &MA, ti, S.trig + (t * TRIG_SLOTS + 9) * 32 + lane, 32);
            s_in[k * 32 + lane] = tp;
            S.trig[(t * TRIG_SLOTS + 8) * 32 + lane] = MA;
#pragma unroll 1
            for (int p = 0; p < m.n_planets; ++p) if (m.idx_tp[p] == k) {
                const bool fin = isfinite(tp);
                s_const[(p * PC_COUNT + PC_tp) * 32 + lane] = fin ? tp : 0.0;
                if (!fin) s_ok[lane] = 0;
            }
        } else {
            const int p = it - T;
            prologue_products(s_const + p * PC_COUNT * 32, lane, (!LEAN && m.any_ti) && m.basis[p] == OCTO_BASIS_THIELE_INNES);
        }
    }
    PTICK(4);
    // ordered sums, "healing" of a non-finite prior term (variables.jl:1229-1236), validity — on the last warp, next to
    // the warps above (the derived tp inputs are checked by the warp that computes them); the caller's barrier follows
    if (w == W - 1) {
        double lp, extra;
        const int fl = prior_sums(S.L + lane, S.aux + lane, s_in + lane, D, n_in, 32, !(S.flags[lane] & 16), lp, extra, &P);
        S.lp[lane] = lp; S.extra[lane] = extra; S.flags[lane] = fl;
        if (!(fl & 4)) s_ok[lane] = 0;                               // K1 then returns -Inf / zero gradient
    }
    PTICK(5);
#ifdef OCTO_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && !g_quiet)
        printf("  forward ns: priors +%lld inputs +%lld trig +%lld tperi +%lld sums +%lld\n", ptk[1] - ptk[0], ptk[2] - ptk[1], ptk[3] - ptk[2], ptk[4] - ptk[3], ptk[5] - ptk[4]);
#endif
}

// reverse stage: S.aux holds d ll / d inputs of the 32 chains; S.flags bit 3 = chain is ok (valid and ll finite)
// the 7 (8) partial derivatives of θ_at_epoch_to_tperi definition t w.r.t. its arguments (hand-derived reverse pass)
__device__ __forceinline__ void param_tperi_partials(const DevParam& P, const DevModel& m, const double* s_in, const ParamSmem& S,
                                                  int t, int lane) {
    using namespace octo_param_dev;
    const OctoInputDef& d = P.defs[P.tperi_k[t]];
    const bool ti = d.op == OCTO_IN_TPERI_TI;
    double arg[8], trig[8], part[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) arg[q] = (q < 7 || ti) ? s_in[d.a[q] * 32 + lane] : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) trig[q] = S.trig[(t * TRIG_SLOTS + q) * 32 + lane];
    part[7] = 0.0;
    tperi_reverse(m.c, arg, trig, S.trig[(t * TRIG_SLOTS + 8) * 32 + lane], part, ti, S.trig + (t * TRIG_SLOTS + 9) * 32 + lane, 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) S.part[(t * 8 + q) * 32 + lane] = part[q];
}

__device__ __forceinline__ void param_backward(const DevParam& P, const DevModel& m, const double* s_in, const ParamSmem& S,
                                            double* __restrict__ g_t, int64_t chain0, bool active, int64_t ldg, int w,
                                            int W, int lane, const HmcLeap& leap, bool partials_done) {
    using namespace octo_param_dev;
    const int D = P.D, T = P.n_tperi;
#ifdef OCTO_TIMING
    long long* ptk = g_ptk[1];
#endif
    PTICK(0);
    if (!partials_done) {              // one warp per definition
#pragma unroll 1
        for (int t = w; t < T; t += W) param_tperi_partials(P, m, s_in, S, t, lane);
        __syncthreads();
    }
    PTICK(1);
    if (w == 0) {                                                    // fold the tperi partials into ∂ll/∂inputs, last definition first
#pragma unroll 1
        for (int t = T - 1; t >= 0; --t) {
            const int k = P.tperi_k[t];
            const OctoInputDef& d = P.defs[k];
            const double gk = S.aux[k * 32 + lane];
            const int na = d.op == OCTO_IN_TPERI_TI ? 8 : 7;
#pragma unroll
            for (int q = 0; q < 8; ++q) if (q < na) S.aux[d.a[q] * 32 + lane] += gk * S.part[(t * 8 + q) * 32 + lane];
        }
    }
    PTICK(2);
    if (T > 0) __syncthreads();
    PTICK(3);
    {                                                                // every parameter gathers its inputs: one warp per parameter
        const int fl = S.flags[lane];
        const bool ok = fl & 8, healed = fl & 2;
#pragma unroll 1
        for (int jj = w; jj < D; jj += W) {
            const int j = P.order_gather[jj];
            const double g = param_gather(P, j, healed ? 0.0 : S.gth[j * 32 + lane], S.th + lane, S.aux + lane, 32, S.cir + lane);
            const double gv = ok ? g * S.dxdy[j * 32 + lane] : 0.0;
            if (active) {
                const int64_t at = chain0 + lane + (int64_t)j * ldg;
                g_t[at] = gv;
                if (leap.p) {                                        // the leapfrog update of this coordinate (octo_hmc.cu)
                    const double pj = fma(leap.kick, gv, leap.p[at]);
                    leap.p[at] = pj;
                    if (leap.drift) leap.q[at] = fma(leap.eps * pj, leap.inv_mass[j], leap.q[at]);
                }
            }
        }
    }
    PTICK(4);
#ifdef OCTO_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0 && !g_quiet)
        printf("  backward ns: tperi +%lld accumulate +%lld barrier +%lld stores +%lld\n", ptk[1] - ptk[0], ptk[2] - ptk[1], ptk[3] - ptk[2], ptk[4] - ptk[3]);
#endif
}

// INLINE OR OUT OF LINE.  A call to an out-of-line device function costs its caller the spill of every live register
// around the call plus a break in the instruction stream (the callee's first lines are fetched only once the call
// issues).  Measured on B200 with everything else equal (profiles/tools/ab.sh): with all stages of the
// evaluation inline the resident explorer went 17.2 -> 13.6 us per leapfrog and C2 12.3 -> 10.2 us per step.  So: the
// lean loops are inline everywhere (throughput instantiation, 4096 x 20000: astrometry / RV+jitter 1.173e11 / 8.46e10 evals/s
// with the astrometry loop out of line, 1.204e11 / 8.20e10 with the RV loops out of line, 1.179e11 / 8.61e10 with both
// inline: within 4 %, -DOCTO_THR_ASTROM_INLINE / -DOCTO_THR_RV_INLINE); the loops of the less common tables are always
// out of line (they would only bloat every kernel).
template <bool GRAD, int NPT, int MODE, int ILP>
__device__ __noinline__ void seg_astrom_ool(const DevModel& m, const DevBlock& B, int k0, int k1, const double* s_const,
                                            double* acc, double2* stage, const double* __restrict__ in, int64_t c, int64_t ld, int lane, int ch) {
    seg_astrom<GRAD, NPT, MODE, ILP>(m, B, k0, k1, s_const, acc, stage, in, c, ld, lane, ch);
}
template <bool GRAD, int NPT, bool MARGIN, bool JIT, bool TREND, int ILP>
__device__ __noinline__ void seg_rv_ool(const DevModel& m, const DevBlock& B, int k0, int k1, const double* s_const,
                                        double* acc, double2* stage, const double* __restrict__ in, int64_t c, int64_t ld, int lane, int ch) {
    seg_rv<GRAD, NPT, MARGIN, JIT, TREND, ILP>(m, B, k0, k1, s_const, acc, stage, in, c, ld, lane, ch);
}
#ifndef OCTO_REDUCE_UNROLL
#define OCTO_REDUCE_UNROLL 8
#endif
constexpr int kReduceUnroll = OCTO_REDUCE_UNROLL;
#ifndef OCTO_ILP1_FROM_NPT
#define OCTO_ILP1_FROM_NPT 2       // table loops of kernels with at least this many planets keep ONE pair in flight per lane: two pairs x
                                   // several orbits spill (C3 20.2 -> 17.5 us per step, 3 lean planets 30.5 -> 26.9, 4: 48.5 -> 46.7)
#endif
#ifndef OCTO_THR_ASTROM_INLINE
#define OCTO_THR_ASTROM_INLINE 1
#endif
#ifndef OCTO_THR_RV_INLINE
#define OCTO_THR_RV_INLINE 1
#endif

// INL: the latency-tuned instantiations (and the resident explorer)
template <bool GRAD, int NPT, int ILP, bool LEAN, bool INL>
__device__ __forceinline__ void run_segment(const DevModel& m, const DevBlock& B, int k0, int k1, const double* s_const,
                                            double* acc, double2* stage, const double* __restrict__ in, int64_t c,
                                            int64_t ld, int lane, int ch) {
#define OCTO_SEG_ARGS m, B, k0, k1, s_const, acc, stage, in, c, ld, lane, ch
#ifdef OCTO_SEG_ALL_OOL      // analysis build (profiles/tools/sass_flops.py): every table loop as a subroutine of its own
    constexpr bool AI = false, RI = false;
#else
    constexpr bool AI = INL || OCTO_THR_ASTROM_INLINE, RI = INL || OCTO_THR_RV_INLINE;
#endif
    const bool astrom = B.kind <= OCTO_KIND_ASTROM_PASEP;
    const bool plain = B.kind == OCTO_KIND_ASTROM_RADEC && B.idx_platescale < 0 && B.idx_northangle < 0 && B.slot_obsprior < 0;
    if (LEAN || (astrom ? (plain && !B.jit) : (B.n_trend == 0 && B.kind != OCTO_KIND_RV_STAR_MARGIN))) {      // the lean loops
        if (astrom) {
            if constexpr (AI) seg_astrom<GRAD, NPT, 0, ILP>(OCTO_SEG_ARGS); else seg_astrom_ool<GRAD, NPT, 0, ILP>(OCTO_SEG_ARGS);
        } else if (B.jit) {
            if constexpr (RI) seg_rv<GRAD, NPT, false, true, false, ILP>(OCTO_SEG_ARGS); else seg_rv_ool<GRAD, NPT, false, true, false, ILP>(OCTO_SEG_ARGS);
        } else {
            if constexpr (RI) seg_rv<GRAD, NPT, false, false, false, ILP>(OCTO_SEG_ARGS); else seg_rv_ool<GRAD, NPT, false, false, false, ILP>(OCTO_SEG_ARGS);
        }
        return;
    }
    if constexpr (!LEAN) {
        if (astrom) {
            if (plain) seg_astrom_ool<GRAD, NPT, 1, 1>(OCTO_SEG_ARGS);
            else seg_astrom_ool<GRAD, NPT, 2, 1>(OCTO_SEG_ARGS);
        } else if (B.n_trend > 0) {
            if (B.kind == OCTO_KIND_RV_STAR_MARGIN) seg_rv_ool<GRAD, NPT, true, true, true, 1>(OCTO_SEG_ARGS);
            else if (B.jit) seg_rv_ool<GRAD, NPT, false, true, true, 1>(OCTO_SEG_ARGS);
            else seg_rv_ool<GRAD, NPT, false, false, true, 1>(OCTO_SEG_ARGS);
        } else {
            seg_rv_ool<GRAD, NPT, true, true, false, 1>(OCTO_SEG_ARGS);
        }
    }
#undef OCTO_SEG_ARGS
}

// Everything one evaluation needs beyond the model.  The same body serves the one-shot kernel (k_kepler_like: one
// CTA per chain group x epoch split, pointers into global memory) and the trajectory-resident explorer
// (octo_resident.cuh: one CTA per chain group loops over leapfrogs, pointers into its own shared memory).
struct EvalArgs {
    const double* in;            // [n_chains x n_in] kernel inputs, or θ_t [n_chains x D] with a parameterisation
    int64_t n_chains, ld;
    double* ll_out; double* g_out; int64_t ldg;
    double* partial; unsigned int* tickets;
    const DevParam* P; int post_mode; const double* pw_const;
    DevParam* P_smem;            // where the CTA keeps its copy of *P (FL 3: already filled by the caller), or nullptr
    HmcLeap leap;
    int ch;                      // chains per CTA: 32 / sub-lanes (see below)
    int64_t chain0;              // first chain of this CTA
    int group, gy, by;           // chain-group index (ticket / partial slot), epoch splits, this CTA's split
    bool lat_weights;            // split the epoch list by the latency cost model (DevBlock::wgt_lat)
};

// SUB-LANES.  The classic mapping is lane = chain (ch = 32).  When a batch has too few chains to fill the SMs that way,
// a warp takes ch = 32/S chains and S "sub-lanes" per chain, each sub-lane walking its own contiguous epoch range: the
// epoch split happens inside the warp instead of across CTAs (no partials through L2, no ticket, no second combine) and
// a chain group shrinks to ch chains, so S times as many CTAs exist.  All per-chain shared-memory arrays keep their
// [slot][32] shape and are REPLICATED across sub-lanes (lane l serves chain l mod ch): prologue, epilogue and the
// parameterisation stages run unchanged, computing S identical copies; only global reads (chain index) and writes
// (sub-lane 0 only) know about it.  Accumulators are folded over sub-lanes in sub-lane order right after the fixed-order
// sum over warps, so results stay run-to-run bit-reproducible.
// returns false when this CTA was not the last of its chain group to arrive (it has nothing more to do)
// FL (flavour): what the caller knows at compile time, so that the instantiation does not carry the other paths' code —
// 0 nothing; 1 no parameterisation stage (A.P == nullptr); 2 with one; 3 the resident explorer: with one, a single epoch
// split (no L2 combine), no pointwise mode, inputs never inline
template <bool GRAD, int NPT, int ILP, bool LEAN, int FL, bool INL>
__device__ __forceinline__ bool eval_cta(const DevModel& m, const EvalArgs& A, double* smem, const double* inl_v) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;                            // 8 unless the model needed a smaller CTA
    const int n_acc = m.n_acc;
    const int ch = A.ch, col_c = lane & (ch - 1);
    const int64_t n_chains = A.n_chains, ld = A.ld, ldg = A.ldg;
    const DevParam* Pg = FL == 1 ? nullptr : A.P;              // in global memory
    if (FL >= 2) __builtin_assume(Pg != nullptr);
    if (FL != 3 && Pg && A.P_smem) {                          // stage it (visible after the barrier below)
        const double* src = reinterpret_cast<const double*>(Pg);
        double* dst = reinterpret_cast<double*>(A.P_smem);
        for (int i = threadIdx.x; i < kParamWords; i += blockDim.x) dst[i] = src[i];
    }
    const DevParam* P = (Pg && A.P_smem) ? A.P_smem : Pg;
    const int post_mode = FL == 3 ? 0 : A.post_mode;
    const double* pw_const = FL == 3 ? nullptr : A.pw_const;
    const int n_split = FL == 3 ? 1 : A.gy;                   // epoch splits of a chain group across CTAs
    double* s_const = smem;                                   // [P][PC_COUNT][32]
    double* s_acc = s_const + m.n_planets * PC_COUNT * 32;    // [W][n_acc][32]
    double* s_red = s_acc + W * n_acc * 32;                   // [n_acc][32]
    double2* s_stage = reinterpret_cast<double2*>(s_red + n_acc * 32);   // [W][32 records][3]
    int* s_ok = reinterpret_cast<int*>(s_stage + W * 96);     // [32]
    double* s_in = reinterpret_cast<double*>(s_ok + 32);       // [n_in][32]: the kernel inputs of this CTA's chains
    __shared__ int s_last;

#ifdef OCTO_TIMING
    long long tm[12]; int tmi = 0;
#define OCTO_TICK() do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tm[tmi++] = (long long)t_; } } while (0)
    // per-warp cycle stamps (clock64 of this SM) at named points, printed by the resident kernel
#define OCTO_WTICK(i) do { if (lane == 0 && w < 16 && A.group == 0) g_wtk[w][i] = clock64(); } while (0)
#else
#define OCTO_TICK() do {} while (0)
#define OCTO_WTICK(i) do {} while (0)
#endif
    OCTO_WTICK(0);
    OCTO_TICK();
    // columns of the shared-memory arrays = chains of this CTA, replicated over sub-lanes
    constexpr int ncol = 32;
    const int64_t chain0 = A.chain0;
    auto chain_of = [&](int col) { const int64_t cc = chain0 + (col & (ch - 1)); return cc < n_chains ? cc : n_chains - 1; };
    const bool active = lane < ch && chain0 + lane < n_chains;      // lanes that own a chain's global outputs

    if (threadIdx.x < 32) {
        s_ok[threadIdx.x] = 1;
        if (Pg) param_smem(s_in + m.n_in * 32, m.n_in, Pg->D, Pg->n_tperi).flags[threadIdx.x] = 0;      // (the copy is not visible yet)
    }
    // the value-only kernel of a model without non-linear folds (marginalised RV, observable prior) needs one sum: ll
    const int n_use = (!GRAD && !(!LEAN && m.has_margin)) ? 1 : n_acc;
    double* acc = s_acc + w * n_acc * 32;
#pragma unroll 4
    for (int s = 0; s < n_use; ++s) acc[s * 32 + lane] = 0.0;
    __syncthreads();

    // ---- inputs of this CTA's chains into shared memory, with the finiteness check of logdensitymodel.jl:120-124;
    //      with a parameterisation `in` is θ_t and the inputs are derived here (param_forward)
    ParamSmem PS;
    const double* inp = (post_mode & OCTO_MODE_INLINE) ? inl_v : A.in;       // tiny batches carry their inputs in the parameters
    if (P) {
        PS = param_smem(s_in + m.n_in * 32, m.n_in, P->D, P->n_tperi);
        param_forward<LEAN>(*P, m, inp, chain_of(lane), ld, s_in, s_const, PS, s_ok, w, W, lane);      // includes K1's prologue
    } else {
#pragma unroll 1
        for (int it = threadIdx.x; it < ncol * m.n_in; it += W * 32) {
            const int col = it % ncol, k = it / ncol;
            const double v = inp[chain_of(col) + (int64_t)k * ld];
            s_in[it] = v;
            if (!isfinite(v)) s_ok[col] = 0;
        }
    }
    OCTO_TICK(); OCTO_WTICK(1);
    // ---- prologue phase 1: the per-planet tasks, spread over all threads as (column, task) items.  Without a
    //      parameterisation they read global memory themselves (no barrier between the staging loop and this one:
    //      the load latency overlaps the first, cold pass through the task code)
    if (!P) {
#pragma unroll 1
        for (int it = threadIdx.x; it < ncol * 5 * m.n_planets; it += W * 32) {
            const int col = it % ncol, task = it / ncol;
            if (!prologue_task<LEAN>(m, task / 5, task % 5, [&](int k) { return inp[chain_of(col) + (int64_t)k * ld]; }, s_const + (task / 5) * PC_COUNT * 32, col)) s_ok[col] = 0;
        }
        __syncthreads();
        // ---- phase 2: Thiele-Innes / RV products
#pragma unroll 1
        for (int it = threadIdx.x; it < ncol * m.n_planets; it += W * 32)
            prologue_products(s_const + (it / ncol) * PC_COUNT * 32, it % ncol, (!LEAN && m.any_ti) && m.basis[it / ncol] == OCTO_BASIS_THIELE_INNES);
    }
    __syncthreads();
    OCTO_TICK(); OCTO_WTICK(2);

    {
        // ---- this warp's contiguous range of the concatenated epoch list: unit u of U.  With sub-lanes, the part of
        //      every table inside the warp's range is divided evenly among the S sub-lanes (a warp whose range crosses
        //      a table boundary keeps all its sub-lanes busy in both tables)
        const int lch = 31 - __clz(ch);                                // ch and S are powers of two: shifts, not divisions
        const int S = 32 >> lch, sub = lane >> lch;
        const int64_t U = (int64_t)n_split * W, u = (int64_t)A.by * W + w;
        // contiguous range of the COST-weighted epoch list (an RV+jitter epoch costs ~1.8 lean astrometry epochs)
        // (latency-bound launches weigh a pair by its dependent chain instead of its instruction count)
        const double wtot = A.lat_weights ? m.wtot_lat : m.wtot;
        const double w_lo = wtot * (double)u / (double)U, w_hi = (u + 1 == U) ? 2.0 * wtot + 1.0 : wtot * (double)(u + 1) / (double)U;
#pragma unroll 1
        for (int b = 0; b < m.n_blocks; ++b) {
            const DevBlock& B = m.blocks[b];
            const double bcum = A.lat_weights ? B.cum_lat : B.cum, bwgt = A.lat_weights ? B.wgt_lat : B.wgt;
            int k0 = B.start + min(B.n, max(0, (int)ceil((w_lo - bcum) / bwgt)));
            int k1 = B.start + min(B.n, max(0, (int)ceil(fmin((w_hi - bcum) / bwgt, 2.0e9))));
            if (S > 1) {
                const int len = max(0, k1 - k0), base = k0;
                const int sh = 5 - lch;                                     // log2 S (no 64-bit division subroutine)
                k0 = base + (int)(((int64_t)len * sub) >> sh);
                k1 = base + (int)(((int64_t)len * (sub + 1)) >> sh);
            }
            if (pw_const) {      // pointwise mode (ch = 32): this CTA evaluates the single epoch `by` (warp 0)
                const int ep = A.by + (post_mode >> 9);                     // + first epoch of this chunk (grid.y <= 65535)
                const bool mine = w == 0 && ep >= B.start && ep < B.start + B.n;
                k0 = mine ? ep : 0; k1 = mine ? ep + 1 : 0;
            }
            if (!__any_sync(0xffffffffu, k0 < k1)) continue;
            run_segment<GRAD, NPT, ILP, LEAN, INL>(m, B, k0, k1, s_const, acc, s_stage + w * 96, s_in, lane, 32, lane, ch);
        }
    }
    OCTO_WTICK(3);
    __syncthreads();
    OCTO_TICK(); OCTO_WTICK(4);

    // ---- CTA reduction over the warps, fixed order; then the fold over sub-lanes, in sub-lane order, replicated
#pragma unroll 1
    for (int idx = threadIdx.x; idx < n_use * 32; idx += W * 32) {
        double v = s_acc[idx];
#pragma unroll (kReduceUnroll)
        for (int ww = 1; ww < W; ++ww) v += s_acc[ww * n_acc * 32 + idx];
        if (ch < 32) {
            double t = __shfl_sync(0xffffffffu, v, col_c);
            for (int sl = ch; sl < 32; sl += ch) t += __shfl_sync(0xffffffffu, v, sl + col_c);
            v = t;
        }
        s_red[idx] = v;
    }
    OCTO_WTICK(5);
    __syncthreads();
    OCTO_TICK(); OCTO_WTICK(6);

    // ---- K2: combine the epoch splits of this chain group: partials through L2 + a ticket; the last CTA to
    //      arrive sums them in split order (run-to-run bit-reproducible)
    if (n_split > 1 && !pw_const) {
        double* mine = A.partial + ((int64_t)A.group * n_split + A.by) * n_acc * 32;
#pragma unroll 2
        for (int idx = threadIdx.x; idx < n_use * 32; idx += W * 32) mine[idx] = s_red[idx];
        __threadfence();
        __syncthreads();
        OCTO_TICK();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(&A.tickets[A.group], 1u);
            s_last = (prev == (unsigned)n_split - 1);
            if (s_last) A.tickets[A.group] = 0;      // ready for the next launch on this workspace
        }
        __syncthreads();
        OCTO_TICK();
#ifdef OCTO_TIMING
        if (!s_last && threadIdx.x == 0 && A.group == 0)
            printf("cta(0,%d) ns: start %lld inputs +%lld prologue +%lld segments +%lld reduce +%lld write+fence +%lld ticket +%lld\n", A.by,
                   tm[0] % 100000000, tm[1] - tm[0], tm[2] - tm[1], tm[3] - tm[2], tm[4] - tm[3], tm[5] - tm[4], tm[6] - tm[5]);
#endif
        if (!s_last) return false;
        __threadfence();
        const double* base = A.partial + (int64_t)A.group * n_split * n_acc * 32;
        // every load of this thread (two accumulator cells x up to 16 splits) is issued before the first add: one
        // L2 round trip instead of four; additions stay in split order => same bits every run
        const int64_t stride = (int64_t)n_acc * 32;
        const int ncell = n_use * 32, gy = n_split;
#pragma unroll 1
        for (int idx0 = threadIdx.x; idx0 < ncell; idx0 += 2 * W * 32) {
            const int idx1 = idx0 + W * 32;
            const bool has1 = idx1 < ncell;
            const double* q0 = base + idx0;
            const double* q1 = base + (has1 ? idx1 : idx0);
            double v0 = 0.0, v1 = 0.0;
#pragma unroll 1
            for (int y0 = 0; y0 < gy; y0 += 16) {
                double t0[16], t1[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int y = y0 + j;
                    if (y < gy) { t0[j] = __ldcg(q0 + y * stride); t1[j] = __ldcg(q1 + y * stride); }
                    else { t0[j] = 0.0; t1[j] = 0.0; }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) if (y0 + j < gy) { v0 += t0[j]; v1 += t1[j]; }
            }
            s_red[idx0] = v0;
            if (has1) s_red[idx1] = v1;
        }
        __syncthreads();
    }

    OCTO_TICK();
    // ---- HGCA tables: evaluated here, on the complete sums (not in pointwise mode: they are not part of the epoch list)
    if (!LEAN && m.n_hg > 0 && !pw_const) {
#pragma unroll 1
        for (int h = 0; h < m.n_hg; ++h) hgca_tail<GRAD>(m, m.hg[h], s_const, s_red, s_acc, s_in, n_acc, w, W, lane);
    }
    // ---- epilogue (see epilogue_part): gradient parts live in the now free per-warp accumulator area
    double* s_gp = s_acc;                                     // [EPI_PARTS][n_g][32]
    const int n_g = m.n_in + m.n_planets;                     // + one virtual column per planet (Thiele-Innes: d/da)
    if (GRAD) {
#pragma unroll 1
        for (int idx = threadIdx.x; idx < EPI_PARTS * n_g * 32; idx += W * 32) s_gp[idx] = 0.0;
    }
    __syncthreads();
    OCTO_WTICK(7);
    if ((!LEAN && m.has_margin)) {
        if (w == 0) epilogue_margin<GRAD>(m, s_const, s_red, s_gp, s_in, lane, 32, lane, pw_const != nullptr);
        __syncthreads();
    }
    // the partial derivatives of every θ_at_epoch_to_tperi depend on forward quantities only: the warps that have no
    // gradient part compute them now instead of after the epilogue (param_backward then only folds them in)
    const bool tperi_early = GRAD && P && P->n_tperi > 0 && W >= EPI_PARTS + 1 + P->n_tperi;
    if (GRAD) {
#pragma unroll 1
        for (int part = w; part < EPI_PARTS; part += W) epilogue_part<LEAN>(part, m, s_const, s_red, s_gp + part * n_g * 32, lane);
        if (tperi_early && w >= EPI_PARTS && w < EPI_PARTS + P->n_tperi) param_tperi_partials(*P, m, s_in, PS, w - EPI_PARTS, lane);
    }
    if (w == W - 1) {                                          // ll: a warp without a gradient part when W = 8
        const double cll = pw_const ? pw_const[A.by] : m.const_ll;
        const double llv = s_ok[lane] ? s_red[lane] + cll : -CUDART_INF;
        if (!P) {
            // pointwise mode: out[chain + epoch * ldg] = ln_like of the model reduced to that one epoch
            if (active) A.ll_out[chain0 + lane + (pw_const ? (int64_t)A.by * ldg : 0)] = llv;
        } else {                                               // log posterior (logdensitymodel.jl:110-146)
            const int fl = PS.flags[lane];
            const bool ok = (fl & 4) && isfinite(llv);
            // tempering (octo_hmc.cu): the likelihood of chain c enters with weight beta[c]
            const double bet = A.leap.beta ? A.leap.beta[chain_of(lane)] : 1.0;
            PS.beta[lane] = bet;
            if (A.leap.ll_raw && active) A.leap.ll_raw[chain0 + lane] = ok ? llv : -CUDART_INF;
            // post_mode 1: the likelihood part alone, ln_like(system, arr2nt(θ)) incl. the UnitLengthPrior terms
            const double like = fma(bet, llv, PS.extra[lane]);
            if (active) A.ll_out[chain0 + lane] = !(fl & 1) ? -CUDART_INF : (ok ? ((post_mode & 255) == 1 ? like : PS.lp[lane] + like) : -CUDART_INF);
            PS.flags[lane] = fl | (ok ? 8 : 0);
        }
    }
    OCTO_WTICK(8);
    if (GRAD) {
        __syncthreads();
        OCTO_WTICK(9);
        const int ng = n_g * 32;
        if ((!LEAN && m.any_ti)) {
            if (w == 0) epilogue_ti_distribute(m, s_const, s_gp, ng, lane);
            __syncthreads();
        }
#pragma unroll 1
        for (int idx = threadIdx.x; idx < m.n_in * 32; idx += W * 32) {
            const int l = idx & 31;
            const double v = ((s_gp[idx] + s_gp[ng + idx]) + s_gp[2 * ng + idx]) + s_gp[3 * ng + idx];
            if (P) PS.aux[idx] = (PS.flags[l] & 8) ? PS.beta[l] * v : 0.0;
            else if (l < ch && chain0 + l < n_chains) A.g_out[chain0 + l + (int64_t)(idx >> 5) * ldg] = s_ok[l] ? v : 0.0;
        }
        if (P) {
            __syncthreads();
            OCTO_TICK(); OCTO_WTICK(10);
            param_backward(*P, m, s_in, PS, A.g_out, chain0, active, ldg, w, W, lane, A.leap, tperi_early);
        }
    }
    OCTO_WTICK(11);

#ifdef OCTO_TIMING
    if (threadIdx.x == 0 && A.group == 0 && A.gy > 1) {
        OCTO_TICK();
        printf("cta(0,%d) LAST ns: start %lld inputs +%lld prologue +%lld segments +%lld reduce +%lld write+fence +%lld ticket +%lld reads +%lld epilogue +%lld (+%lld) end %lld\n",
               A.by, tm[0] % 100000000, tm[1] - tm[0], tm[2] - tm[1], tm[3] - tm[2], tm[4] - tm[3], tm[5] - tm[4],
               tm[6] - tm[5], tm[7] - tm[6], tm[8] - tm[7], P && GRAD ? tm[9] - tm[8] : 0LL, tm[tmi - 1] % 100000000);
    }
#endif
#ifdef OCTO_TIMING
    if (threadIdx.x == 0 && A.group == 0 && A.gy == 1) { OCTO_TICK(); for (int i = 0; i < tmi && i < 12; ++i) g_tm_last[i] = tm[i]; g_tm_last[12] = tmi; }
#endif
    return true;
}

// LEAN = the model is made of the lean tables only (plain RA/Dec astrometry without jitter, RV with or without jitter;
// no marginalised RV, trend, observable prior, HGCA or Thiele-Innes planet — DevModel::lean): those kernels do not carry the
// code of the other tables.  Not a matter of tidiness: in the latency regime the SAME executed instructions ran 6-13 %
// slower or faster depending on how much never-executed code the kernel's text section held (instruction fetch).
// LAT = latency-tuned instantiation: no register cap (one CTA per SM, no spills) for launches whose whole grid is a
// single wave of at most one CTA per SM; the other instantiation keeps two CTAs per SM resident for throughput.
template <bool GRAD, int NPT, bool LAT, bool LEAN, int FL>
__global__ void __launch_bounds__((LAT ? OCTO_LAT_WARPS : WMAX) * 32, LAT ? 1 : OCTO_MIN_CTAS)
k_kepler_like(const __grid_constant__ DevModel m, const double* __restrict__ in, int64_t n_chains, int64_t ld,
              double* __restrict__ ll_out, double* __restrict__ g_out, int64_t ldg, double* __restrict__ partial,
              unsigned int* __restrict__ tickets, const DevParam* __restrict__ P, int post_mode,
              const double* __restrict__ pw_const, const HmcLeap leap, int ch, const __grid_constant__ InlineIn inl) {
    extern __shared__ double smem[];
#ifndef OCTO_NO_PDL
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
#ifdef OCTO_EMPTY
    if (n_chains > 0) { if (threadIdx.x == 0 && blockIdx.y == 0) ll_out[blockIdx.x] = 0.0; return; }   // launch-floor probe
#endif
    EvalArgs A;
    A.in = in; A.n_chains = n_chains; A.ld = ld; A.ll_out = ll_out; A.g_out = g_out; A.ldg = ldg;
    A.partial = partial; A.tickets = tickets; A.P = P; A.post_mode = post_mode; A.pw_const = pw_const; A.leap = leap;
    A.P_smem = nullptr;
    if (FL != 1 && P) {
        unsigned dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        A.P_smem = reinterpret_cast<DevParam*>(smem + (dyn >> 3) - kParamWords);
    }
    A.ch = ch; A.chain0 = (int64_t)blockIdx.x * ch; A.group = blockIdx.x; A.gy = gridDim.y; A.by = blockIdx.y; A.lat_weights = LAT;
    eval_cta<GRAD, NPT, (NPT >= OCTO_ILP1_FROM_NPT) ? 1 : (LAT ? OCTO_LAT_ILP : OCTO_THR_ILP), LEAN, FL, LAT>(m, A, smem, inl.v);
#ifdef OCTO_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0 && gridDim.y == 1 && !g_quiet) {      // single-split launch: phase stamps of CTA 0
        const long long* t = g_tm_last;
        printf("cta 0 ns (gy = 1): inputs/forward +%lld prologue +%lld segments +%lld reduce +%lld [hgca] +%lld epilogue +%lld%s +%lld\n",
               t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], P && GRAD ? " backward" : "", t[12] > 7 ? t[7] - t[6] : 0LL);
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// Trajectory-resident HMC explorer (the loop of src/sampling.jl:412-423, batched).  Chain groups are
// independent, so ONE CTA keeps the state of its ch chains (position, momentum, gradient, proposal) in shared memory
// and runs n_iter transitions x n_leapfrog leapfrogs inside one launch, calling the same evaluation body as the
// one-shot kernel on shared-memory pointers (the fused parameterisation applies the kick and the drift itself).  No
// launch and no L2 round trip per leapfrog.  The epoch split of a chain group happens inside the warps (sub-lanes),
// never across CTAs.  Same per-coordinate arithmetic as k_hmc_turn (octo_hmc_dev.cuh), same summation orders as a
// one-shot launch of the same geometry: both explorers produce the same trajectories bit for bit (the log-posterior
// values they report may differ in the last bit: each kernel inlines the evaluation and is contracted on its own).
// ---------------------------------------------------------------------------------------------
template <int NPT, bool LEAN>
__global__ void __launch_bounds__(OCTO_LAT_WARPS * 32, 1)
k_hmc_resident(const __grid_constant__ DevModel m, const DevParam* __restrict__ P, const ResidentArgs R, int ch, int eval_doubles) {
    using namespace octo_hmc_dev;
    extern __shared__ double smem[];
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int D = R.D, tid = threadIdx.x, nthr = blockDim.x;
#ifdef OCTO_TIMING
    g_quiet = 1;
#endif
    // state behind the evaluation body's own shared memory, [j][32] like everything else
    double* s_q = smem + eval_doubles;
    double* s_g = s_q + D * 32;
    double* s_qp = s_g + D * 32;
    double* s_gp = s_qp + D * 32;
    double* s_p = s_gp + D * 32;
    double* s_kin = s_p + D * 32;            // per-coordinate kinetic terms, summed per chain in coordinate order
    double* s_lp = s_kin + D * 32;
    double* s_lpp = s_lp + 32;
    double* s_h0 = s_lpp + 32;
    double* s_ll = s_h0 + 32;
    double* s_llp = s_ll + 32;
    double* s_beta = s_llp + 32;
    double* s_acc = s_beta + 32;
    double* s_im = s_acc + 32;               // [D]
    const int64_t chain0 = (int64_t)blockIdx.x * ch;
    const int nvalid = (int)min((int64_t)ch, R.n - chain0);
    const bool tempered = R.beta != nullptr;
    for (int it = tid; it < D * 32; it += nthr) {
        const int col = it & 31, j = it >> 5;
        s_q[it] = col < nvalid ? R.q[chain0 + col + (int64_t)j * R.n] : 0.0;
    }
    if (tid < D) s_im[tid] = R.inv_mass[tid];
    DevParam* sP = reinterpret_cast<DevParam*>(smem + eval_doubles - kParamWords);      // staged once for the whole run
    for (int i = tid; i < kParamWords; i += nthr) reinterpret_cast<double*>(sP)[i] = reinterpret_cast<const double*>(P)[i];
    if (tid < 32) {
        s_beta[tid] = (tempered && tid < nvalid) ? R.beta[chain0 + tid] : 1.0;
        s_acc[tid] = 0.0; s_ll[tid] = 0.0; s_llp[tid] = 0.0;
    }
    __syncthreads();

    EvalArgs A;
    A.n_chains = nvalid; A.ld = 32; A.ldg = 32; A.partial = nullptr; A.tickets = nullptr; A.P = P; A.P_smem = sP; A.post_mode = 0;
    A.pw_const = nullptr; A.ch = ch; A.chain0 = 0; A.group = blockIdx.x; A.gy = 1; A.by = 0; A.lat_weights = true;
    // fresh momentum, first half kick and drift: one (chain, coordinate) item per thread
    auto start_transition = [&](int it) {
        const uint64_t key = hmc_key(R.seed, (uint64_t)it);
        for (int item = tid; item < D * 32; item += nthr) {
            const int col = item & 31, j = item >> 5;
            if (col >= nvalid) continue;
            const HmcStart r = hmc_start(key, (uint64_t)(R.chain_offset + chain0 + col), j, s_im[j], R.eps, s_g[item], s_q[item]);
            s_kin[item] = r.kin; s_p[item] = r.p; s_qp[item] = r.q;
        }
        __syncthreads();
        if (tid < nvalid) {
            double kin = 0.0;
            for (int j = 0; j < D; ++j) kin = __dadd_rn(kin, s_kin[j * 32 + tid]);
            s_h0[tid] = hmc_h0(s_lp[tid], kin);
        }
    };
    // Metropolis step, sample store
    auto finish_transition = [&](int it) {
        if (tid < nvalid) {
            double kin = 0.0;
            for (int j = 0; j < D; ++j) kin = __dadd_rn(kin, hmc_kin(s_p[j * 32 + tid], s_im[j]));
            const double lpp = s_lpp[tid];
            const int64_t c = chain0 + tid;
            if (hmc_accept(R.seed, it, (uint64_t)(R.chain_offset + c), D, s_h0[tid], lpp, kin)) {
                for (int j = 0; j < D; ++j) { s_q[j * 32 + tid] = s_qp[j * 32 + tid]; s_g[j * 32 + tid] = s_gp[j * 32 + tid]; }
                s_lp[tid] = lpp; s_acc[tid] += 1.0; s_ll[tid] = s_llp[tid];
            }
            if (R.out_lp) R.out_lp[(int64_t)(R.it0 + it) * R.n + c] = s_lp[tid];
            if (R.out_theta) for (int j = 0; j < D; ++j) R.out_theta[((int64_t)(R.it0 + it) * D + j) * R.n + c] = s_q[j * 32 + tid];
        }
        __syncthreads();
    };
    // one call site of the evaluation body for every evaluation of the run (it = -1: the current states at the current
    // weights, then the leapfrogs of every transition): its code is fetched once and stays in the instruction cache
    for (int it = -1, l = 0;;) {
        if (it < 0) {
            A.in = s_q; A.ll_out = s_lp; A.g_out = s_g;
            A.leap = HmcLeap{nullptr, nullptr, nullptr, 0.0, 0.0, 0, 0, tempered ? s_beta : nullptr, tempered ? s_ll : nullptr};
        } else {
            const bool last = l == R.n_leapfrog - 1;
            A.in = s_qp; A.ll_out = s_lpp; A.g_out = s_gp;
            A.leap = HmcLeap{s_p, s_qp, s_im, R.eps, last ? 0.5 * R.eps : R.eps, last ? 0 : 1, 0, tempered ? s_beta : nullptr, tempered ? s_llp : nullptr};
        }
        eval_cta<true, NPT, (NPT >= OCTO_ILP1_FROM_NPT) ? 1 : OCTO_LAT_ILP, LEAN, 3, true>(m, A, smem, nullptr);
        __syncthreads();
#ifdef OCTO_TIMING
        if (blockIdx.x == 0 && tid == 0 && it == R.n_iter - 1 && l == R.n_leapfrog - 1) {
            printf("resident eval ns:");
            for (int i = 1; i < (int)g_tm_last[12]; ++i) printf(" +%lld", g_tm_last[i] - g_tm_last[i - 1]);
            printf("  (forward, prologue, segments, reduce, [hgca], epilogue, backward)\n");
            printf("  forward ns: priors +%lld inputs +%lld trig +%lld tperi +%lld sums +%lld\n", g_ptk[0][1] - g_ptk[0][0], g_ptk[0][2] - g_ptk[0][1], g_ptk[0][3] - g_ptk[0][2], g_ptk[0][4] - g_ptk[0][3], g_ptk[0][5] - g_ptk[0][4]);
            printf("  backward ns: tperi +%lld accumulate +%lld barrier +%lld stores +%lld\n", g_ptk[1][1] - g_ptk[1][0], g_ptk[1][2] - g_ptk[1][1], g_ptk[1][3] - g_ptk[1][2], g_ptk[1][4] - g_ptk[1][3]);
            printf("  per-warp cycles since warp 0's start: forward-done prologue-done segments-done barrier reduce-done barrier epilogue-start epilogue-done barrier backward-start end\n");
            for (int ww = 0; ww < (int)(blockDim.x >> 5) && ww < 16; ++ww) {
                printf("   warp %2d:", ww);
                for (int i = 0; i < 12; ++i) printf(" %6lld", g_wtk[ww][i] - g_wtk[0][0]);
                printf("\n");
            }
        }
#endif
        if (it >= 0 && ++l < R.n_leapfrog) continue;
        if (it >= 0) finish_transition(it);
        if (++it >= R.n_iter) break;
        start_transition(it);
        l = 0;
    }
    for (int item = tid; item < D * 32; item += nthr) {
        const int col = item & 31, j = item >> 5;
        if (col >= nvalid) continue;
        R.q[chain0 + col + (int64_t)j * R.n] = s_q[item];
        R.g[chain0 + col + (int64_t)j * R.n] = s_g[item];
    }
    if (tid < nvalid) {
        R.lp[chain0 + tid] = s_lp[tid];
        R.acc[chain0 + tid] += s_acc[tid];
        if (R.ll) R.ll[chain0 + tid] = s_ll[tid];
        if (R.pair_out) {                 // what this chain contributes to the swap round that follows
            double l_ref, l_target;
            pt_pair(s_lp[tid], s_beta[tid], s_ll[tid], l_ref, l_target);
            R.pair_out[2 * (chain0 + tid)] = l_ref; R.pair_out[2 * (chain0 + tid) + 1] = l_target;
        }
    }
}


__global__ void k_selftest_kepler(const double* __restrict__ MA, const double* __restrict__ e, int64_t n,
                                  double* __restrict__ sE, double* __restrict__ cE) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Orb o;
    o.nd = 1.0; o.tp = 0.0; o.e = e[i];
    o.ef = (float)o.e; o.omef = (float)(1.0 - o.e); o.ca1f = (float)(kA1 / (1.0 + o.e));
    double dt, s, c;
    kepler_sincos(o, MA[i], dt, s, c);
    sE[i] = s; cE[i] = c;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host side.  build.py compiles this file once per planet-count instantiation (-DOCTO_NPT=1, 2, 3, 4: four objects, in
// parallel — one object with every instantiation takes > 5 minutes) and per lean / full kernel family (-DOCTO_LEANSEL); each
// object exports its entry points as an OctoNptEntry, and the (1 planet, full) object also holds the dispatchers.  Without
// -DOCTO_NPT everything is one object.
// ---------------------------------------------------------------------------------------------
namespace {
template <bool GRAD, int NPT, bool LAT, bool LEAN, int FL>
cudaError_t launch_t(const DevModel& m, const LaunchGeom& g, const double* d_in, int64_t n, int64_t ld,
                     double* d_ll, double* d_g, int64_t ldg, double* d_partial, unsigned int* d_tickets,
                     const DevParam* d_param, int post_mode, const double* d_pw_const, const HmcLeap& leap,
                     cudaStream_t st, const InlineIn* inl) {
    // programmatic dependent launch: the kernel lets the next launch on the stream be scheduled while it is still
    // running (griddepcontrol.launch_dependents) and itself waits for everything before it in the stream to complete
    // and become visible before it touches memory (griddepcontrol.wait) — stream semantics, minus the launch gap
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.gx, g.gy); cfg.blockDim = dim3(g.block); cfg.dynamicSmemBytes = g.smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
#ifdef OCTO_NO_PDL
    cfg.numAttrs = 0;
#endif
    static const InlineIn none{};
    return cudaLaunchKernelEx(&cfg, k_kepler_like<GRAD, NPT, LAT, LEAN, FL>, m, d_in, n, ld, d_ll, d_g, ldg, d_partial, d_tickets, d_param,
                              inl ? (post_mode | OCTO_MODE_INLINE) : post_mode, d_pw_const, leap, g.ch, inl ? *inl : none);
}

// opt every instantiation in to the device's full dynamic shared memory (a per-function, process-wide attribute:
// contexts of different sizes must not shrink it for each other)
template <int NPT, bool LEAN>
cudaError_t npt_attr(size_t smem_optin) {
    cudaError_t e;
#define OCTO_ATTR1(G, L, F)                                                                                          \
    e = cudaFuncSetAttribute(k_kepler_like<G, NPT, L, LEAN, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin); \
    if (e != cudaSuccess) return e
    if constexpr (LEAN) { OCTO_ATTR1(true, false, 1); OCTO_ATTR1(true, true, 1); OCTO_ATTR1(true, false, 2); OCTO_ATTR1(true, true, 2);
                          OCTO_ATTR1(false, false, 1); OCTO_ATTR1(false, true, 1); OCTO_ATTR1(false, false, 2); OCTO_ATTR1(false, true, 2); }
    else { OCTO_ATTR1(true, false, 0); OCTO_ATTR1(true, true, 0); OCTO_ATTR1(false, false, 0); OCTO_ATTR1(false, true, 0); }
#undef OCTO_ATTR1
    return cudaFuncSetAttribute(k_hmc_resident<NPT, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
}
// resident CTAs per SM of the gradient kernel (throughput instantiation) this model dispatches to: drives the launch geometry
template <int NPT, bool LEAN>
cudaError_t npt_occupancy(const DevModel&, int W, size_t smem_bytes, int* ctas_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_kepler_like<true, NPT, false, LEAN, LEAN ? 1 : 0>, W * 32, smem_bytes);
}
template <int NPT, bool LEAN>
cudaError_t npt_launch(const DevModel& m, const LaunchGeom& g, bool grad, const double* d_in, int64_t n_chains,
                       int64_t ld, double* d_ll, double* d_g, int64_t ldg, double* d_partial,
                       unsigned int* d_tickets, const DevParam* d_param, int post_mode, const double* d_pw_const,
                       const HmcLeap& leap, cudaStream_t st, const InlineIn* inl) {
#define OCTO_ARGS m, g, d_in, n_chains, ld, d_ll, d_g, ldg, d_partial, d_tickets, d_param, post_mode, d_pw_const, leap, st, inl
#define OCTO_DISPATCH2(F)                                                                                         \
    if (g.lat) return grad ? launch_t<true, NPT, true, LEAN, F>(OCTO_ARGS) : launch_t<false, NPT, true, LEAN, F>(OCTO_ARGS);   \
    return grad ? launch_t<true, NPT, false, LEAN, F>(OCTO_ARGS) : launch_t<false, NPT, false, LEAN, F>(OCTO_ARGS)
    // lean models: separate instantiations with and without the parameterisation stage (see eval_cta, FL)
    if constexpr (LEAN) {
        if (d_param) { OCTO_DISPATCH2(2); }
        OCTO_DISPATCH2(1);
    } else {
        OCTO_DISPATCH2(0);
    }
#undef OCTO_DISPATCH2
#undef OCTO_ARGS
}
template <int NPT, bool LEAN>
cudaError_t npt_resident(const DevModel& m, const cudaLaunchConfig_t* cfg, const DevParam* d_param, const ResidentArgs& R, int ch, int eval_doubles) {
    return cudaLaunchKernelEx(cfg, k_hmc_resident<NPT, LEAN>, m, d_param, R, ch, eval_doubles);
}
#define OCTO_ENTRY(N, L) {&npt_attr<N, L>, &npt_occupancy<N, L>, &npt_launch<N, L>, &npt_resident<N, L>}
}  // namespace

// which instantiations this object holds: all of them, or (build.py) -DOCTO_NPT=1|2|3|4 -DOCTO_LEANSEL=0|1
#ifdef OCTO_NPT
#define OCTO_HAS(N, L) (OCTO_NPT == N && OCTO_LEANSEL == L)
#else
#define OCTO_HAS(N, L) 1
#endif
#ifdef OCTO_NPT_ONLY1                // tuning builds (profiles/tools/ab.sh): one-planet kernels only
#define OCTO_HAS_MULTI(N, L) 0
#else
#define OCTO_HAS_MULTI(N, L) OCTO_HAS(N, L)
#endif
#if OCTO_HAS(1, 0)
extern const OctoNptEntry octo_entry_n1_full = OCTO_ENTRY(1, false);
#endif
#if OCTO_HAS(1, 1)
extern const OctoNptEntry octo_entry_n1_lean = OCTO_ENTRY(1, true);
#endif
#if OCTO_HAS_MULTI(2, 0)
extern const OctoNptEntry octo_entry_n2_full = OCTO_ENTRY(2, false);
#endif
#if OCTO_HAS_MULTI(2, 1)
extern const OctoNptEntry octo_entry_n2_lean = OCTO_ENTRY(2, true);
#endif
#if OCTO_HAS_MULTI(3, 0)
extern const OctoNptEntry octo_entry_n3_full = OCTO_ENTRY(3, false);
#endif
#if OCTO_HAS_MULTI(3, 1)
extern const OctoNptEntry octo_entry_n3_lean = OCTO_ENTRY(3, true);
#endif
#if OCTO_HAS_MULTI(4, 0)
extern const OctoNptEntry octo_entry_n4_full = OCTO_ENTRY(4, false);
#endif
#if OCTO_HAS_MULTI(4, 1)
extern const OctoNptEntry octo_entry_n4_lean = OCTO_ENTRY(4, true);
#endif

#if OCTO_HAS(1, 0)                   // this object also holds the dispatchers
#ifdef OCTO_NPT_ONLY1
extern const OctoNptEntry octo_entry_n2_full = {nullptr, nullptr, nullptr, nullptr}, octo_entry_n2_lean = {nullptr, nullptr, nullptr, nullptr};
extern const OctoNptEntry octo_entry_n3_full = {nullptr, nullptr, nullptr, nullptr}, octo_entry_n3_lean = {nullptr, nullptr, nullptr, nullptr};
extern const OctoNptEntry octo_entry_n4_full = {nullptr, nullptr, nullptr, nullptr}, octo_entry_n4_lean = {nullptr, nullptr, nullptr, nullptr};
#endif
extern const OctoNptEntry octo_entry_n1_lean, octo_entry_n2_full, octo_entry_n2_lean, octo_entry_n3_full, octo_entry_n3_lean, octo_entry_n4_full,
    octo_entry_n4_lean;
static const OctoNptEntry& entry_of(const DevModel& m) {
    if (m.n_planets == 1) return m.lean ? octo_entry_n1_lean : octo_entry_n1_full;
    if (m.n_planets == 2) return m.lean ? octo_entry_n2_lean : octo_entry_n2_full;
    if (m.n_planets == 3) return m.lean ? octo_entry_n3_lean : octo_entry_n3_full;
    return m.lean ? octo_entry_n4_lean : octo_entry_n4_full;
}

cudaError_t octo_selftest_kepler_launch(const double* d_MA, const double* d_e, int64_t n, double* d_s, double* d_c) {
    k_selftest_kepler<<<(unsigned)((n + 255) / 256), 256>>>(d_MA, d_e, n, d_s, d_c);
    return cudaGetLastError();
}

// D > 0: with the fused parameterisation stage (D parameters, T θ_at_epoch_to_tperi definitions)
size_t octo_smem_bytes(const DevModel& m, int W, int D, int T) {
    size_t acc = (size_t)W * m.n_acc * 32, gp = (size_t)EPI_PARTS * (m.n_in + m.n_planets) * 32;      // the epilogue's gradient parts reuse the accumulator area
    size_t d = (size_t)m.n_planets * PC_COUNT * 32 + (acc > gp ? acc : gp) + (size_t)m.n_acc * 32 + (size_t)m.n_in * 32;
    if (D > 0) d += param_smem_doubles(m.n_in, D, T) + 32 + kParamWords;         // + the CTA's own copy of DevParam (at the end)
    return d * sizeof(double) + (size_t)W * 96 * sizeof(double2) + (size_t)32 * sizeof(int);
}

cudaError_t octo_kernels_init(const DevModel& m, size_t smem_bytes, size_t smem_optin, int W, int* ctas_per_sm) {
    for (const OctoNptEntry* en : {&octo_entry_n1_full, &octo_entry_n1_lean, &octo_entry_n2_full, &octo_entry_n2_lean, &octo_entry_n3_full,
                                   &octo_entry_n3_lean, &octo_entry_n4_full, &octo_entry_n4_lean}) {
        if (!en->attr) continue;
        const cudaError_t e = en->attr(smem_optin);
        if (e != cudaSuccess) return e;
    }
    const OctoNptEntry& en = entry_of(m);
    return en.occupancy ? en.occupancy(m, W, smem_bytes, ctas_per_sm) : cudaErrorNotSupported;
}

// d_param != nullptr: d_in is θ_t [n x D], d_ll receives the log posterior (post_mode 1: its likelihood part) and d_g
// its gradient [n x D].  d_pw_const != nullptr (value-only): pointwise mode, grid.y = epochs, d_ll is [n x E] with
// leading dimension ldg.
cudaError_t octo_launch(const DevModel& m, const LaunchGeom& g, bool grad, const double* d_in, int64_t n_chains,
                        int64_t ld, double* d_ll, double* d_g, int64_t ldg, double* d_partial,
                        unsigned int* d_tickets, const DevParam* d_param, int post_mode, const double* d_pw_const,
                        const HmcLeap& leap, cudaStream_t st, const InlineIn* inl) {
    const OctoNptEntry& en = entry_of(m);
    if (!en.launch) return cudaErrorNotSupported;
    return en.launch(m, g, grad, d_in, n_chains, ld, d_ll, d_g, ldg, d_partial, d_tickets, d_param, post_mode, d_pw_const, leap, st, inl);
}

// ---- trajectory-resident explorer
size_t octo_resident_smem_bytes(const DevModel& m, int D, int T) {
    return octo_smem_bytes(m, OCTO_LAT_WARPS, D, T) + ((size_t)(6 * D + 7) * 32 + (size_t)D) * sizeof(double);
}
cudaError_t octo_resident_init(const DevModel&, size_t) { return cudaSuccess; }      // attributes: octo_kernels_init
cudaError_t octo_resident_launch(const DevModel& m, const DevParam* d_param, int T, const ResidentArgs& R, int ch, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    const size_t eval_bytes = octo_smem_bytes(m, OCTO_LAT_WARPS, R.D, T);
    cfg.gridDim = dim3((unsigned)((R.n + ch - 1) / ch)); cfg.blockDim = dim3(OCTO_LAT_WARPS * 32);
    cfg.dynamicSmemBytes = octo_resident_smem_bytes(m, R.D, T); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const OctoNptEntry& en = entry_of(m);
    if (!en.resident) return cudaErrorNotSupported;
    return en.resident(m, &cfg, d_param, R, ch, (int)((eval_bytes + 7) / 8));
}
#endif   // dispatchers
