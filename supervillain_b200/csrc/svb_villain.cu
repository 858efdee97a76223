// svb_villain.cu -- batched checkerboard Metropolis sweeps of the Villain action
//   S = kappa/2 sum_l (d phi - 2 pi n)_l^2
// replacing NeighborhoodUpdate.step (supervillain/generator/villain/neighborhood.py:59-137),
// Villain.__call__ (supervillain/action/villain.py:51-66) and the scalar Villain observables
// (supervillain/observable/{action,energy,winding,wrapping}.py).
//
// Per site x, with the four links touching x
//   f0 = (0,x)  b0 = (0,x-e0)  f1 = (1,x)  b1 = (1,x-e1)
//   r_l   = (phi[head] - phi[tail]) - 2 pi n_l
//   dr_f  = (0 - dphi) - 2 pi dn_f          dr_b = (dphi - 0) - 2 pi dn_b
//   dS    = 0 + s_f0 + s_b0 + s_f1 + s_b1,  s_l = ((kappa/2) dr_l) ((2 r_l) + dr_l)
//   accept iff u < min(1, exp(-dS)); then phi[x] += dphi, n_l += dn_l.
// Same-colour sites touch disjoint links and never neighbour each other, so one colour is
// updated concurrently without atomics; colours are separated by a barrier.
//
// The kernel is instruction-issue bound (Philox + fp64 dS + exp per site against 32 B of HBM
// traffic), so the hot instantiations fix the lattice size and block size at compile time: all
// index arithmetic folds to shifts/immediates, loops unroll, integer->double conversions go
// through the fp64 adder instead of the conversion unit, and integer reductions use REDUX.

#include <type_traits>

#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)

#include "svb_common.cuh"

#ifndef SVB_SITE_UNROLL
#define SVB_SITE_UNROLL 1
#endif
#ifdef SVB_CVT_I2F
#define SVB_CVT(n) ((double)(n))      /* I2F.F64 on the conversion unit: 16 lanes/clk/SM, measured 6 % slower */
#else
#define SVB_CVT(n) int_to_double(n)   /* LOP3 + DADD on the alu / fp64 pipes */
#endif
#ifndef SVB_PHILOX_ROUNDS
#define SVB_PHILOX_ROUNDS 10
#endif
#ifndef SVB_MINB32
#define SVB_MINB32 6
#endif

namespace svb {

constexpr int kSiteUnroll = SVB_SITE_UNROLL;

struct VillainArgs {
    void* phi;
    int32_t* n;
    long long chains;
    int N;
    double kappa;
    const double* kappa_chain;
    int W;
    double interval_phi;
    int interval_n;
    int n_sweeps;
    unsigned long long seed, sweep0, chain0;
    uint32_t round_key[20];   // Philox key schedule (k0 + r W0, k1 + r W1), r = 0..9, precomputed on the host
    uint32_t stream_hi;       // the proposal stream of the generator kind, << 24 (counter word 3)
    uint32_t refine_stream;   // its refinement stream
    int wide;                 // K^4 > 256: the uniform's leading bits come from the refinement block (draw mapping below)
    const double* inj_u;
    const double* inj_dphi;
    const int32_t* inj_dn_fwd;
    const int32_t* inj_dn_bwd;
    double* obs;
    uint8_t* accept_mask;
    double* dS_out;
    // ExactUpdate as a mode of the site kernels (svb_villain_decoupled): proposals are (dphi = 0, dn = d z restricted to x)
    int exact_mode;
    const int32_t* inj_z;
    // Philox launches of the decoupled updates: the fp32-filtered kernel with the STRICT cold path (svb_villain_filtered.cuh)
    int filtered_strict;
    // overlapped launches (svb_villain_sweep_overlapped); epochs == nullptr otherwise
    double* obs_in;
    uint32_t* epochs;
    uint32_t wait_epoch, signal_epoch;
    int grid_wait;
};

// exact int32 -> double on the fp64 add pipe: (2^52 + 2^31 + n) - (2^52 + 2^31)
__device__ __forceinline__ double int_to_double(int n) {
    return __hiloint2double(0x43300000, n ^ 0x80000000) - 4503601774854144.0;
}

// ------------------------------------------------------------------------------------------
// arithmetic policies
// ------------------------------------------------------------------------------------------
template <typename T, bool STRICT>
struct Arith;

// STRICT: the reference's operation order, one rounding per numpy operation, no FMA contraction.
// FAST:   the same expressions with multiply-adds fused (agrees to ~1e-15 relative).
template <>
struct Arith<double, true> {
    static __device__ __forceinline__ double cvt(int n) { return (double)n; }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    // x - c * t:  r = dphi - (2 pi) n  (neighborhood.py:91),  dr = d(dphi) - (2 pi) dn  (:110)
    static __device__ __forceinline__ double resid(double x, double c, double t) { return __dsub_rn(x, __dmul_rn(c, t)); }
    // s = ((kappa/2) dr) ((2 r) + dr)   (neighborhood.py:111)
    static __device__ __forceinline__ double link(double hk, double dr, double r) {
        return __dmul_rn(__dmul_rn(hk, dr), __dadd_rn(__dmul_rn(2.0, r), dr));
    }
    static __device__ __forceinline__ double accept_prob(double dS) { return exp_clipped(-dS); }
    static __device__ __forceinline__ bool metropolis(double dS, double u, double& prob) { prob = exp_clipped(-dS); return u < prob; }
    static __device__ __forceinline__ double twice_plus(double r, double dr) { return fma(2.0, r, dr); }
    static __device__ __forceinline__ double mad(double a, double b, double c) { return fma(a, b, c); }
};
template <>
struct Arith<double, false> {
    static __device__ __forceinline__ double cvt(int n) { return SVB_CVT(n); }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double resid(double x, double c, double t) { return fma(-c, t, x); }
    static __device__ __forceinline__ double link(double hk, double dr, double r) { return (hk * dr) * fma(2.0, r, dr); }
    static __device__ __forceinline__ double accept_prob(double dS) { return exp_clipped(-dS); }
    static __device__ __forceinline__ bool metropolis(double dS, double u, double& prob) { return metropolis_filtered(dS, u, prob); }
    static __device__ __forceinline__ double twice_plus(double r, double dr) { return fma(2.0, r, dr); }
    static __device__ __forceinline__ double mad(double a, double b, double c) { return fma(a, b, c); }
};
template <>
struct Arith<float, true> {
    static __device__ __forceinline__ float cvt(int n) { return (float)n; }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float resid(float x, float c, float t) { return __fsub_rn(x, __fmul_rn(c, t)); }
    static __device__ __forceinline__ float link(float hk, float dr, float r) {
        return __fmul_rn(__fmul_rn(hk, dr), __fadd_rn(__fmul_rn(2.0f, r), dr));
    }
    static __device__ __forceinline__ double accept_prob(float dS) { return fmin((double)expf(-dS), 1.0); }
    static __device__ __forceinline__ bool metropolis(float dS, double u, double& prob) { prob = fmin((double)expf(-dS), 1.0); return u < prob; }
    static __device__ __forceinline__ float twice_plus(float r, float dr) { return fmaf(2.0f, r, dr); }
    static __device__ __forceinline__ float mad(float a, float b, float c) { return fmaf(a, b, c); }
};
template <>
struct Arith<float, false> {
    static __device__ __forceinline__ float cvt(int n) { return (float)n; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float resid(float x, float c, float t) { return fmaf(-c, t, x); }
    static __device__ __forceinline__ float link(float hk, float dr, float r) { return (hk * dr) * fmaf(2.0f, r, dr); }
    static __device__ __forceinline__ double accept_prob(float dS) { return fmin((double)expf(-dS), 1.0); }
    static __device__ __forceinline__ bool metropolis(float dS, double u, double& prob) { prob = fmin((double)expf(-dS), 1.0); return u < prob; }
    static __device__ __forceinline__ float twice_plus(float r, float dr) { return fmaf(2.0f, r, dr); }
    static __device__ __forceinline__ float mad(float a, float b, float c) { return fmaf(a, b, c); }
};

// ------------------------------------------------------------------------------------------
// proposals
// ------------------------------------------------------------------------------------------
// A proposal for one site: the Metropolis uniform, dphi, and the four link changes in UNITS
// dg[i] (links f0, b0, f1, b1); the integer change of n is unit * dg[i] and its residual change is
// c * dg[i], where (unit, c) = (W, fl(2 pi W)) for Philox draws in FAST arithmetic and (1, 2 pi)
// otherwise (then dg already carries the factor W and 2 pi * dn is rounded exactly as numpy does).
struct VillainDraw {
    double u;          // INJECTED: the Metropolis uniform.  Philox: the midpoint (f + 1/2) 2^-32 of the known bracket
    double dphi;
    int dg[4];
    uint32_t f;        // Philox: leading 32 bits of the uniform, u in [f, f + 1] 2^-32; the rest is drawn lazily
    uint32_t c0;       // Philox: counter word 0 of the site's pair
    uint32_t half;     // Philox: which half of the 128-bit block belongs to this site
};

struct VillainConsts {
    double c;        // residual change per unit of dg
    int unit;        // integer change of n per unit of dg
};

// The Philox draw mapping (version 2): 64 bits per site per sweep, and more only when a decision needs them.
//   Sites (x0, x1) and (x0 ^ 8, x1) -- same colour -- share one Philox4x32-10 block with counter word 0
//   c0 = (x0 & ~8) N + x1; the site with bit 3 of x0 clear owns words (0, 1), the other one words (2, 3):
//   word A -> dphi = -I + (2 I) * ((A + 1/2) 2^-32)       [numpy: lo + (hi - lo) * U, multiply then add, no FMA]
//   word B -> the four leading base-K digits (K = 2 interval_n + 1) of the fraction B / 2^32:
//               p = f K;  digit = p >> 32;  f = p mod 2^32;   dg_i = digit_i - interval_n   (links f0, b0, f1, b1)
//             and the remainder f after the fourth digit (uniform on [0, 2^32), independent of the digits) is the
//             LEADING 32 bits of the Metropolis uniform:
//               u = min(fl(f + (e + 1/2) 2^-32) 2^-32, 1 - 2^-53),  e = word `2 half` of the block with the same
//               counter in stream STREAM_VILLAIN_REFINE.
//   u is known to lie in [f, f + 1] 2^-32 without e, which decides u < A unless A falls in that bracket
//   (probability 2^-32 per proposal); only then is e generated.  Every kernel and the oracle implement exactly this
//   rule, so decisions are those of the full 64-bit uniform.
//   The remainder f is uniform, but GIVEN the digits it only takes the values of a lattice of step K^4 (it is K^4 B mod 2^32
//   with B confined to an interval of length 2^32 / K^4): P(u < A | proposal) is quantised in units of K^4 2^-32.  For
//   K = 3 (interval_n = 1, the reference's default) that is 1.9e-8, below the 2^-24 resolution of an fp32 uniform, and
//   the mapping stands.  For K^4 > 256 (interval_n >= 2; "wide") f is NOT taken from word B: the leading 32 bits are word
//   `2 half` of the refinement block and the trailing bits e its word `2 half + 1` -- independent of the proposal.
//   Wide launches are served by the generic kernels only (a second Philox block per site pair).
//   The streams (proposal, refinement) belong to the generator kind: NeighborhoodUpdate (1, 4), SiteUpdate (9, 10),
//   ExactUpdate (11, 12) -- generators that share a seed never share a Philox block.
__host__ __device__ __forceinline__ uint32_t villain_pair_counter(int x0, int x1, int N) { return (uint32_t)((x0 & ~8) * N + x1); }
__host__ __device__ __forceinline__ uint32_t villain_pair_half(int x0) { return (uint32_t)((x0 >> 3) & 1); }

__device__ __forceinline__ double villain_dphi_from_word(uint32_t A, double interval_phi) {
    const double bias = 4503599627370495.5;   // 2^52 - 1/2
    const double U = (__hiloint2double(0x43300000, (int)A) - bias) * 2.3283064365386963e-10;   // (A + 1/2) 2^-32, exact
    return __dadd_rn(-interval_phi, __dmul_rn(2.0 * interval_phi, U));
}

__device__ __forceinline__ VillainDraw villain_draw_from_words(uint32_t A, uint32_t B, double interval_phi, int interval_n) {
    VillainDraw d;
    d.dphi = villain_dphi_from_word(A, interval_phi);
    const uint32_t K = (uint32_t)(2 * interval_n + 1);
    uint32_t f = B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t prod = (uint64_t)f * K;
        f = (uint32_t)prod;
        d.dg[i] = (int)(prod >> 32) - interval_n;
    }
    d.f = f;
    d.u = (__hiloint2double(0x43300000, (int)f) - 4503599627370495.5) * 2.3283064365386963e-10;   // (f + 1/2) 2^-32
    return d;
}

// u < A decided from the bracket [f, f + 1] 2^-32 of u, refining only when A falls inside it (cold path; svb_common.cuh).
__device__ __forceinline__ bool villain_decide_lazy(double A, const VillainDraw& d, const RefineCtx& rc) {
    LazyUniform lu;
    lu.f = d.f; lu.c0 = d.c0; lu.word = 2 * d.half + (rc.wide ? 1u : 0u);
    return decide_lazy(A, lu, rc.stream, rc);
}

// Philox4x32-10 with the key schedule read from kernel parameters (constant bank operands).
__device__ __forceinline__ Philox4 philox_site_keys(const VillainArgs& a, uint64_t chain, uint64_t sweep, uint32_t site) {
    uint32_t c0 = site, c1 = (uint32_t)chain, c2 = (uint32_t)sweep;
    uint32_t c3 = a.stream_hi | ((uint32_t)((chain >> 32) & 0xFFu) << 16) | (uint32_t)((sweep >> 32) & 0xFFFFu);
#pragma unroll
    for (int r = 0; r < SVB_PHILOX_ROUNDS; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c0, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ a.round_key[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ a.round_key[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}

struct SiteOut {
    double A;
    bool ok;
    double dS;
};

// Running sums of the observables that can be collected while the LAST colour of the last sweep is
// processed: every link has exactly one end of that colour (even N), so summing the four links of
// each of its sites covers each link once.
struct LinkSums {
    double r2;     // sum of final residuals squared
    int w0, w1;    // sum of final n_0, n_1
};

// u < min(1, e^-dS) for a Philox draw: the fp32 log filter (FAST fp64 only) on the bracket of u, else the exact test.
template <typename real, bool STRICT>
__device__ __forceinline__ bool villain_metropolis_lazy(real dS, const VillainDraw& d, const RefineCtx& rc, double& prob) {
    if (!STRICT && sizeof(real) == 8) {
        if (d.f >= 65536u) {     // half-width of the bracket relative to u: 2^-33 / u <= 2^-17
            const int r = metropolis_log_filter((double)dS, d.u, 7.62939453125e-06f, prob);
            if (r >= 0) return r != 0;
        }
        prob = exp_clipped(-(double)dS);
        return villain_decide_lazy(prob, d, rc);
    }
    prob = Arith<real, STRICT>::accept_prob(dS);
    return villain_decide_lazy(prob, d, rc);
}

// One Metropolis proposal at site (x0, x1) of one chain, in place.  `phi`, `n0`, `n1` point at the
// chain's fields (shared or global memory).  NT > 0: N == NT is a compile-time power of two.
// LAZY: the draw is a Philox draw whose uniform is known by its leading 32 bits (see villain_decide_lazy).
template <typename real, bool STRICT, int NT, bool LAZY>
__device__ __forceinline__ SiteOut villain_site_update(real* __restrict__ phi, int32_t* __restrict__ n0,
                                                       int32_t* __restrict__ n1, int Nrt, int x0, int x1, real half_kappa,
                                                       const VillainConsts& k, const VillainDraw& d, bool collect,
                                                       LinkSums& sums, const RefineCtx& rc) {
    using A = Arith<real, STRICT>;
    const int N = NT ? NT : Nrt;
    int xp0, xm0, xp1, xm1;
    if (NT) {
        xp0 = (x0 + 1) & (NT - 1); xm0 = (x0 - 1) & (NT - 1);
        xp1 = (x1 + 1) & (NT - 1); xm1 = (x1 - 1) & (NT - 1);
    } else {
        xp0 = (x0 + 1 == N) ? 0 : x0 + 1; xm0 = (x0 == 0) ? N - 1 : x0 - 1;
        xp1 = (x1 + 1 == N) ? 0 : x1 + 1; xm1 = (x1 == 0) ? N - 1 : x1 - 1;
    }
    const int row = x0 * N;
    const int i_c = row + x1;
    const int i_b0 = xm0 * N + x1;
    const int i_b1 = row + xm1;

    const real pc = phi[i_c];
    const real pf0 = phi[xp0 * N + x1];
    const real pb0 = phi[i_b0];
    const real pf1 = phi[row + xp1];
    const real pb1 = phi[i_b1];
    const int nf0 = n0[i_c], nb0 = n0[i_b0], nf1 = n1[i_c], nb1 = n1[i_b1];

    // residuals of the four links, recomputed from the current fields (neighborhood.py:91)
    const real two_pi = (real)SVB_TWO_PI;
    const real r_f0 = A::resid(A::sub(pf0, pc), two_pi, A::cvt(nf0));
    const real r_b0 = A::resid(A::sub(pc, pb0), two_pi, A::cvt(nb0));
    const real r_f1 = A::resid(A::sub(pf1, pc), two_pi, A::cvt(nf1));
    const real r_b1 = A::resid(A::sub(pc, pb1), two_pi, A::cvt(nb1));

    // change of the residuals (neighborhood.py:110): d(dphi) is -dphi on forward links, +dphi on backward links
    const real dphi = (real)d.dphi;
    const real c = (real)k.c;
    const real dr_f0 = A::resid(-dphi, c, A::cvt(d.dg[0]));
    const real dr_b0 = A::resid(dphi, c, A::cvt(d.dg[1]));
    const real dr_f1 = A::resid(-dphi, c, A::cvt(d.dg[2]));
    const real dr_b1 = A::resid(dphi, c, A::cvt(d.dg[3]));

    real dS;
    if (STRICT) {
        // dS in the reference's face_sum order (neighborhood.py:111-112; lattice/_kernels.py:37-45)
        dS = A::link(half_kappa, dr_f0, r_f0);
        dS = A::add(dS, A::link(half_kappa, dr_b0, r_b0));
        dS = A::add(dS, A::link(half_kappa, dr_f1, r_f1));
        dS = A::add(dS, A::link(half_kappa, dr_b1, r_b1));
    } else {
        // same sum with kappa/2 factored out and the multiply-adds fused: 8 FMAs + 1 multiply
        real acc2 = dr_f0 * A::twice_plus(r_f0, dr_f0);
        acc2 = A::mad(dr_b0, A::twice_plus(r_b0, dr_b0), acc2);
        acc2 = A::mad(dr_f1, A::twice_plus(r_f1, dr_f1), acc2);
        acc2 = A::mad(dr_b1, A::twice_plus(r_b1, dr_b1), acc2);
        dS = half_kappa * acc2;
    }

    double acc;                                                 // clip(exp(-dS), 0, 1)   (:115)
    const bool ok = LAZY ? villain_metropolis_lazy<real, STRICT>(dS, d, rc, acc)
                         : A::metropolis(dS, d.u, acc);         // u < acc                (:116)
    if (ok) {                                                   // (:121-128)
        phi[i_c] = A::add(pc, dphi);
        n0[i_c] = nf0 + k.unit * d.dg[0];
        n0[i_b0] = nb0 + k.unit * d.dg[1];
        n1[i_c] = nf1 + k.unit * d.dg[2];
        n1[i_b1] = nb1 + k.unit * d.dg[3];
    }
    if (collect) {
        const double s = ok ? 1.0 : 0.0;
        const double q0 = fma(s, (double)dr_f0, (double)r_f0), q1 = fma(s, (double)dr_b0, (double)r_b0);
        const double q2 = fma(s, (double)dr_f1, (double)r_f1), q3 = fma(s, (double)dr_b1, (double)r_b1);
        sums.r2 = fma(q0, q0, sums.r2);
        sums.r2 = fma(q1, q1, sums.r2);
        sums.r2 = fma(q2, q2, sums.r2);
        sums.r2 = fma(q3, q3, sums.r2);
        const int m = ok ? k.unit : 0;
        sums.w0 += nf0 + nb0 + m * (d.dg[0] + d.dg[1]);
        sums.w1 += nf1 + nb1 + m * (d.dg[2] + d.dg[3]);
    }
    SiteOut o;
    o.A = acc;
    o.ok = ok;
    o.dS = (double)dS;
    return o;
}

template <bool INJECTED, bool KEYS>
__device__ __forceinline__ VillainDraw villain_get_draw(const VillainArgs& a, long long chain, int sweep, int x0, int x1,
                                                        int site, int dg_scale) {
    if (INJECTED) {
        const long long V = (long long)a.N * a.N;
        const long long base = ((long long)sweep * a.chains + chain) * V + site;
        const long long lbase = ((long long)sweep * a.chains + chain) * 2 * V + site;
        VillainDraw d;
        d.u = a.inj_u[base];
        d.f = 0; d.c0 = 0; d.half = 0;
        if (a.exact_mode) {
            // ExactUpdate (exact.py:94-102): z on the proposing site alone, n += d z: forward links -z, backward links +z
            const int z = a.inj_z[base];
            d.dphi = 0.0;
            d.dg[0] = -z; d.dg[1] = z; d.dg[2] = -z; d.dg[3] = z;
            return d;
        }
        d.dphi = a.inj_dphi[base];
        d.dg[0] = a.inj_dn_fwd ? a.inj_dn_fwd[lbase] : 0;          // SiteUpdate: no dn proposals at all
        d.dg[1] = a.inj_dn_bwd ? a.inj_dn_bwd[lbase] : 0;
        d.dg[2] = a.inj_dn_fwd ? a.inj_dn_fwd[lbase + V] : 0;
        d.dg[3] = a.inj_dn_bwd ? a.inj_dn_bwd[lbase + V] : 0;
        return d;
    } else {
        const unsigned long long gc = a.chain0 + (unsigned long long)chain, gs = a.sweep0 + (unsigned long long)sweep;
        const uint32_t c0 = villain_pair_counter(x0, x1, a.N), half = villain_pair_half(x0);
        const Philox4 p = KEYS ? philox_site_keys(a, gc, gs, c0) : philox_site(a.seed, gc, gs, c0, a.stream_hi >> 24);
        if (a.exact_mode) {
            // word B: z = one of the 2 I nonzero values (exact.py:38, :94), the remainder leads the uniform; word A unused
            const uint64_t pz = (uint64_t)(half ? p.w : p.y) * (uint64_t)(2 * a.interval_n);
            const int idx = (int)(pz >> 32);
            const int z = (idx < a.interval_n) ? idx - a.interval_n : idx - a.interval_n + 1;
            VillainDraw d;
            d.dphi = 0.0;
            d.dg[0] = -z; d.dg[1] = z; d.dg[2] = -z; d.dg[3] = z;
            d.f = (uint32_t)pz;
            d.u = (__hiloint2double(0x43300000, (int)d.f) - 4503599627370495.5) * 2.3283064365386963e-10;
            d.c0 = c0; d.half = half;
            return d;
        }
        VillainDraw d = villain_draw_from_words(half ? p.z : p.x, half ? p.w : p.y, a.interval_phi, a.interval_n);
        d.c0 = c0; d.half = half;
        if (a.wide) {
            // wide dn intervals: the uniform's leading bits must not be the digits' remainder (mapping above)
            const Philox4 r = philox_site(a.seed, gc, gs, c0, a.refine_stream);
            d.f = half ? r.z : r.x;
            d.u = (__hiloint2double(0x43300000, (int)d.f) - 4503599627370495.5) * 2.3283064365386963e-10;
        }
        if (dg_scale != 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) d.dg[i] *= dg_scale;
        }
        return d;
    }
}

__device__ __forceinline__ RefineCtx villain_refine_ctx(const VillainArgs& a, long long chain, int sweep) {
    RefineCtx rc;
    rc.seed = a.seed; rc.chain = a.chain0 + (unsigned long long)chain; rc.sweep = a.sweep0 + (unsigned long long)sweep;
    rc.stream = a.refine_stream; rc.wide = (uint32_t)a.wide;
    return rc;
}

// (unit, c, dg_scale) for a launch: FAST Philox keeps dg in units of W; everything else folds W into dg
template <bool INJECTED, bool STRICT>
__device__ __forceinline__ void villain_consts(const VillainArgs& a, VillainConsts& k, int& dg_scale) {
    if (a.exact_mode) {
        k.unit = 1;
        k.c = SVB_TWO_PI;
        dg_scale = 1;
    } else if (!INJECTED && !STRICT) {
        k.unit = a.W;
        k.c = SVB_TWO_PI * (double)a.W;
        dg_scale = 1;
    } else {
        k.unit = 1;
        k.c = SVB_TWO_PI;
        dg_scale = INJECTED ? 1 : a.W;
    }
}

// Observables of one site's forward links and plaquette (full pass; used for odd N and by the obs kernel).
template <typename real, int NT>
__device__ __forceinline__ void villain_obs_site(const real* __restrict__ phi, const int32_t* __restrict__ n0,
                                                 const int32_t* __restrict__ n1, int Nrt, int x0, int x1, double& s_action,
                                                 long long& i_dn2, int& i_w0, int& i_w1) {
    const int N = NT ? NT : Nrt;
    int xp0, xp1;
    if (NT) {
        xp0 = (x0 + 1) & (NT - 1); xp1 = (x1 + 1) & (NT - 1);
    } else {
        xp0 = (x0 + 1 == N) ? 0 : x0 + 1; xp1 = (x1 + 1 == N) ? 0 : x1 + 1;
    }
    const int i = x0 * N + x1, i0 = xp0 * N + x1, i1 = x0 * N + xp1;
    const double pc = (double)phi[i];
    const int a0 = n0[i], a1 = n1[i];
    const double r0 = __dsub_rn(__dsub_rn((double)phi[i0], pc), __dmul_rn(SVB_TWO_PI, (double)a0));
    const double r1 = __dsub_rn(__dsub_rn((double)phi[i1], pc), __dmul_rn(SVB_TWO_PI, (double)a1));
    s_action = fma(r0, r0, s_action);
    s_action = fma(r1, r1, s_action);
    // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x])      (compact.py d,1 rows)
    const int dn = (n1[i0] - a1) - (n0[i1] - a0);
    i_dn2 += (long long)dn * dn;
    i_w0 += a0;
    i_w1 += a1;
}

// (dn)^2 of the plaquette based at (x0, x1): the only observable that needs a pass after the last colour.
template <int NT>
__device__ __forceinline__ long long villain_dn2_site(const int32_t* __restrict__ n0, const int32_t* __restrict__ n1, int Nrt,
                                                      int x0, int x1) {
    const int N = NT ? NT : Nrt;
    int xp0, xp1;
    if (NT) {
        xp0 = (x0 + 1) & (NT - 1); xp1 = (x1 + 1) & (NT - 1);
    } else {
        xp0 = (x0 + 1 == N) ? 0 : x0 + 1; xp1 = (x1 + 1 == N) ? 0 : x1 + 1;
    }
    const int i = x0 * N + x1;
    const int dn = (n1[xp0 * N + x1] - n1[i]) - (n0[x0 * N + xp1] - n0[i]);
    return (long long)dn * dn;
}

// Block reduction of the per-chain record.  Doubles go through shuffles, integers through REDUX.
struct ChainSums {
    double action, sumA;
    long long dn2;
    int w0, w1, accepted;
};

__device__ __forceinline__ ChainSums block_reduce_chain(ChainSums s, double* scratch /* 6*32 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    s.action = warp_sum(s.action);
    s.sumA = warp_sum(s.sumA);
    unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(s.dn2 & 0xFFFFFF));
    unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)((unsigned long long)s.dn2 >> 24));
    s.dn2 = (long long)lo + ((long long)hi << 24);
    s.w0 = __reduce_add_sync(0xffffffffu, s.w0);
    s.w1 = __reduce_add_sync(0xffffffffu, s.w1);
    s.accepted = __reduce_add_sync(0xffffffffu, s.accepted);
    if (nwarps == 1) return s;
    long long* iscr = reinterpret_cast<long long*>(scratch + 2 * 32);
    __syncthreads();
    if (lane == 0) {
        scratch[warp] = s.action;
        scratch[32 + warp] = s.sumA;
        iscr[warp] = s.dn2;
        iscr[32 + warp] = (long long)s.w0;
        iscr[64 + warp] = (long long)s.w1;
        iscr[96 + warp] = (long long)s.accepted;
    }
    __syncthreads();
    if (warp == 0) {
        const bool in = lane < nwarps;
        s.action = warp_sum(in ? scratch[lane] : 0.0);
        s.sumA = warp_sum(in ? scratch[32 + lane] : 0.0);
        s.dn2 = warp_sum(in ? iscr[lane] : 0LL);
        s.w0 = (int)warp_sum(in ? iscr[32 + lane] : 0LL);
        s.w1 = (int)warp_sum(in ? iscr[64 + lane] : 0LL);
        s.accepted = (int)warp_sum(in ? iscr[96 + lane] : 0LL);
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// SMEM path: one CTA per chain (grid-stride over chains), whole lattice in shared memory.
// NT/TT > 0 fix the lattice size (power of two) and the block size at compile time; MINB is the
// occupancy the register allocation is held to.
// ------------------------------------------------------------------------------------------
template <typename real, bool INJECTED, bool STRICT, int NT, int TT, int MINB>
__global__ void __launch_bounds__(TT ? TT : 256, MINB) villain_smem_kernel(const __grid_constant__ VillainArgs a, int use_bulk) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int N = NT ? NT : a.N, V = N * N;
    const int tid = threadIdx.x, T = TT ? TT : (int)blockDim.x;
    const size_t bytes_phi = (size_t)V * sizeof(real);
    const size_t bytes_n = (size_t)2 * V * sizeof(int32_t);
    const size_t off_n = (bytes_phi + 15) & ~(size_t)15;
    const size_t off_scr = (off_n + bytes_n + 15) & ~(size_t)15;
    real* sphi = reinterpret_cast<real*>(smem_raw);
    int32_t* sn0 = reinterpret_cast<int32_t*>(smem_raw + off_n);
    int32_t* sn1 = sn0 + V;
    double* scratch = reinterpret_cast<double*>(smem_raw + off_scr);   // 6 * 32 doubles
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + off_scr + 6 * 32 * sizeof(double));

    if (use_bulk) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    uint32_t phase = 0;
    const int ncol = n_colours(N);
    const int halfN = N >> 1, nhalf = V >> 1;
    VillainConsts kc;
    int dg_scale;
    villain_consts<INJECTED, STRICT>(a, kc, dg_scale);
    const bool fuse_obs = (a.obs != nullptr) && (ncol == 2);

    for (long long chain = blockIdx.x; chain < a.chains; chain += gridDim.x) {
        real* gphi = reinterpret_cast<real*>(a.phi) + chain * V;
        int32_t* gn = a.n + chain * 2 * V;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const real half_kappa = (real)(kappa / 2);

        // ---- load the chain: HBM -> shared ----
        if (use_bulk) {
            if (tid == 0) {
                mbar_expect_tx(bar, (uint32_t)(bytes_phi + bytes_n));
                bulk_g2s(sphi, gphi, (uint32_t)bytes_phi, bar);
                bulk_g2s(sn0, gn, (uint32_t)bytes_n, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            for (int i = tid; i < V; i += T) sphi[i] = gphi[i];
            for (int i = tid; i < 2 * V; i += T) sn0[i] = gn[i];
            __syncthreads();
        }

        int n_acc = 0;
        double sum_A = 0.0;
        LinkSums ls;
        ls.r2 = 0.0; ls.w0 = 0; ls.w1 = 0;
        for (int s = 0; s < a.n_sweeps; ++s) {
            const bool last = (s == a.n_sweeps - 1);
            const bool debug = last && (a.accept_mask != nullptr || a.dS_out != nullptr);
            for (int c = 0; c < ncol; ++c) {
                if (ncol == 2) {
                    const bool collect = fuse_obs && last && (c == 1);
#pragma unroll kSiteUnroll
                    for (int j = tid; j < nhalf; j += T) {
                        const int x0 = j / halfN;
                        const int x1 = 2 * (j - x0 * halfN) + ((x0 + c) & 1);
                        const int site = x0 * N + x1;
                        const VillainDraw d = villain_get_draw<INJECTED, true>(a, chain, s, x0, x1, site, dg_scale);
                        const SiteOut o = villain_site_update<real, STRICT, NT, !INJECTED>(sphi, sn0, sn1, N, x0, x1, half_kappa, kc, d,
                                                                                           collect, ls, villain_refine_ctx(a, chain, s));
                        n_acc += o.ok ? 1 : 0;
                        sum_A += o.A;
                        if (debug) {
                            if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                            if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
                        }
                    }
                } else {
                    for (int site = tid; site < V; site += T) {
                        const int x0 = site / N, x1 = site - x0 * N;
                        if (site_colour(x0, x1, N) != c) continue;
                        const VillainDraw d = villain_get_draw<INJECTED, true>(a, chain, s, x0, x1, site, dg_scale);
                        const SiteOut o = villain_site_update<real, STRICT, NT, !INJECTED>(sphi, sn0, sn1, N, x0, x1, half_kappa, kc, d,
                                                                                           false, ls, villain_refine_ctx(a, chain, s));
                        n_acc += o.ok ? 1 : 0;
                        sum_A += o.A;
                        if (debug) {
                            if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                            if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
                        }
                    }
                }
                __syncthreads();
            }
        }

        // ---- observables of the final state ----
        if (a.obs) {
            ChainSums cs;
            cs.sumA = sum_A; cs.accepted = n_acc; cs.dn2 = 0;
            if (fuse_obs) {
                // action and wrapping were collected link by link during the last colour pass
                cs.action = ls.r2; cs.w0 = ls.w0; cs.w1 = ls.w1;
#pragma unroll
                for (int i = tid; i < V; i += T) {
                    const int x0 = i / N, x1 = i - x0 * N;
                    cs.dn2 += villain_dn2_site<NT>(sn0, sn1, N, x0, x1);
                }
            } else {
                cs.action = 0.0; cs.w0 = 0; cs.w1 = 0;
                for (int i = tid; i < V; i += T) {
                    const int x0 = i / N, x1 = i - x0 * N;
                    villain_obs_site<real, NT>(sphi, sn0, sn1, N, x0, x1, cs.action, cs.dn2, cs.w0, cs.w1);
                }
            }
            cs = block_reduce_chain(cs, scratch);
            if (tid == 0) {
                double* o = a.obs + chain * SVB_VOBS_COUNT;
                o[SVB_VOBS_ACTION] = (kappa / 2) * cs.action;
                o[SVB_VOBS_SUM_DN2] = (double)cs.dn2;
                o[SVB_VOBS_WRAP0] = (double)cs.w0;
                o[SVB_VOBS_WRAP1] = (double)cs.w1;
                o[SVB_VOBS_ACCEPTED] = (double)cs.accepted;
                o[SVB_VOBS_ACCEPTANCE] = cs.sumA;
            }
        }

        // ---- store the chain: shared -> HBM ----
        if (use_bulk) {
            fence_proxy_async();   // make this thread's generic-proxy smem writes visible to the bulk copy engine
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(gphi, sphi, (uint32_t)bytes_phi);
                bulk_s2g(gn, sn0, (uint32_t)bytes_n);
                bulk_commit();
                bulk_wait_read0();  // smem may be overwritten by the next chain's load after this
            }
            __syncthreads();
        } else {
            for (int i = tid; i < V; i += T) gphi[i] = sphi[i];
            for (int i = tid; i < 2 * V; i += T) gn[i] = sn0[i];
            __syncthreads();
        }
    }
}

// Sum of (dn)^2 over the plaquettes of PER consecutive sites of one row, with 128-bit shared loads.
template <int NT, int TT>
__device__ __forceinline__ long long villain_dn2_rows(const int32_t* __restrict__ n0, const int32_t* __restrict__ n1, int tid) {
    constexpr int PER = NT * NT / TT;           // consecutive plaquettes per thread (a multiple of 4)
    constexpr int SEGS = NT / PER;               // such segments per row
    static_assert(PER % 4 == 0 && NT % PER == 0, "villain_dn2_rows: unsupported geometry");
    const int x0 = tid / SEGS, seg = (tid % SEGS) * PER;
    const int xp0 = (x0 + 1) & (NT - 1);
    const int4* up = reinterpret_cast<const int4*>(n1 + xp0 * NT + seg);   // n1[x + e0]
    const int4* me = reinterpret_cast<const int4*>(n1 + x0 * NT + seg);    // n1[x]
    const int4* ho = reinterpret_cast<const int4*>(n0 + x0 * NT + seg);    // n0[x], n0[x + e1]
    const int wrap = n0[x0 * NT + ((seg + PER) & (NT - 1))];
    long long acc = 0;
    int4 C = ho[0];
#pragma unroll
    for (int q = 0; q < PER / 4; ++q) {
        const int4 A = up[q], B = me[q];
        const int4 Cn = (q + 1 < PER / 4) ? ho[q + 1] : make_int4(wrap, 0, 0, 0);
        // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x])      (compact.py d,1 rows)
        const int d0 = (A.x - B.x) - (C.y - C.x), d1 = (A.y - B.y) - (C.z - C.y);
        const int d2 = (A.z - B.z) - (C.w - C.z), d3 = (A.w - B.w) - (Cn.x - C.w);
        acc += (long long)d0 * d0;
        acc += (long long)d1 * d1;
        acc += (long long)d2 * d2;
        acc += (long long)d3 * d3;
        C = Cn;
    }
    return acc;
}

// ------------------------------------------------------------------------------------------
// SMEM path, production instantiation: Philox draws, compile-time geometry, and a two-stage
// TMA pipeline per CTA.  While a CTA sweeps the chain in one shared-memory buffer, the bulk load
// of its next chain lands in the other buffer and the bulk store of the previous chain drains,
// so HBM latency never sits on the CTA's critical path.
// ------------------------------------------------------------------------------------------
template <typename real, bool STRICT, int NT, int TT, int MINB>
__global__ void __launch_bounds__(TT, MINB) villain_smem_pipelined_kernel(const __grid_constant__ VillainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = NT * NT, T = TT;
    constexpr int halfN = N / 2, nhalf = V / 2;
    constexpr uint32_t bytes_phi = V * sizeof(real);
    constexpr uint32_t bytes_n = 2 * V * sizeof(int32_t);
    constexpr uint32_t stage_bytes = bytes_phi + bytes_n;          // a multiple of 16
    const int tid = threadIdx.x;
    double* scratch = reinterpret_cast<double*>(smem_raw + 2 * stage_bytes);   // 6 * 32 doubles
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * stage_bytes + 6 * 32 * sizeof(double));

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    VillainConsts kc;
    int dg_scale;
    villain_consts<false, STRICT>(a, kc, dg_scale);
    const bool want_obs = a.obs != nullptr;

    auto issue_load = [&](long long chain, int b) {
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        mbar_expect_tx(&bar[b], stage_bytes);
        bulk_g2s(stage, reinterpret_cast<const real*>(a.phi) + chain * V, bytes_phi, &bar[b]);
        bulk_g2s(stage + bytes_phi, a.n + chain * 2 * V, bytes_n, &bar[b]);
    };

    long long chain = blockIdx.x;
    if (tid == 0 && chain < a.chains) issue_load(chain, 0);

    for (int it = 0; chain < a.chains; chain += gridDim.x, ++it) {
        const int b = it & 1;
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        real* sphi = reinterpret_cast<real*>(stage);
        int32_t* sn0 = reinterpret_cast<int32_t*>(stage + bytes_phi);
        int32_t* sn1 = sn0 + V;
        const long long next = chain + gridDim.x;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const real half_kappa = (real)(kappa / 2);

        mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));

        int n_acc = 0;
        double sum_A = 0.0;
        LinkSums ls;
        ls.r2 = 0.0; ls.w0 = 0; ls.w1 = 0;
        for (int s = 0; s < a.n_sweeps; ++s) {
            const bool last = (s == a.n_sweeps - 1);
            const bool debug = last && (a.accept_mask != nullptr || a.dS_out != nullptr);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const bool collect = want_obs && last && (c == 1);
#pragma unroll kSiteUnroll
                for (int j = tid; j < nhalf; j += T) {
                    const int x0 = j / halfN;
                    const int x1 = 2 * (j - x0 * halfN) + ((x0 + c) & 1);
                    const int site = x0 * N + x1;
                    const VillainDraw d = villain_get_draw<false, true>(a, chain, s, x0, x1, site, dg_scale);
                    const SiteOut o = villain_site_update<real, STRICT, NT, true>(sphi, sn0, sn1, N, x0, x1, half_kappa, kc, d,
                                                                                  collect, ls, villain_refine_ctx(a, chain, s));
                    n_acc += o.ok ? 1 : 0;
                    sum_A += o.A;
                    if (debug) {
                        if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                        if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
                    }
                }
                __syncthreads();
                if (s == 0 && c == 0 && tid == 0 && next < a.chains) {
                    // The other buffer was stored from at the end of the previous iteration; once that bulk
                    // store has finished reading shared memory the next chain can be fetched into it.
                    bulk_wait_read0();
                    issue_load(next, b ^ 1);
                }
            }
        }

        if (want_obs) {
            ChainSums cs;
            cs.sumA = sum_A; cs.accepted = n_acc;
            cs.action = ls.r2; cs.w0 = ls.w0; cs.w1 = ls.w1;      // collected during the last colour pass
            cs.dn2 = villain_dn2_rows<NT, TT>(sn0, sn1, tid);
            cs = block_reduce_chain(cs, scratch);
            if (tid == 0) {
                double* o = a.obs + chain * SVB_VOBS_COUNT;
                o[SVB_VOBS_ACTION] = (kappa / 2) * cs.action;
                o[SVB_VOBS_SUM_DN2] = (double)cs.dn2;
                o[SVB_VOBS_WRAP0] = (double)cs.w0;
                o[SVB_VOBS_WRAP1] = (double)cs.w1;
                o[SVB_VOBS_ACCEPTED] = (double)cs.accepted;
                o[SVB_VOBS_ACCEPTANCE] = cs.sumA;
            }
        }

        fence_proxy_async();   // generic-proxy writes of this thread -> visible to the bulk copy engine
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(reinterpret_cast<real*>(a.phi) + chain * V, sphi, bytes_phi);
            bulk_s2g(a.n + chain * 2 * V, sn0, bytes_n);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait0();   // all stores complete before the CTA (and its shared memory) retires
}

// ------------------------------------------------------------------------------------------
// SMEM path with the residuals resident (production instantiation, Philox draws).
//
// The reference keeps r = d(phi) - 2 pi n as a field and patches it after every accepted colour
// (neighborhood.py:91, :129).  Doing the same in shared memory removes, from every proposal, the five
// phi loads, four n loads, four int->double conversions and eight fp64 operations that recomputing r
// costs: a proposal reads four doubles of r, and phi / n are touched only when it is accepted.  In STRICT
// arithmetic r is rebuilt at the start of every sweep and patched as (r + d(dphi)) - 2 pi dn, exactly the
// reference's roundings, so dS is bit-equal to the reference; in FAST arithmetic r is carried across the
// sweeps of a launch and patched with one add.  2 pi dn comes from a K-entry table in shared memory
// (K = 2 interval_n + 1 distinct proposals), so no integer is converted to fp64 on the hot path.
// ------------------------------------------------------------------------------------------
struct VillainDigits {
    VillainDraw d;     // u bracket, dphi; d.dg[] unused
    int digit[4];      // proposal index in [0, K) for links f0, b0, f1, b1;  dn = W * (digit - interval_n)
};

__device__ __forceinline__ VillainDigits villain_digits_from_words(uint32_t A, uint32_t B, double interval_phi, uint32_t K) {
    VillainDigits d;
    d.d.dphi = villain_dphi_from_word(A, interval_phi);
    uint32_t f = B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t prod = (uint64_t)f * K;
        f = (uint32_t)prod;
        d.digit[i] = (int)(prod >> 32);
    }
    d.d.f = f;
    d.d.u = (__hiloint2double(0x43300000, (int)f) - 4503599627370495.5) * 2.3283064365386963e-10;
    return d;
}

template <bool STRICT, int NT>
__device__ __forceinline__ SiteOut villain_site_update_resid(double* __restrict__ phi, int32_t* __restrict__ n0,
                                                             int32_t* __restrict__ n1, double* __restrict__ r0,
                                                             double* __restrict__ r1, const double* __restrict__ tpdn, int x0,
                                                             int x1, double half_kappa, int W, int interval_n,
                                                             const VillainDigits& d, const RefineCtx& rc) {
    using A = Arith<double, STRICT>;
    const int i_c = x0 * NT + x1;
    const int i_b0 = ((x0 - 1) & (NT - 1)) * NT + x1;
    const int i_b1 = x0 * NT + ((x1 - 1) & (NT - 1));
    const double r_f0 = r0[i_c], r_b0 = r0[i_b0], r_f1 = r1[i_c], r_b1 = r1[i_b1];
    const double t0 = tpdn[d.digit[0]], t1 = tpdn[d.digit[1]], t2 = tpdn[d.digit[2]], t3 = tpdn[d.digit[3]];   // 2 pi dn
    const double dphi = d.d.dphi;
    // dr = d(dphi) - 2 pi dn   (neighborhood.py:110)
    const double dr_f0 = A::sub(-dphi, t0), dr_b0 = A::sub(dphi, t1), dr_f1 = A::sub(-dphi, t2), dr_b1 = A::sub(dphi, t3);
    double dS;
    if (STRICT) {
        dS = A::link(half_kappa, dr_f0, r_f0);
        dS = A::add(dS, A::link(half_kappa, dr_b0, r_b0));
        dS = A::add(dS, A::link(half_kappa, dr_f1, r_f1));
        dS = A::add(dS, A::link(half_kappa, dr_b1, r_b1));
    } else {
        double acc2 = dr_f0 * fma(2.0, r_f0, dr_f0);
        acc2 = fma(dr_b0, fma(2.0, r_b0, dr_b0), acc2);
        acc2 = fma(dr_f1, fma(2.0, r_f1, dr_f1), acc2);
        acc2 = fma(dr_b1, fma(2.0, r_b1, dr_b1), acc2);
        dS = half_kappa * acc2;
    }
    double acc;
    const bool ok = villain_metropolis_lazy<double, STRICT>(dS, d.d, rc, acc);
    if (ok) {
        phi[i_c] = A::add(phi[i_c], dphi);
        n0[i_c] += W * (d.digit[0] - interval_n);
        n0[i_b0] += W * (d.digit[1] - interval_n);
        n1[i_c] += W * (d.digit[2] - interval_n);
        n1[i_b1] += W * (d.digit[3] - interval_n);
        if (STRICT) {       // r = r + d(dphi) - 2 pi dn, left to right   (neighborhood.py:129)
            r0[i_c] = __dsub_rn(__dadd_rn(r_f0, -dphi), t0);
            r0[i_b0] = __dsub_rn(__dadd_rn(r_b0, dphi), t1);
            r1[i_c] = __dsub_rn(__dadd_rn(r_f1, -dphi), t2);
            r1[i_b1] = __dsub_rn(__dadd_rn(r_b1, dphi), t3);
        } else {
            r0[i_c] = r_f0 + dr_f0;
            r0[i_b0] = r_b0 + dr_b0;
            r1[i_c] = r_f1 + dr_f1;
            r1[i_b1] = r_b1 + dr_b1;
        }
    }
    SiteOut o;
    o.A = acc;
    o.ok = ok;
    o.dS = dS;
    return o;
}

template <bool STRICT, int NT, int TT, int MINB, int STAGES>
__global__ void __launch_bounds__(TT, MINB) villain_smem_resid_kernel(const __grid_constant__ VillainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = NT * NT, T = TT;
    constexpr int halfN = N / 2, nhalf = V / 2;
    constexpr uint32_t bytes_phi = V * sizeof(double);
    constexpr uint32_t bytes_n = 2 * V * sizeof(int32_t);
    constexpr uint32_t stage_bytes = bytes_phi + bytes_n;
    static_assert(V % (2 * T) == 0, "unsupported geometry");
    const int tid = threadIdx.x;
    double* sr0 = reinterpret_cast<double*>(smem_raw + STAGES * stage_bytes);
    double* sr1 = sr0 + V;
    double* tpdn = sr1 + V;                                                   // 64 doubles: 2 pi W (k - interval_n)
    double* scratch = tpdn + 64;                                              // 6 * 32 doubles
    uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + 6 * 32);

    if (tid == 0) {
        for (int b = 0; b < STAGES; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
    }
    const uint32_t K = (uint32_t)(2 * a.interval_n + 1);
    if (tid < (int)K) tpdn[tid] = __dmul_rn(SVB_TWO_PI, (double)(a.W * (tid - a.interval_n)));   // numpy: 2*np.pi*change_n
    __syncthreads();
    const bool want_obs = a.obs != nullptr;

    auto issue_load = [&](long long chain, int b) {
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        mbar_expect_tx(&bar[b], stage_bytes);
        bulk_g2s(stage, reinterpret_cast<const double*>(a.phi) + chain * V, bytes_phi, &bar[b]);
        bulk_g2s(stage + bytes_phi, a.n + chain * 2 * V, bytes_n, &bar[b]);
    };

    long long chain = blockIdx.x;
    if (tid == 0 && chain < a.chains) issue_load(chain, 0);

    for (int it = 0; chain < a.chains; chain += gridDim.x, ++it) {
        const int b = (STAGES == 2) ? (it & 1) : 0;
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        double* sphi = reinterpret_cast<double*>(stage);
        int32_t* sn0 = reinterpret_cast<int32_t*>(stage + bytes_phi);
        int32_t* sn1 = sn0 + V;
        const long long next = chain + gridDim.x;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double half_kappa = kappa / 2;

        mbar_wait(&bar[b], (uint32_t)((STAGES == 2 ? (it >> 1) : it) & 1));

        int n_acc = 0;
        double sum_A = 0.0;
        for (int s = 0; s < a.n_sweeps; ++s) {
            const bool last = (s == a.n_sweeps - 1);
            const bool debug = last && (a.accept_mask != nullptr || a.dS_out != nullptr);
            if (s == 0 || STRICT) {
                // r = d(phi) - 2 pi n   (neighborhood.py:91), two horizontally adjacent sites per thread per step:
                // consecutive threads touch consecutive 16-byte pairs, so every access is conflict-free
#pragma unroll
                for (int q = 0; q < V / (2 * T); ++q) {
                    const int i = 2 * (tid + T * q);                       // sites i, i + 1 (same row)
                    const int x0 = i / N, x1 = i - x0 * N;
                    const int iup = ((x0 + 1) & (N - 1)) * N + x1;
                    const double2 pc = *reinterpret_cast<const double2*>(sphi + i);
                    const double2 pu = *reinterpret_cast<const double2*>(sphi + iup);
                    const double pr = sphi[x0 * N + ((x1 + 2) & (N - 1))];
                    const int2 a0 = *reinterpret_cast<const int2*>(sn0 + i);
                    const int2 a1 = *reinterpret_cast<const int2*>(sn1 + i);
                    double2 o0, o1;
                    o0.x = __dsub_rn(__dsub_rn(pu.x, pc.x), __dmul_rn(SVB_TWO_PI, int_to_double(a0.x)));
                    o0.y = __dsub_rn(__dsub_rn(pu.y, pc.y), __dmul_rn(SVB_TWO_PI, int_to_double(a0.y)));
                    o1.x = __dsub_rn(__dsub_rn(pc.y, pc.x), __dmul_rn(SVB_TWO_PI, int_to_double(a1.x)));
                    o1.y = __dsub_rn(__dsub_rn(pr, pc.y), __dmul_rn(SVB_TWO_PI, int_to_double(a1.y)));
                    *reinterpret_cast<double2*>(sr0 + i) = o0;
                    *reinterpret_cast<double2*>(sr1 + i) = o1;
                }
                __syncthreads();
            }
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
#pragma unroll kSiteUnroll
                for (int j = tid; j < nhalf; j += T) {
                    const int x0 = j / halfN;
                    const int x1 = 2 * (j - x0 * halfN) + ((x0 + c) & 1);
                    const int site = x0 * N + x1;
                    const uint32_t pc0 = villain_pair_counter(x0, x1, N), phalf = villain_pair_half(x0);
                    const Philox4 bits = philox_site_keys(a, a.chain0 + (unsigned long long)chain,
                                                          a.sweep0 + (unsigned long long)s, pc0);
                    VillainDigits d = villain_digits_from_words(phalf ? bits.z : bits.x, phalf ? bits.w : bits.y, a.interval_phi, K);
                    d.d.c0 = pc0; d.d.half = phalf;
                    const SiteOut o = villain_site_update_resid<STRICT, NT>(sphi, sn0, sn1, sr0, sr1, tpdn, x0, x1, half_kappa,
                                                                            a.W, a.interval_n, d, villain_refine_ctx(a, chain, s));
                    n_acc += o.ok ? 1 : 0;
                    sum_A += o.A;
                    if (debug) {
                        if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                        if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
                    }
                }
                __syncthreads();
                if (STAGES == 2 && s == 0 && c == 0 && tid == 0 && next < a.chains) {
                    bulk_wait_read0();
                    issue_load(next, b ^ 1);
                }
            }
        }

        if (want_obs) {
            // one pass over site pairs: sum r^2 (the residuals are resident), sum n, sum (dn)^2; conflict-free vector loads
            ChainSums cs;
            cs.sumA = sum_A; cs.accepted = n_acc; cs.action = 0.0; cs.w0 = 0; cs.w1 = 0;
            long long dn2 = 0;
#pragma unroll
            for (int q = 0; q < V / (2 * T); ++q) {
                const int i = 2 * (tid + T * q);
                const int x0 = i / N, x1 = i - x0 * N;
                const int iup = ((x0 + 1) & (N - 1)) * N + x1;
                const double2 A0 = *reinterpret_cast<const double2*>(sr0 + i);
                const double2 A1 = *reinterpret_cast<const double2*>(sr1 + i);
                cs.action = fma(A0.x, A0.x, cs.action);
                cs.action = fma(A0.y, A0.y, cs.action);
                cs.action = fma(A1.x, A1.x, cs.action);
                cs.action = fma(A1.y, A1.y, cs.action);
                const int2 h = *reinterpret_cast<const int2*>(sn0 + i);          // n0[x], n0[x + e1]
                const int hr = sn0[x0 * N + ((x1 + 2) & (N - 1))];                // n0[x + 2 e1]
                const int2 me = *reinterpret_cast<const int2*>(sn1 + i);         // n1[x]
                const int2 up = *reinterpret_cast<const int2*>(sn1 + iup);       // n1[x + e0]
                // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x])      (compact.py d,1 rows)
                const int d0 = (up.x - me.x) - (h.y - h.x), d1 = (up.y - me.y) - (hr - h.y);
                dn2 += (long long)d0 * d0 + (long long)d1 * d1;
                cs.w0 += h.x + h.y;
                cs.w1 += me.x + me.y;
            }
            cs.dn2 = dn2;
            cs = block_reduce_chain(cs, scratch);
            if (tid == 0) {
                double* o = a.obs + chain * SVB_VOBS_COUNT;
                o[SVB_VOBS_ACTION] = (kappa / 2) * cs.action;
                o[SVB_VOBS_SUM_DN2] = (double)cs.dn2;
                o[SVB_VOBS_WRAP0] = (double)cs.w0;
                o[SVB_VOBS_WRAP1] = (double)cs.w1;
                o[SVB_VOBS_ACCEPTED] = (double)cs.accepted;
                o[SVB_VOBS_ACCEPTANCE] = cs.sumA;
            }
        }

        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(reinterpret_cast<double*>(a.phi) + chain * V, sphi, bytes_phi);
            bulk_s2g(a.n + chain * 2 * V, sn0, bytes_n);
            bulk_commit();
            if (STAGES == 1) {
                bulk_wait_read0();
                if (next < a.chains) issue_load(next, 0);
            }
        }
        if (STAGES == 1) __syncthreads();
    }
    if (tid == 0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------
// TILED path: lattices that do not fit shared memory (config 4: L=128, config 5: L=4096).
//
// One CTA owns a 32 x 32 tile of one chain and produces its final state in ONE pass: it loads the tile plus a ghost
// zone (2 sites on the low sides, 3 on the high sides: 37 x 38 with padding) from the INPUT buffers, runs the colour-0
// updates on the tile grown by one ring more than the colour-1 updates need, then the colour-1 updates, and writes only
// the sites and links it owns to the OUTPUT buffers (ping-pong: a neighbouring CTA may still be reading the inputs).
// The Philox counter is keyed by the GLOBAL site and chain, so the redundant ghost updates are bit-identical in every
// tile that computes them and the result does not depend on the tiling (tests: identical to the global path and to
// the oracle).  HBM traffic: (37 x 38 / 1024) reads + 1 write of the state = 38 B per site-update instead of the
// >= 56 B of the per-colour global path; compute redundancy (35^2 + 33^2) / (2 x 1024) = 1.13.
// Why these sizes: a link (mu, x) owned by the tile is last touched by x or x + e_mu, so final decisions are needed
// on the tile grown by +1 on the high sides (33 x 33, local [2, 35)); the colour-1 sites there read phi and links that
// colour-0 sites one ring further out (35 x 35, local [1, 36)) may have changed; those read initial data one more
// ring out (37 x 37, local [0, 37)).
// ------------------------------------------------------------------------------------------
constexpr int kTile = 32;
constexpr int kRegRows = kTile + 5;      // 37
constexpr int kRegCols = kTile + 6;      // 38: even, so rows are pairs of 16-byte aligned phi
constexpr int kRegSize = kRegRows * kRegCols;

template <bool STRICT>
__global__ void __launch_bounds__(128, 8) villain_tiled_kernel(const __grid_constant__ VillainArgs a, const double* __restrict__ phi_in,
                                                               const int32_t* __restrict__ n_in, double* __restrict__ phi_out,
                                                               int32_t* __restrict__ n_out, int sweep, int tiles_per_side,
                                                               int fuse_obs) {
    __shared__ __align__(16) double sphi[kRegSize];
    __shared__ __align__(16) int32_t sn0[kRegSize];
    __shared__ __align__(16) int32_t sn1[kRegSize];
    __shared__ double red[2 * 32];
    const int N = a.N;
    const long long V = (long long)N * N;
    const int tiles = tiles_per_side * tiles_per_side;
    const long long chain = blockIdx.x / tiles;
    const int tile = blockIdx.x - (int)(chain * tiles);
    const int a0 = (tile / tiles_per_side) * kTile, a1 = (tile % tiles_per_side) * kTile;
    const int tid = threadIdx.x;
    const double* gphi = phi_in + chain * V;
    const int32_t* gn0 = n_in + chain * 2 * V;
    const int32_t* gn1 = gn0 + V;

    // ---- load the region, two sites at a time (N and the origins are even: a pair never straddles the wrap) ----
    for (int p = tid; p < kRegRows * (kRegCols / 2); p += 128) {
        const int i = p / (kRegCols / 2), jj = 2 * (p - i * (kRegCols / 2));
        int x0 = a0 - 2 + i;  x0 += (x0 < 0) ? N : 0;  x0 -= (x0 >= N) ? N : 0;
        int x1 = a1 - 2 + jj; x1 += (x1 < 0) ? N : 0;  x1 -= (x1 >= N) ? N : 0;
        const long long g = (long long)x0 * N + x1;
        const int l = i * kRegCols + jj;
        *reinterpret_cast<double2*>(sphi + l) = *reinterpret_cast<const double2*>(gphi + g);
        *reinterpret_cast<int2*>(sn0 + l) = *reinterpret_cast<const int2*>(gn0 + g);
        *reinterpret_cast<int2*>(sn1 + l) = *reinterpret_cast<const int2*>(gn1 + g);
    }
    __syncthreads();

    const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
    const double half_kappa = kappa / 2;
    VillainConsts kc;
    int dg_scale;
    villain_consts<false, STRICT>(a, kc, dg_scale);
    LinkSums unused;
    double n_acc = 0.0, sum_A = 0.0;

#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        const int lo = (c == 0) ? 1 : 2;                      // colour 0 on local [1, 36), colour 1 on local [2, 35)
        const int side = (c == 0) ? kTile + 3 : kTile + 1;    // 35, 33
        const int per_row = (side + 1) / 2;
        for (int idx = tid; idx < side * per_row; idx += 128) {
            const int i = lo + idx / per_row;
            const int j = lo + 2 * (idx % per_row) + ((i + lo + c) & 1);       // (i + j) & 1 == c  (origins are even)
            if (j >= lo + side) continue;
            int x0 = a0 - 2 + i;  x0 += (x0 < 0) ? N : 0;  x0 -= (x0 >= N) ? N : 0;
            int x1 = a1 - 2 + j;  x1 += (x1 < 0) ? N : 0;  x1 -= (x1 >= N) ? N : 0;
            const int site = x0 * N + x1;
            const VillainDraw d = villain_get_draw<false, true>(a, chain, sweep, x0, x1, site, dg_scale);
            const SiteOut o = villain_site_update<double, STRICT, 0, true>(sphi, sn0, sn1, kRegCols, i, j, half_kappa, kc, d, false, unused,
                                                                           villain_refine_ctx(a, chain, sweep));
            const bool owned = (i >= 2) && (i < 2 + kTile) && (j >= 2) && (j < 2 + kTile);
            if (owned) {
                n_acc += o.ok ? 1.0 : 0.0;
                sum_A += o.A;
                if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
            }
        }
        __syncthreads();
    }

    // ---- write the owned tile ----
    double* ophi = phi_out + chain * V;
    int32_t* on0 = n_out + chain * 2 * V;
    int32_t* on1 = on0 + V;
    for (int p = tid; p < kTile * (kTile / 2); p += 128) {
        const int i = p / (kTile / 2), jj = 2 * (p - i * (kTile / 2));
        const long long g = (long long)(a0 + i) * N + (a1 + jj);
        const int l = (i + 2) * kRegCols + (jj + 2);
        *reinterpret_cast<double2*>(ophi + g) = *reinterpret_cast<const double2*>(sphi + l);
        *reinterpret_cast<int2*>(on0 + g) = *reinterpret_cast<const int2*>(sn0 + l);
        *reinterpret_cast<int2*>(on1 + g) = *reinterpret_cast<const int2*>(sn1 + l);
    }
    if (a.obs) {
        double sred[2] = {n_acc, sum_A};
        block_sum<2>(sred, red);
        if (tid == 0) {
            atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, sred[0]);
            atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, sred[1]);
        }
        if (fuse_obs) {
            // Observables of the tile's owned sites from the final state in shared memory.  Every link an owned
            // plaquette touches has both of its end sites inside the region where final decisions were made
            // (local [2, 35)), so (dn)^2 is final here as well.  Partials are combined with atomics.
            __syncthreads();
            ChainSums cs;
            cs.action = 0.0; cs.sumA = 0.0; cs.dn2 = 0; cs.w0 = 0; cs.w1 = 0; cs.accepted = 0;
            for (int p = tid; p < kTile * kTile; p += 128) {
                const int i = 2 + p / kTile, j = 2 + (p % kTile);
                villain_obs_site<double, 0>(sphi, sn0, sn1, kRegCols, i, j, cs.action, cs.dn2, cs.w0, cs.w1);
            }
            __shared__ double scratch[6 * 32];
            cs = block_reduce_chain(cs, scratch);
            if (tid == 0) {
                double* o = a.obs + chain * SVB_VOBS_COUNT;
                atomicAdd(o + SVB_VOBS_ACTION, (kappa / 2) * cs.action);
                atomicAdd(o + SVB_VOBS_SUM_DN2, (double)cs.dn2);
                atomicAdd(o + SVB_VOBS_WRAP0, (double)cs.w0);
                atomicAdd(o + SVB_VOBS_WRAP1, (double)cs.w1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// GLOBAL path: one launch per colour pass, straight out of HBM / L2 (any N).
// ------------------------------------------------------------------------------------------
template <typename real, bool INJECTED, bool STRICT>
__global__ void __launch_bounds__(256) villain_colour_pass_kernel(const __grid_constant__ VillainArgs a, int sweep, int colour,
                                                                  int blocks_per_chain, int write_debug) {
    const int N = a.N, V = N * N;
    const long long chain = blockIdx.x / blocks_per_chain;
    const int blk = blockIdx.x - (int)(chain * blocks_per_chain);
    real* gphi = reinterpret_cast<real*>(a.phi) + chain * V;
    int32_t* gn0 = a.n + chain * 2 * V;
    int32_t* gn1 = gn0 + V;
    const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
    const real half_kappa = (real)(kappa / 2);
    double n_acc = 0.0, sum_A = 0.0;

    int site = -1, x0 = 0, x1 = 0;
    if ((N & 1) == 0) {
        const int j = blk * blockDim.x + threadIdx.x;
        if (j < (V >> 1)) {
            const int halfN = N >> 1;
            x0 = j / halfN;
            x1 = 2 * (j - x0 * halfN) + ((x0 + colour) & 1);
            site = x0 * N + x1;
        }
    } else {
        const int i = blk * blockDim.x + threadIdx.x;
        if (i < V) {
            x0 = i / N;
            x1 = i - x0 * N;
            if (site_colour(x0, x1, N) == colour) site = i;
        }
    }
    if (site >= 0) {
        VillainConsts kc;
        int dg_scale;
        villain_consts<INJECTED, STRICT>(a, kc, dg_scale);
        LinkSums unused;
        const VillainDraw d = villain_get_draw<INJECTED, true>(a, chain, sweep, x0, x1, site, dg_scale);
        const SiteOut o = villain_site_update<real, STRICT, 0, !INJECTED>(gphi, gn0, gn1, N, x0, x1, half_kappa, kc, d, false, unused,
                                                                          villain_refine_ctx(a, chain, sweep));
        n_acc = o.ok ? 1.0 : 0.0;
        sum_A = o.A;
        if (write_debug) {
            if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
            if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
        }
    }
    if (a.obs) {
        __shared__ double scratch[2 * 32];
        double s[2] = {n_acc, sum_A};
        block_sum<2>(s, scratch);
        if (threadIdx.x == 0) {
            atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, s[0]);
            atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, s[1]);
        }
    }
}

// Observables of the current state, one CTA per chain.  keep_counters: leave ACCEPTED/ACCEPTANCE alone.
template <typename real>
__global__ void __launch_bounds__(256) villain_obs_kernel(const real* __restrict__ phi, const int32_t* __restrict__ n,
                                                          long long chains, int N, double kappa_scalar,
                                                          const double* __restrict__ kappa_chain, double* __restrict__ obs,
                                                          int keep_counters) {
    __shared__ double scratch[6 * 32];
    const int V = N * N;
    for (long long chain = blockIdx.x; chain < chains; chain += gridDim.x) {
        const real* gphi = phi + chain * V;
        const int32_t* gn0 = n + chain * 2 * V;
        ChainSums cs;
        cs.action = 0.0; cs.sumA = 0.0; cs.dn2 = 0; cs.w0 = 0; cs.w1 = 0; cs.accepted = 0;
        for (int i = threadIdx.x; i < V; i += blockDim.x) {
            const int x0 = i / N, x1 = i - x0 * N;
            villain_obs_site<real, 0>(gphi, gn0, gn0 + V, N, x0, x1, cs.action, cs.dn2, cs.w0, cs.w1);
        }
        cs = block_reduce_chain(cs, scratch);
        if (threadIdx.x == 0) {
            const double kappa = kappa_chain ? kappa_chain[chain] : kappa_scalar;
            double* o = obs + chain * SVB_VOBS_COUNT;
            o[SVB_VOBS_ACTION] = (kappa / 2) * cs.action;
            o[SVB_VOBS_SUM_DN2] = (double)cs.dn2;
            o[SVB_VOBS_WRAP0] = (double)cs.w0;
            o[SVB_VOBS_WRAP1] = (double)cs.w1;
            if (!keep_counters) {
                o[SVB_VOBS_ACCEPTED] = 0.0;
                o[SVB_VOBS_ACCEPTANCE] = 0.0;
            }
        }
        __syncthreads();
    }
}

// Observables of large lattices: many CTAs per chain, block partials combined with atomics (the summation order of
// the fp64 partials is then not fixed: results agree to ~1e-16 relative between runs, not bitwise).
template <typename real>
__global__ void __launch_bounds__(256) villain_obs_tiled_kernel(const real* __restrict__ phi, const int32_t* __restrict__ n,
                                                                long long chains, int N, double kappa_scalar,
                                                                const double* __restrict__ kappa_chain, double* __restrict__ obs,
                                                                int blocks_per_chain, int sites_per_block) {
    __shared__ double scratch[6 * 32];
    const int V = N * N;
    const long long chain = blockIdx.x / blocks_per_chain;
    const int blk = blockIdx.x - (int)(chain * blocks_per_chain);
    const real* gphi = phi + chain * V;
    const int32_t* gn0 = n + chain * 2 * V;
    ChainSums cs;
    cs.action = 0.0; cs.sumA = 0.0; cs.dn2 = 0; cs.w0 = 0; cs.w1 = 0; cs.accepted = 0;
    const int lo = blk * sites_per_block;
    const int hi = min(V, lo + sites_per_block);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int x0 = i / N, x1 = i - x0 * N;
        villain_obs_site<real, 0>(gphi, gn0, gn0 + V, N, x0, x1, cs.action, cs.dn2, cs.w0, cs.w1);
    }
    cs = block_reduce_chain(cs, scratch);
    if (threadIdx.x == 0) {
        const double kappa = kappa_chain ? kappa_chain[chain] : kappa_scalar;
        double* o = obs + chain * SVB_VOBS_COUNT;
        atomicAdd(o + SVB_VOBS_ACTION, (kappa / 2) * cs.action);
        atomicAdd(o + SVB_VOBS_SUM_DN2, (double)cs.dn2);
        atomicAdd(o + SVB_VOBS_WRAP0, (double)cs.w0);
        atomicAdd(o + SVB_VOBS_WRAP1, (double)cs.w1);
    }
}

__global__ void villain_zero_record_kernel(double* obs, long long chains, int keep_counters) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < chains) {
        double* o = obs + c * SVB_VOBS_COUNT;
        o[SVB_VOBS_ACTION] = 0.0; o[SVB_VOBS_SUM_DN2] = 0.0; o[SVB_VOBS_WRAP0] = 0.0; o[SVB_VOBS_WRAP1] = 0.0;
        if (!keep_counters) { o[SVB_VOBS_ACCEPTED] = 0.0; o[SVB_VOBS_ACCEPTANCE] = 0.0; }
    }
}

// one launch for both records of svb_villain_sweep_inplace: the state columns of `state` (if any), the counters of `counters`
__global__ void villain_zero_inplace_records_kernel(double* state, double* counters, long long chains) {
    // (programmatic dependent launch: see launch_pdl in svb_villain_stream.cuh)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < chains) {
        if (state) {
            double* o = state + c * SVB_VOBS_COUNT;
            o[SVB_VOBS_ACTION] = 0.0; o[SVB_VOBS_SUM_DN2] = 0.0; o[SVB_VOBS_WRAP0] = 0.0; o[SVB_VOBS_WRAP1] = 0.0;
        }
        if (counters) {
            counters[c * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED] = 0.0;
            counters[c * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE] = 0.0;
        }
    }
}

__global__ void villain_zero_counters_kernel(double* obs, long long chains) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < chains) {
        obs[c * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED] = 0.0;
        obs[c * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE] = 0.0;
    }
}

__global__ void villain_draws_kernel(long long chains, int N, int W, double interval_phi, int interval_n,
                                     unsigned long long seed, unsigned long long sweep, unsigned long long chain0,
                                     double* u, double* dphi, int32_t* dn) {
    const long long V = (long long)N * N;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= chains * V) return;
    const long long chain = i / V;
    const int site = (int)(i - chain * V);
    const int x0 = site / N, x1 = site - x0 * N;
    const uint32_t c0 = villain_pair_counter(x0, x1, N), half = villain_pair_half(x0);
    const Philox4 p = philox_site(seed, chain0 + chain, sweep, c0, STREAM_VILLAIN_NEIGHBORHOOD);
    const VillainDraw d = villain_draw_from_words(half ? p.z : p.x, half ? p.w : p.y, interval_phi, interval_n);
    const long long K = 2LL * interval_n + 1;
    if (K * K * K * K > 256) {                      // wide: leading bits = refinement word 2 half, trailing bits = word 2 half + 1
        const Philox4 r = philox_site(seed, chain0 + chain, sweep, c0, STREAM_VILLAIN_REFINE);
        u[i] = refined_uniform(half ? r.z : r.x, c0, 2 * half + 1, STREAM_VILLAIN_REFINE, seed, chain0 + chain, sweep);
    } else {
        u[i] = refined_uniform(d.f, c0, 2 * half, STREAM_VILLAIN_REFINE, seed, chain0 + chain, sweep);
    }
    dphi[i] = d.dphi;
    for (int k = 0; k < 4; ++k) dn[(chain * 4 + k) * V + site] = W * d.dg[k];
}

// ------------------------------------------------------------------------------------------
// LinkUpdate (supervillain/generator/villain/link.py:53-101): every link independently against the frozen phi,
//   change = W * choice([-I..-1, 1..I]);  dS = ((-2 pi) kappa change) ((dphi - (2 pi) n) - pi change);  n += change if u < e^-dS.
// One thread per link; STRICT arithmetic (one rounding per numpy operation).  Philox draws: the block with counter word 0 =
// site index in stream STREAM_VILLAIN_LINK, word mu for link (mu, x): p = (2 I) w, choice index = p >> 32, the remainder
// leads the lazily refined uniform (stream STREAM_VILLAIN_LINK_REFINE).
// ------------------------------------------------------------------------------------------
struct LinkArgs {
    const double* phi;
    int32_t* n;
    long long chains;
    int N;
    double kappa;
    const double* kappa_chain;
    int W, interval;
    unsigned long long seed, sweep, chain0;
    const double* inj_u;        // (chains, 2, N, N)
    const int32_t* inj_c;       // (chains, 2, N, N), already times W
    double* obs;                // ACCEPTED / ACCEPTANCE accumulate here (atomics)
    double* dS_out;             // (chains, 2, N, N)
};

template <bool INJECTED>
__global__ void __launch_bounds__(256) villain_link_kernel(LinkArgs a) {
    __shared__ double scratch[2 * 32];
    const int N = a.N, V = N * N;
    const int blocks_per_chain = (2 * V + 255) / 256;
    const long long chain = blockIdx.x / blocks_per_chain;
    const int l = (blockIdx.x - (int)(chain * blocks_per_chain)) * 256 + threadIdx.x;
    double acc_n = 0.0, acc_A = 0.0;
    if (l < 2 * V) {
        const int mu = l / V, site = l - mu * V;
        const int x0 = site / N, x1 = site - x0 * N;
        const int head = (mu == 0) ? ((x0 + 1 == N ? 0 : x0 + 1) * N + x1) : (x0 * N + (x1 + 1 == N ? 0 : x1 + 1));
        const double* gphi = a.phi + chain * V;
        int32_t* gn = a.n + chain * 2 * V;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double dphi = __dsub_rn(gphi[head], gphi[site]);                          // d(phi)   (link.py:74)
        int c;
        double u = 0.0;
        LazyUniform lu;
        lu.f = 0; lu.c0 = (uint32_t)site; lu.word = (uint32_t)mu;
        if (INJECTED) {
            c = a.inj_c[chain * 2 * V + l];
            u = a.inj_u[chain * 2 * V + l];
        } else {
            const Philox4 p = philox_site(a.seed, a.chain0 + (unsigned long long)chain, a.sweep, (uint32_t)site, STREAM_VILLAIN_LINK);
            const uint64_t pz = (uint64_t)(mu ? p.y : p.x) * (uint64_t)(2 * a.interval);
            const int idx = (int)(pz >> 32);
            c = a.W * ((idx < a.interval) ? idx - a.interval : idx - a.interval + 1);
            lu.f = (uint32_t)pz;
        }
        const int nl = gn[l];
        // dS = -2 pi kappa change (dphi - 2 pi n - pi change), numpy's left-to-right order   (link.py:83-86)
        const double t1 = __dmul_rn(__dmul_rn(-SVB_TWO_PI, kappa), (double)c);
        const double t2 = __dsub_rn(__dsub_rn(dphi, __dmul_rn(SVB_TWO_PI, (double)nl)), __dmul_rn(3.141592653589793116, (double)c));
        const double dS = __dmul_rn(t1, t2);
        const double A = exp_clipped(-dS);
        bool ok;
        if (INJECTED) ok = u < A;
        else {
            RefineCtx rc;
            rc.seed = a.seed; rc.chain = a.chain0 + (unsigned long long)chain; rc.sweep = a.sweep;
            ok = decide_lazy(A, lu, STREAM_VILLAIN_LINK_REFINE, rc);
        }
        if (ok) gn[l] = nl + c;
        if (a.dS_out) a.dS_out[chain * 2 * V + l] = dS;
        acc_n = ok ? 1.0 : 0.0;
        acc_A = A;
    }
    if (a.obs) {
        double s[2] = {acc_n, acc_A};
        block_sum<2>(s, scratch);
        if (threadIdx.x == 0) {
            atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, s[0]);
            atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, s[1]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// CohomologyUpdate (supervillain/generator/villain/cohomology.py:64-117): per direction mu one proposal h added to n_mu on
// the whole slice x_mu = 0;  change_r = -2 pi h;  dS = sum over the slice of ((kappa/2) change_r) ((2 r) + change_r).
// The two directions touch different components, so they are decided concurrently: one warp per (chain, mu), lanes striding
// over the N links of the slice.  The N terms are summed by a warp tree where numpy sums pairwise: dS agrees to ~1e-15
// relative, not bitwise.  Philox: block with counter word 0 = mu in stream STREAM_VILLAIN_COHOMOLOGY; word 0 -> h (index =
// (2 I) w >> 32), words 1, 2 -> a 52-bit uniform (k + 1/2) 2^-52.
// ------------------------------------------------------------------------------------------
struct CohomologyArgs {
    const double* phi;
    int32_t* n;
    long long chains;
    int N;
    double kappa;
    const double* kappa_chain;
    int interval;
    unsigned long long seed, sweep, chain0;
    const double* inj_u;       // (chains, 2)
    const int32_t* inj_h;      // (chains, 2)
    double* counters;          // (chains, 2): accepted, sum of acceptance (accumulated)
    double* dS_out;            // (chains, 2)
};

template <bool INJECTED>
__global__ void __launch_bounds__(128) villain_cohomology_kernel(CohomologyArgs a) {
    const int lane = threadIdx.x & 31;
    const long long job = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);          // (chain, mu)
    if (job >= 2 * a.chains) return;
    const long long chain = job >> 1;
    const int mu = (int)(job & 1), N = a.N, V = N * N;
    const double* gphi = a.phi + chain * V;
    int32_t* gn = a.n + chain * 2 * V + (long long)mu * V;
    const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
    int h;
    double u;
    if (INJECTED) {
        h = a.inj_h[job];
        u = a.inj_u[job];
    } else {
        const Philox4 p = philox_site(a.seed, a.chain0 + (unsigned long long)chain, a.sweep, (uint32_t)mu, STREAM_VILLAIN_COHOMOLOGY);
        const int idx = (int)(((uint64_t)p.x * (uint64_t)(2 * a.interval)) >> 32);
        h = (idx < a.interval) ? idx - a.interval : idx - a.interval + 1;
        const uint64_t ku = ((uint64_t)(p.y & 0xFFFFFu) << 32) | (uint64_t)p.z;
        u = __dmul_rn(__dadd_rn((double)ku, 0.5), 2.220446049250313e-16);
    }
    const double change_r = __dmul_rn(-SVB_TWO_PI, (double)h);                         // -2 * np.pi * h
    const double hk_cr = __dmul_rn(kappa / 2, change_r);
    double dS = 0.0;
    for (int j = lane; j < N; j += 32) {
        // slice x_mu = 0: link (mu, x) with x = (0, j) for mu = 0, (j, 0) for mu = 1
        const int tail = (mu == 0) ? j : j * N;
        const int head = (mu == 0) ? N + j : j * N + 1;
        const double r = __dsub_rn(__dsub_rn(gphi[head], gphi[tail]), __dmul_rn(SVB_TWO_PI, (double)gn[tail]));
        dS = __dadd_rn(dS, __dmul_rn(hk_cr, __dadd_rn(__dmul_rn(2.0, r), change_r)));
    }
    dS = warp_sum(dS);
    const double A = exp_clipped(-dS);
    const bool ok = u < A;
    if (ok)
        for (int j = lane; j < N; j += 32) gn[(mu == 0) ? j : j * N] += h;
    if (lane == 0) {
        if (a.counters) {
            atomicAdd(a.counters + chain * 2, ok ? 1.0 : 0.0);
            atomicAdd(a.counters + chain * 2 + 1, A);
        }
        if (a.dS_out) a.dS_out[job] = dS;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static size_t villain_smem_bytes(int N, size_t real_size) {
    const size_t V = (size_t)N * N;
    const size_t off_n = (V * real_size + 15) & ~(size_t)15;
    const size_t off_scr = (off_n + 2 * V * sizeof(int32_t) + 15) & ~(size_t)15;
    return off_scr + 6 * 32 * sizeof(double) + 16;
}

static int villain_threads_for(int N) {
    const int V = N * N;
    int t = (V / 8 + 31) / 32 * 32;   // ~4 sites per colour per thread
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
}

struct DeviceInfo {
    int sm_count;
    int max_smem_optin;
    int device;
};
static int get_device_info(DeviceInfo& info) {
    // device attributes never change: cached per device (launch-bound loops call this every step)
    static DeviceInfo cache[64];
    static bool have[64];
    int dev = 0;
    SVB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && have[dev]) {
        info = cache[dev];
        return 0;
    }
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    info.device = dev;
    if (dev >= 0 && dev < 64) {
        cache[dev] = info;
        have[dev] = true;
    }
    return 0;
}

template <typename real, bool INJECTED, bool STRICT, int NT, int TT, int MINB>
static int launch_villain_smem_inst(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    auto kern = villain_smem_kernel<real, INJECTED, STRICT, NT, TT, MINB>;
    const size_t smem = villain_smem_bytes(a.N, sizeof(real));
    const int threads = TT ? TT : villain_threads_for(a.N);
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "villain smem kernel does not fit an SM at N=%d", a.N);
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > a.chains) grid = a.chains;
    // bulk (TMA 1-D) copies need 16-byte sizes and addresses
    const size_t bytes_phi = (size_t)a.N * a.N * sizeof(real);
    const size_t bytes_n = (size_t)2 * a.N * a.N * sizeof(int32_t);
    const int use_bulk = (bytes_phi % 16 == 0) && (bytes_n % 16 == 0) && ((uintptr_t)a.phi % 16 == 0) &&
                         ((uintptr_t)a.n % 16 == 0);
    kern<<<(unsigned)grid, threads, smem, stream>>>(a, use_bulk);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename real, bool STRICT, int NT, int TT, int MINB>
static int launch_villain_pipelined(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    auto kern = villain_smem_pipelined_kernel<real, STRICT, NT, TT, MINB>;
    const size_t stage = (size_t)NT * NT * (sizeof(real) + 2 * sizeof(int32_t));
    const size_t smem = 2 * stage + 6 * 32 * sizeof(double) + 16;
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TT, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "pipelined villain kernel does not fit an SM at N=%d", NT);
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > a.chains) grid = a.chains;
    kern<<<(unsigned)grid, TT, smem, stream>>>(a);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <bool STRICT, int NT, int TT, int MINB, int STAGES>
static int launch_villain_resid(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    auto kern = villain_smem_resid_kernel<STRICT, NT, TT, MINB, STAGES>;
    const size_t V = (size_t)NT * NT;
    const size_t smem = STAGES * V * 16 + 2 * V * sizeof(double) + 64 * sizeof(double) + 6 * 32 * sizeof(double) + 16;
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TT, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "resident-residual villain kernel does not fit an SM at N=%d", NT);
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > a.chains) grid = a.chains;
    kern<<<(unsigned)grid, TT, smem, stream>>>(a);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

#include "svb_villain_filtered.cuh"
#include "svb_villain_stream.cuh"
#include "svb_villain_cluster.cuh"
#include "svb_villain_strips.cuh"
#include "svb_villain_link.cuh"
#ifndef SVB_CLUSTER_TPB
#define SVB_CLUSTER_TPB 512
#endif
#ifndef SVB_CLUSTER_CL
#define SVB_CLUSTER_CL 4
#endif
#ifndef SVB_CLUSTER_STAGES
#define SVB_CLUSTER_STAGES 1
#endif

#ifndef SVB_FILT_MINB32
#ifndef SVB_STREAM_MINB32
#define SVB_STREAM_MINB32 8     /* CTAs per SM the streaming chain kernel is compiled for: 4 N threads each */
#endif
#ifndef SVB_STREAM_MINB64
#define SVB_STREAM_MINB64 4
#endif
#ifndef SVB_STREAM_MINB128
#define SVB_STREAM_MINB128 2
#endif
#define SVB_FILT_MINB32 8        /* 64 registers per thread: 8 CTAs = 32 warps per SM (28.5 vs 30.0 us at config 2) */
#endif
#define SVB_FILT_STAGES 1
#ifndef SVB_RESID_STAGES
#define SVB_RESID_STAGES 1
#endif
#ifndef SVB_RESID_T32
#define SVB_RESID_T32 128
#endif

// Which kernel family serves the production sweeps (Philox, fp64, FAST): the streaming kernels (svb_villain_stream.cuh,
// default) or the shared-memory-resident ones (svb_villain_filtered.cuh / svb_villain_cluster.cuh; SVB_VILLAIN_KERNEL=smem in
// the environment, read per call so that one process can time both).
static bool villain_stream_enabled() {
    const char* e = getenv("SVB_VILLAIN_KERNEL");
    return !(e && e[0] == 's' && e[1] == 'm');
}
// The chain-per-CTA streaming kernel is an experiment kept for A/B runs (SVB_VILLAIN_KERNEL=stream): it is bound by L1
// wavefronts (nine global loads per site), 48.7 us per step at config 2 against 28.2 us for the shared-memory kernel and
// 171 us at L = 128 against 157 us for the SPARSE cluster kernel.
static bool villain_stream_chain_enabled(int N) {
    const char* e = getenv("SVB_VILLAIN_KERNEL");
    (void)N;
    return e && e[0] == 's' && e[1] == 't';
}
static int launch_villain_stream_by_size(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    switch (a.N) {
        case 16: return launch_villain_stream_chain<16, 16>(a, stream, info);
        case 32: return launch_villain_stream_chain<32, SVB_STREAM_MINB32>(a, stream, info);
        case 64: return launch_villain_stream_chain<64, SVB_STREAM_MINB64>(a, stream, info);
        default: return launch_villain_stream_chain<128, SVB_STREAM_MINB128>(a, stream, info);
    }
}
static bool villain_stream_serves(const VillainArgs& a, bool injected, bool strict, size_t real_size) {
    return !injected && !strict && !a.filtered_strict && !a.exact_mode && !a.wide && real_size == 8 && !a.accept_mask && !a.dS_out &&
           (a.N == 16 || a.N == 32 || a.N == 64 || a.N == 128) && ((uintptr_t)a.phi % 16 == 0) && ((uintptr_t)a.n % 16 == 0) &&
           villain_stream_chain_enabled(a.N);
}

template <typename real, bool INJECTED, bool STRICT>
static int launch_villain_smem(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    const bool aligned = ((uintptr_t)a.phi % 16 == 0) && ((uintptr_t)a.n % 16 == 0);
    if (villain_stream_serves(a, INJECTED, STRICT, sizeof(real))) return launch_villain_stream_by_size(a, stream, info);
#ifndef SVB_NO_FILTERED_KERNEL
    if (!INJECTED && (!STRICT || a.filtered_strict) && aligned && sizeof(real) == 8 && !a.accept_mask && !a.dS_out &&
        (!a.exact_mode || a.filtered_strict) && !a.wide) {
        // production path: fp32-filtered decisions on resident fp32 residuals (svb_villain_filtered.cuh)
        switch (a.N) {
            case 16: return launch_villain_filtered<16, 16, 1>(a, stream, info);
            case 32: return launch_villain_filtered<32, SVB_FILT_MINB32, SVB_FILT_STAGES>(a, stream, info);
            case 64: return launch_villain_filtered<64, 2, 1>(a, stream, info);
            default: break;
        }
    }
#endif
    if (!INJECTED && aligned && sizeof(real) == 8 && !a.exact_mode && !a.wide) {
        // One sweep per launch in FAST arithmetic: recomputing the residuals is as cheap as building the resident copy,
        // and the two-stage pipeline hides the loads (52.0 vs 53.5 us at config 2).  Fused sweeps, and STRICT
        // arithmetic always (bit-exact dS), keep the residuals resident (34.9 vs 38.3 us per sweep).
        const bool resident = STRICT || a.n_sweeps > 1;
        if (resident) {
            switch (a.N) {
                case 16: return launch_villain_resid<STRICT, 16, 32, 16, 2>(a, stream, info);
                case 32: return launch_villain_resid<STRICT, 32, SVB_RESID_T32, SVB_MINB32 * 128 / SVB_RESID_T32, SVB_RESID_STAGES>(a, stream, info);
                case 64: return launch_villain_resid<STRICT, 64, 256, 1, 1>(a, stream, info);
                default: break;
            }
        } else {
            switch (a.N) {
                case 16: return launch_villain_pipelined<real, STRICT, 16, 32, 16>(a, stream, info);
                case 32: return launch_villain_pipelined<real, STRICT, 32, 128, SVB_MINB32>(a, stream, info);
                case 64: return launch_villain_pipelined<real, STRICT, 64, 256, 1>(a, stream, info);
                default: break;
            }
        }
    }
    return launch_villain_smem_inst<real, INJECTED, STRICT, 0, 0, 1>(a, stream, info);
}

// Observables of the current state.  Small lattices: one CTA per chain (deterministic order).  Large lattices with few
// chains: several CTAs per chain + atomics, so a single L=4096 chain still fills the GPU.
template <typename real>
static int launch_villain_obs(const real* phi, const int32_t* n, long long chains, int N, double kappa, const double* kappa_chain,
                              double* obs, int keep_counters, cudaStream_t stream) {
    const long long V = (long long)N * N;
    const long long want_ctas = 148LL * 8;
    if (sizeof(real) == 8 && chains >= 64 && ((uintptr_t)phi % 16 == 0) && ((uintptr_t)n % 16 == 0)) {
        const double* p64 = reinterpret_cast<const double*>(phi);
        switch (N) {          // the chain staged in shared memory by TMA, one vectorised pass (svb_villain_filtered.cuh)
            case 16: return launch_villain_obs_smem<16>(p64, n, chains, kappa, kappa_chain, obs, keep_counters, stream);
            case 32: return launch_villain_obs_smem<32>(p64, n, chains, kappa, kappa_chain, obs, keep_counters, stream);
            case 64: return launch_villain_obs_smem<64>(p64, n, chains, kappa, kappa_chain, obs, keep_counters, stream);
            default: break;
        }
    }
    if (V <= 16384 || chains >= want_ctas) {
        long long grid = chains < want_ctas ? chains : want_ctas;
        villain_obs_kernel<real><<<(unsigned)grid, 256, 0, stream>>>(phi, n, chains, N, kappa, kappa_chain, obs, keep_counters);
        SVB_CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (sizeof(real) == 8 && N % 16 == 0 && ((uintptr_t)phi % 16 == 0) && ((uintptr_t)n % 16 == 0)) {
        // vectorised streaming pass (svb_villain_stream.cuh): four sites per thread from 16-byte loads
        villain_zero_record_kernel<<<(unsigned)((chains + 255) / 256), 256, 0, stream>>>(obs, chains, keep_counters);
        SVB_CUDA_TRY(cudaGetLastError());
        villain_stream_obs_kernel<<<(unsigned)want_ctas, 256, 0, stream>>>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs);
        SVB_CUDA_TRY(cudaGetLastError());
        return 0;
    }
    int bpc = (int)((want_ctas + chains - 1) / chains);
    int spb = (int)((V + bpc - 1) / bpc);
    if (spb < 2048) spb = 2048;
    bpc = (int)((V + spb - 1) / spb);
    villain_zero_record_kernel<<<(unsigned)((chains + 255) / 256), 256, 0, stream>>>(obs, chains, keep_counters);
    SVB_CUDA_TRY(cudaGetLastError());
    villain_obs_tiled_kernel<real><<<(unsigned)(bpc * chains), 256, 0, stream>>>(phi, n, chains, N, kappa, kappa_chain, obs, bpc, spb);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename real, bool INJECTED, bool STRICT>
static int launch_villain_global(const VillainArgs& a, cudaStream_t stream) {
    const int N = a.N, V = N * N;
    const int threads = 256;
    const int work = (N & 1) ? V : V / 2;
    const int bpc = (work + threads - 1) / threads;
    const long long blocks = (long long)bpc * a.chains;
    if (blocks > 0x7fffffffLL) return fail(SVB_E_SHAPE, "too many blocks (%lld) for the global path", blocks);
    if (a.obs) {
        villain_zero_counters_kernel<<<(unsigned)((a.chains + 255) / 256), 256, 0, stream>>>(a.obs, a.chains);
        SVB_CUDA_TRY(cudaGetLastError());
    }
    const int ncol = n_colours(N);
    for (int s = 0; s < a.n_sweeps; ++s) {
        for (int c = 0; c < ncol; ++c) {
            villain_colour_pass_kernel<real, INJECTED, STRICT>
                <<<(unsigned)blocks, threads, 0, stream>>>(a, s, c, bpc, s == a.n_sweeps - 1);
            SVB_CUDA_TRY(cudaGetLastError());
        }
    }
    if (a.obs) return launch_villain_obs<real>(reinterpret_cast<const real*>(a.phi), a.n, a.chains, N, a.kappa, a.kappa_chain, a.obs, 1, stream);
    return 0;
}

// Philox key schedule and the stream pair of the generator kind; `wide`: see the draw mapping at villain_pair_counter.
static int villain_wide(int interval_n) {
    const long long K = 2LL * interval_n + 1;
    return (K * K * K * K > 256) ? 1 : 0;
}
static void villain_rng_setup(VillainArgs& a, uint32_t stream, uint32_t refine_stream, int wide) {
    for (int r = 0; r < 10; ++r) {
        a.round_key[2 * r] = (uint32_t)a.seed + (uint32_t)r * 0x9E3779B9u;
        a.round_key[2 * r + 1] = (uint32_t)(a.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    a.stream_hi = stream << 24;
    a.refine_stream = refine_stream;
    a.wide = wide;
}

template <typename real>
static int dispatch_villain(const VillainArgs& a, int rng_mode, int arith_mode, int path, cudaStream_t stream) {
    DeviceInfo info;
    int rc = get_device_info(info);
    if (rc) return rc;
    if (path != SVB_PATH_GLOBAL && villain_stream_serves(a, rng_mode == SVB_RNG_INJECTED, arith_mode == SVB_ARITH_STRICT, sizeof(real)))
        return launch_villain_stream_by_size(a, stream, info);
#ifndef SVB_NO_CLUSTER_KERNEL
    if (path != SVB_PATH_GLOBAL && a.N == 128 && sizeof(real) == 8 && rng_mode != SVB_RNG_INJECTED &&
        (arith_mode != SVB_ARITH_STRICT || a.filtered_strict) && !a.accept_mask && !a.dS_out && (!a.exact_mode || a.filtered_strict) &&
        !a.wide && ((uintptr_t)a.phi % 16 == 0) && ((uintptr_t)a.n % 16 == 0)) {
        // single sparse sweeps: one chain per CTA, phi and n streamed through a ring of strips (svb_villain_strips.cuh); everything
        // else: one chain per cluster of four CTAs, a 32-row strip each (svb_villain_cluster.cuh)
        if (arith_mode != SVB_ARITH_STRICT && villain_strips_serves(a)) return launch_villain_strips(a, stream, info);
        return launch_villain_cluster<128, SVB_CLUSTER_CL, SVB_CLUSTER_TPB, SVB_CLUSTER_STAGES>(a, stream, info);
    }
#endif
    if (path == SVB_PATH_AUTO) {
        path = (villain_smem_bytes(a.N, sizeof(real)) <= (size_t)info.max_smem_optin) ? SVB_PATH_SMEM : SVB_PATH_GLOBAL;
    }
    if (path == SVB_PATH_SMEM) {
        if (villain_smem_bytes(a.N, sizeof(real)) > (size_t)info.max_smem_optin)
            return fail(SVB_E_UNSUPPORTED, "N=%d does not fit shared memory; use SVB_PATH_GLOBAL", a.N);
        if (rng_mode == SVB_RNG_INJECTED) return launch_villain_smem<real, true, true>(a, stream, info);
        if (arith_mode == SVB_ARITH_STRICT) return launch_villain_smem<real, false, true>(a, stream, info);
        return launch_villain_smem<real, false, false>(a, stream, info);
    }
    if (rng_mode == SVB_RNG_INJECTED) return launch_villain_global<real, true, true>(a, stream);
    if (arith_mode == SVB_ARITH_STRICT) return launch_villain_global<real, false, true>(a, stream);
    return launch_villain_global<real, false, false>(a, stream);
}

}  // namespace svb

using namespace svb;

extern "C" int svb_villain_sweep(void* phi, int phi_dtype, int32_t* n, int64_t chains, int N, double kappa,
                                 const double* kappa_chain, int W, double interval_phi, int interval_n, int n_sweeps,
                                 uint64_t seed, uint64_t sweep0, uint64_t chain0, int rng_mode, int arith_mode, int path,
                                 const double* inj_u, const double* inj_dphi, const int32_t* inj_dn_fwd,
                                 const int32_t* inj_dn_bwd, double* obs, uint8_t* accept_mask, double* dS_out,
                                 void* stream) {
    if (!phi || !n) return fail(SVB_E_NULL, "svb_villain_sweep: phi and n are required");
    if (chains < 0 || N < 3 || N > 32768) return fail(SVB_E_SHAPE, "svb_villain_sweep: chains=%lld N=%d", (long long)chains, N);
    if (phi_dtype != SVB_F64 && phi_dtype != SVB_F32) return fail(SVB_E_DTYPE, "svb_villain_sweep: phi dtype %d", phi_dtype);
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_villain_sweep: kappa must be positive");
    if (W < 1) return fail(SVB_E_PARAM, "svb_villain_sweep: W must be a finite integer >= 1 (got %d)", W);
    if (interval_n < 0 || interval_n > 31) return fail(SVB_E_PARAM, "svb_villain_sweep: interval_n must be in [0, 31]");
    if (!(interval_phi >= 0)) return fail(SVB_E_PARAM, "svb_villain_sweep: interval_phi must be >= 0");
    if (n_sweeps < 0) return fail(SVB_E_PARAM, "svb_villain_sweep: n_sweeps < 0");
    if (rng_mode != SVB_RNG_PHILOX && rng_mode != SVB_RNG_INJECTED) return fail(SVB_E_PARAM, "svb_villain_sweep: rng_mode");
    if (rng_mode == SVB_RNG_INJECTED && (!inj_u || !inj_dphi || !inj_dn_fwd || !inj_dn_bwd))
        return fail(SVB_E_NULL, "svb_villain_sweep: injected mode needs inj_u, inj_dphi, inj_dn_fwd, inj_dn_bwd");
    if (arith_mode != SVB_ARITH_STRICT && arith_mode != SVB_ARITH_FAST) return fail(SVB_E_PARAM, "svb_villain_sweep: arith_mode");
    if (path < SVB_PATH_AUTO || path > SVB_PATH_GLOBAL) return fail(SVB_E_PARAM, "svb_villain_sweep: path");
    if (chains == 0 || n_sweeps == 0) return SVB_OK;

    VillainArgs a;
    a.phi = phi; a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W;
    a.interval_phi = interval_phi; a.interval_n = interval_n; a.n_sweeps = n_sweeps;
    a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    villain_rng_setup(a, STREAM_VILLAIN_NEIGHBORHOOD, STREAM_VILLAIN_REFINE, villain_wide(interval_n));
    a.inj_u = inj_u; a.inj_dphi = inj_dphi; a.inj_dn_fwd = inj_dn_fwd; a.inj_dn_bwd = inj_dn_bwd;
    a.obs = obs; a.accept_mask = accept_mask; a.dS_out = dS_out;
    a.exact_mode = 0; a.inj_z = nullptr; a.filtered_strict = 0; a.obs_in = nullptr; a.epochs = nullptr; a.wait_epoch = 0; a.signal_epoch = 0; a.grid_wait = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (phi_dtype == SVB_F64) return dispatch_villain<double>(a, rng_mode, arith_mode, path, st);
    return dispatch_villain<float>(a, rng_mode, arith_mode, path, st);
}

extern "C" int svb_villain_sweep_overlapped(void* phi, int32_t* n, int64_t chains, int N, double kappa, const double* kappa_chain,
                                            int W, double interval_phi, int interval_n, int n_sweeps, uint64_t seed,
                                            uint64_t sweep0, uint64_t chain0, double* obs, double* obs_in, uint32_t* epochs,
                                            uint32_t wait_epoch, uint32_t signal_epoch, int flags, void* stream) {
    if (!phi || !n || !epochs) return fail(SVB_E_NULL, "svb_villain_sweep_overlapped: phi, n and epochs are required");
    if (chains < 0) return fail(SVB_E_SHAPE, "svb_villain_sweep_overlapped: chains=%lld", (long long)chains);
    if (N != 16 && N != 32 && N != 64 && N != 128)
        return fail(SVB_E_UNSUPPORTED, "svb_villain_sweep_overlapped: N must be 16, 32, 64 or 128 (got %d)", N);
    if (((uintptr_t)phi % 16) || ((uintptr_t)n % 16)) return fail(SVB_E_ALIGN, "svb_villain_sweep_overlapped: fields must be 16-byte aligned");
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_villain_sweep_overlapped: kappa must be positive");
    if (W < 1) return fail(SVB_E_PARAM, "svb_villain_sweep_overlapped: W must be a finite integer >= 1 (got %d)", W);
    if (interval_n < 0 || interval_n > 31) return fail(SVB_E_PARAM, "svb_villain_sweep_overlapped: interval_n must be in [0, 31]");
    if (!(interval_phi >= 0)) return fail(SVB_E_PARAM, "svb_villain_sweep_overlapped: interval_phi must be >= 0");
    if (n_sweeps < 1) return fail(SVB_E_PARAM, "svb_villain_sweep_overlapped: n_sweeps must be >= 1 (every launch signals its epoch)");
    if (villain_wide(interval_n))
        return fail(SVB_E_UNSUPPORTED, "svb_villain_sweep_overlapped: interval_n > 1 is served by svb_villain_sweep (generic kernels) only");
    if (flags & ~SVB_OVERLAP_PREDECESSOR) return fail(SVB_E_PARAM, "svb_villain_sweep_overlapped: flags");
    if (chains == 0) return SVB_OK;
    VillainArgs a;
    a.phi = phi; a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W;
    a.interval_phi = interval_phi; a.interval_n = interval_n; a.n_sweeps = n_sweeps;
    a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    villain_rng_setup(a, STREAM_VILLAIN_NEIGHBORHOOD, STREAM_VILLAIN_REFINE, 0);
    a.inj_u = nullptr; a.inj_dphi = nullptr; a.inj_dn_fwd = nullptr; a.inj_dn_bwd = nullptr;
    a.obs = obs; a.accept_mask = nullptr; a.dS_out = nullptr;
    a.exact_mode = 0; a.inj_z = nullptr; a.filtered_strict = 0;
    a.obs_in = obs_in; a.epochs = epochs; a.wait_epoch = wait_epoch; a.signal_epoch = signal_epoch;
    a.grid_wait = (flags & SVB_OVERLAP_PREDECESSOR) ? 0 : 1;
    DeviceInfo info;
    int rc = get_device_info(info);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (villain_stream_serves(a, false, false, 8)) return launch_villain_stream_by_size(a, st, info);
    switch (N) {
        case 16: return launch_villain_filtered<16, 16, 1>(a, st, info);
        case 32: return launch_villain_filtered<32, SVB_FILT_MINB32, SVB_FILT_STAGES>(a, st, info);
        case 64: return launch_villain_filtered<64, 2, 1>(a, st, info);
        default:
            if (villain_strips_serves(a)) return launch_villain_strips(a, st, info);
            return launch_villain_cluster<128, SVB_CLUSTER_CL, SVB_CLUSTER_TPB, SVB_CLUSTER_STAGES>(a, st, info);
    }
}

// SiteUpdate, LinkUpdate, ExactUpdate (generator/villain/{site,link,exact}.py): see include/svb200.h
extern "C" int svb_villain_decoupled(int kind, void* phi, int32_t* n, int64_t chains, int N, double kappa, const double* kappa_chain,
                                     int W, double interval_phi, int interval, int n_sweeps, uint64_t seed, uint64_t sweep0,
                                     uint64_t chain0, int rng_mode, int path, const double* inj_u, const double* inj_dphi,
                                     const int32_t* inj_a, double* obs, uint8_t* accept_mask, double* dS_out, void* stream) {
    if (kind < SVB_VU_SITE || kind > SVB_VU_EXACT) return fail(SVB_E_PARAM, "svb_villain_decoupled: kind %d", kind);
    if (!phi || !n) return fail(SVB_E_NULL, "svb_villain_decoupled: phi and n are required");
    if (chains < 0 || N < 3 || N > 32768) return fail(SVB_E_SHAPE, "svb_villain_decoupled: chains=%lld N=%d", (long long)chains, N);
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_villain_decoupled: kappa must be positive");
    if (W < 1) return fail(SVB_E_PARAM, "svb_villain_decoupled: W must be a finite integer >= 1 (got %d)", W);
    // 2 * interval <= 256: the choice's remainder leads the uniform, whose resolution given the choice is 2 interval 2^-32 <= 2^-24
    if (kind != SVB_VU_SITE && (interval < 1 || interval > 128)) return fail(SVB_E_PARAM, "svb_villain_decoupled: interval must be in [1, 128]");
    if (kind == SVB_VU_SITE && !(interval_phi >= 0)) return fail(SVB_E_PARAM, "svb_villain_decoupled: interval_phi must be >= 0");
    if (n_sweeps < 0) return fail(SVB_E_PARAM, "svb_villain_decoupled: n_sweeps < 0");
    if (rng_mode != SVB_RNG_PHILOX && rng_mode != SVB_RNG_INJECTED) return fail(SVB_E_PARAM, "svb_villain_decoupled: rng_mode");
    if (rng_mode == SVB_RNG_INJECTED && (!inj_u || (kind == SVB_VU_SITE ? !inj_dphi : !inj_a)))
        return fail(SVB_E_NULL, "svb_villain_decoupled: injected mode needs inj_u and inj_dphi (SITE) or inj_a (LINK, EXACT)");
    if (path < SVB_PATH_AUTO || path > SVB_PATH_GLOBAL) return fail(SVB_E_PARAM, "svb_villain_decoupled: path");
    if (chains == 0 || n_sweeps == 0) return SVB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    if (kind == SVB_VU_LINK) {
        if (accept_mask) return fail(SVB_E_UNSUPPORTED, "svb_villain_decoupled: no accept_mask for LINK");
        const long long V = (long long)N * N;
        if (rng_mode == SVB_RNG_PHILOX && !dS_out && path != SVB_PATH_GLOBAL && (N == 16 || N == 32 || N == 64) &&
            ((uintptr_t)phi % 16 == 0) && ((uintptr_t)n % 16 == 0)) {
            // the chain staged in shared memory, all sweeps and the observables in one launch (svb_villain_link.cuh)
            LinkArgs la;
            la.phi = reinterpret_cast<const double*>(phi); la.n = n; la.chains = chains; la.N = N; la.kappa = kappa; la.kappa_chain = kappa_chain;
            la.W = W; la.interval = interval; la.seed = seed; la.chain0 = chain0; la.obs = obs; la.sweep = sweep0;
            la.inj_u = nullptr; la.inj_c = nullptr; la.dS_out = nullptr;
            DeviceInfo info;
            const int rc = get_device_info(info);
            if (rc) return rc;
            switch (N) {
                case 16: return launch_villain_link_smem<16, 16>(la, n_sweeps, st, info);
                case 32: return launch_villain_link_smem<32, 8>(la, n_sweeps, st, info);
                default: return launch_villain_link_smem<64, 3>(la, n_sweeps, st, info);
            }
        }
        if (rng_mode == SVB_RNG_PHILOX && !dS_out && path != SVB_PATH_GLOBAL && N == 128 && ((uintptr_t)phi % 16 == 0) &&
            ((uintptr_t)n % 16 == 0)) {
            // strips of 32 rows (svb_villain_link.cuh): counters by atomics into a zeroed record, state columns afterwards
            LinkArgs la;
            la.phi = reinterpret_cast<const double*>(phi); la.n = n; la.chains = chains; la.N = N; la.kappa = kappa; la.kappa_chain = kappa_chain;
            la.W = W; la.interval = interval; la.seed = seed; la.chain0 = chain0; la.obs = obs; la.sweep = sweep0;
            la.inj_u = nullptr; la.inj_c = nullptr; la.dS_out = nullptr;
            DeviceInfo info;
            int rc = get_device_info(info);
            if (rc) return rc;
            if (obs) {
                villain_zero_counters_kernel<<<(unsigned)((chains + 255) / 256), 256, 0, st>>>(obs, chains);
                SVB_CUDA_TRY(cudaGetLastError());
            }
            rc = launch_villain_link_strip<128, 32, 2>(la, n_sweeps, st, info);
            if (rc) return rc;
            if (obs) return launch_villain_obs<double>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs, 1, st);
            return SVB_OK;
        }
        const long long blocks = chains * ((2 * V + 255) / 256);
        if (blocks > 0x7fffffffLL) return fail(SVB_E_SHAPE, "svb_villain_decoupled: too many blocks");
        if (obs) {
            villain_zero_counters_kernel<<<(unsigned)((chains + 255) / 256), 256, 0, st>>>(obs, chains);
            SVB_CUDA_TRY(cudaGetLastError());
        }
        LinkArgs a;
        a.phi = reinterpret_cast<const double*>(phi); a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain;
        a.W = W; a.interval = interval; a.seed = seed; a.chain0 = chain0; a.obs = obs;
        for (int s = 0; s < n_sweeps; ++s) {
            a.sweep = sweep0 + (uint64_t)s;
            a.dS_out = (s == n_sweeps - 1) ? dS_out : nullptr;
            if (rng_mode == SVB_RNG_INJECTED) {
                a.inj_u = inj_u + (long long)s * chains * 2 * V;
                a.inj_c = inj_a + (long long)s * chains * 2 * V;
                villain_link_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(a);
            } else {
                a.inj_u = nullptr; a.inj_c = nullptr;
                villain_link_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(a);
            }
            SVB_CUDA_TRY(cudaGetLastError());
        }
        if (obs) return launch_villain_obs<double>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs, 1, st);
        return SVB_OK;
    }

    // SITE = the neighbourhood move with no dn proposals; EXACT = the site kernels with (dphi = 0, dn = d z) proposals.
    // STRICT arithmetic: fp64, one rounding per numpy operation.
    VillainArgs a;
    a.phi = phi; a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W;
    a.interval_phi = (kind == SVB_VU_SITE) ? interval_phi : 0.0;
    a.interval_n = (kind == SVB_VU_SITE) ? 0 : interval;
    a.n_sweeps = n_sweeps; a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    if (kind == SVB_VU_SITE) villain_rng_setup(a, STREAM_VILLAIN_SITE, STREAM_VILLAIN_SITE_REFINE, 0);
    else villain_rng_setup(a, STREAM_VILLAIN_EXACT, STREAM_VILLAIN_EXACT_REFINE, 0);
    a.inj_u = inj_u; a.inj_dphi = inj_dphi; a.inj_dn_fwd = nullptr; a.inj_dn_bwd = nullptr;
    a.obs = obs; a.accept_mask = accept_mask; a.dS_out = dS_out;
    a.exact_mode = (kind == SVB_VU_EXACT) ? 1 : 0;
    a.inj_z = inj_a;
    // Philox draws without debug outputs: the fp32-filtered kernel whose cold path decides in STRICT arithmetic -- the same
    // decisions, fields and records (the acceptance sum to 1e-5) as the STRICT kernels below, at 2-3 times their speed
    a.filtered_strict = (rng_mode == SVB_RNG_PHILOX && !accept_mask && !dS_out) ? 1 : 0;
    a.obs_in = nullptr; a.epochs = nullptr; a.wait_epoch = 0; a.signal_epoch = 0; a.grid_wait = 1;
    return dispatch_villain<double>(a, rng_mode, SVB_ARITH_STRICT, path, st);
}

extern "C" int svb_villain_cohomology(const void* phi, int32_t* n, int64_t chains, int N, double kappa, const double* kappa_chain,
                                      int interval, uint64_t seed, uint64_t sweep, uint64_t chain0, int rng_mode,
                                      const double* inj_u, const int32_t* inj_h, double* counters, double* dS_out, void* stream) {
    if (!phi || !n) return fail(SVB_E_NULL, "svb_villain_cohomology: phi and n are required");
    if (chains < 0 || N < 3 || N > 32768) return fail(SVB_E_SHAPE, "svb_villain_cohomology: chains=%lld N=%d", (long long)chains, N);
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_villain_cohomology: kappa must be positive");
    if (interval < 1 || interval > 1024) return fail(SVB_E_PARAM, "svb_villain_cohomology: interval");
    if (rng_mode != SVB_RNG_PHILOX && rng_mode != SVB_RNG_INJECTED) return fail(SVB_E_PARAM, "svb_villain_cohomology: rng_mode");
    if (rng_mode == SVB_RNG_INJECTED && (!inj_u || !inj_h)) return fail(SVB_E_NULL, "svb_villain_cohomology: injected mode needs inj_u, inj_h");
    if (chains == 0) return SVB_OK;
    CohomologyArgs a;
    a.phi = reinterpret_cast<const double*>(phi); a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain;
    a.interval = interval; a.seed = seed; a.sweep = sweep; a.chain0 = chain0; a.inj_u = inj_u; a.inj_h = inj_h;
    a.counters = counters; a.dS_out = dS_out;
    const long long blocks = (2 * chains + 3) / 4;
    if (blocks > 0x7fffffffLL) return fail(SVB_E_SHAPE, "svb_villain_cohomology: too many blocks");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (rng_mode == SVB_RNG_INJECTED) villain_cohomology_kernel<true><<<(unsigned)blocks, 128, 0, st>>>(a);
    else villain_cohomology_kernel<false><<<(unsigned)blocks, 128, 0, st>>>(a);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

extern "C" int svb_villain_observables(const void* phi, int phi_dtype, const int32_t* n, int64_t chains, int N, double kappa,
                                       const double* kappa_chain, double* obs, void* stream) {
    if (!phi || !n || !obs) return fail(SVB_E_NULL, "svb_villain_observables: phi, n, obs are required");
    if (chains < 0 || N < 3 || N > 32768) return fail(SVB_E_SHAPE, "svb_villain_observables: chains=%lld N=%d", (long long)chains, N);
    if (phi_dtype != SVB_F64 && phi_dtype != SVB_F32) return fail(SVB_E_DTYPE, "svb_villain_observables: phi dtype %d", phi_dtype);
    if (chains == 0) return SVB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (phi_dtype == SVB_F64)
        return launch_villain_obs<double>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs, 0, st);
    return launch_villain_obs<float>(reinterpret_cast<const float*>(phi), n, chains, N, kappa, kappa_chain, obs, 0, st);
}

extern "C" int svb_villain_draws(int64_t chains, int N, int W, double interval_phi, int interval_n, uint64_t seed,
                                 uint64_t sweep, uint64_t chain0, double* u, double* dphi, int32_t* dn, void* stream) {
    if (!u || !dphi || !dn) return fail(SVB_E_NULL, "svb_villain_draws: outputs required");
    if (chains <= 0 || N < 1) return fail(SVB_E_SHAPE, "svb_villain_draws: shape");
    if (interval_n < 0 || interval_n > 31) return fail(SVB_E_PARAM, "svb_villain_draws: interval_n");
    const long long total = (long long)chains * N * N;
    villain_draws_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        chains, N, W, interval_phi, interval_n, seed, sweep, chain0, u, dphi, dn);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}


// Host-buffer form of svb_villain_sweep: the reference's `step(cfg)` contract (host arrays in, host arrays out) at
// batch scale.  The chains are cut into `n_chunks` chunks; chunk i is copied host->device, swept and copied back on
// stream i % n_streams, so the H2D copy of one chunk, the sweep of another and the D2H copy of a third overlap
// (PCIe is full duplex).  Host buffers must be pinned for the copies to be asynchronous.  Nothing is synchronised
// here: the caller synchronises the streams it passed.
extern "C" int svb_villain_sweep_host(void* phi_host, int phi_dtype, int32_t* n_host, double* obs_host,
                                      void* phi_dev, int32_t* n_dev, double* obs_dev,
                                      int64_t chains, int N, double kappa, int W, double interval_phi, int interval_n,
                                      int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0, int arith_mode,
                                      int n_chunks, void* const* streams, int n_streams) {
    if (!phi_host || !n_host || !phi_dev || !n_dev) return fail(SVB_E_NULL, "svb_villain_sweep_host: field buffers are required");
    if (!streams || n_streams < 1 || n_chunks < 1) return fail(SVB_E_PARAM, "svb_villain_sweep_host: streams / chunks");
    if ((obs_host == nullptr) != (obs_dev == nullptr)) return fail(SVB_E_NULL, "svb_villain_sweep_host: obs_host and obs_dev go together");
    if (phi_dtype != SVB_F64 && phi_dtype != SVB_F32) return fail(SVB_E_DTYPE, "svb_villain_sweep_host: phi dtype %d", phi_dtype);
    if (chains <= 0) return SVB_OK;
    if (n_chunks > chains) n_chunks = (int)chains;
    const size_t V = (size_t)N * N;
    const size_t esz = (phi_dtype == SVB_F64) ? 8 : 4;
    for (int i = 0; i < n_chunks; ++i) {
        const int64_t lo = chains * i / n_chunks, hi = chains * (i + 1) / n_chunks;
        if (hi <= lo) continue;
        const int64_t cnt = hi - lo;
        cudaStream_t st = reinterpret_cast<cudaStream_t>(streams[i % n_streams]);
        char* ph = reinterpret_cast<char*>(phi_host) + lo * V * esz;
        char* pd = reinterpret_cast<char*>(phi_dev) + lo * V * esz;
        int32_t* nh = n_host + lo * 2 * V;
        int32_t* nd = n_dev + lo * 2 * V;
        SVB_CUDA_TRY(cudaMemcpyAsync(pd, ph, cnt * V * esz, cudaMemcpyHostToDevice, st));
        SVB_CUDA_TRY(cudaMemcpyAsync(nd, nh, cnt * 2 * V * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        int rc = svb_villain_sweep(pd, phi_dtype, nd, cnt, N, kappa, nullptr, W, interval_phi, interval_n, n_sweeps, seed, sweep0,
                                   chain0 + (uint64_t)lo, SVB_RNG_PHILOX, arith_mode, SVB_PATH_AUTO, nullptr, nullptr, nullptr,
                                   nullptr, obs_dev ? obs_dev + lo * SVB_VOBS_COUNT : nullptr, nullptr, nullptr, st);
        if (rc) return rc;
        SVB_CUDA_TRY(cudaMemcpyAsync(ph, pd, cnt * V * esz, cudaMemcpyDeviceToHost, st));
        SVB_CUDA_TRY(cudaMemcpyAsync(nh, nd, cnt * 2 * V * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (obs_dev)
            SVB_CUDA_TRY(cudaMemcpyAsync(obs_host + lo * SVB_VOBS_COUNT, obs_dev + lo * SVB_VOBS_COUNT,
                                         cnt * SVB_VOBS_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    return SVB_OK;
}


// Tiled single-pass sweeps for lattices too large for shared memory (N a multiple of 32): ping-pong between the fields
// and a caller-provided workspace of the same size; the result always ends in (phi, n).
// allow_swap: an odd number of sweeps on the fast path may END in the workspace (*state_in_workspace = 1) instead of
// being copied back -- for callers that own both buffer pairs and exchange their roles (svb_villain_sweep_tiled_swap).
static int villain_sweep_tiled_impl(void* phi, int32_t* n, void* phi_ws, int32_t* n_ws, int64_t chains, int N, double kappa,
                                    const double* kappa_chain, int W, double interval_phi, int interval_n, int n_sweeps,
                                    uint64_t seed, uint64_t sweep0, uint64_t chain0, int arith_mode, double* obs,
                                    uint8_t* accept_mask, double* dS_out, void* stream, bool allow_swap, int* state_in_workspace) {
    if (state_in_workspace) *state_in_workspace = 0;
    if (!phi || !n || !phi_ws || !n_ws) return fail(SVB_E_NULL, "svb_villain_sweep_tiled: fields and workspace are required");
    if (chains < 0 || N < kTile || (N % kTile) != 0) return fail(SVB_E_SHAPE, "svb_villain_sweep_tiled: N=%d must be a multiple of %d", N, kTile);
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_villain_sweep_tiled: kappa must be positive");
    if (W < 1 || interval_n < 0 || interval_n > 31 || n_sweeps < 0) return fail(SVB_E_PARAM, "svb_villain_sweep_tiled: W / interval_n / n_sweeps");
    if (chains == 0 || n_sweeps == 0) return SVB_OK;
    const int tps = N / kTile;
    const long long blocks = (long long)chains * tps * tps;
    if (blocks > 0x7fffffffLL) return fail(SVB_E_SHAPE, "svb_villain_sweep_tiled: too many tiles (%lld)", blocks);
    VillainArgs a;
    a.phi = phi; a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W;
    a.interval_phi = interval_phi; a.interval_n = interval_n; a.n_sweeps = n_sweeps;
    a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    villain_rng_setup(a, STREAM_VILLAIN_NEIGHBORHOOD, STREAM_VILLAIN_REFINE, villain_wide(interval_n));
    a.inj_u = nullptr; a.inj_dphi = nullptr; a.inj_dn_fwd = nullptr; a.inj_dn_bwd = nullptr;
    a.obs = obs; a.accept_mask = nullptr; a.dS_out = nullptr;
    a.exact_mode = 0; a.inj_z = nullptr; a.filtered_strict = 0; a.obs_in = nullptr; a.epochs = nullptr; a.wait_epoch = 0; a.signal_epoch = 0; a.grid_wait = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (arith_mode != SVB_ARITH_STRICT && !accept_mask && !dS_out && !a.wide && N % 16 == 0 && ((uintptr_t)phi % 16 == 0) &&
        ((uintptr_t)n % 16 == 0) && villain_stream_enabled()) {
        // the streaming colour passes (svb_villain_stream.cuh): in place, no workspace, the state stays in (phi, n)
        DeviceInfo info;
        int rc = get_device_info(info);
        if (rc) return rc;
        if (obs) {
            villain_zero_record_kernel<<<(unsigned)((chains + 255) / 256), 256, 0, st>>>(obs, chains, 0);
            SVB_CUDA_TRY(cudaGetLastError());
        }
        // TMA-staged tiles where the tile shape divides the lattice (SVB_VILLAIN_PASS=stream: straight from global memory)
        const char* ep = getenv("SVB_VILLAIN_PASS");
        const bool tma = N % kTileCols == 0 && chains * 2 < 0x7fffffffLL && !(ep && ep[0] == 's');
        rc = tma ? launch_villain_tile_passes(a, nullptr, obs, st, info) : launch_villain_stream_passes(a, nullptr, obs, st, info);
        if (rc) return rc;
        if (obs) return launch_villain_obs<double>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs, 1, st);
        return SVB_OK;
    }
    // An even number of ping-pong sweeps ends in (phi, n); an odd count ends in the workspace.  The filtered kernel is fast
    // enough that copying the state back (two device-to-device copies) beats doing the last sweep in place with the
    // per-colour global path (330 vs 385 us for a config-4 shard); the fp64 kernels (STRICT, debug outputs) keep the latter.
    const bool odd = (n_sweeps & 1) != 0;
#ifndef SVB_NO_FILTERED_KERNEL
    const bool all_tiled = odd && arith_mode != SVB_ARITH_STRICT && !accept_mask && !dS_out && !a.wide;
#else
    const bool all_tiled = false;
#endif
    const bool copy_back = all_tiled && !allow_swap;
    const int n_tiled = all_tiled ? n_sweeps : (n_sweeps & ~1);
    const bool tail_global = odd && !all_tiled;
    if (all_tiled && allow_swap && state_in_workspace) *state_in_workspace = 1;
    if (obs) {
        villain_zero_record_kernel<<<(unsigned)((chains + 255) / 256), 256, 0, st>>>(obs, chains, 0);
        SVB_CUDA_TRY(cudaGetLastError());
    }
    double* bufp[2] = {reinterpret_cast<double*>(phi), reinterpret_cast<double*>(phi_ws)};
    int32_t* bufn[2] = {n, n_ws};
    for (int s = 0; s < n_tiled; ++s) {
        const int src = s & 1, dst = src ^ 1;
        const bool last = (s == n_sweeps - 1);
        a.accept_mask = last ? accept_mask : nullptr;
        a.dS_out = last ? dS_out : nullptr;
        const int fuse = (obs && last) ? 1 : 0;
        if (arith_mode == SVB_ARITH_STRICT)
            villain_tiled_kernel<true><<<(unsigned)blocks, 128, 0, st>>>(a, bufp[src], bufn[src], bufp[dst], bufn[dst], s, tps, fuse);
        else if (a.accept_mask || a.dS_out || a.wide)       // debug outputs, wide dn intervals: the fp64 kernel
            villain_tiled_kernel<false><<<(unsigned)blocks, 128, 0, st>>>(a, bufp[src], bufn[src], bufp[dst], bufn[dst], s, tps, fuse);
        else
#ifndef SVB_NO_FILTERED_KERNEL
        {
            const int rc = launch_villain_tiled_filtered(a, make_filter_consts(interval_phi, W, interval_n), bufp[src], bufn[src],
                                                         bufp[dst], bufn[dst], s, tps, fuse, blocks, st);
            if (rc) return rc;
        }
#else
            villain_tiled_kernel<false><<<(unsigned)blocks, 128, 0, st>>>(a, bufp[src], bufn[src], bufp[dst], bufn[dst], s, tps, fuse);
#endif
        SVB_CUDA_TRY(cudaGetLastError());
    }
    if (copy_back) {
        const size_t V = (size_t)N * N;
        SVB_CUDA_TRY(cudaMemcpyAsync(phi, phi_ws, (size_t)chains * V * sizeof(double), cudaMemcpyDeviceToDevice, st));
        SVB_CUDA_TRY(cudaMemcpyAsync(n, n_ws, (size_t)chains * 2 * V * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    if (tail_global) {
        const int V = N * N;
        const int bpc = (V / 2 + 255) / 256;
        const long long gblocks = (long long)bpc * chains;
        if (gblocks > 0x7fffffffLL) return fail(SVB_E_SHAPE, "svb_villain_sweep_tiled: too many blocks");
        a.accept_mask = accept_mask;
        a.dS_out = dS_out;
        for (int c = 0; c < 2; ++c) {
            if (arith_mode == SVB_ARITH_STRICT)
                villain_colour_pass_kernel<double, false, true><<<(unsigned)gblocks, 256, 0, st>>>(a, n_sweeps - 1, c, bpc, 1);
            else
                villain_colour_pass_kernel<double, false, false><<<(unsigned)gblocks, 256, 0, st>>>(a, n_sweeps - 1, c, bpc, 1);
            SVB_CUDA_TRY(cudaGetLastError());
        }
        if (obs) return launch_villain_obs<double>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs, 1, st);
    }
    return SVB_OK;
}

extern "C" int svb_villain_sweep_tiled(void* phi, int32_t* n, void* phi_ws, int32_t* n_ws, int64_t chains, int N, double kappa,
                                       const double* kappa_chain, int W, double interval_phi, int interval_n, int n_sweeps,
                                       uint64_t seed, uint64_t sweep0, uint64_t chain0, int arith_mode, double* obs,
                                       uint8_t* accept_mask, double* dS_out, void* stream) {
    return villain_sweep_tiled_impl(phi, n, phi_ws, n_ws, chains, N, kappa, kappa_chain, W, interval_phi, interval_n, n_sweeps, seed,
                                    sweep0, chain0, arith_mode, obs, accept_mask, dS_out, stream, false, nullptr);
}

// The same sweeps for a caller that owns both buffer pairs: after an odd number of sweeps on the fast path the state is
// left in (phi_ws, n_ws) and *state_in_workspace = 1 -- no copy back (two device-to-device copies of the whole state, a
// third of a single-sweep call at L=4096); otherwise the state is in (phi, n) and *state_in_workspace = 0.
extern "C" int svb_villain_sweep_tiled_swap(void* phi, int32_t* n, void* phi_ws, int32_t* n_ws, int64_t chains, int N, double kappa,
                                            const double* kappa_chain, int W, double interval_phi, int interval_n, int n_sweeps,
                                            uint64_t seed, uint64_t sweep0, uint64_t chain0, int arith_mode, double* obs,
                                            int* state_in_workspace, void* stream) {
    if (!state_in_workspace) return fail(SVB_E_NULL, "svb_villain_sweep_tiled_swap: state_in_workspace is required");
    return villain_sweep_tiled_impl(phi, n, phi_ws, n_ws, chains, N, kappa, kappa_chain, W, interval_phi, interval_n, n_sweeps, seed,
                                    sweep0, chain0, arith_mode, obs, nullptr, nullptr, stream, true, state_in_workspace);
}

// In-place sweeps by colour passes for lattices beyond a CTA (svb_villain_stream.cuh): see include/svb200.h.
extern "C" int svb_villain_sweep_inplace(void* phi, int32_t* n, int64_t chains, int N, double kappa, const double* kappa_chain, int W,
                                         double interval_phi, int interval_n, int n_sweeps, uint64_t seed, uint64_t sweep0,
                                         uint64_t chain0, double* obs, double* obs_in, void* stream) {
    if (!phi || !n) return fail(SVB_E_NULL, "svb_villain_sweep_inplace: phi and n are required");
    if (chains < 0 || N < 16 || N > 32768 || (N % 16) != 0)
        return fail(SVB_E_SHAPE, "svb_villain_sweep_inplace: N=%d must be a multiple of 16 (chains=%lld)", N, (long long)chains);
    if (((uintptr_t)phi % 16) || ((uintptr_t)n % 16)) return fail(SVB_E_ALIGN, "svb_villain_sweep_inplace: fields must be 16-byte aligned");
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_villain_sweep_inplace: kappa must be positive");
    if (W < 1 || !(interval_phi >= 0) || n_sweeps < 0) return fail(SVB_E_PARAM, "svb_villain_sweep_inplace: W / interval_phi / n_sweeps");
    if (interval_n < 0 || villain_wide(interval_n))
        return fail(SVB_E_UNSUPPORTED, "svb_villain_sweep_inplace: interval_n must be 0 or 1 (wider proposals: svb_villain_sweep)");
    if (obs_in && !obs) return fail(SVB_E_NULL, "svb_villain_sweep_inplace: obs_in needs obs (this launch's counters go there)");
    if (chains == 0 || n_sweeps == 0) return SVB_OK;
    VillainArgs a;
    a.phi = phi; a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W;
    a.interval_phi = interval_phi; a.interval_n = interval_n; a.n_sweeps = n_sweeps;
    a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    villain_rng_setup(a, STREAM_VILLAIN_NEIGHBORHOOD, STREAM_VILLAIN_REFINE, 0);
    a.inj_u = nullptr; a.inj_dphi = nullptr; a.inj_dn_fwd = nullptr; a.inj_dn_bwd = nullptr;
    a.obs = obs; a.accept_mask = nullptr; a.dS_out = nullptr;
    a.exact_mode = 0; a.inj_z = nullptr; a.filtered_strict = 0; a.obs_in = obs_in; a.epochs = nullptr; a.wait_epoch = 0; a.signal_epoch = 0; a.grid_wait = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceInfo info;
    int rc = get_device_info(info);
    if (rc) return rc;
    if (obs || obs_in) {    // the state columns of the arriving state are accumulated into obs_in, this launch's counters into obs
        SVB_CUDA_TRY(launch_pdl(villain_zero_inplace_records_kernel, (unsigned)((chains + 255) / 256), 256, 0, st, obs_in, obs, (long long)chains));
    }
    const char* ep = getenv("SVB_VILLAIN_PASS");
    const bool tma = N % kTileCols == 0 && chains * 2 < 0x7fffffffLL && !(ep && ep[0] == 's');
    rc = tma ? launch_villain_tile_passes(a, obs_in, obs, st, info) : launch_villain_stream_passes(a, obs_in, obs, st, info);
    if (rc) return rc;
    if (obs && !obs_in)     // the full record of the state after the sweeps: one more read of the state
        return launch_villain_obs<double>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs, 1, st);
    return SVB_OK;
}

// The same in-place sweeps as ONE launch per step: the phases follow one another down the lattice as a wavefront through L2
// (villain_tile_wave_kernel, svb_villain_stream.cuh): see include/svb200.h.
extern "C" long long svb_villain_wavefront_workspace(int64_t chains, int N, int n_sweeps, int with_obs_in) {
    if (chains < 0 || N < 128 || (N % svb::kTileCols) != 0 || n_sweeps < 1) return 0;
    return svb::villain_wave_workspace_ints(chains, N, n_sweeps, with_obs_in != 0);
}

extern "C" int svb_villain_sweep_wavefront(void* phi, int32_t* n, int64_t chains, int N, double kappa, const double* kappa_chain, int W,
                                           double interval_phi, int interval_n, int n_sweeps, uint64_t seed, uint64_t sweep0,
                                           uint64_t chain0, double* obs, double* obs_in, int32_t* workspace, int64_t workspace_ints,
                                           void* stream) {
    if (!phi || !n || !workspace) return fail(SVB_E_NULL, "svb_villain_sweep_wavefront: phi, n and the workspace are required");
    if (chains < 0 || N < 128 || N > 32768 || (N % kTileCols) != 0 || chains * 2 >= 0x7fffffffLL)
        return fail(SVB_E_SHAPE, "svb_villain_sweep_wavefront: N=%d must be a multiple of %d (chains=%lld)", N, kTileCols, (long long)chains);
    if (((uintptr_t)phi % 16) || ((uintptr_t)n % 16) || ((uintptr_t)workspace % 4))
        return fail(SVB_E_ALIGN, "svb_villain_sweep_wavefront: fields must be 16-byte aligned");
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_villain_sweep_wavefront: kappa must be positive");
    if (W < 1 || !(interval_phi >= 0) || n_sweeps < 0) return fail(SVB_E_PARAM, "svb_villain_sweep_wavefront: W / interval_phi / n_sweeps");
    if (interval_n < 0 || villain_wide(interval_n))
        return fail(SVB_E_UNSUPPORTED, "svb_villain_sweep_wavefront: interval_n must be 0 or 1 (wider proposals: svb_villain_sweep)");
    if (obs_in && !obs) return fail(SVB_E_NULL, "svb_villain_sweep_wavefront: obs_in needs obs (this launch's counters go there)");
    if (chains == 0 || n_sweeps == 0) return SVB_OK;
    VillainArgs a;
    a.phi = phi; a.n = n; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W;
    a.interval_phi = interval_phi; a.interval_n = interval_n; a.n_sweeps = n_sweeps;
    a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    villain_rng_setup(a, STREAM_VILLAIN_NEIGHBORHOOD, STREAM_VILLAIN_REFINE, 0);
    a.inj_u = nullptr; a.inj_dphi = nullptr; a.inj_dn_fwd = nullptr; a.inj_dn_bwd = nullptr;
    a.obs = obs; a.accept_mask = nullptr; a.dS_out = nullptr;
    a.exact_mode = 0; a.inj_z = nullptr; a.filtered_strict = 0; a.obs_in = obs_in; a.epochs = nullptr; a.wait_epoch = 0; a.signal_epoch = 0; a.grid_wait = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceInfo info;
    int rc = get_device_info(info);
    if (rc) return rc;
    if (obs || obs_in) {
        SVB_CUDA_TRY(launch_pdl(villain_zero_inplace_records_kernel, (unsigned)((chains + 255) / 256), 256, 0, st, obs_in, obs, (long long)chains));
    }
    rc = launch_villain_tile_wave(a, obs_in, obs, workspace, (long long)workspace_ints, st, info);
    if (rc) return rc;
    if (obs && !obs_in)     // the full record of the state after the sweeps: one more read of the state
        return launch_villain_obs<double>(reinterpret_cast<const double*>(phi), n, chains, N, kappa, kappa_chain, obs, 1, st);
    return SVB_OK;
}

#ifdef SVB_TRACE
// Evidence build only: copy the CTA time stamps of the overlapped launches to the host (launches x ctas x {start, end} u64).
extern "C" int svb_debug_trace_read(unsigned long long* out_host, int* launches, int* ctas) {
    *launches = svb::kTraceLaunches; *ctas = svb::kTraceCtas;
    if (out_host) SVB_CUDA_TRY(cudaMemcpyFromSymbol(out_host, svb::g_svb_trace, sizeof(unsigned long long) * svb::kTraceLaunches * svb::kTraceCtas * 2));
    return 0;
}
#endif
