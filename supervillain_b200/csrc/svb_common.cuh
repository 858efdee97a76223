// svb_common.cuh -- shared device/host helpers for libsvb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "svb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsvb200 is written for sm_100a (B200) only"
#endif

namespace svb {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SVB_CUDA_TRY(expr)                                             \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return ::svb::cuda_fail(_e, #expr);     \
    } while (0)

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------
// numpy's 2*np.pi as a double: 2 * 0x400921FB54442D18 (exact doubling)
#define SVB_TWO_PI 6.283185307179586476925286766559

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Counter-based: the draw for
// (site, chain, sweep) never depends on launch geometry, tiling, or the number of GPUs.
// ------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;   // one IMAD.WIDE.U32 on the device
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
}

template <int ROUNDS = 10>
__host__ __device__ __forceinline__ Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, hi0, lo0);
        mulhilo32(M1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// Stream ids folded into the top byte of counter word 3.  EVERY generator kind has its own pair (proposal stream,
// refinement stream), so generators that are given the same seed and advance their sweep counters in lockstep (the natural
// setup inside Sequentially) still consume disjoint Philox blocks: SiteUpdate never re-proposes NeighborhoodUpdate's dphi,
// ExactUpdate's z is not a function of SiteUpdate's uniform, and so on.
enum : uint32_t { STREAM_VILLAIN_NEIGHBORHOOD = 1u, STREAM_WORLDLINE_PLAQUETTE = 2u, STREAM_WORLDLINE_WRAPPING = 3u,
                  STREAM_VILLAIN_REFINE = 4u, STREAM_WORLDLINE_REFINE = 5u,
                  STREAM_VILLAIN_LINK = 6u, STREAM_VILLAIN_LINK_REFINE = 7u, STREAM_VILLAIN_COHOMOLOGY = 8u,
                  STREAM_VILLAIN_SITE = 9u, STREAM_VILLAIN_SITE_REFINE = 10u,
                  STREAM_VILLAIN_EXACT = 11u, STREAM_VILLAIN_EXACT_REFINE = 12u,
                  STREAM_WORLDLINE_VORTEX = 13u, STREAM_WORLDLINE_VORTEX_REFINE = 14u,
                  STREAM_WORLDLINE_COEXACT = 15u, STREAM_WORLDLINE_COEXACT_REFINE = 16u };

__host__ __device__ __forceinline__ Philox4 philox_site(uint64_t seed, uint64_t chain, uint64_t sweep,
                                                         uint32_t site, uint32_t stream_id) {
    // counter = (site, chain[31:0], sweep[31:0], stream<<24 | chain[39:32]<<16 | sweep[47:32])
    uint32_t c3 = (stream_id << 24) | ((uint32_t)((chain >> 32) & 0xFFu) << 16) | (uint32_t)((sweep >> 32) & 0xFFFFu);
    return philox4x32<10>(site, (uint32_t)chain, (uint32_t)sweep, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// ------------------------------------------------------------------------------------------
// min(1, exp(x)): the Metropolis acceptance probability.
// Degree-11 near-minimax polynomial on [-ln2/2, ln2/2] (Chebyshev-node interpolation, 0.15 ulp
// approximation error; tests/test_gpu_villain.py checks it against libm) after the usual
// x = n ln2 + z reduction.  The coefficients live in constant memory so the FMAs read them as
// constant-bank operands instead of rebuilding 64-bit immediates in registers per call.
// Arguments are clamped to [-708, 0]: above 0 the clipped result is exactly 1; below -708 the true
// value (< 4e-308) is far under the smallest uniform the generators can draw (2^-53), so such a
// proposal is rejected either way and its contribution to acceptance statistics is negligible.
// ------------------------------------------------------------------------------------------
static __constant__ double SVB_EXP_C[12] = {
    1.0, 1.0, 0.5000000000000019, 0.1666666666666668, 0.0416666666664881, 0.008333333333319601,
    0.0013888888952314775, 0.00019841269890047113, 2.4801485482328494e-05, 2.755724091857897e-06,
    2.763263963904103e-07, 2.5110037605963777e-08};

__device__ __forceinline__ double exp_clipped(double x) {
    x = fmin(fmax(x, -708.0), 0.0);
    const double magic = 6755399441055744.0;                     // 1.5 * 2^52: rounds x log2(e) to an integer
    double t = fma(x, 1.4426950408889634, magic);
    const int n = __double2loint(t);
    t -= magic;
    double z = fma(t, -0.6931471805599453, x);
    z = fma(t, -2.3190468138462996e-17, z);
    double p = SVB_EXP_C[11];
#pragma unroll
    for (int k = 10; k >= 0; --k) p = fma(p, z, SVB_EXP_C[k]);
    // scale by 2^n (n in [-1022, 0], p in [0.70, 1.42]): add n to the exponent field
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// ------------------------------------------------------------------------------------------
// The Metropolis test u < min(1, e^-dS) without evaluating the fp64 exponential for (almost) every proposal.
// u < e^-dS  <=>  dS < -ln u.  -ln u is formed in fp32 (one MUFU.LG2); whenever dS is farther from it than a guard band
// that covers every fp32 error in the comparison by a factor > 20, the decision is already certain.  Only inside the band
// (probability ~1e-6 per proposal) is the exact fp64 test evaluated, so every decision equals the fp64 decision.
// The acceptance probability itself, wanted only for the generator's report() statistic, comes from MUFU.EX2 in fp32
// (relative accuracy ~1e-6); STRICT arithmetic keeps the fp64 exponential for both.
// ------------------------------------------------------------------------------------------
// Returns +1 (accept), 0 (reject) or -1 (undecided: evaluate exactly).  `u_mid` is the midpoint of a bracket of
// half-width `u_halfwidth_rel` * u_mid (relative) that is known to contain u; pass 0 for an exactly known u.
#ifndef SVB_NO_LOG_FILTER
__device__ __forceinline__ int metropolis_log_filter(double dS, double u_mid, float u_halfwidth_rel, double& prob) {
    const float dSf = (float)dS, uf = (float)u_mid;
    prob = (double)fminf(exp2f(-1.4426950408889634f * dSf), 1.0f);
    const float L = -0.6931471805599453f * __log2f(uf);
    const float band = 1e-5f * (1.0f + L + fabsf(dSf)) + 1.5f * u_halfwidth_rel;
    if (dSf > L + band) return 0;
    if (dSf < L - band) return 1;
    return -1;
}
__device__ __forceinline__ bool metropolis_filtered(double dS, double u, double& prob) {
    const int r = metropolis_log_filter(dS, u, 0.0f, prob);
    if (r >= 0) return r != 0;
    return u < exp_clipped(-dS);
}
#else
__device__ __forceinline__ int metropolis_log_filter(double dS, double u_mid, float u_halfwidth_rel, double& prob) {
    prob = exp_clipped(-dS);
    return -1;
}
__device__ __forceinline__ bool metropolis_filtered(double dS, double u, double& prob) {
    prob = exp_clipped(-dS);
    return u < prob;
}
#endif

// ------------------------------------------------------------------------------------------
// Lazily refined Metropolis uniforms.  A proposal carries only the LEADING 32 bits f of its uniform, so u is known to lie
// in [f, f + 1] 2^-32; that decides u < A unless A falls inside the bracket (probability 2^-32 per proposal).  Only then
// are the trailing bits generated: e = word `word` of the Philox block with the same counter in the refinement stream,
//   u = min(fl(f + (e + 1/2) 2^-32) 2^-32, 1 - 2^-53).
// Every kernel and the oracle implement exactly this rule, so decisions are those of the full 64-bit uniform.
// ------------------------------------------------------------------------------------------
struct LazyUniform {
    uint32_t f;        // leading 32 bits
    uint32_t c0;       // counter word 0 of the block the proposal came from
    uint32_t word;     // which word of the refinement block belongs to this proposal
};
struct RefineCtx {
    unsigned long long seed, chain, sweep;
    uint32_t stream;   // the refinement stream of the generator kind that made the proposal
    uint32_t wide;     // Villain proposals with wide dn intervals: the leading bits came from the refinement block too (svb_villain.cu)
};

static __device__ __noinline__ double refined_uniform(uint32_t f, uint32_t c0, uint32_t word, uint32_t stream_id, unsigned long long seed,
                                               unsigned long long chain, unsigned long long sweep) {
    const Philox4 p = philox_site(seed, chain, sweep, c0, stream_id);
    const uint32_t e = (word == 0) ? p.x : (word == 1) ? p.y : (word == 2) ? p.z : p.w;
    const double frac = __dmul_rn(__dadd_rn((double)e, 0.5), 2.3283064365386963e-10);          // (e + 1/2) 2^-32
    const double u = __dmul_rn(__dadd_rn((double)f, frac), 2.3283064365386963e-10);
    return fmin(u, 0.99999999999999988898);                                                     // 1 - 2^-53
}

// u < A decided from the bracket of u, refining only when A falls inside it.
__device__ __forceinline__ bool decide_lazy(double A, const LazyUniform& lu, uint32_t stream_id, const RefineCtx& rc) {
    const double u_lo = __dmul_rn((double)lu.f, 2.3283064365386963e-10);
    const double u_hi = __dadd_rn(u_lo, 2.3283064365386963e-10);
    if (A > u_hi) return true;
    if (A <= u_lo) return false;
    return refined_uniform(lu.f, lu.c0, lu.word, stream_id, rc.seed, rc.chain, rc.sweep) < A;
}

// ------------------------------------------------------------------------------------------
// checkerboard colouring (supervillain/lattice/compact.py:192-239, D = 2)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int n_colours(int N) { return (N & 1) ? 4 : 2; }

__host__ __device__ __forceinline__ int site_colour(int x0, int x1, int N) {
    if ((N & 1) == 0) return (x0 + x1) & 1;
    // FFT coordinates: index i -> i for i <= N/2, else i - N   (lattice/__init__.py:4-9)
    int h = N >> 1;
    int c0 = (x0 <= h) ? x0 : x0 - N;
    int c1 = (x1 <= h) ? x1 : x1 - N;
    int parity = (c0 + c1) & 1;                       // two's complement: correct for negatives
    int mixed = ((c0 >= 0) != (c1 >= 0)) ? 1 : 0;     // compact.py:36-53 with D = 2
    return 2 * mixed + parity;
}

// ------------------------------------------------------------------------------------------
// warp / block reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_i32(int v) { return __reduce_add_sync(0xffffffffu, v); }      // REDUX
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum K doubles per thread across the block; result valid in thread 0.  scratch: K * 32 doubles.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    if (nwarps == 1) return;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) scratch[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = (lane < nwarps) ? scratch[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// ------------------------------------------------------------------------------------------
// TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier.  Sizes and addresses must be
// multiples of 16 bytes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, completion tracked by the bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// all but the most recent bulk group of this thread have completed (their global writes are performed)
__device__ __forceinline__ void bulk_wait1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// Overlapped launches (svb_*_sweep_overlapped): a launch may begin while its predecessor in the stream is still running
// (programmatic dependent launch) and the data is ordered per chain through a caller-owned epoch word per chain.
//   prologue      every thread, first thing: let the next launch start once all CTAs of this one are resident, and --
//                 unless the caller vouches for the predecessor -- wait for everything before this launch
//   peek / wait   one thread: a chain may be loaded once its epoch reads wait_epoch; `peek` early so the common case
//                 costs no round trip at the point of the load
//   publish_all   one warp, when the CTA is done and behind a block barrier before which the thread that issued the bulk
//                 stores has seen them complete: ONE gpu-scope release fence, then a relaxed store per chain
// ------------------------------------------------------------------------------------------
struct OverlapArgs {
    uint32_t* epochs;          // nullptr: an ordinary launch
    uint32_t wait_epoch, signal_epoch;
    int grid_wait;
};

__device__ __forceinline__ void overlap_prologue(const OverlapArgs& o) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (o.grid_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ uint32_t overlap_peek(const OverlapArgs& o, long long chain) {
    uint32_t e = o.wait_epoch;
    if (!o.grid_wait) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(o.epochs + chain) : "memory");
    return e;
}
__device__ __forceinline__ void overlap_wait(const OverlapArgs& o, long long chain, uint32_t seen) {
    if (o.grid_wait) return;
    uint32_t e = seen;
    unsigned ns = 32, naps = 0;
    while (e != o.wait_epoch) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(o.epochs + chain) : "memory");
        if (e == o.wait_epoch) break;
        __nanosleep(ns);
        if (ns < 1024) ns *= 2;
        if (++naps > (1u << 21)) __trap();      // a producer that never comes is a caller error: fail, do not hang the GPU
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void overlap_publish_all(const OverlapArgs& o, int lane, int count, long long first, long long stride) {
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    for (int i = lane; i < count; i += 32)
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(o.epochs + first + (long long)i * stride), "r"(o.signal_epoch)
                     : "memory");
}

}  // namespace svb
