// svb_observables.cu -- two-point observables that are not simple per-chain scalars.
//
// Lattice.correlation (supervillain/lattice/compact.py:465-536), C[r] = N^-2 sum_x conj(s[x]) s[x - r], for
//   Spin_Spin.Villain        s = exp(i phi)             supervillain/observable/spin.py:28-42
//   Winding_Winding.Villain  s = dn                     supervillain/observable/winding.py:77-86
//   Vortex_Vortex.Worldline  s = exp(2 pi i v / W)      supervillain/observable/vortex.py:22-37
// The reference evaluates it with three FFTs.  Here: the direct O(N^4) sum out of shared memory for small or
// non-power-of-two lattices (exact to rounding), a shared-memory FFT for N = 16, 32, 64, and a three-launch FFT for
// power-of-two lattices up to N = 4096.

#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)
#include <cstring>
#include <type_traits>
#include "svb_common.cuh"

namespace svb {

// KIND: SVB_CORR_SPIN    s = exp(i phi)                       field = phi (chains,1,N,N)
//       SVB_CORR_WINDING s = dn = d(n) (real)                  field = n   (chains,2,N,N) int32
//       SVB_CORR_VORTEX  s = exp(2 pi i v / W)                 field = v   (chains,1,N,N) int32
template <typename real, int KIND>
__global__ void __launch_bounds__(256) correlation_kernel(const real* __restrict__ field, long long chains, int N, int W,
                                                          double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int V = N * N;
    double* sre = reinterpret_cast<double*>(smem_raw);
    double* sim = sre + V;
    const double inv_V = 1.0 / (double)V;
    for (long long chain = blockIdx.x; chain < chains; chain += gridDim.x) {
        const real* g = field + chain * (KIND == SVB_CORR_WINDING ? 2 : 1) * V;
        for (int i = threadIdx.x; i < V; i += blockDim.x) {
            if (KIND == SVB_CORR_WINDING) {
                const int x0 = i / N, x1 = i - x0 * N;
                const int i0 = ((x0 + 1 == N) ? 0 : x0 + 1) * N + x1, i1 = x0 * N + ((x1 + 1 == N) ? 0 : x1 + 1);
                // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x])
                sre[i] = (double)(((long long)g[V + i0] - (long long)g[V + i]) - ((long long)g[i1] - (long long)g[i]));
                sim[i] = 0.0;
            } else {
                double s, c;
                // np.exp(2j * np.pi * v / W): the phase is rounded as (2 pi v) / W   (observable/vortex.py:34)
                const double ang = (KIND == SVB_CORR_VORTEX) ? (SVB_TWO_PI * (double)g[i]) / (double)W : (double)g[i];
                sincos(ang, &s, &c);
                sre[i] = c;
                sim[i] = s;
            }
        }
        __syncthreads();
        for (int r = threadIdx.x; r < V; r += blockDim.x) {
            const int r0 = r / N, r1 = r - r0 * N;
            double are = 0.0, aim = 0.0;
            for (int x0 = 0; x0 < N; ++x0) {
                int y0 = x0 - r0;
                if (y0 < 0) y0 += N;
                const double* rowx_re = sre + x0 * N;
                const double* rowx_im = sim + x0 * N;
                const double* rowy_re = sre + y0 * N;
                const double* rowy_im = sim + y0 * N;
                int y1 = (r1 == 0) ? 0 : N - r1;      // y1 = (0 - r1) mod N, then walks forward with x1
                for (int x1 = 0; x1 < N; ++x1) {
                    const double a = rowx_re[x1], b = rowx_im[x1];     // s[x]
                    const double cc = rowy_re[y1], dd = rowy_im[y1];   // s[x - r]
                    // conj(a + ib)(cc + i dd) = (a cc + b dd) + i (a dd - b cc)
                    are += a * cc + b * dd;
                    aim += a * dd - b * cc;
                    y1 = (y1 + 1 == N) ? 0 : y1 + 1;
                }
            }
            double* o = out + (chain * V + r) * 2;
            o[0] = are * inv_V;
            o[1] = aim * inv_V;
        }
        __syncthreads();
    }
}

// 64-point transforms (the two-step transforms of the big lattices below, and both dimensions of an L = 64 lattice) as TWO radix-8 passes with the butterflies of a pass in
// registers: a radix-2 stage costs about 18 instructions per element and a shared-memory round trip, three of them in registers
// cost a third of that.  Position p = 8 a + b.  Decimation in frequency: DFT8 over a (stride 8) for every b, twiddle
// W_64^{a b}, DFT8 over b -- position p then holds frequency (p >> 3) + 8 (p & 7) (digit reversal in base 8, fft64_freq).
// Decimation in time takes that order back to the natural one: DFT8 over b, the same twiddle, DFT8 over a.
__device__ __forceinline__ int fft64_freq(int p) { return ((p & 7) << 3) | (p >> 3); }
__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(const double2 a, const double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(const double2 a, const double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(const double2 z) { return make_double2(z.y, -z.x); }                 // z (-i)
// X[k] = sum_n x[n] e^{-2 pi i n k / 8}, natural order in and out
__device__ __forceinline__ void dft8(double2 (&x)[8]) {
    constexpr double S = 0.70710678118654752440;
    const double2 a0 = cadd(x[0], x[4]), a1 = cadd(x[1], x[5]), a2 = cadd(x[2], x[6]), a3 = cadd(x[3], x[7]);
    const double2 b0 = csub(x[0], x[4]), t1 = csub(x[1], x[5]), t2 = csub(x[2], x[6]), t3 = csub(x[3], x[7]);
    const double2 b1 = make_double2((t1.x + t1.y) * S, (t1.y - t1.x) * S);                                      // w
    const double2 b2 = mul_mi(t2);                                                                              // w^2 = -i
    const double2 b3 = make_double2((t3.y - t3.x) * S, -(t3.x + t3.y) * S);                                     // w^3
    const double2 c0 = cadd(a0, a2), c1 = cadd(a1, a3), d0 = csub(a0, a2), d1 = mul_mi(csub(a1, a3));
    const double2 e0 = cadd(b0, b2), e1 = cadd(b1, b3), f0 = csub(b0, b2), f1 = mul_mi(csub(b1, b3));
    x[0] = cadd(c0, c1); x[2] = cadd(d0, d1); x[4] = csub(c0, c1); x[6] = csub(d0, d1);
    x[1] = cadd(e0, e1); x[3] = cadd(f0, f1); x[5] = csub(e0, e1); x[7] = csub(f0, f1);
}
// X[k] = sum_n x[n] e^{-2 pi i n k / 4}, natural order in and out
__device__ __forceinline__ void dft4(double2 (&x)[4]) {
    const double2 c0 = cadd(x[0], x[2]), c1 = cadd(x[1], x[3]), d0 = csub(x[0], x[2]), d1 = mul_mi(csub(x[1], x[3]));
    x[0] = cadd(c0, c1); x[1] = cadd(d0, d1); x[2] = csub(c0, c1); x[3] = csub(d0, d1);
}
// X[k] = sum_n x[n] e^{-2 pi i n k / 16}, natural order in and out: n = 4 a + b, DFT4 over a, W_16^{a' b}, DFT4 over b, k = a' + 4 b'
__device__ __forceinline__ void dft16(double2 (&x)[16]) {
    constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, S = 0.70710678118654752440;
    double2 y[4][4];                                                  // y[a'][b]
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double2 t[4] = {x[b], x[4 + b], x[8 + b], x[12 + b]};
        dft4(t);
#pragma unroll
        for (int a = 0; a < 4; ++a) y[a][b] = t[a];
    }
    // W_16^{a' b}: exponents 1, 2, 3 (a' = 1), 2, 4, 6 (a' = 2), 3, 6, 9 (a' = 3)
    y[1][1] = cmul(y[1][1], make_double2(C1, -S1));
    y[1][2] = make_double2((y[1][2].x + y[1][2].y) * S, (y[1][2].y - y[1][2].x) * S);
    y[1][3] = cmul(y[1][3], make_double2(S1, -C1));
    y[2][1] = make_double2((y[2][1].x + y[2][1].y) * S, (y[2][1].y - y[2][1].x) * S);
    y[2][2] = mul_mi(y[2][2]);
    y[2][3] = make_double2((y[2][3].y - y[2][3].x) * S, -(y[2][3].x + y[2][3].y) * S);
    y[3][1] = cmul(y[3][1], make_double2(S1, -C1));
    y[3][2] = make_double2((y[3][2].y - y[3][2].x) * S, -(y[3][2].x + y[3][2].y) * S);
    y[3][3] = cmul(y[3][3], make_double2(-C1, S1));
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double2 t[4] = {y[a][0], y[a][1], y[a][2], y[a][3]};
        dft4(t);
#pragma unroll
        for (int b = 0; b < 4; ++b) x[a + 4 * b] = t[b];
    }
}
template <int R> __device__ __forceinline__ void dft_r(double2 (&x)[R]);
template <> __device__ __forceinline__ void dft_r<16>(double2 (&x)[16]) { dft16(x); }
template <> __device__ __forceinline__ void dft_r<8>(double2 (&x)[8]) { dft8(x); }
template <> __device__ __forceinline__ void dft_r<4>(double2 (&x)[4]) { dft4(x); }

// Transforms of length RA RB (64 = 8 x 8, 32 = 8 x 4, 16 = 4 x 4) of 2^lanes_log2 independent lines as two passes with the
// butterflies of a pass in registers: position p = RB a + b; DFT_RA over a (stride RB), twiddle W^{a b}, DFT_RB over b
// (decimation in frequency), or the two passes in the opposite order (decimation in time: takes the mixed order back to
// the natural one).  Element (line, position) at d[line * lane_stride + position * pos_stride]; wn[t] = e^{-2 pi i t / (RA RB)}.
// Consecutive threads take consecutive lines.
// `scale(line, position)` (optional): a factor for every element, applied to the OUTPUT of a DIF transform and to the INPUT of a
// DIT one -- the twiddle between the two steps of the split rides along instead of costing a pass over shared memory.
#ifndef SVB_FFT_UNROLL
#define SVB_FFT_UNROLL 1
#endif
constexpr int kFftUnroll = SVB_FFT_UNROLL;                            // butterflies of a pass in flight per thread
struct NoScale {};
struct Norm2 {};                                                      // as the Scale of a DIF transform: every output element x -> |x|^2
// `oscale(line, position)` (optional, DIT only): a factor for every element of the OUTPUT of a decimation-in-time transform.
template <bool DIF, int RA, int RB, class Scale = NoScale, class OutScale = NoScale>
__device__ __forceinline__ void fft_rr(double2* __restrict__ d, int pos_stride, int lane_stride, int lanes_log2,
                                       const double2* __restrict__ wn, Scale scale = Scale(), OutScale oscale = OutScale()) {
    constexpr bool NORM2 = std::is_same<Scale, Norm2>::value;
    constexpr bool SCALED = !std::is_same<Scale, NoScale>::value && !NORM2;
    static_assert(!(NORM2 && !DIF), "|.|^2 is taken on the output of a decimation-in-frequency transform");
    constexpr bool OSCALED = !std::is_same<OutScale, NoScale>::value;
    static_assert(!(OSCALED && DIF), "an output factor is for decimation-in-time transforms");
    const int lmask = (1 << lanes_log2) - 1;
    auto pass_a = [&](bool first) {                               // RA-point transforms over a, one per (line, b)
#pragma unroll kFftUnroll
        for (int i = threadIdx.x; i < (RB << lanes_log2); i += blockDim.x) {
            const int g = i >> lanes_log2, lane = i & lmask;
            double2* base = d + lane * lane_stride + g * pos_stride;
            const int step = RB * pos_stride;
            double2 x[RA];
#pragma unroll
            for (int j = 0; j < RA; ++j) x[j] = base[j * step];
            dft_r<RA>(x);
            if (first && g != 0) {                                // (first only in a DIF transform; a scale never applies here)
#pragma unroll
                for (int k = 1; k < RA; ++k) x[k] = cmul(x[k], wn[g * k]);
            }
            if constexpr (OSCALED) {
                if (!first) {
#pragma unroll
                    for (int k = 0; k < RA; ++k) x[k] = cmul(x[k], oscale(lane, g + RB * k));
                }
            }
#pragma unroll
            for (int k = 0; k < RA; ++k) base[k * step] = x[k];
        }
        __syncthreads();
    };
    auto pass_b = [&](bool first) {                               // RB-point transforms over b, one per (line, a)
#pragma unroll kFftUnroll
        for (int i = threadIdx.x; i < (RA << lanes_log2); i += blockDim.x) {
            const int g = i >> lanes_log2, lane = i & lmask;
            double2* base = d + lane * lane_stride + RB * g * pos_stride;
            double2 x[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) x[j] = base[j * pos_stride];
            if constexpr (SCALED && !DIF) {
                if (first) {
#pragma unroll
                    for (int j = 0; j < RB; ++j) x[j] = cmul(x[j], scale(lane, RB * g + j));
                }
            }
            dft_r<RB>(x);
            if (first && g != 0) {
#pragma unroll
                for (int k = 1; k < RB; ++k) x[k] = cmul(x[k], wn[g * k]);
            }
            if constexpr (SCALED && DIF) {
                if (!first) {
#pragma unroll
                    for (int k = 0; k < RB; ++k) x[k] = cmul(x[k], scale(lane, RB * g + k));
                }
            }
            if constexpr (NORM2) {
                if (!first) {
#pragma unroll
                    for (int k = 0; k < RB; ++k) x[k] = make_double2(x[k].x * x[k].x + x[k].y * x[k].y, 0.0);
                }
            }
#pragma unroll
            for (int k = 0; k < RB; ++k) base[k * pos_stride] = x[k];
        }
        __syncthreads();
    };
    if (DIF) { pass_a(true); pass_b(false); } else { pass_b(true); pass_a(false); }
}
template <bool DIF, class Scale = NoScale, class OutScale = NoScale>
__device__ __forceinline__ void fft64(double2* __restrict__ d, int pos_stride, int lane_stride, int lanes_log2,
                                      const double2* __restrict__ w64, Scale scale = Scale(), OutScale oscale = OutScale()) {
    fft_rr<DIF, 8, 8, Scale, OutScale>(d, pos_stride, lane_stride, lanes_log2, w64, scale, oscale);
}
template <int NN>
__device__ __forceinline__ void fft_rr_twiddles(double2* wn) {
    for (int k = threadIdx.x; k < NN; k += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)k / (double)NN, &sn, &cs);
        wn[k] = make_double2(cs, sn);
    }
}
__device__ __forceinline__ void fft64_twiddles(double2* w64) {
    for (int k = threadIdx.x; k < 64; k += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)k / 64.0, &sn, &cs);
        w64[k] = make_double2(cs, sn);
    }
}
// the frequency that position p of an n1-point transform of the split holds: two radix passes at n1 = 64, 32, 16 (position
// RB k_a + k_b holds k_a + RA k_b), radix-2 stages (bit reversal) below
__device__ __forceinline__ int split_freq(int p, int log2n1) {
    switch (log2n1) {
        case 6: return fft64_freq(p);
        case 5: return (p >> 2) + 8 * (p & 3);
        case 4: return (p >> 2) + 4 * (p & 3);
        default: return (int)(__brev((unsigned)p) >> (32 - log2n1));
    }
}
// the n1-point transforms of the split with radix passes (n1 = 64: 8 x 8, 32: 8 x 4, 16: 4 x 4); wn[t] = e^{-2 pi i t / n1}, t < n1
template <bool DIF, class Scale = NoScale, class OutScale = NoScale>
__device__ __forceinline__ void fft_n1(double2* __restrict__ d, int pos_stride, int lane_stride, int lanes_log2, int log2n1,
                                       const double2* __restrict__ wn, Scale scale = Scale(), OutScale oscale = OutScale()) {
    if (log2n1 == 6) fft_rr<DIF, 8, 8, Scale, OutScale>(d, pos_stride, lane_stride, lanes_log2, wn, scale, oscale);
    else if (log2n1 == 5) fft_rr<DIF, 8, 4, Scale, OutScale>(d, pos_stride, lane_stride, lanes_log2, wn, scale, oscale);
    else fft_rr<DIF, 4, 4, Scale, OutScale>(d, pos_stride, lane_stride, lanes_log2, wn, scale, oscale);
}
__device__ __forceinline__ void fft_n1_twiddles(double2* wn, int n1) {
    for (int k = threadIdx.x; k < n1; k += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)k / (double)n1, &sn, &cs);
        wn[k] = make_double2(cs, sn);
    }
}

// ------------------------------------------------------------------------------------------
// The same correlators by FFT for power-of-two lattices that fit shared memory (N = 16, 32, 64), the reference's own route
// (compact.py:465-536):  sum_x conj(s[x]) s[x - r] = N^-2 DFT[ |DFT s|^2 ](r), so C = N^-4 fft2(|fft2 s|^2).
// Two transforms per dimension out of shared memory: decimation in frequency (natural order in, mixed order out), the
// pointwise |.|^2 in that order, then decimation in time (mixed order in, natural out) -- no permutation pass.
// O(N^2 log N) per chain instead of O(N^4): 8192 chains of 32^2 take 6.4 ms by direct summation.
// ------------------------------------------------------------------------------------------
template <typename real, int KIND, int NT>
__global__ void __launch_bounds__(256) correlation_fft_kernel(const real* __restrict__ field, long long chains, int W,
                                                              double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N;
    static_assert(NT == 64 || NT == 32 || NT == 16, "the radix passes of fft_rr cover 16, 32 and 64");
    const double scale = 1.0 / ((double)V * (double)V);
    {
        // both dimensions as two register-resident radix passes each (fft_rr: 64 = 8 x 8, 32 = 8 x 4, 16 = 4 x 4) on a tile padded
        // to N + 1 columns: L = 64 x 1024 chains 175 -> 74 us against the radix-2 stages below
        constexpr int RS = N + 1, L2 = NT == 64 ? 6 : NT == 32 ? 5 : 4, RA = NT == 16 ? 4 : 8, RB = NT == 64 ? 8 : 4;
        double2* d = reinterpret_cast<double2*>(smem_raw);             // [N][N + 1]
        double2* wn = d + N * RS;
        fft_rr_twiddles<NT>(wn);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (long long chain = blockIdx.x; chain < chains; chain += gridDim.x) {
            const real* g = field + chain * (KIND == SVB_CORR_WINDING ? 2 : 1) * V;
            if (warp == 0) bulk_wait_read0();                           // the previous chain's result has left the tile
            __syncthreads();
            for (int i = threadIdx.x; i < V; i += blockDim.x) {
                const int x0 = i >> L2, x1 = i & (N - 1);
                if (KIND == SVB_CORR_WINDING) {
                    const int i0 = ((x0 + 1) & (N - 1)) * N + x1, i1 = x0 * N + ((x1 + 1) & (N - 1));
                    d[x0 * RS + x1] = make_double2((double)(((long long)g[V + i0] - (long long)g[V + i]) - ((long long)g[i1] - (long long)g[i])), 0.0);
                } else {
                    double sn, cs;
                    const double ang = (KIND == SVB_CORR_VORTEX) ? (SVB_TWO_PI * (double)g[i]) / (double)W : (double)g[i];
                    sincos(ang, &sn, &cs);
                    d[x0 * RS + x1] = make_double2(cs, sn);
                }
            }
            __syncthreads();
            fft_rr<true, RA, RB>(d, 1, RS, L2, wn);                     // along x1, a thread per row and group
            fft_rr<true, RA, RB>(d, RS, 1, L2, wn, Norm2());            // along x0, a thread per column and group; |.|^2 on the way out
            fft_rr<false, RA, RB>(d, RS, 1, L2, wn);
            fft_rr<false, RA, RB>(d, 1, RS, L2, wn, NoScale(), [&](int, int) { return make_double2(scale, 0.0); });
            // the result leaves as bulk copies (TMA), a row each, while the next chain's field is being read
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {
                double2* o = reinterpret_cast<double2*>(out) + chain * V;
                for (int r = lane; r < N; r += 32) bulk_s2g(o + r * N, d + r * RS, (uint32_t)(N * sizeof(double2)));
                bulk_commit();
            }
        }
        if (warp == 0) bulk_wait0();
    }
}

// ------------------------------------------------------------------------------------------
// L = 32 (config 2's lattices) with a WARP per chain and a whole line of 32 elements per thread: three visits of shared
// memory instead of eight -- rows; columns, |.|^2 and the columns of the second transform without leaving the registers;
// rows -- and no block-wide barrier anywhere.  dft32 = 8 x 4 in registers with compile-time twiddles.
// Measured SLOWER than correlation_fft_kernel<32> (122 against 102 us for 8192 chains: 150 registers leave 13 warps per SM to
// hide the latencies of 32 sincos and of the fp64 butterflies), so it is opt-in (SVB_CORR_FFT32_WARP=1) and kept as an
// independent implementation the tests compare with.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 w32(int m) {                       // e^{-2 pi i m / 32}, m < 32 (folded after unrolling)
    constexpr double C[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                             0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0};
    const int q = m >> 3, r = m & 7;
    const double c = C[r], sn = C[8 - r];
    return q == 0 ? make_double2(c, -sn) : q == 1 ? make_double2(-sn, -c) : q == 2 ? make_double2(-c, sn) : make_double2(sn, c);
}
// natural order in (n = 4 a + b); position p = 4 k_a + k_b of the result holds frequency k_a + 8 k_b
__device__ __forceinline__ void dft32(double2 (&x)[32]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double2 t[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) t[a] = x[4 * a + b];
        dft8(t);
#pragma unroll
        for (int a = 0; a < 8; ++a) x[4 * a + b] = (a * b == 0) ? t[a] : cmul(t[a], w32(a * b));
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        double2 t[4] = {x[4 * a], x[4 * a + 1], x[4 * a + 2], x[4 * a + 3]};
        dft4(t);
#pragma unroll
        for (int b = 0; b < 4; ++b) x[4 * a + b] = t[b];
    }
}
__host__ __device__ constexpr int dft32_freq(int p) { return (p >> 2) + 8 * (p & 3); }

template <typename real, int KIND>
__global__ void __launch_bounds__(32) correlation_fft32_warp_kernel(const real* __restrict__ field, long long chains, int W,
                                                                    double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = 32, V = N * N, RS = N + 1;
    double2* d = reinterpret_cast<double2*>(smem_raw);                 // [32][33]
    const int lane = threadIdx.x;
    const double scale = 1.0 / ((double)V * (double)V);
    for (long long chain = blockIdx.x; chain < chains; chain += gridDim.x) {
        const real* g = field + chain * (KIND == SVB_CORR_WINDING ? 2 : 1) * V;
        bulk_wait_read0();                                              // this lane's row of the previous result has left the tile
        __syncwarp();
#pragma unroll 4
        for (int x0 = 0; x0 < N; ++x0) {
            const int i = x0 * N + lane;
            if (KIND == SVB_CORR_WINDING) {
                const int i0 = ((x0 + 1) & (N - 1)) * N + lane, i1 = x0 * N + ((lane + 1) & (N - 1));
                d[x0 * RS + lane] = make_double2((double)(((long long)g[V + i0] - (long long)g[V + i]) - ((long long)g[i1] - (long long)g[i])), 0.0);
            } else {
                double sn, cs;
                const double ang = (KIND == SVB_CORR_VORTEX) ? (SVB_TWO_PI * (double)g[i]) / (double)W : (double)g[i];
                sincos(ang, &sn, &cs);
                d[x0 * RS + lane] = make_double2(cs, sn);
            }
        }
        __syncwarp();
        double2 x[32];
        // along x1: this lane's row
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = d[lane * RS + j];
        dft32(x);
#pragma unroll
        for (int p = 0; p < N; ++p) d[lane * RS + dft32_freq(p)] = x[p];
        __syncwarp();
        // along x0: this lane's column; |.|^2; the second transform along the same axis straight away
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = d[j * RS + lane];
        dft32(x);
        {
            double2 y[32];
#pragma unroll
            for (int p = 0; p < N; ++p) y[dft32_freq(p)] = make_double2(x[p].x * x[p].x + x[p].y * x[p].y, 0.0);
            dft32(y);
#pragma unroll
            for (int p = 0; p < N; ++p) d[dft32_freq(p) * RS + lane] = y[p];
        }
        __syncwarp();
        // along k1 -> r1: this lane's row of the result
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = d[lane * RS + j];
        dft32(x);
#pragma unroll
        for (int p = 0; p < N; ++p) d[lane * RS + dft32_freq(p)] = make_double2(x[p].x * scale, x[p].y * scale);
        fence_proxy_async();
        __syncwarp();
        bulk_s2g(reinterpret_cast<double2*>(out) + chain * V + lane * N, d + lane * RS, (uint32_t)(N * sizeof(double2)));
        bulk_commit();
    }
    bulk_wait0();
}

// ------------------------------------------------------------------------------------------
// Power-of-two lattices beyond shared memory (128 <= N <= 4096: configs 4 and 5) -- the same transform in three launches
// that use `out` (chains, N, N) complex128 itself as the workspace:
//   1. rows:    s from the field, decimation-in-frequency FFT of R rows per CTA, written in bit-reversed column order;
//   2. columns: C columns per CTA (padded in shared memory): DIF, |.|^2, decimation-in-time -- back in natural row order;
//   3. rows:    DIT of the bit-reversed rows, scaling by N^-4.
// No transposes and no bit-reversal passes; 16 + 32 + 32 B of traffic per site on top of reading the field.
// ------------------------------------------------------------------------------------------
template <bool DIF>
__device__ __forceinline__ void fft_lines(double2* __restrict__ d, int n, int log2n, int lines, int line_stride,
                                          const double2* __restrict__ tw) {
    const int half_n = n >> 1;
    for (int st = 0; st < log2n; ++st) {
        const int lh = DIF ? (log2n - 1 - st) : st;               // log2 of the butterfly half-span
        const int half = 1 << lh;
        const int tw_shift = log2n - 1 - lh;                      // twiddle e^{-2 pi i j / (2 half)} = tw[j << tw_shift]
        // a thread's butterflies of a stage share j, hence the twiddle, while half <= blockDim.x: loaded once -- tw[j << tw_shift] is
        // a strided read (up to 32 lanes on four banks), and per butterfly it made these kernels wait on shared memory
        const bool fixed_w = half <= (int)blockDim.x && (blockDim.x & (blockDim.x - 1)) == 0;
        const double2 w_fixed = tw[(threadIdx.x & (half - 1)) << tw_shift];
#pragma unroll 4
        for (int b = threadIdx.x; b < lines * half_n; b += blockDim.x) {
            const int line = b >> (log2n - 1), bb = b & (half_n - 1);
            const int group = bb >> lh, j = bb & (half - 1);
            double2* p0 = d + line * line_stride + (group << (lh + 1)) + j;
            double2* p1 = p0 + half;
            const double2 w = fixed_w ? w_fixed : tw[j << tw_shift];
            const double2 a = *p0, c = *p1;
            if (DIF) {
                const double dr = a.x - c.x, di = a.y - c.y;
                *p0 = make_double2(a.x + c.x, a.y + c.y);
                *p1 = make_double2(dr * w.x - di * w.y, dr * w.y + di * w.x);
            } else {
                const double tr = c.x * w.x - c.y * w.y, ti = c.x * w.y + c.y * w.x;
                *p0 = make_double2(a.x + tr, a.y + ti);
                *p1 = make_double2(a.x - tr, a.y - ti);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void fft_twiddles(double2* tw, int n) {
    for (int k = threadIdx.x; k < n / 2; k += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)k / (double)n, &sn, &cs);
        tw[k] = make_double2(cs, sn);
    }
}

// launch 1: R rows per work item
template <typename real, int KIND>
__global__ void __launch_bounds__(256) correlation_rows_forward_kernel(const real* __restrict__ field, long long chains, int N, int log2n,
                                                                       int W, int R, double2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* d = reinterpret_cast<double2*>(smem_raw);
    double2* tw = d + (size_t)R * N;
    fft_twiddles(tw, N);
    const long long V = (long long)N * N, groups = N / R, items = chains * groups;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long chain = item / groups;
        const int row0 = (int)(item - chain * groups) * R;
        const real* g = field + chain * (KIND == SVB_CORR_WINDING ? 2 : 1) * V;
        __syncthreads();
        for (int i = threadIdx.x; i < R * N; i += blockDim.x) {
            const int x0 = row0 + (i >> log2n), x1 = i & (N - 1);
            const long long at = (long long)x0 * N + x1;
            if (KIND == SVB_CORR_WINDING) {
                const long long i0 = (long long)((x0 + 1) & (N - 1)) * N + x1, i1 = (long long)x0 * N + ((x1 + 1) & (N - 1));
                d[i] = make_double2((double)(((long long)g[V + i0] - (long long)g[V + at]) - ((long long)g[i1] - (long long)g[at])), 0.0);
            } else {
                double sn, cs;
                const double ang = (KIND == SVB_CORR_VORTEX) ? (SVB_TWO_PI * (double)g[at]) / (double)W : (double)g[at];
                sincos(ang, &sn, &cs);
                d[i] = make_double2(cs, sn);
            }
        }
        __syncthreads();
        fft_lines<true>(d, N, log2n, R, N, tw);
        double2* o = out + chain * V + (long long)row0 * N;
        for (int i = threadIdx.x; i < R * N; i += blockDim.x) o[i] = d[i];
    }
}

// launch 2: C columns per work item, both column transforms and the pointwise square in one visit
__global__ void __launch_bounds__(256) correlation_columns_kernel(long long chains, int N, int log2n, int C, int log2c,
                                                                  double2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* d = reinterpret_cast<double2*>(smem_raw);
    const int stride = N + 1;                                     // padded: a column's rows are contiguous, columns a bank apart
    double2* tw = d + (size_t)C * stride;
    fft_twiddles(tw, N);
    const long long V = (long long)N * N, groups = N / C, items = chains * groups;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long chain = item / groups;
        const int col0 = (int)(item - chain * groups) * C;
        double2* o = out + chain * V + col0;
        __syncthreads();
        for (int i = threadIdx.x; i < C * N; i += blockDim.x) {
            const int r = i >> log2c, c = i & (C - 1);
            d[c * stride + r] = o[(long long)r * N + c];
        }
        __syncthreads();
        fft_lines<true>(d, N, log2n, C, stride, tw);
        for (int i = threadIdx.x; i < C * N; i += blockDim.x) {
            double2* q = d + (i >> log2n) * stride + (i & (N - 1));
            const double2 v = *q;
            *q = make_double2(v.x * v.x + v.y * v.y, 0.0);
        }
        __syncthreads();
        fft_lines<false>(d, N, log2n, C, stride, tw);
        for (int i = threadIdx.x; i < C * N; i += blockDim.x) {
            const int r = i >> log2c, c = i & (C - 1);
            o[(long long)r * N + c] = d[c * stride + r];
        }
    }
}

// launch 3
__global__ void __launch_bounds__(256) correlation_rows_inverse_kernel(long long chains, int N, int log2n, int R, double scale,
                                                                       double2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* d = reinterpret_cast<double2*>(smem_raw);
    double2* tw = d + (size_t)R * N;
    fft_twiddles(tw, N);
    const long long V = (long long)N * N, groups = N / R, items = chains * groups;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long chain = item / groups;
        double2* o = out + chain * V + (long long)((int)(item - chain * groups) * R) * N;
        __syncthreads();
        for (int i = threadIdx.x; i < R * N; i += blockDim.x) d[i] = o[i];
        __syncthreads();
        fft_lines<false>(d, N, log2n, R, N, tw);
        for (int i = threadIdx.x; i < R * N; i += blockDim.x) {
            const double2 v = d[i];
            o[i] = make_double2(v.x * scale, v.y * scale);
        }
    }
}

// ------------------------------------------------------------------------------------------
// The column transforms of the BIGGEST lattices (N >= 2048: config 5) in two steps.  A whole column of N complex128 in shared
// memory leaves room for 8192 / N columns per CTA -- at N = 4096 two, so correlation_columns_kernel moves 32-byte pieces of a
// row and runs at a tenth of the memory bandwidth.  Here the column index is split, r = 64 r1 + r2, k = k1 + n1 k2 (n1 = N / 64):
//   A.  (r2 fixed; rows 64 r1 + r2, strided)  n1-point DFT over r1, then the twiddle W_N^{r2 k1}
//   B.  (k1 fixed; 64 consecutive rows)       64-point DFT over r2 -- |.|^2 -- 64-point DFT over k2, then the twiddle W_N^{k1 r2'}
//   A'. (r2' fixed; rows 64 p + r2')          n1-point DFT over k1: row 64 r1' + r2' holds displacement r' in natural order
// three visits of the array instead of one, but every one of them in tiles of 32 columns: 512 contiguous bytes per row.
// Forward transforms are decimation in frequency (natural in, bit-reversed out), the second ones decimation in time (bit-
// reversed in, natural out), so position p of a block holds k1 = bitrev(p) and nothing is ever permuted.
// ------------------------------------------------------------------------------------------
constexpr int kSplitCols = 32, kSplitLog2Cols = 5, kSplitN2 = 64, kSplitLog2N2 = 6;

// Column tiles are rows of 128 .. 512 bytes a row of the lattice (or 64 of them) apart: as 1-D bulk copies a tile is 64 .. 512
// operations each way, and the copy engine of an SM retires one in about 10 ns -- 70 us of a 115 us kernel at N = 4096.  A
// 3-D tensor map (the array as doubles: (2 column, ...)) moves a tile with ONE operation each way.
__device__ __forceinline__ void box_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void box_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
                 "r"(smem_u32(smem_src))
                 : "memory");
}

// FFT of length n along the ROWS of a tile d[row][kSplitCols] (a thread per column and butterfly: conflict-free); tw = the
// n / 2 twiddles of the length-n transform
template <bool DIF, int LOG2W = kSplitLog2Cols>
__device__ __forceinline__ void fft_tile_rows(double2* __restrict__ d, int n, int log2n, const double2* __restrict__ tw,
                                              int row_stride = (1 << LOG2W)) {
    const int half_n = n >> 1;
    constexpr int WIDTH = 1 << LOG2W;
    for (int st = 0; st < log2n; ++st) {
        const int lh = DIF ? (log2n - 1 - st) : st;
        const int half = 1 << lh;
        const int tw_shift = log2n - 1 - lh;
#pragma unroll 4
        for (int b = threadIdx.x; b < WIDTH * half_n; b += blockDim.x) {
            const int c = b & (WIDTH - 1), bb = b >> LOG2W;
            const int group = bb >> lh, j = bb & (half - 1);
            double2* p0 = d + ((group << (lh + 1)) + j) * row_stride + c;
            double2* p1 = p0 + half * row_stride;
            const double2 w = tw[j << tw_shift];
            const double2 a = *p0, q = *p1;
            if (DIF) {
                const double dr = a.x - q.x, di = a.y - q.y;
                *p0 = make_double2(a.x + q.x, a.y + q.y);
                *p1 = make_double2(dr * w.x - di * w.y, dr * w.y + di * w.x);
            } else {
                const double tr = q.x * w.x - q.y * w.y, ti = q.x * w.y + q.y * w.x;
                *p0 = make_double2(a.x + tr, a.y + ti);
                *p1 = make_double2(a.x - tr, a.y - ti);
            }
        }
        __syncthreads();
    }
}

// The ROW transforms of the same lattices, split the same way inside a row held in shared memory: position r = 64 r1 + r2 of
// the row is element (r1, r2) of an n1 x 64 tile, so step A is a transform down the tile's rows (a thread per column:
// conflict-free, twiddles broadcast) and step B one of 64 contiguous elements per line (a warp per line).  The 4096-point
// transform with strided twiddles it replaces ran at a seventh of this.  W_N^{t}, t = r2 k1 < 4096, comes from two tables of 64:
// W_N^t = W_N^{64 (t >> 6)} W_N^{t & 63}.  FIRST: s from the field, A, twiddle, B, written in the mixed order that every
// later pass works in; !FIRST: B', twiddle, A', scaling -- natural order out.
// An item is R = 2^log2r consecutive rows (R N = 4096 elements for N <= 512, one row beyond): n1 R lines of 64.  Step A is two
// radix passes for n1 >= 16 (R = 1) and ONE pass of n1-point transforms in registers for n1 <= 8 (natural order in and out).
template <int N1>
__device__ __forceinline__ void dft_small(double2 (&x)[N1]) {
    if constexpr (N1 == 2) { const double2 a = x[0], b = x[1]; x[0] = cadd(a, b); x[1] = csub(a, b); }
    else if constexpr (N1 == 4) dft4(x);
    else if constexpr (N1 == 8) dft8(x);
    else if constexpr (N1 == 16) dft16(x);
}
// n1-point transforms over elements `stride` apart, one per work item: item i starts at d[start(i)], tw(i, k) multiplies
// output k of a decimation-in-frequency transform / input k of a decimation-in-time one.  PowerTw{w}: the factors of item i
// are the powers w(i)^k, built by repeated multiplication (k <= 15: within 2e-15) -- one look-up per item instead of one or two
// per element, which was a quarter of the row kernels' shared-memory traffic.
template <class F> struct PowerTw { F base; };
template <class F> __device__ __forceinline__ PowerTw<F> power_tw(F f) { return PowerTw<F>{f}; }
template <class T> struct is_power_tw : std::false_type {};
template <class F> struct is_power_tw<PowerTw<F>> : std::true_type {};
template <int N1, class Tw>
__device__ __forceinline__ void small_dft_twiddle(double2 (&x)[N1], int i, Tw& tw) {
    if constexpr (is_power_tw<Tw>::value) {
        const double2 w1 = tw.base(i);
        double2 wk = w1;
#pragma unroll
        for (int k = 1; k < N1; ++k) {
            x[k] = cmul(x[k], wk);
            if (k + 1 < N1) wk = cmul(wk, w1);
        }
    } else {
#pragma unroll
        for (int k = 0; k < N1; ++k) x[k] = cmul(x[k], tw(i, k));
    }
}
template <int N1, bool DIF, class Start, class Tw>
__device__ __forceinline__ void small_dft_pass(double2* __restrict__ d, int items, int stride, Start start, Tw tw) {
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        double2* base = d + start(i);
        double2 x[N1];
#pragma unroll
        for (int j = 0; j < N1; ++j) x[j] = base[j * stride];
        if constexpr (!DIF && !std::is_same<Tw, NoScale>::value) small_dft_twiddle<N1>(x, i, tw);
        dft_small<N1>(x);
        if constexpr (DIF && !std::is_same<Tw, NoScale>::value) small_dft_twiddle<N1>(x, i, tw);
#pragma unroll
        for (int k = 0; k < N1; ++k) base[k * stride] = x[k];
    }
    __syncthreads();
}
template <bool DIF, class Start, class Tw>
__device__ __forceinline__ void small_dft_pass_n1(int log2n1, double2* __restrict__ d, int items, int stride, Start start, Tw tw) {
    if (log2n1 == 1) small_dft_pass<2, DIF>(d, items, stride, start, tw);
    else if (log2n1 == 2) small_dft_pass<4, DIF>(d, items, stride, start, tw);
    else if (log2n1 == 3) small_dft_pass<8, DIF>(d, items, stride, start, tw);
    else small_dft_pass<16, DIF>(d, items, stride, start, tw);
}

template <typename real, int KIND, bool FIRST>
__global__ void __launch_bounds__(256, 3) correlation_rows_split_kernel(const real* __restrict__ field, long long chains, int N, int log2n1, int log2r,
                                                                        int W, double scale, double2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n1 = 1 << log2n1, lines = n1 << log2r, log2lines = log2n1 + log2r;
    const int E = N << log2r;                                           // elements of an item
    constexpr int RS = kSplitN2 + 1;                                    // row stride of the tile: lines a bank group apart
    double2* d = reinterpret_cast<double2*>(smem_raw);                 // [lines][64 (+1)]: the rows
    double2* tw1 = d + (size_t)lines * RS;                              // W_n1^t, t < n1 (radix passes, n1 >= 16)
    double2* w64 = tw1 + n1;                                            // W_64^t, t < 64
    double2* th = w64 + kSplitN2;                                       // W_N^{64 a}, a < 64
    double2* tl = th + kSplitN2;                                        // W_N^b, b < 64
    if (log2n1 >= 4) fft_n1_twiddles(tw1, n1);
    fft64_twiddles(w64);
    for (int i = threadIdx.x; i < kSplitN2; i += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)((i * kSplitN2) & (N - 1)) / (double)N, &sn, &cs);
        th[i] = make_double2(cs, sn);
        sincospi(-2.0 * (double)i / (double)N, &sn, &cs);
        tl[i] = make_double2(cs * scale, sn * scale);                   // every element meets W_N^t exactly once: the scaling rides along
    }
    // the rows travel between global memory and the padded tile as bulk copies (TMA), a line of 64 elements each, issued by
    // warp 0: nothing of them passes through registers, and the next item is on its way while this one is being written back
    __shared__ __align__(8) uint64_t ld_bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&ld_bar, 1); fence_mbar_init(); }
    uint32_t ld_parity = 0;
    constexpr uint32_t kLineBytes = kSplitN2 * sizeof(double2);
    auto load_rows = [&](const double2* src) {
        if (lane == 0) mbar_expect_tx(&ld_bar, (uint32_t)E * (uint32_t)sizeof(double2));
        __syncwarp();
        for (int l = lane; l < lines; l += 32) bulk_g2s(d + l * RS, src + l * kSplitN2, kLineBytes, &ld_bar);
    };
    auto store_rows = [&](double2* dst) {
        for (int l = lane; l < lines; l += 32) bulk_s2g(dst + l * kSplitN2, d + l * RS, kLineBytes);
        bulk_commit();
    };
    const long long V = (long long)N * N, items = (chains * N) >> log2r;
    const int items_per_chain = N >> log2r;
    auto pad = [&](int i) { return i + (i >> kSplitLog2N2); };          // element i of the item (64 per line) in the padded tile
    // W_N^{r2 k1}; with radix passes position p holds k1 = split_freq(p)
    auto tw_nat = [&](int r2, int k1) {
        const int t = r2 * k1;
        return cmul(th[t >> kSplitLog2N2], tl[t & (kSplitN2 - 1)]);
    };
    auto tw_of = [&](int r2, int p) { return tw_nat(r2, split_freq(p, log2n1)); };
    // the n1-point transforms down the tile (one per row and r2) together with the twiddle between the two steps
    auto outer = [&](auto dif_tag) {
        constexpr bool DIF = decltype(dif_tag)::value;
        if (log2n1 >= 4) fft_n1<DIF>(d, RS, 1, kSplitLog2N2, log2n1, tw1, tw_of);
        else
            small_dft_pass_n1<DIF>(log2n1, d, kSplitN2 << log2r, RS,
                                   [&](int i) { return (i >> kSplitLog2N2) * n1 * RS + (i & (kSplitN2 - 1)); },
                                   [&](int i, int k) { return tw_nat(i & (kSplitN2 - 1), k); });
    };
    __syncthreads();
    if (!FIRST && warp == 0 && blockIdx.x < items) load_rows(out + (long long)blockIdx.x * E);
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long chain = item / items_per_chain, next = item + gridDim.x;
        const int x0 = (int)(item - chain * items_per_chain) << log2r;  // the first row of the item
        double2* o = out + item * E;
        if (FIRST) {
            const real* g = field + chain * (KIND == SVB_CORR_WINDING ? 2 : 1) * V;
            if (KIND != SVB_CORR_WINDING && threadIdx.x == 0 && next < items)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(field + next * E), "r"((uint32_t)E * (uint32_t)sizeof(real)) : "memory");
            if (warp == 0) bulk_wait_read0();                            // the previous item has left the tile
            __syncthreads();
            if (KIND == SVB_CORR_WINDING) {
                for (int e = threadIdx.x; e < E; e += blockDim.x) {
                    const int r = x0 + (e >> (log2n1 + kSplitLog2N2)), x1 = e & (N - 1);
                    const long long at = (long long)r * N + x1;
                    const long long i0 = (long long)((r + 1) & (N - 1)) * N + x1, i1 = (long long)r * N + ((x1 + 1) & (N - 1));
                    d[pad(e)] = make_double2((double)(((long long)g[V + i0] - (long long)g[V + at]) - ((long long)g[i1] - (long long)g[at])), 0.0);
                }
            } else {
                constexpr int kBatch = 8;                                // loads in flight per thread ahead of the sincos
                const real* rows = g + (long long)x0 * N;
                for (int eb = threadIdx.x; eb < E; eb += kBatch * blockDim.x) {
                    real v[kBatch];
#pragma unroll
                    for (int j = 0; j < kBatch; ++j) {
                        const int e = eb + j * blockDim.x;
                        v[j] = e < E ? rows[e] : (real)0;
                    }
#pragma unroll
                    for (int j = 0; j < kBatch; ++j) {
                        const int e = eb + j * blockDim.x;
                        if (e < E) {
                            double sn, cs;
                            const double ang = (KIND == SVB_CORR_VORTEX) ? (SVB_TWO_PI * (double)v[j]) / (double)W : (double)v[j];
                            sincos(ang, &sn, &cs);
                            d[pad(e)] = make_double2(cs, sn);
                        }
                    }
                }
            }
            __syncthreads();
            outer(std::true_type());
            fft64<true>(d, 1, RS, log2lines, w64);                      // 64 contiguous elements per line, a thread per line and group
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) store_rows(o);
        } else {
            if (threadIdx.x == 0 && next < items)                        // (its copy into the tile is issued once this item has left)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(out + next * E), "r"((uint32_t)E * (uint32_t)sizeof(double2)) : "memory");
            mbar_wait(&ld_bar, ld_parity);
            ld_parity ^= 1u;
            fft64<false>(d, 1, RS, log2lines, w64);
            outer(std::false_type());
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {
                store_rows(o);
                if (next < items) {
                    bulk_wait_read0();                                   // ... and only then may the next item land in the tile
                    load_rows(out + next * E);
                }
            }
        }
    }
    if (warp == 0) bulk_wait0();
}

// visits A (FIRST: DIF over r1, then W_N^{r2 bitrev(p)}) and A' (!FIRST: DIT over k1): a tile = (chain, r2, 32 columns)
template <bool FIRST>
__global__ void __launch_bounds__(256, 4) correlation_split_outer_kernel(const __grid_constant__ CUtensorMap map, long long chains, int N, int log2n1) {
    extern __shared__ __align__(128) unsigned char box_smem[];
    const int n1 = 1 << log2n1;
    double2* d = reinterpret_cast<double2*>(box_smem);                 // [n1][32]: one box (2 x 32 doubles, 1, n1) of the map
    double2* tw = d + (size_t)n1 * kSplitCols;                          // W_n1^t, t < n1 (radix passes, n1 >= 16) or t < n1 / 2 (radix-2 stages)
    double2* tq = tw + n1;                                              // n1 twiddles W_N^{r2 k1(p)}
    if (log2n1 >= 4) fft_n1_twiddles(tw, n1); else fft_twiddles(tw, n1);
    const long long V = (long long)N * N;
    const int col_blocks = N / kSplitCols;
    const long long items = chains * kSplitN2 * col_blocks;
    // a tile travels as one box of the tensor map (columns, r2, chain n1 + r1), issued by one thread; the next tile is requested
    // the moment this one has left shared memory
    __shared__ __align__(8) uint64_t ld_bar;
    if (threadIdx.x == 0) { mbar_init(&ld_bar, 1); fence_mbar_init(); }
    uint32_t ld_parity = 0;
    auto load_tile = [&](long long item) {                              // thread 0
        const long long chain = item / (kSplitN2 * col_blocks);
        const int rem = (int)(item - chain * (kSplitN2 * col_blocks));
        const int r2 = rem / col_blocks, c0 = (rem - r2 * col_blocks) * kSplitCols;
        mbar_expect_tx(&ld_bar, (uint32_t)(n1 * kSplitCols * sizeof(double2)));
        box_load_3d(d, &map, 2 * c0, r2, (int)(chain * n1), &ld_bar);
    };
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < items) load_tile(blockIdx.x);
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long next = item + gridDim.x;
        const long long chain = item / (kSplitN2 * col_blocks);
        const int rem = (int)(item - chain * (kSplitN2 * col_blocks));
        const int r2 = rem / col_blocks, c0 = (rem - r2 * col_blocks) * kSplitCols;
        if (FIRST)
            for (int p = threadIdx.x; p < n1; p += blockDim.x) {        // (the previous tile's last pass ended in a barrier)
                const int k1 = split_freq(p, log2n1);
                double sn, cs;
                sincospi(-2.0 * (double)((r2 * k1) & (N - 1)) / (double)N, &sn, &cs);
                tq[p] = make_double2(cs, sn);
            }
        mbar_wait(&ld_bar, ld_parity);
        ld_parity ^= 1u;
        if (log2n1 >= 4) {
            if constexpr (FIRST) fft_n1<true>(d, kSplitCols, 1, kSplitLog2Cols, log2n1, tw, [&](int, int p) { return tq[p]; });
            else fft_n1<false>(d, kSplitCols, 1, kSplitLog2Cols, log2n1, tw);
        } else {
            if (FIRST) __syncthreads();                                 // tq
            fft_tile_rows<FIRST>(d, n1, log2n1, tw);
            if (FIRST) {
                for (int i = threadIdx.x; i < n1 * kSplitCols; i += blockDim.x) d[i] = cmul(d[i], tq[i >> kSplitLog2Cols]);
                __syncthreads();
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            box_store_3d(&map, 2 * c0, r2, (int)(chain * n1), d);
            bulk_commit();
            if (next < items) {
                bulk_wait_read0();
                load_tile(next);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait0();
}

// visit B: a tile = (chain, p, 32 columns) = 64 consecutive rows: DIF over r2, |.|^2, DIT over k2, W_N^{bitrev(p) r2'}
__global__ void __launch_bounds__(256, 4) correlation_split_inner_kernel(const __grid_constant__ CUtensorMap map, long long chains, int N, int log2n1) {
    extern __shared__ __align__(128) unsigned char box_smem[];
    double2* d = reinterpret_cast<double2*>(box_smem);                 // [64][32]: one box (2 x 32 doubles, 64, 1) of the map
    double2* w64 = d + (size_t)kSplitN2 * kSplitCols;                   // W_64^t, t < 64
    double2* tq = w64 + kSplitN2;                                       // 64 twiddles W_N^{k1 r2'}
    fft64_twiddles(w64);
    const long long V = (long long)N * N;
    const int n1 = 1 << log2n1, col_blocks = N / kSplitCols;
    const long long items = chains * n1 * col_blocks;
    __shared__ __align__(8) uint64_t ld_bar;
    if (threadIdx.x == 0) { mbar_init(&ld_bar, 1); fence_mbar_init(); }
    uint32_t ld_parity = 0;
    auto load_tile = [&](long long item) {                              // thread 0
        const long long chain = item / (n1 * col_blocks);
        const int rem = (int)(item - chain * (n1 * col_blocks));
        const int p = rem / col_blocks, c0 = (rem - p * col_blocks) * kSplitCols;
        mbar_expect_tx(&ld_bar, (uint32_t)(kSplitN2 * kSplitCols * sizeof(double2)));
        box_load_3d(d, &map, 2 * c0, 0, (int)(chain * n1) + p, &ld_bar);
    };
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < items) load_tile(blockIdx.x);
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long next = item + gridDim.x;
        const long long chain = item / (n1 * col_blocks);
        const int rem = (int)(item - chain * (n1 * col_blocks));
        const int p = rem / col_blocks, c0 = (rem - p * col_blocks) * kSplitCols;
        const int k1 = split_freq(p, log2n1);
        for (int r = threadIdx.x; r < kSplitN2; r += blockDim.x) {       // (the previous tile's last pass ended in a barrier)
            double sn, cs;
            sincospi(-2.0 * (double)((k1 * r) & (N - 1)) / (double)N, &sn, &cs);
            tq[r] = make_double2(cs, sn);
        }
        mbar_wait(&ld_bar, ld_parity);
        ld_parity ^= 1u;
        fft64<true>(d, kSplitCols, 1, kSplitLog2Cols, w64, Norm2());    // ... |.|^2 on the way out of its second pass
        fft64<false>(d, kSplitCols, 1, kSplitLog2Cols, w64, NoScale(), [&](int, int r) { return tq[r]; });
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            box_store_3d(&map, 2 * c0, 0, (int)(chain * n1) + p, d);
            bulk_commit();
            if (next < items) {
                bulk_wait_read0();
                load_tile(next);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait0();
}

// The columns of 128 <= N <= 512 in ONE visit: a tile = (chain, CW = 4096 / N columns) holds whole columns, 4096 elements.
// Row r = 64 r1 + r2 of the tile lives in slot (r2, r1) of a 64 x 64 array whose line index is L = r1 CW + c, so that the steps
// of the split are: A, n1-point transforms in registers over r1 (elements CW apart) and W_N^{r2 k1}; B, 64-point transforms
// over r2 for all 64 lines at once (two radix-8 passes, |.|^2 on the way out), the same back (W_N^{k1 r2'} on the way out);
// A', n1-point transforms over k1.  Rows travel as bulk copies (TMA) of CW elements, straight into and out of their slots.
__global__ void __launch_bounds__(256, 3) correlation_columns_fused_kernel(const __grid_constant__ CUtensorMap map, long long chains, int N, int log2n1) {
    extern __shared__ __align__(128) unsigned char box_smem[];
    const int log2cw = kSplitLog2N2 - log2n1, CW = 1 << log2cw, n1 = 1 << log2n1;
    double2* d = reinterpret_cast<double2*>(box_smem);                 // [64 r2][64 lines]: one box (2 CW doubles, n1, 64) of the map
    double2* w64 = d + kSplitN2 * kSplitN2;                             // W_64^t, t < 64
    double2* tN = w64 + kSplitN2;                                       // W_N^t, t < N
    fft64_twiddles(w64);
    fft_n1_twiddles(tN, N);
    __shared__ __align__(8) uint64_t ld_bar;
    if (threadIdx.x == 0) { mbar_init(&ld_bar, 1); fence_mbar_init(); }
    uint32_t ld_parity = 0;
    const int col_blocks = N >> log2cw;
    const long long items = chains * col_blocks;
    auto load_tile = [&](long long item) {                              // thread 0: the map's dimensions are (column, chain n1 + r1, r2)
        const long long chain = item / col_blocks;
        mbar_expect_tx(&ld_bar, (uint32_t)(kSplitN2 * kSplitN2 * sizeof(double2)));
        box_load_3d(d, &map, (int)(item - chain * col_blocks) << (log2cw + 1), (int)(chain * n1), 0, &ld_bar);
    };
    auto start = [&](int i) { return ((i >> log2cw) << kSplitLog2N2) + (i & (CW - 1)); };      // (r2, c): line c of slot row r2
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < items) load_tile(blockIdx.x);
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long next = item + gridDim.x;
        mbar_wait(&ld_bar, ld_parity);
        ld_parity ^= 1u;
        small_dft_pass_n1<true>(log2n1, d, kSplitN2 << log2cw, CW, start, [&](int i, int k) { return tN[(i >> log2cw) * k]; });
        fft64<true>(d, kSplitN2, 1, kSplitLog2N2, w64, Norm2());
        fft64<false>(d, kSplitN2, 1, kSplitLog2N2, w64, NoScale(), [&](int L, int q) { return tN[(L >> log2cw) * q]; });
        small_dft_pass_n1<false>(log2n1, d, kSplitN2 << log2cw, CW, start, NoScale());
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long chain = item / col_blocks;
            box_store_3d(&map, (int)(item - chain * col_blocks) << (log2cw + 1), (int)(chain * n1), 0, d);
            bulk_commit();
            if (next < items) {
                bulk_wait_read0();
                load_tile(next);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------
// The same two kernels with radix-16 butterflies: lines of LL = RA RB = 256 (16 x 16; 128 = 16 x 8 at N = 128) elements
// instead of 64, so a row of N = n1 LL elements is an n1-point register pass (n1 = N / LL <= 16, none at n1 = 1) and TWO radix
// passes -- three visits of shared memory where the kernels above make three (N <= 512) or four (N >= 1024), at twice the
// registers per thread and two CTAs per SM.  An item is always 4096 elements (4096 / N rows, 4096 / LL lines).
// ------------------------------------------------------------------------------------------
template <typename real, int KIND, bool FIRST, int RA, int RB>
__global__ void __launch_bounds__(256, 2) correlation_rows_r16_kernel(const real* __restrict__ field, long long chains, int N, int log2n1, int W,
                                                                      double scale, double2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LL = RA * RB, LOG2LL = LL == 256 ? 8 : 7, RS = LL + 1, E = 4096, lines = E / LL, LOG2LINES = 12 - LOG2LL;
    static_assert(LL == 256 || LL == 128, "lines of 256 or 128 elements");
    const int n1 = 1 << log2n1, log2n = log2n1 + LOG2LL;
    double2* d = reinterpret_cast<double2*>(smem_raw);                 // [lines][LL (+1)]
    double2* wl = d + (size_t)lines * RS;                               // W_LL^t, t < LL
    double2* th = wl + LL;                                              // W_N^{64 a}, a < 64
    double2* tl = th + kSplitN2;                                        // W_N^b, b < 64
    real* stage = reinterpret_cast<real*>(tl + kSplitN2);               // FIRST, spin fields: the NEXT item's 4096 field values (TMA)
    constexpr bool STAGED = FIRST && KIND != SVB_CORR_WINDING;
    fft_n1_twiddles(wl, LL);
    for (int i = threadIdx.x; i < kSplitN2; i += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)((i * kSplitN2) & (N - 1)) / (double)N, &sn, &cs);
        th[i] = make_double2(cs, sn);
        sincospi(-2.0 * (double)i / (double)N, &sn, &cs);
        tl[i] = make_double2(cs, sn);
    }
    __shared__ __align__(8) uint64_t ld_bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&ld_bar, 1); fence_mbar_init(); }
    uint32_t ld_parity = 0;
    constexpr uint32_t kLineBytes = LL * sizeof(double2);
    auto load_field = [&](long long item) {                             // thread 0: an item's field values are contiguous
        mbar_expect_tx(&ld_bar, (uint32_t)(E * sizeof(real)));
        bulk_g2s(stage, field + item * E, (uint32_t)(E * sizeof(real)), &ld_bar);
    };
    auto load_rows = [&](const double2* src) {
        if (lane == 0) mbar_expect_tx(&ld_bar, (uint32_t)E * (uint32_t)sizeof(double2));
        __syncwarp();
        for (int l = lane; l < lines; l += 32) bulk_g2s(d + l * RS, src + l * LL, kLineBytes, &ld_bar);
    };
    auto store_rows = [&](double2* dst) {
        for (int l = lane; l < lines; l += 32) bulk_s2g(dst + l * LL, d + l * RS, kLineBytes);
        bulk_commit();
    };
    const long long V = (long long)N * N, items = (chains * V) >> 12;
    const int items_per_chain = (int)(V >> 12), log2r = 12 - log2n;
    auto pad = [&](int i) { return i + (i >> LOG2LL); };
    auto start = [&](int i) { return (i >> LOG2LL) * n1 * RS + (i & (LL - 1)); };        // (row, r2): line row n1, element r2
    auto tw = power_tw([&](int i) {                                     // W_N^{r2 k1} = (W_N^{r2})^{k1}
        const int t = i & (LL - 1);
        return cmul(th[t >> kSplitLog2N2], tl[t & (kSplitN2 - 1)]);
    });
    __syncthreads();
    if (!FIRST && warp == 0 && blockIdx.x < items) load_rows(out + (long long)blockIdx.x * E);
    if (STAGED && threadIdx.x == 0 && blockIdx.x < items) load_field(blockIdx.x);
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long chain = item / items_per_chain, next = item + gridDim.x;
        const int x0 = (int)(item - chain * items_per_chain) << log2r;
        double2* o = out + item * E;
        if (FIRST) {
            const real* g = field + chain * (KIND == SVB_CORR_WINDING ? 2 : 1) * V;
            if (warp == 0) bulk_wait_read0();
            __syncthreads();
            if (KIND == SVB_CORR_WINDING) {
                for (int e = threadIdx.x; e < E; e += blockDim.x) {
                    const int r = x0 + (e >> log2n), x1 = e & (N - 1);
                    const long long at = (long long)r * N + x1;
                    const long long i0 = (long long)((r + 1) & (N - 1)) * N + x1, i1 = (long long)r * N + ((x1 + 1) & (N - 1));
                    d[pad(e)] = make_double2((double)(((long long)g[V + i0] - (long long)g[V + at]) - ((long long)g[i1] - (long long)g[at])), 0.0);
                }
            } else {
                // the field values arrived by TMA while the previous item was being transformed; the next item's are requested
                // as soon as these have been read
                mbar_wait(&ld_bar, ld_parity);
                ld_parity ^= 1u;
                for (int e = threadIdx.x; e < E; e += blockDim.x) {
                    double sn, cs;
                    const double ang = (KIND == SVB_CORR_VORTEX) ? (SVB_TWO_PI * (double)stage[e]) / (double)W : (double)stage[e];
                    sincos(ang, &sn, &cs);
                    d[pad(e)] = make_double2(cs, sn);
                }
                fence_proxy_async();
            }
            __syncthreads();
            if (STAGED && threadIdx.x == 0 && next < items) load_field(next);
            if (log2n1 > 0) small_dft_pass_n1<true>(log2n1, d, E >> log2n1, RS, start, tw);
            fft_rr<true, RA, RB>(d, 1, RS, LOG2LINES, wl);
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) store_rows(o);
        } else {
            if (threadIdx.x == 0 && next < items)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(out + next * E), "r"((uint32_t)E * (uint32_t)sizeof(double2)) : "memory");
            mbar_wait(&ld_bar, ld_parity);
            ld_parity ^= 1u;
            fft_rr<false, RA, RB>(d, 1, RS, LOG2LINES, wl, [&](int, int) { return make_double2(scale, 0.0); });
            if (log2n1 > 0) small_dft_pass_n1<false>(log2n1, d, E >> log2n1, RS, start, tw);
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {
                store_rows(o);
                if (next < items) {
                    bulk_wait_read0();
                    load_rows(out + next * E);
                }
            }
        }
    }
    if (warp == 0) bulk_wait0();
}

// columns of 128 <= N <= 512, whole columns of a 4096 / N-column tile: row r = LL r1 + r2 in slot (r2, r1) of an LL x lines array
template <int RA, int RB>
__global__ void __launch_bounds__(256, 2) correlation_columns_r16_kernel(const __grid_constant__ CUtensorMap map, long long chains, int N, int log2n1) {
    extern __shared__ __align__(128) unsigned char box_smem[];
    constexpr int LL = RA * RB, LOG2LL = LL == 256 ? 8 : 7, E = 4096, lines = E / LL, LOG2LINES = 12 - LOG2LL;
    const int log2cw = LOG2LINES - log2n1, CW = 1 << log2cw, n1 = 1 << log2n1;
    double2* d = reinterpret_cast<double2*>(box_smem);                 // [LL r2][lines]: one box (2 CW doubles, n1, LL) of the map
    double2* wl = d + E;                                                // W_LL^t, t < LL
    double2* tN = wl + LL;                                              // W_N^t, t < N
    fft_n1_twiddles(wl, LL);
    fft_n1_twiddles(tN, N);
    __shared__ __align__(8) uint64_t ld_bar;
    if (threadIdx.x == 0) { mbar_init(&ld_bar, 1); fence_mbar_init(); }
    uint32_t ld_parity = 0;
    const int col_blocks = N >> log2cw;
    const long long items = chains * col_blocks;
    auto load_tile = [&](long long item) {                              // thread 0: the map's dimensions are (column, chain n1 + r1, r2)
        const long long chain = item / col_blocks;
        mbar_expect_tx(&ld_bar, (uint32_t)(E * sizeof(double2)));
        box_load_3d(d, &map, (int)(item - chain * col_blocks) << (log2cw + 1), (int)(chain * n1), 0, &ld_bar);
    };
    auto start = [&](int i) { return ((i >> log2cw) << LOG2LINES) + (i & (CW - 1)); };         // (r2, c)
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < items) load_tile(blockIdx.x);
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long next = item + gridDim.x;
        mbar_wait(&ld_bar, ld_parity);
        ld_parity ^= 1u;
        if (log2n1 > 0) {
            small_dft_pass_n1<true>(log2n1, d, LL << log2cw, CW, start, power_tw([&](int i) { return tN[i >> log2cw]; }));
            fft_rr<true, RA, RB>(d, lines, 1, LOG2LINES, wl, Norm2());
            fft_rr<false, RA, RB>(d, lines, 1, LOG2LINES, wl, NoScale(), [&](int L, int q) { return tN[(L >> log2cw) * q]; });
            small_dft_pass_n1<false>(log2n1, d, LL << log2cw, CW, start, NoScale());
        } else {
            fft_rr<true, RA, RB>(d, lines, 1, LOG2LINES, wl, Norm2());
            fft_rr<false, RA, RB>(d, lines, 1, LOG2LINES, wl);
        }
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long chain = item / col_blocks;
            box_store_3d(&map, (int)(item - chain * col_blocks) << (log2cw + 1), (int)(chain * n1), 0, d);
            bulk_commit();
            if (next < items) {
                bulk_wait_read0();
                load_tile(next);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait0();
}

}  // namespace svb

using namespace svb;

// tensor map over the complex array `base` taken as doubles: dims[0] = 2 N (contiguous), two more dimensions with byte strides
typedef CUresult (*corr_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int corr_tensor_map(CUtensorMap& map, void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1, cuuint64_t s2,
                           cuuint32_t b0, cuuint32_t b1, cuuint32_t b2) {
    static corr_map_encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult status;
        SVB_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status));
        if (status != cudaDriverEntryPointSuccess || !fn) return fail(SVB_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<corr_map_encode_fn>(fn);
    }
    const cuuint64_t dims[3] = {d0, d1, d2}, strides[2] = {s1, s2};
    const cuuint32_t box[3] = {b0, b1, b2}, ones[3] = {1, 1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(SVB_E_UNSUPPORTED, "cuTensorMapEncodeTiled(correlator tiles: dims %llu %llu %llu, box %u %u %u) failed: %d", (unsigned long long)d0,
                    (unsigned long long)d1, (unsigned long long)d2, b0, b1, b2, (int)r);
    return SVB_OK;
}

template <typename real, int KIND, bool FIRST, int RA, int RB>
static int launch_rows_r16(const void* field, long long chains, int N, int log2n, int W, double scale, double2* o, int sms, cudaStream_t st) {
    constexpr int LL = RA * RB, LOG2LL = LL == 256 ? 8 : 7;
    auto kern = correlation_rows_r16_kernel<real, KIND, FIRST, RA, RB>;
    const size_t smem = ((size_t)(4096 / LL) * (LL + 1) + LL + 2 * kSplitN2) * sizeof(double2) +
                        (FIRST && KIND != SVB_CORR_WINDING ? 4096 * sizeof(real) : 0);      // + the staged field values of the next item
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "svb_correlation: N=%d does not fit the radix-16 row kernel", N);
    const long long items = (chains * N * N) >> 12, cap = (long long)per_sm * sms;
    kern<<<(unsigned)(items < cap ? items : cap), 256, smem, st>>>(reinterpret_cast<const real*>(field), chains, N, log2n - LOG2LL, W, scale, o);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}
template <int RA, int RB>
static int launch_columns_r16(long long chains, int N, int log2n, double2* o, int sms, cudaStream_t st) {
    constexpr int LL = RA * RB, LOG2LL = LL == 256 ? 8 : 7;
    auto kern = correlation_columns_r16_kernel<RA, RB>;
    const size_t smem = ((size_t)4096 + LL + N) * sizeof(double2);
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "svb_correlation: N=%d does not fit the radix-16 column kernel", N);
    const long long items = (chains * N * N) >> 12, cap = (long long)per_sm * sms;
    // (column, chain n1 + r1, r2): a box of all r2 and r1 lands as [r2][r1][column], the layout the passes work in
    const int n1 = N / LL;
    CUtensorMap map;
    const int rc = corr_tensor_map(map, o, 2 * (cuuint64_t)N, (cuuint64_t)n1 * chains, LL, (cuuint64_t)LL * N * sizeof(double2), (cuuint64_t)N * sizeof(double2),
                                   2 * (4096 / N), n1, LL);
    if (rc != SVB_OK) return rc;
    kern<<<(unsigned)(items < cap ? items : cap), 256, smem, st>>>(map, chains, N, log2n - LOG2LL);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

template <typename real, int KIND>
static int launch_correlation_fft_large(const void* field, long long chains, int N, int W, double* out, int sms, cudaStream_t st) {
    int log2n = 0;
    while ((1 << log2n) < N) ++log2n;
    const int R = N >= 2048 ? 1 : 2048 / N;                       // 32 KiB of rows per work item
    int C = 8192 / N;                                             // <= 128 KiB of columns per work item
    if (C > 8) C = 8;
    int log2c = 0;
    while ((1 << log2c) < C) ++log2c;
    const size_t smem_rows = ((size_t)R * N + N / 2) * sizeof(double2);
    const size_t smem_cols = ((size_t)C * (N + 1) + N / 2) * sizeof(double2);
    auto k1 = correlation_rows_forward_kernel<real, KIND>;
    SVB_CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    SVB_CUDA_TRY(cudaFuncSetAttribute(correlation_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
    SVB_CUDA_TRY(cudaFuncSetAttribute(correlation_rows_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    int per_sm_rows = 0, per_sm_cols = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_rows, k1, 256, smem_rows));
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_cols, correlation_columns_kernel, 256, smem_cols));
    if (per_sm_rows < 1 || per_sm_cols < 1) return fail(SVB_E_UNSUPPORTED, "svb_correlation: N=%d does not fit the FFT kernels", N);
    const long long row_items = chains * (N / R), col_items = chains * (N / C);
    const long long cap_rows = (long long)per_sm_rows * sms, cap_cols = (long long)per_sm_cols * sms;
    double2* o = reinterpret_cast<double2*>(out);
    // whole lines in shared memory up to N = 512; beyond, every transform in two steps,
    // r = 64 r1 + r2 (SVB_CORR_SPLIT_MIN_N lowers the threshold, for tests of the split at sizes that can be checked element by
    // element)
    int split_min = 1024;                                          // 8.4 M sites: 629 against 745 us at N = 1024, 669 against 594 us at N = 512
    if (const char* e = getenv("SVB_CORR_SPLIT_MIN_N")) split_min = atoi(e);
    // 128 <= N <= 512: the row kernels of the split with 4096 / N rows per item and ONE column kernel (whole columns of a
    // 4096 / N-column tile in shared memory): three launches, every element read and written once by each
    // (SVB_CORR_ROUTE=legacy: the radix-2 kernels they replace; SVB_CORR_SPLIT_MIN_N set: the five-launch split or legacy)
    const char* route = getenv("SVB_CORR_ROUTE");
    int mid_max = 2048;                                            // (8.4 M sites: 240 against 350 us at N = 1024, 276 against 322 us at N = 2048)
    if (const char* e = getenv("SVB_CORR_MID_MAX_N")) mid_max = atoi(e);
    int r16 = 1;
    if (const char* e = getenv("SVB_CORR_R16")) r16 = atoi(e);
    if ((uintptr_t)field % 16) r16 = 0;                            // (the radix-16 row kernel stages the field by bulk copies)
    if (r16 <= 0 && mid_max > 512) mid_max = 512;                  // (the 64-line kernels take several rows per item only for n1 <= 8)
    const bool mid = N <= mid_max && !getenv("SVB_CORR_SPLIT_MIN_N") && !(route && !strcmp(route, "legacy"));
    const bool split = mid || (N >= split_min && N >= 2 * kSplitN2);
    const double V = (double)N * (double)N;
    const int log2n1_rows = log2n - kSplitLog2N2, n1_rows = split ? (1 << log2n1_rows) : 1;
    const int log2r = mid ? 12 - log2n : 0, lines_rows = n1_rows << log2r;
    const size_t smem_rsplit = ((size_t)lines_rows * (kSplitN2 + 1) + n1_rows + 3 * kSplitN2) * sizeof(double2);
    // radix-16 kernels (lines of 256 or 128 elements: a visit of shared memory fewer): the rows of every split route -- per
    // launch at 8.4 M sites 48-62 against 61-72 us (N = 128, 256), 57-72 against 62-75 us (N = 512), 125 + 158 against
    // 137 + 164 us at N = 4096 -- and the columns of the fused route except N = 512 (with tensor-map tiles, per correlator:
    // 194 against 204 us at N = 256, 229 against 222 us at N = 512; N = 2048 has n1 = 32, beyond the 64-line kernel's register
    // pass).  SVB_CORR_R16=0: the radix-8 kernels throughout (fused columns up to N = 512), =2: radix-16 columns at N = 512 too
    const bool rows16 = split && r16 > 0, cols16 = mid && rows16 && (N != 512 || r16 > 1);
    if (rows16) {
        const int rc = N == 128 ? launch_rows_r16<real, KIND, true, 16, 8>(field, chains, N, log2n, W, 1.0, o, sms, st)
                                : launch_rows_r16<real, KIND, true, 16, 16>(field, chains, N, log2n, W, 1.0, o, sms, st);
        if (rc != SVB_OK) return rc;
    }
    auto kr1 = correlation_rows_split_kernel<real, KIND, true>;
    auto kr2 = correlation_rows_split_kernel<real, KIND, false>;
    long long cap_rsplit = 0;
    const int rows_threads = 256;                                 // (512, an item of a radix-8 pass per thread at N = 4096: 1039 against 860 us)
    if (split && !rows16) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kr1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rsplit));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kr2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rsplit));
        int per_sm = 0;
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kr1, rows_threads, smem_rsplit));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "svb_correlation: N=%d does not fit the split row kernels", N);
        cap_rsplit = (long long)per_sm * sms;
        const long long items = (chains * N) >> log2r;
        kr1<<<(unsigned)(items < cap_rsplit ? items : cap_rsplit), rows_threads, smem_rsplit, st>>>(reinterpret_cast<const real*>(field), chains, N,
                                                                                           log2n1_rows, log2r, W, 1.0, o);
    } else if (!split) {
        k1<<<(unsigned)(row_items < cap_rows ? row_items : cap_rows), 256, smem_rows, st>>>(reinterpret_cast<const real*>(field), chains, N,
                                                                                             log2n, W, R, o);
    }
    SVB_CUDA_TRY(cudaGetLastError());
    if (cols16) {
        const int rc = N == 128 ? launch_columns_r16<16, 8>(chains, N, log2n, o, sms, st) : launch_columns_r16<16, 16>(chains, N, log2n, o, sms, st);
        if (rc != SVB_OK) return rc;
    } else if (mid) {
        const size_t smem_fused = ((size_t)kSplitN2 * kSplitN2 + kSplitN2 + N) * sizeof(double2);
        SVB_CUDA_TRY(cudaFuncSetAttribute(correlation_columns_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused));
        int per_sm = 0;
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, correlation_columns_fused_kernel, 256, smem_fused));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "svb_correlation: N=%d does not fit the fused column kernel", N);
        const long long items = chains * (N >> (kSplitLog2N2 - (log2n - kSplitLog2N2))), cap = (long long)per_sm * sms;
        const int n1 = N / kSplitN2;
        CUtensorMap map;                                           // (column, chain n1 + r1, r2) -> [r2][r1][column]
        const int rc = corr_tensor_map(map, o, 2 * (cuuint64_t)N, (cuuint64_t)n1 * chains, kSplitN2, (cuuint64_t)kSplitN2 * N * sizeof(double2),
                                       (cuuint64_t)N * sizeof(double2), 2 * (4096 / N), n1, kSplitN2);
        if (rc != SVB_OK) return rc;
        correlation_columns_fused_kernel<<<(unsigned)(items < cap ? items : cap), 256, smem_fused, st>>>(map, chains, N, log2n - kSplitLog2N2);
        SVB_CUDA_TRY(cudaGetLastError());
    } else if (split) {
        const int log2n1 = log2n - kSplitLog2N2, n1 = 1 << log2n1;
        const size_t smem_outer = ((size_t)n1 * kSplitCols + 2 * n1) * sizeof(double2);
        const size_t smem_inner = ((size_t)kSplitN2 * kSplitCols + 2 * kSplitN2) * sizeof(double2);
        SVB_CUDA_TRY(cudaFuncSetAttribute(correlation_split_outer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_outer));
        SVB_CUDA_TRY(cudaFuncSetAttribute(correlation_split_outer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_outer));
        SVB_CUDA_TRY(cudaFuncSetAttribute(correlation_split_inner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_inner));
        int per_sm_outer = 0, per_sm_inner = 0;
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_outer, correlation_split_outer_kernel<true>, 256, smem_outer));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_inner, correlation_split_inner_kernel, 256, smem_inner));
        if (per_sm_outer < 1 || per_sm_inner < 1) return fail(SVB_E_UNSUPPORTED, "svb_correlation: N=%d does not fit the split column kernels", N);
        const long long outer_items = chains * kSplitN2 * (N / kSplitCols), inner_items = chains * n1 * (N / kSplitCols);
        const long long cap_outer = (long long)per_sm_outer * sms, cap_inner = (long long)per_sm_inner * sms;
        // (column, r2, chain n1 + r1): an outer tile is the box (32 columns, one r2, all r1), an inner tile (32 columns, all r2, one r1)
        CUtensorMap map_outer, map_inner;
        int rc = corr_tensor_map(map_outer, o, 2 * (cuuint64_t)N, kSplitN2, (cuuint64_t)n1 * chains, (cuuint64_t)N * sizeof(double2),
                                 (cuuint64_t)kSplitN2 * N * sizeof(double2), 2 * kSplitCols, 1, n1);
        if (rc != SVB_OK) return rc;
        rc = corr_tensor_map(map_inner, o, 2 * (cuuint64_t)N, kSplitN2, (cuuint64_t)n1 * chains, (cuuint64_t)N * sizeof(double2),
                             (cuuint64_t)kSplitN2 * N * sizeof(double2), 2 * kSplitCols, kSplitN2, 1);
        if (rc != SVB_OK) return rc;
        correlation_split_outer_kernel<true><<<(unsigned)(outer_items < cap_outer ? outer_items : cap_outer), 256, smem_outer, st>>>(map_outer, chains, N, log2n1);
        SVB_CUDA_TRY(cudaGetLastError());
        correlation_split_inner_kernel<<<(unsigned)(inner_items < cap_inner ? inner_items : cap_inner), 256, smem_inner, st>>>(map_inner, chains, N, log2n1);
        SVB_CUDA_TRY(cudaGetLastError());
        correlation_split_outer_kernel<false><<<(unsigned)(outer_items < cap_outer ? outer_items : cap_outer), 256, smem_outer, st>>>(map_outer, chains, N, log2n1);
        SVB_CUDA_TRY(cudaGetLastError());
    } else {
        correlation_columns_kernel<<<(unsigned)(col_items < cap_cols ? col_items : cap_cols), 256, smem_cols, st>>>(chains, N, log2n, C, log2c, o);
        SVB_CUDA_TRY(cudaGetLastError());
    }
    if (rows16) {
        return N == 128 ? launch_rows_r16<real, KIND, false, 16, 8>(field, chains, N, log2n, W, 1.0 / (V * V), o, sms, st)
                        : launch_rows_r16<real, KIND, false, 16, 16>(field, chains, N, log2n, W, 1.0 / (V * V), o, sms, st);
    } else if (split) {
        const long long items = (chains * N) >> log2r;
        kr2<<<(unsigned)(items < cap_rsplit ? items : cap_rsplit), rows_threads, smem_rsplit, st>>>(reinterpret_cast<const real*>(field), chains, N,
                                                                                           log2n1_rows, log2r, W, 1.0 / (V * V), o);
    } else {
        correlation_rows_inverse_kernel<<<(unsigned)(row_items < cap_rows ? row_items : cap_rows), 256, smem_rows, st>>>(chains, N, log2n, R,
                                                                                                                         1.0 / (V * V), o);
    }
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

template <typename real, int KIND, int NT>
static int launch_correlation_fft(const void* field, long long chains, int W, double* out, int sms, cudaStream_t st) {
    int warp32 = 0;                                                // SVB_CORR_FFT32_WARP=1: L = 32 with a warp per chain and a line per thread -- three visits of
                                                                   // shared memory instead of eight, but 150 registers and 13 warps per SM: 122 against 102 us
                                                                   // for 8192 chains, so the CTA kernel stays the default
    if (const char* e = getenv("SVB_CORR_FFT32_WARP")) warp32 = atoi(e);
    if (NT == 32 && warp32) {
        auto kw = correlation_fft32_warp_kernel<real, KIND>;
        const size_t smem_w = (size_t)32 * 33 * sizeof(double2);
        SVB_CUDA_TRY(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
        int per_sm = 0;
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kw, 32, smem_w));
        if (per_sm < 1) per_sm = 1;
        const long long cap = (long long)per_sm * sms;
        kw<<<(unsigned)(chains < cap ? chains : cap), 32, smem_w, st>>>(reinterpret_cast<const real*>(field), chains, W, out);
        SVB_CUDA_TRY(cudaGetLastError());
        return SVB_OK;
    }
    auto kern = correlation_fft_kernel<real, KIND, NT>;
    const size_t smem = (size_t)(NT * (NT + 1) + NT) * sizeof(double2);      // the padded tile and the N twiddles (>= the radix-2 layout)
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // a thread per radix-8 butterfly of a pass (N^2 / 8 of them): more threads than that idle through half of the passes
    int threads = NT == 64 ? 256 : NT == 32 ? 128 : 64;
    if (const char* e = getenv("SVB_CORR_FFT_THREADS")) threads = atoi(e);
    if (threads < 32 || threads > 256 || threads % 32) return fail(SVB_E_PARAM, "SVB_CORR_FFT_THREADS=%d", threads);
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)per_sm * sms;
    kern<<<(unsigned)(chains < cap ? chains : cap), threads, smem, st>>>(reinterpret_cast<const real*>(field), chains, W, out);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

template <typename real, int KIND>
static int launch_correlation(const void* field, long long chains, int N, int W, double* out, size_t smem, long long grid,
                              cudaStream_t st) {
    auto kern = correlation_kernel<real, KIND>;
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, 256, smem, st>>>(reinterpret_cast<const real*>(field), chains, N, W, out);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

extern "C" int svb_correlation(int kind, const void* field, int dtype, int64_t chains, int N, int W, double* out, void* stream) {
    if (!field || !out) return fail(SVB_E_NULL, "svb_correlation: field and out are required");
    if (chains < 0 || N < 1) return fail(SVB_E_SHAPE, "svb_correlation: shape");
    if (kind == SVB_CORR_SPIN && dtype != SVB_F64 && dtype != SVB_F32) return fail(SVB_E_DTYPE, "svb_correlation: phi dtype %d", dtype);
    if (kind != SVB_CORR_SPIN && dtype != SVB_I32) return fail(SVB_E_DTYPE, "svb_correlation: integer fields must be int32");
    if (kind == SVB_CORR_VORTEX && W < 1) return fail(SVB_E_PARAM, "svb_correlation: W");
    if (chains == 0) return SVB_OK;
    const size_t smem = (size_t)2 * N * N * sizeof(double);
    int dev = 0, max_smem = 0, sms = 0;
    SVB_CUDA_TRY(cudaGetDevice(&dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (N >= 128 && N <= 4096 && (N & (N - 1)) == 0 && (uintptr_t)out % 16 == 0) {
        // power-of-two lattices beyond shared memory: three-launch FFT with `out` as the workspace
        if (kind == SVB_CORR_SPIN && dtype == SVB_F64) return launch_correlation_fft_large<double, SVB_CORR_SPIN>(field, chains, N, W, out, sms, st);
        if (kind == SVB_CORR_SPIN && dtype == SVB_F32) return launch_correlation_fft_large<float, SVB_CORR_SPIN>(field, chains, N, W, out, sms, st);
        if (kind == SVB_CORR_WINDING) return launch_correlation_fft_large<int32_t, SVB_CORR_WINDING>(field, chains, N, W, out, sms, st);
        if (kind == SVB_CORR_VORTEX) return launch_correlation_fft_large<int32_t, SVB_CORR_VORTEX>(field, chains, N, W, out, sms, st);
        return fail(SVB_E_PARAM, "svb_correlation: kind %d", kind);
    }
    if (smem > (size_t)max_smem)
        return fail(SVB_E_UNSUPPORTED,
                    "svb_correlation: N=%d is neither a power of two in [16, 4096] (FFT) nor small enough for the direct sum out of shared memory", N);
#define SVB_CORR_FFT(real, KIND)                                                                             \
    switch (N) {                                                                                             \
        case 16: return launch_correlation_fft<real, KIND, 16>(field, chains, W, out, sms, st);             \
        case 32: return launch_correlation_fft<real, KIND, 32>(field, chains, W, out, sms, st);             \
        case 64: return launch_correlation_fft<real, KIND, 64>(field, chains, W, out, sms, st);             \
        default: break;                                                                                      \
    }
    if ((uintptr_t)out % 16 == 0) {
        if (kind == SVB_CORR_SPIN && dtype == SVB_F64) { SVB_CORR_FFT(double, SVB_CORR_SPIN) }
        if (kind == SVB_CORR_SPIN && dtype == SVB_F32) { SVB_CORR_FFT(float, SVB_CORR_SPIN) }
        if (kind == SVB_CORR_WINDING) { SVB_CORR_FFT(int32_t, SVB_CORR_WINDING) }
        if (kind == SVB_CORR_VORTEX) { SVB_CORR_FFT(int32_t, SVB_CORR_VORTEX) }
    }
#undef SVB_CORR_FFT
    const long long grid = chains < (long long)sms * 4 ? chains : (long long)sms * 4;
    switch (kind) {
        case SVB_CORR_SPIN:
            return dtype == SVB_F64 ? launch_correlation<double, SVB_CORR_SPIN>(field, chains, N, W, out, smem, grid, st)
                                    : launch_correlation<float, SVB_CORR_SPIN>(field, chains, N, W, out, smem, grid, st);
        case SVB_CORR_WINDING: return launch_correlation<int32_t, SVB_CORR_WINDING>(field, chains, N, W, out, smem, grid, st);
        case SVB_CORR_VORTEX: return launch_correlation<int32_t, SVB_CORR_VORTEX>(field, chains, N, W, out, smem, grid, st);
        default: return fail(SVB_E_PARAM, "svb_correlation: kind %d", kind);
    }
}

extern "C" int svb_villain_spin_spin(const void* phi, int phi_dtype, int64_t chains, int N, double* out, void* stream) {
    return svb_correlation(SVB_CORR_SPIN, phi, phi_dtype, chains, N, 1, out, stream);
}


// ------------------------------------------------------------------------------------------
// Autocorrelation of scalar columns (supervillain/analysis/autocorrelation.py:7-66), one series per chain:
//   Delta(t) = data(t) - mean;   C(tau) = <Delta(t + tau) Delta(t)> / <Delta(t)^2>  with t + tau wrapping around the
//   series (the reference evaluates exactly this circular form with FFTs, :48-52);
//   tau_int = ceil(sum_{tau < tau_0} C(tau)),  tau_0 = argmin(clip(C, 0, None))   (:57-59).
// Direct O(T^2) sums out of shared memory, one CTA per series: exact to rounding (the reference's FFT result agrees to
// ~1e-13) and cheap next to the sweeps that produced the series.  A series whose fluctuations are below the reference's
// cutoff (:53) gets tau = -1 and C = 0 (the Python wrapper raises the reference's ValueError).
// ------------------------------------------------------------------------------------------
namespace svb {

__global__ void __launch_bounds__(256) autocorrelation_kernel(const double* __restrict__ data, long long series, int T,
                                                              const double* __restrict__ mean_in, double cutoff,
                                                              double* __restrict__ C_out, int32_t* __restrict__ tau_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* delta = reinterpret_cast<double*>(smem_raw);           // T
    double* corr = delta + T;                                      // T
    __shared__ double scratch[32];
    __shared__ double s_mean, s_c0;
    const int tid = threadIdx.x;
    for (long long c = blockIdx.x; c < series; c += gridDim.x) {
        const double* x = data + c * T;
        double part[1] = {0.0};
        if (mean_in) {
            if (tid == 0) s_mean = mean_in[c];
        } else {
            for (int t = tid; t < T; t += blockDim.x) part[0] += x[t];
            block_sum<1>(part, scratch);
            if (tid == 0) s_mean = part[0] / (double)T;
        }
        __syncthreads();
        const double mean = s_mean;
        for (int t = tid; t < T; t += blockDim.x) delta[t] = x[t] - mean;
        __syncthreads();
        for (int tau = tid; tau < T; tau += blockDim.x) {
            double acc = 0.0;
            int u = tau;
            for (int t = 0; t < T; ++t) {
                acc = fma(delta[t], delta[u], acc);
                u = (u + 1 == T) ? 0 : u + 1;
            }
            corr[tau] = acc / (double)T;
        }
        __syncthreads();
        if (tid == 0) s_c0 = corr[0];
        __syncthreads();
        const double c0 = s_c0;
        const bool flat = fabs(c0) < cutoff;
        for (int tau = tid; tau < T; tau += blockDim.x) {
            const double v = flat ? 0.0 : corr[tau] / c0;
            corr[tau] = v;
            if (C_out) C_out[c * T + tau] = v;
        }
        __syncthreads();
        if (tid == 0 && tau_out) {
            int tau = -1;
            if (!flat) {
                // argmin of the clamped function: its first zero if there is one, else its smallest positive value
                int arg = 0;
                double best = fmax(corr[0], 0.0);
                for (int t = 1; t < T && best > 0.0; ++t) {
                    const double v = fmax(corr[t], 0.0);
                    if (v < best) { best = v; arg = t; }
                }
                // np.sum over C[:arg] is pairwise in numpy; tau is its ceiling, insensitive to the order unless the sum
                // sits within rounding of an integer
                double sum = 0.0;
                for (int t = 0; t < arg; ++t) sum += corr[t];
                tau = (int)ceil(sum);
            }
            tau_out[c] = tau;
        }
        __syncthreads();
    }
}

}  // namespace svb

extern "C" int svb_autocorrelation(const double* data, int64_t series, int T, const double* mean, double* C, int32_t* tau,
                                   void* stream) {
    using namespace svb;
    if (!data || (!C && !tau)) return fail(SVB_E_NULL, "svb_autocorrelation: data and at least one output are required");
    if (series < 0 || T < 1) return fail(SVB_E_SHAPE, "svb_autocorrelation: series=%lld T=%d", (long long)series, T);
    const size_t smem = (size_t)2 * T * sizeof(double);
    if (smem > 200 * 1024) return fail(SVB_E_UNSUPPORTED, "svb_autocorrelation: T=%d does not fit shared memory (T <= 12800)", T);
    if (series == 0) return SVB_OK;
    SVB_CUDA_TRY(cudaFuncSetAttribute(autocorrelation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = series < 148LL * 8 ? series : 148LL * 8;
    autocorrelation_kernel<<<(unsigned)grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(data, series, T, mean, 1e-16, C, tau);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}


// ------------------------------------------------------------------------------------------
// The taxicab reweighting observables: Spin_Spin.Worldline (supervillain/observable/spin.py:50-224) and Vortex_Vortex.Villain
// (supervillain/observable/vortex.py:63-189).  For every displacement D = (Dt, Dx) (FFT coordinates) and every starting site x
// a fixed path P -- |Dt| steps in time, then |Dx| steps in space -- picks +-1 on |D| links, and
//     out[D] = mean_x exp(alpha P(x; D).links + beta |D|)
//   SPIN   links = m - delta(v)/W:   time part on (0, .) links, rows x0 .. x0 + Dt - 1; space part on (1, .) links of row
//          x0 + Dt, columns x1 .. x1 + Dx - 1; + along, - against;  alpha = -1/kappa, beta = -1/(2 kappa)
//   VORTEX links = d(phi) - 2 pi n:  time part on (1, .) links, rows x0 + 1 .. x0 + Dt; space part on (0, .) links of row
//          x0 + Dt, columns x1 + 1 .. x1 + Dx with the opposite sign;  alpha = 2 pi kappa, beta = -2 pi^2 kappa
// The reference gathers the |D| links of all N^2 translates for each of the N^2 displacements: O(N^5).  Here circular prefix
// sums of the two components along their path directions live in shared memory and a path sum is two differences:
// O(N^4) exponentials per chain, one thread per displacement, 256 displacements per CTA.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double circular_range(const double* __restrict__ P, int stride, int N, int start, int count) {
    const int end = start + count;                                // P[k * stride] = sum of the first k elements
    if (end <= N) return P[end * stride] - P[start * stride];
    return (P[N * stride] - P[start * stride]) + P[(end - N) * stride];
}

__global__ void __launch_bounds__(256) taxicab_kernel(const double* __restrict__ links, long long chains, int N, int kind, double kappa,
                                                      const double* __restrict__ kappa_chain, int tiles, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* PT = reinterpret_cast<double*>(smem_raw);            // [N + 1][N]: prefix over rows of the time component
    double* PS = PT + (size_t)(N + 1) * N;                        // [N][N + 1]: prefix over columns of the space component
    const int V = N * N;
    const int comp_t = kind == SVB_TAXI_SPIN ? 0 : 1, comp_s = 1 - comp_t, offset = kind == SVB_TAXI_SPIN ? 0 : 1;
    const double sign_s = kind == SVB_TAXI_SPIN ? 1.0 : -1.0;
    const long long items = chains * tiles;
    long long staged = -1;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long chain = item / tiles;
        const int tile = (int)(item - chain * tiles);
        const double k = kappa_chain ? kappa_chain[chain] : kappa;
        const double alpha = kind == SVB_TAXI_SPIN ? -1.0 / k : SVB_TWO_PI * k;
        const double beta = kind == SVB_TAXI_SPIN ? -0.5 / k : -0.5 * SVB_TWO_PI * SVB_TWO_PI * k;
        if (chain != staged) {
            __syncthreads();
            const double* g = links + chain * 2 * V;
            for (int x = threadIdx.x; x < N; x += blockDim.x) {
                double acc = 0.0;
                PT[x] = 0.0;
                for (int t = 0; t < N; ++t) { acc += g[comp_t * V + t * N + x]; PT[(t + 1) * N + x] = acc; }
            }
            for (int r = threadIdx.x; r < N; r += blockDim.x) {
                double acc = 0.0;
                PS[r * (N + 1)] = 0.0;
                for (int x = 0; x < N; ++x) { acc += g[comp_s * V + r * N + x]; PS[r * (N + 1) + x + 1] = acc; }
            }
            staged = chain;
            __syncthreads();
        }
        const int d = tile * 256 + threadIdx.x;
        if (d < V) {
            const int a = d / N, b = d - a * N;
            const int Dt = a <= N / 2 ? a : a - N, Dx = b <= N / 2 ? b : b - N;
            const int T = Dt < 0 ? -Dt : Dt, X = Dx < 0 ? -Dx : Dx;
            const double st = Dt < 0 ? -1.0 : 1.0, ss = Dx < 0 ? -sign_s : sign_s;
            const double shift = beta * (double)(T + X);
            double acc = 0.0;
            for (int x0 = 0; x0 < N; ++x0) {
                int t_start = x0 + offset - (Dt < 0 ? T : 0);
                t_start %= N; if (t_start < 0) t_start += N;
                int row = (x0 + Dt) % N; if (row < 0) row += N;
                const double* prow = PS + row * (N + 1);
                for (int x1 = 0; x1 < N; ++x1) {
                    int s_start = x1 + offset - (Dx < 0 ? X : 0);
                    s_start %= N; if (s_start < 0) s_start += N;
                    const double S = st * circular_range(PT + x1, N, N, t_start, T) + ss * circular_range(prow, 1, N, s_start, X);
                    acc += exp(fma(alpha, S, shift));
                }
            }
            out[chain * V + d] = acc / (double)V;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Blocking and Bootstrap of scalar columns (supervillain/analysis/blocking.py:54-66, bootstrap.py:57-67), one column per
// series (e.g. one observable of every chain), T samples each, data (series, T):
//   block_mean      out[s, b] = mean_{i < width} w[drop + b width + i] data[s, drop + b width + i]
//   bootstrap_mean  out[s, d] = mean_c (w[idx[c, d]] data[s, idx[c, d]]) / mean_c w[idx[c, d]],  idx (T, draws) as numpy drew it
// (w == nullptr: unit weights).  Sums run in sample order, which is numpy's order for the bootstrap (a reduction over the
// leading axis); numpy blocks along the contiguous axis with its pairwise scheme, so block means agree to ~1e-15 relative.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) block_mean_kernel(const double* __restrict__ data, const double* __restrict__ w, long long series,
                                                         long long T, int width, long long drop, long long blocks,
                                                         double* __restrict__ out) {
    const long long total = series * blocks;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / blocks, b = i - s * blocks;
        const double* p = data + s * T + drop + b * width;
        const double* q = w ? w + drop + b * width : nullptr;
        double acc = 0.0;
        for (int k = 0; k < width; ++k) acc += q ? __dmul_rn(p[k], q[k]) : p[k];
        out[i] = acc / (double)width;
    }
}

__global__ void __launch_bounds__(256) bootstrap_mean_kernel(const double* __restrict__ data, const double* __restrict__ w,
                                                             long long series, long long T, const long long* __restrict__ idx,
                                                             int draws, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* col = reinterpret_cast<double*>(smem_raw);           // the series' column, staged when it fits
    const bool staged = (size_t)T * sizeof(double) <= 160 * 1024;
    for (long long s = blockIdx.x; s < series; s += gridDim.x) {
        const double* g = data + s * T;
        __syncthreads();
        if (staged)
            for (long long c = threadIdx.x; c < T; c += blockDim.x) col[c] = w ? __dmul_rn(w[c], g[c]) : g[c];
        __syncthreads();
        for (int d = threadIdx.x; d < draws; d += blockDim.x) {
            double acc = 0.0, wsum = 0.0;
            for (long long c = 0; c < T; ++c) {
                const long long j = idx[c * draws + d];
                acc += staged ? col[j] : (w ? __dmul_rn(w[j], g[j]) : g[j]);
                if (w) wsum += w[j];
            }
            const double mean = acc / (double)T;
            out[s * draws + d] = w ? mean / (wsum / (double)T) : mean;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Test hook for the lazily refined uniform (svb_common.cuh): the refinement branch of decide_lazy is taken with probability
// 2^-32 per proposal, so no sweep test ever reaches it.  This evaluates decide_lazy(A, (f, c0, word)) and the refined
// uniform itself for caller-chosen inputs, to be compared with the oracle's statement of the same rule.
// ------------------------------------------------------------------------------------------
namespace svb {
__global__ void debug_decide_lazy_kernel(const double* A, const uint32_t* f, const uint32_t* c0, const uint32_t* word, long long n,
                                         uint32_t stream_id, unsigned long long seed, unsigned long long chain,
                                         unsigned long long sweep, uint8_t* decision, double* u_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LazyUniform lu;
    lu.f = f[i]; lu.c0 = c0[i]; lu.word = word[i];
    RefineCtx rc;
    rc.seed = seed; rc.chain = chain; rc.sweep = sweep;
    decision[i] = decide_lazy(A[i], lu, stream_id, rc) ? 1 : 0;
    u_out[i] = refined_uniform(lu.f, lu.c0, lu.word, stream_id, seed, chain, sweep);
}
}  // namespace svb

extern "C" int svb_debug_decide_lazy(const double* A, const uint32_t* f, const uint32_t* c0, const uint32_t* word, int64_t n,
                                     uint32_t stream_id, uint64_t seed, uint64_t chain, uint64_t sweep, uint8_t* decision,
                                     double* u_out, void* stream) {
    using namespace svb;
    if (!A || !f || !c0 || !word || !decision || !u_out) return fail(SVB_E_NULL, "svb_debug_decide_lazy: all arrays are required");
    if (n <= 0) return SVB_OK;
    debug_decide_lazy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        A, f, c0, word, n, stream_id, seed, chain, sweep, decision, u_out);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

extern "C" int svb_block_mean(const double* data, const double* weight, int64_t series, int64_t T, int width, int64_t drop, double* out,
                              void* stream) {
    if (!data || !out) return fail(SVB_E_NULL, "svb_block_mean: data and out are required");
    if (series < 0 || T < 0 || width < 1 || drop < 0 || drop > T || (T - drop) % width != 0)
        return fail(SVB_E_SHAPE, "svb_block_mean: (T - drop) must be a multiple of width (T=%lld drop=%lld width=%d)", (long long)T,
                    (long long)drop, width);
    const long long blocks = (T - drop) / width;
    if (series == 0 || blocks == 0) return SVB_OK;
    const long long total = series * blocks;
    const long long grid = (total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16;
    block_mean_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(data, weight, series, T, width, drop, blocks, out);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

extern "C" int svb_bootstrap_mean(const double* data, const double* weight, int64_t series, int64_t T, const int64_t* idx, int draws,
                                  double* out, void* stream) {
    if (!data || !idx || !out) return fail(SVB_E_NULL, "svb_bootstrap_mean: data, idx and out are required");
    if (series < 0 || T < 1 || draws < 1) return fail(SVB_E_SHAPE, "svb_bootstrap_mean: shape");
    if (series == 0) return SVB_OK;
    const size_t smem = ((size_t)T * sizeof(double) <= 160 * 1024) ? (size_t)T * sizeof(double) : 0;
    SVB_CUDA_TRY(cudaFuncSetAttribute(bootstrap_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    const long long grid = series < 148 * 8 ? series : 148 * 8;
    bootstrap_mean_kernel<<<(unsigned)grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        data, weight, series, T, reinterpret_cast<const long long*>(idx), draws, out);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

extern "C" int svb_taxicab_correlator(int kind, const double* links, int64_t chains, int N, double kappa, const double* kappa_chain,
                                      double* out, void* stream) {
    if (kind != SVB_TAXI_SPIN && kind != SVB_TAXI_VORTEX) return fail(SVB_E_PARAM, "svb_taxicab_correlator: kind %d", kind);
    if (!links || !out) return fail(SVB_E_NULL, "svb_taxicab_correlator: links and out are required");
    if (chains < 0 || N < 2) return fail(SVB_E_SHAPE, "svb_taxicab_correlator: shape");
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_taxicab_correlator: kappa must be positive");
    if (chains == 0) return SVB_OK;
    const size_t smem = (size_t)2 * N * (N + 1) * sizeof(double);
    int dev = 0, max_smem = 0, sms = 0;
    SVB_CUDA_TRY(cudaGetDevice(&dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > (size_t)max_smem) return fail(SVB_E_UNSUPPORTED, "svb_taxicab_correlator: the prefix sums of N=%d do not fit shared memory", N);
    SVB_CUDA_TRY(cudaFuncSetAttribute(taxicab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (N * N + 255) / 256;
    const long long items = chains * tiles, cap = (long long)sms * 8;
    taxicab_kernel<<<(unsigned)(items < cap ? items : cap), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        links, chains, N, kind, kappa, kappa_chain, tiles, out);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}
