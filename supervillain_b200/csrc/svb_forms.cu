// svb_forms.cu -- the lattice form operators d, delta, face_sum, coface_sum (D = 2) for batches
// of forms, dtype-preserving, replacing supervillain/lattice/compact.py:954-1037 and the numba
// kernels of supervillain/lattice/_kernels.py:19-46.
//
// The arithmetic follows the reference kernels literally -- accumulate into a zero, one
// incidence-table row at a time, with face/coface sums as two separate adds per row -- so float
// results are bit-equal to the reference (its own tests compare with ==,
// test/test_lattice_kernels.py:17-34).  Incidence rows (out, in, axis, sign), compact.py:144-174:
//   d,0: (0,0,0,+)(1,0,1,+)   d,1: (0,1,0,+)(0,0,1,-)   delta,1: (0,0,0,+)(0,1,1,+)   delta,2: (0,0,1,-)(1,0,0,+)
//
// One thread produces ALL output components of one site from one pass over its inputs; threads of
// a warp walk the contiguous axis, neighbouring rows come through L1/L2, so HBM traffic is one
// read of the input form and one write of the output form.

#include "svb_common.cuh"

namespace svb {

template <typename T>
struct Acc {
    // res += s * (a - b) with s = +/-1, exactly as `res[oi] += sign * (F[sn] - F[s0])`
    static __device__ __forceinline__ T plus_diff(T res, int s, T a, T b) {
        T dlt = a - b;
        return res + (s > 0 ? dlt : -dlt);
    }
    static __device__ __forceinline__ T minus_diff(T res, int s, T a, T b) {
        T dlt = a - b;
        return res - (s > 0 ? dlt : -dlt);
    }
};

template <int OP, int DEG, typename T>
__global__ void __launch_bounds__(256) form_op_kernel(const T* __restrict__ in, T* __restrict__ out, long long chains, int N) {
    constexpr int CIN = (DEG == 1) ? 2 : 1;
    constexpr bool UP = (OP == SVB_OP_D || OP == SVB_OP_COFACE_SUM);
    constexpr int COUT = UP ? ((DEG == 0) ? 2 : 1) : ((DEG == 2) ? 2 : 1);
    const long long V = (long long)N * N;
    const long long total = chains * V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long chain = i / V;
        const int site = (int)(i - chain * V);
        const int x0 = site / N, x1 = site - x0 * N;
        const int xp0 = (x0 + 1 == N) ? 0 : x0 + 1, xm0 = (x0 == 0) ? N - 1 : x0 - 1;
        const int xp1 = (x1 + 1 == N) ? 0 : x1 + 1, xm1 = (x1 == 0) ? N - 1 : x1 - 1;
        const T* f = in + chain * CIN * V;
        T* g = out + chain * COUT * V;
        const int c = site, p0 = xp0 * N + x1, m0 = xm0 * N + x1, p1 = x0 * N + xp1, m1 = x0 * N + xm1;
        const T zero = (T)0;
        if (OP == SVB_OP_D && DEG == 0) {
            g[c] = Acc<T>::plus_diff(zero, +1, f[p0], f[c]);
            g[V + c] = Acc<T>::plus_diff(zero, +1, f[p1], f[c]);
        } else if (OP == SVB_OP_D && DEG == 1) {
            T r = Acc<T>::plus_diff(zero, +1, f[V + p0], f[V + c]);
            r = Acc<T>::plus_diff(r, -1, f[p1], f[c]);
            g[c] = r;
        } else if (OP == SVB_OP_DELTA && DEG == 1) {
            T r = Acc<T>::minus_diff(zero, +1, f[c], f[m0]);
            r = Acc<T>::minus_diff(r, +1, f[V + c], f[V + m1]);
            g[c] = r;
        } else if (OP == SVB_OP_DELTA && DEG == 2) {
            g[c] = Acc<T>::minus_diff(zero, -1, f[c], f[m1]);
            g[V + c] = Acc<T>::minus_diff(zero, +1, f[c], f[m0]);
        } else if (OP == SVB_OP_FACE_SUM && DEG == 1) {
            T r = zero + f[c];
            r = r + f[m0];
            r = r + f[V + c];
            r = r + f[V + m1];
            g[c] = r;
        } else if (OP == SVB_OP_FACE_SUM && DEG == 2) {
            g[c] = (zero + f[c]) + f[m1];
            g[V + c] = (zero + f[c]) + f[m0];
        } else if (OP == SVB_OP_COFACE_SUM && DEG == 0) {
            g[c] = (zero + f[c]) + f[p0];
            g[V + c] = (zero + f[c]) + f[p1];
        } else if (OP == SVB_OP_COFACE_SUM && DEG == 1) {
            T r = zero + f[V + c];
            r = r + f[V + p0];
            r = r + f[c];
            r = r + f[p1];
            g[c] = r;
        }
    }
}

// Vectorised form of the same kernel: a thread produces VEC = 16 B / sizeof(T) consecutive sites of a row from 128-bit loads
// (the rows above / below come as whole vectors, the in-row neighbours as the vector shifted by one plus one scalar), so a
// 4-byte form moves 16 B per thread-step instead of 4.  The arithmetic per element is the scalar kernel's, bit for bit.
template <typename T, int VEC>
struct alignas(16) FormVec {
    T v[VEC];
};

template <int OP, int DEG, typename T, int VEC>
__global__ void __launch_bounds__(256) form_op_vec_kernel(const T* __restrict__ in, T* __restrict__ out, long long chains, int N) {
    constexpr int CIN = (DEG == 1) ? 2 : 1;
    constexpr bool UP = (OP == SVB_OP_D || OP == SVB_OP_COFACE_SUM);
    constexpr int COUT = UP ? ((DEG == 0) ? 2 : 1) : ((DEG == 2) ? 2 : 1);
    using Vec = FormVec<T, VEC>;
    const long long V = (long long)N * N;
    const int groups_per_row = N / VEC;
    const long long groups = chains * (long long)N * groups_per_row;
    const T zero = (T)0;
    for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (long long)gridDim.x * blockDim.x) {
        const long long row = gi / groups_per_row;                 // chain * N + x0
        const int x1 = (int)(gi - row * groups_per_row) * VEC;
        const long long chain = row / N;
        const int x0 = (int)(row - chain * N);
        const int xp0 = (x0 + 1 == N) ? 0 : x0 + 1, xm0 = (x0 == 0) ? N - 1 : x0 - 1;
        const int xl = (x1 == 0) ? N - 1 : x1 - 1, xr = (x1 + VEC == N) ? 0 : x1 + VEC;
        const T* f = in + chain * CIN * V;
        T* g = out + chain * COUT * V;
        const int c = x0 * N + x1;
        auto ld = [&](int comp, int r0) { return *reinterpret_cast<const Vec*>(f + comp * V + r0 * N + x1); };
        // in-row neighbours of component `comp`: element k+1 (right) and k-1 (left)
        auto right_of = [&](const Vec& C, int comp) { Vec R;
#pragma unroll
            for (int k = 0; k + 1 < VEC; ++k) R.v[k] = C.v[k + 1];
            R.v[VEC - 1] = f[comp * V + x0 * N + xr]; return R; };
        auto left_of = [&](const Vec& C, int comp) { Vec Lf;
#pragma unroll
            for (int k = 1; k < VEC; ++k) Lf.v[k] = C.v[k - 1];
            Lf.v[0] = f[comp * V + x0 * N + xl]; return Lf; };
        Vec o0, o1;
        if (OP == SVB_OP_D && DEG == 0) {
            const Vec C = ld(0, x0), P0 = ld(0, xp0), p1 = right_of(C, 0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { o0.v[k] = Acc<T>::plus_diff(zero, +1, P0.v[k], C.v[k]); o1.v[k] = Acc<T>::plus_diff(zero, +1, p1.v[k], C.v[k]); }
        } else if (OP == SVB_OP_D && DEG == 1) {
            const Vec C0 = ld(0, x0), C1 = ld(1, x0), P1 = ld(1, xp0), p0 = right_of(C0, 0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { T r = Acc<T>::plus_diff(zero, +1, P1.v[k], C1.v[k]); o0.v[k] = Acc<T>::plus_diff(r, -1, p0.v[k], C0.v[k]); }
        } else if (OP == SVB_OP_DELTA && DEG == 1) {
            const Vec C0 = ld(0, x0), C1 = ld(1, x0), M0 = ld(0, xm0), m1 = left_of(C1, 1);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { T r = Acc<T>::minus_diff(zero, +1, C0.v[k], M0.v[k]); o0.v[k] = Acc<T>::minus_diff(r, +1, C1.v[k], m1.v[k]); }
        } else if (OP == SVB_OP_DELTA && DEG == 2) {
            const Vec C = ld(0, x0), M0 = ld(0, xm0), m1 = left_of(C, 0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { o0.v[k] = Acc<T>::minus_diff(zero, -1, C.v[k], m1.v[k]); o1.v[k] = Acc<T>::minus_diff(zero, +1, C.v[k], M0.v[k]); }
        } else if (OP == SVB_OP_FACE_SUM && DEG == 1) {
            const Vec C0 = ld(0, x0), C1 = ld(1, x0), M0 = ld(0, xm0), m1 = left_of(C1, 1);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { T r = zero + C0.v[k]; r = r + M0.v[k]; r = r + C1.v[k]; o0.v[k] = r + m1.v[k]; }
        } else if (OP == SVB_OP_FACE_SUM && DEG == 2) {
            const Vec C = ld(0, x0), M0 = ld(0, xm0), m1 = left_of(C, 0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { o0.v[k] = (zero + C.v[k]) + m1.v[k]; o1.v[k] = (zero + C.v[k]) + M0.v[k]; }
        } else if (OP == SVB_OP_COFACE_SUM && DEG == 0) {
            const Vec C = ld(0, x0), P0 = ld(0, xp0), p1 = right_of(C, 0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { o0.v[k] = (zero + C.v[k]) + P0.v[k]; o1.v[k] = (zero + C.v[k]) + p1.v[k]; }
        } else if (OP == SVB_OP_COFACE_SUM && DEG == 1) {
            const Vec C0 = ld(0, x0), C1 = ld(1, x0), P1 = ld(1, xp0), p0 = right_of(C0, 0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { T r = zero + C1.v[k]; r = r + P1.v[k]; r = r + C0.v[k]; o0.v[k] = r + p0.v[k]; }
        }
        *reinterpret_cast<Vec*>(g + c) = o0;
        if (COUT == 2) *reinterpret_cast<Vec*>(g + V + c) = o1;
    }
}

template <int OP, int DEG, typename T>
static int launch_form_op(const void* in, void* out, long long chains, int N, cudaStream_t stream) {
    const long long total = chains * (long long)N * N;
    const long long cap = 148LL * 32;
    constexpr int VEC = 16 / (int)sizeof(T);
    if (N % VEC == 0 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
        long long vblocks = (total / VEC + 255) / 256;
        if (vblocks > cap) vblocks = cap;
        form_op_vec_kernel<OP, DEG, T, VEC><<<(unsigned)vblocks, 256, 0, stream>>>(reinterpret_cast<const T*>(in),
                                                                                 reinterpret_cast<T*>(out), chains, N);
        SVB_CUDA_TRY(cudaGetLastError());
        return 0;
    }
    long long blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    form_op_kernel<OP, DEG, T><<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out),
                                                                     chains, N);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T>
static int dispatch_form_op(int op, int degree, const void* in, void* out, long long chains, int N, cudaStream_t st) {
    if (op == SVB_OP_D && degree == 0) return launch_form_op<SVB_OP_D, 0, T>(in, out, chains, N, st);
    if (op == SVB_OP_D && degree == 1) return launch_form_op<SVB_OP_D, 1, T>(in, out, chains, N, st);
    if (op == SVB_OP_DELTA && degree == 1) return launch_form_op<SVB_OP_DELTA, 1, T>(in, out, chains, N, st);
    if (op == SVB_OP_DELTA && degree == 2) return launch_form_op<SVB_OP_DELTA, 2, T>(in, out, chains, N, st);
    if (op == SVB_OP_FACE_SUM && degree == 1) return launch_form_op<SVB_OP_FACE_SUM, 1, T>(in, out, chains, N, st);
    if (op == SVB_OP_FACE_SUM && degree == 2) return launch_form_op<SVB_OP_FACE_SUM, 2, T>(in, out, chains, N, st);
    if (op == SVB_OP_COFACE_SUM && degree == 0) return launch_form_op<SVB_OP_COFACE_SUM, 0, T>(in, out, chains, N, st);
    if (op == SVB_OP_COFACE_SUM && degree == 1) return launch_form_op<SVB_OP_COFACE_SUM, 1, T>(in, out, chains, N, st);
    return fail(SVB_E_PARAM, "svb_form_op: op %d is the scalar 0 on a %d-form in D=2 (or unknown op)", op, degree);
}

}  // namespace svb

using namespace svb;

extern "C" int svb_form_op(int op, int degree, int dtype, const void* in, void* out, int64_t chains, int N, void* stream) {
    if (!in || !out) return fail(SVB_E_NULL, "svb_form_op: in and out are required");
    if (in == out) return fail(SVB_E_PARAM, "svb_form_op: in-place operation is not supported");
    if (chains < 0 || N < 1 || N > 32768) return fail(SVB_E_SHAPE, "svb_form_op: chains=%lld N=%d", (long long)chains, N);
    if (degree < 0 || degree > 2) return fail(SVB_E_PARAM, "svb_form_op: degree %d", degree);
    if (chains == 0) return SVB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case SVB_F64: return dispatch_form_op<double>(op, degree, in, out, chains, N, st);
        case SVB_F32: return dispatch_form_op<float>(op, degree, in, out, chains, N, st);
        case SVB_I32: return dispatch_form_op<int32_t>(op, degree, in, out, chains, N, st);
        case SVB_I64: return dispatch_form_op<long long>(op, degree, in, out, chains, N, st);
        default: return fail(SVB_E_DTYPE, "svb_form_op: dtype %d", dtype);
    }
}
