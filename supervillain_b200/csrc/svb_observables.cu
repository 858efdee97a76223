// svb_observables.cu -- two-point observables that are not simple per-chain scalars.
//
// Spin_Spin.Villain (supervillain/observable/spin.py:28-42) through Lattice.correlation
// (supervillain/lattice/compact.py:465-536):
//     C[r] = N^-2 sum_x conj(s[x]) s[x - r],   s = exp(i phi)
// The reference evaluates it with three FFTs; here it is the direct O(N^4) sum out of shared
// memory, which is exact to rounding and cheap for the lattices that fit an SM (N <= 64).

#include "svb_common.cuh"

namespace svb {

template <typename real>
__global__ void __launch_bounds__(256) villain_spin_spin_kernel(const real* __restrict__ phi, long long chains, int N,
                                                                double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int V = N * N;
    double* sre = reinterpret_cast<double*>(smem_raw);
    double* sim = sre + V;
    const double inv_V = 1.0 / (double)V;
    for (long long chain = blockIdx.x; chain < chains; chain += gridDim.x) {
        const real* g = phi + chain * V;
        for (int i = threadIdx.x; i < V; i += blockDim.x) {
            double s, c;
            sincos((double)g[i], &s, &c);
            sre[i] = c;
            sim[i] = s;
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

}  // namespace svb

using namespace svb;

extern "C" int svb_villain_spin_spin(const void* phi, int phi_dtype, int64_t chains, int N, double* out, void* stream) {
    if (!phi || !out) return fail(SVB_E_NULL, "svb_villain_spin_spin: phi and out are required");
    if (chains < 0 || N < 1) return fail(SVB_E_SHAPE, "svb_villain_spin_spin: shape");
    if (phi_dtype != SVB_F64 && phi_dtype != SVB_F32) return fail(SVB_E_DTYPE, "svb_villain_spin_spin: dtype %d", phi_dtype);
    if (chains == 0) return SVB_OK;
    const size_t smem = (size_t)2 * N * N * sizeof(double);
    int dev = 0, max_smem = 0, sms = 0;
    SVB_CUDA_TRY(cudaGetDevice(&dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > (size_t)max_smem)
        return fail(SVB_E_UNSUPPORTED, "svb_villain_spin_spin: direct evaluation needs the lattice in shared memory (N=%d too large)", N);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long grid = chains < (long long)sms * 4 ? chains : (long long)sms * 4;
    if (phi_dtype == SVB_F64) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(villain_spin_spin_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        villain_spin_spin_kernel<double><<<(unsigned)grid, 256, smem, st>>>(reinterpret_cast<const double*>(phi), chains, N, out);
    } else {
        SVB_CUDA_TRY(cudaFuncSetAttribute(villain_spin_spin_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        villain_spin_spin_kernel<float><<<(unsigned)grid, 256, smem, st>>>(reinterpret_cast<const float*>(phi), chains, N, out);
    }
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}
