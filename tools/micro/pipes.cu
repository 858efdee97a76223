// Pipe-throughput microbenchmarks for sm_100a: thread-ops per clock per SM for a few instruction kinds.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int KIND>
__global__ void bench(double* out, long long* clocks, int n) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3, u4 = u0 + 5, u5 = u0 + 7, u6 = u0 + 11, u7 = u0 + 13;
    int i0 = threadIdx.x, i1 = i0 * 3, i2 = i0 * 5, i3 = i0 * 7;
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
        if (KIND == 0) {  // DFMA, 8 independent chains
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        } else if (KIND == 1) {  // IMAD.WIDE.U32
            unsigned long long p0 = (unsigned long long)u0 * 0xD2511F53u, p1 = (unsigned long long)u1 * 0xCD9E8D57u;
            unsigned long long p2 = (unsigned long long)u2 * 0xD2511F53u, p3 = (unsigned long long)u3 * 0xCD9E8D57u;
            unsigned long long p4 = (unsigned long long)u4 * 0xD2511F53u, p5 = (unsigned long long)u5 * 0xCD9E8D57u;
            unsigned long long p6 = (unsigned long long)u6 * 0xD2511F53u, p7 = (unsigned long long)u7 * 0xCD9E8D57u;
            u0 = (unsigned)(p0 >> 32) ^ (unsigned)p0; u1 = (unsigned)(p1 >> 32) ^ (unsigned)p1;
            u2 = (unsigned)(p2 >> 32) ^ (unsigned)p2; u3 = (unsigned)(p3 >> 32) ^ (unsigned)p3;
            u4 = (unsigned)(p4 >> 32) ^ (unsigned)p4; u5 = (unsigned)(p5 >> 32) ^ (unsigned)p5;
            u6 = (unsigned)(p6 >> 32) ^ (unsigned)p6; u7 = (unsigned)(p7 >> 32) ^ (unsigned)p7;
        } else if (KIND == 2) {  // I2F.F64 + DADD
            a0 += (double)i0; a1 += (double)i1; a2 += (double)i2; a3 += (double)i3;
            i0 += 1; i1 += 3; i2 += 5; i3 += 7;
        } else if (KIND == 3) {  // LOP3 / IADD mix (alu pipe)
            u0 = (u0 ^ u1) + u2; u1 = (u1 ^ u2) + u3; u2 = (u2 ^ u3) + u4; u3 = (u3 ^ u4) + u5;
            u4 = (u4 ^ u5) + u6; u5 = (u5 ^ u6) + u7; u6 = (u6 ^ u7) + u0; u7 = (u7 ^ u0) + u1;
        } else if (KIND == 4) {  // DFMA + independent integer work interleaved (can they co-issue at full rate?)
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            u0 = (u0 ^ u1) + u2; u1 = (u1 ^ u2) + u3; u2 = (u2 ^ u3) + u4; u3 = (u3 ^ u4) + u5;
            unsigned long long p0 = (unsigned long long)u4 * 0xD2511F53u, p1 = (unsigned long long)u5 * 0xCD9E8D57u;
            u4 = (unsigned)(p0 >> 32) ^ (unsigned)p0; u5 = (unsigned)(p1 >> 32) ^ (unsigned)p1;
        } else if (KIND == 5) {  // dependent DFMA chain (latency)
            a0 = fma(a0, m, c); a0 = fma(a0, m, c); a0 = fma(a0, m, c); a0 = fma(a0, m, c);
            a0 = fma(a0, m, c); a0 = fma(a0, m, c); a0 = fma(a0, m, c); a0 = fma(a0, m, c);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + u0 + u1 + u2 + u3 + u4 + u5 + u6 + u7 + i0 + i1 + i2 + i3;
}

template <int KIND>
void run(const char* name, int ops_per_iter, int threads, int blocks_per_sm) {
    int sms = 148;
    double* out; long long* clk;
    cudaMalloc(&out, sizeof(double) * sms * blocks_per_sm * threads);
    cudaMalloc(&clk, sizeof(long long) * sms * blocks_per_sm);
    bench<KIND><<<sms * blocks_per_sm, threads>>>(out, clk, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<KIND><<<sms * blocks_per_sm, threads>>>(out, clk, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double ops_per_clk_sm = (double)ops_per_iter * ITERS * threads * blocks_per_sm / (double)h;
    printf("%-28s threads/SM=%4d  %7.1f thread-ops/clk/SM   (%.3f ms, %lld clk, %.2f GHz eff)\n", name, threads * blocks_per_sm,
           ops_per_clk_sm, ms, h, h / (ms * 1e6));
    cudaFree(out); cudaFree(clk);
}

int main() {
    for (int w : {256, 512, 1024}) {
        run<0>("DFMA x8 indep", 8, 256, w / 256);
        run<1>("IMAD.WIDE+LOP x8", 8, 256, w / 256);
        run<2>("I2F.F64+DADD x4", 4, 256, w / 256);
        run<3>("LOP3/IADD x8 (2 instr each)", 8, 256, w / 256);
        run<4>("mix 4 DFMA+8 alu+2 wide", 14, 256, w / 256);
        run<5>("DFMA dependent x8", 8, 256, w / 256);
    }
    return 0;
}
