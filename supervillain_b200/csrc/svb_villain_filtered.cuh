// svb_villain_filtered.cuh -- the production Villain sweep kernel (included by svb_villain.cu).
//
// NeighborhoodUpdate.step (supervillain/generator/villain/neighborhood.py:59-137) for Philox draws, fp64 phi, FAST
// arithmetic and N in {16, 32, 64}, organised so that almost no fp64 instruction and only half a Philox block is spent
// per proposal:
//
//  * The reference's residual field r = d(phi) - 2 pi n (neighborhood.py:91) is built once per sweep in fp64 and kept
//    in shared memory ROUNDED TO fp32, in colour-separated arrays (conflict-free for every access of a colour pass).
//  * A proposal is decided in fp32: dS32 from the four fp32 residuals, L32 = -ln u from one MUFU.LG2, and the
//    decision u < e^-dS <=> dS < -ln u is taken whenever |dS32 - L32| exceeds a band that bounds every fp32 error of
//    the comparison (derivation at FilterConsts) and the width of the bracket in which u is known.  Inside the band
//    (a few 1e-5 of the proposals) the exact test is evaluated: dS in fp64 recomputed from phi and n, the fp64
//    exponential, and the lazily refined uniform.  Every decision therefore equals the exact fp64 decision: the
//    full-size tests are bit-exact against the oracle.
//  * phi and n are touched only when a proposal is accepted (phi += dphi in fp64, the reference's one rounding).
//  * One Philox4x32-10 block serves the two sites (x0, x1), (x0 ^ 8, x1) that a thread owns (draw mapping version 2).
//
// Geometry: T = 4 N threads per CTA (so a thread's sites of one colour are rows x0, x0 + 8, x0 + 16, ... of one
// column), one CTA per chain at a time, grid-stride over chains, chains moved by 1-D TMA bulk copies.
#pragma once
// (included inside namespace svb)

// Band of the fp32 decision.  With D = interval_phi + 2 pi W interval_n >= |dr|, R = max |r| over the four links,
// eps = 2^-24 and
//   |dphi32 - dphi| <= 2 I 2^-24 (23 leading bits, centred) + rounding      -> |delta dr| <= 1.0e-6 (1 + D/10)
//   |r32 - r| <= eps |r| + 1.2e-6 (one fp32 patch by an accepted neighbour of the other colour)
// the error of one term dr (2 r + dr) is bounded by (|2r + dr| + |dr|) delta_dr + |dr| (2 delta_r) + 3 eps |dr||2r + dr|,
// and four terms plus their accumulation give
//   |dS32 - dS| <= (kappa/2) (bA R + bB),   bA = 8e-6 + 3.9e-6 D,   bB = 1.8e-5 D + 1.7e-6 D^2
// (constants rounded up).  The kernel uses twice that, plus 2e-5 + 4e-6 L for the logarithm (MUFU.LG2: < 1e-6 (1 + L))
// and for the bracket of u (relative half-width <= 2^-17 once f >= 2^16; smaller f always take the exact path).
struct FilterConsts {
    float I, two_I;          // interval_phi, 2 interval_phi
    float c;                 // 2 pi W
    float bA, bB;            // band coefficients (already doubled)
    float g_bias;            // 12582912 + interval_n: subtracting it from the planted digit gives dg as a float
};

static FilterConsts make_filter_consts(double interval_phi, int W, int interval_n) {
    FilterConsts fc;
    const double D = interval_phi + SVB_TWO_PI * W * interval_n;
    fc.I = (float)interval_phi;
    fc.two_I = (float)(2.0 * interval_phi);
    fc.c = (float)(SVB_TWO_PI * W);
    fc.bA = (float)(2.0 * (8e-6 + 3.9e-6 * D));
    fc.bB = (float)(2.0 * (1.8e-5 * D + 1.7e-6 * D * D));
    fc.g_bias = 12582912.0f + (float)interval_n;
    return fc;
}

// Everything the exact (cold) path needs about one proposal.
struct ExactProposal {
    const double* phi;
    const int32_t* n0;
    const int32_t* n1;
    int i_c, i_f0, i_b0, i_f1, i_b1;
    double half_kappa, c, dphi;
    int g[4];
    VillainDraw d;
    RefineCtx rc;
};

// The exact decision: dS in fp64 from the current phi and n (FAST arithmetic of villain_site_update), the fp64
// exponential, the lazily refined uniform.
__device__ __noinline__ bool villain_exact_decision(const ExactProposal& p) {
    const double pc = p.phi[p.i_c];
    const double r_f0 = fma(-SVB_TWO_PI, (double)p.n0[p.i_c], p.phi[p.i_f0] - pc);
    const double r_b0 = fma(-SVB_TWO_PI, (double)p.n0[p.i_b0], pc - p.phi[p.i_b0]);
    const double r_f1 = fma(-SVB_TWO_PI, (double)p.n1[p.i_c], p.phi[p.i_f1] - pc);
    const double r_b1 = fma(-SVB_TWO_PI, (double)p.n1[p.i_b1], pc - p.phi[p.i_b1]);
    const double dr_f0 = fma(-p.c, (double)p.g[0], -p.dphi), dr_b0 = fma(-p.c, (double)p.g[1], p.dphi);
    const double dr_f1 = fma(-p.c, (double)p.g[2], -p.dphi), dr_b1 = fma(-p.c, (double)p.g[3], p.dphi);
    double acc2 = dr_f0 * fma(2.0, r_f0, dr_f0);
    acc2 = fma(dr_b0, fma(2.0, r_b0, dr_b0), acc2);
    acc2 = fma(dr_f1, fma(2.0, r_f1, dr_f1), acc2);
    acc2 = fma(dr_b1, fma(2.0, r_b1, dr_b1), acc2);
    const double dS = p.half_kappa * acc2;
    return villain_decide_lazy(exp_clipped(-dS), p.d, p.rc);
}

// fp64 residuals of the forward links of the sites (x0, 2k) and (x0, 2k + 1); jj = x0 N/2 + k.
template <int N>
struct PairResiduals {
    double r0e, r0o, r1e, r1o;
    int2 a0, a1;
};
template <int N>
__device__ __forceinline__ PairResiduals<N> villain_pair_residuals(const double* __restrict__ sphi, const int32_t* __restrict__ sn0,
                                                                    const int32_t* __restrict__ sn1, int jj) {
    constexpr int HN = N / 2;
    const int i = 2 * jj;
    const int x0 = jj / HN, x1 = 2 * (jj - x0 * HN);
    const int iup = ((x0 + 1) & (N - 1)) * N + x1;
    const double2 pc = *reinterpret_cast<const double2*>(sphi + i);
    const double2 pu = *reinterpret_cast<const double2*>(sphi + iup);
    const double pr = sphi[x0 * N + ((x1 + 2) & (N - 1))];
    PairResiduals<N> o;
    o.a0 = *reinterpret_cast<const int2*>(sn0 + i);
    o.a1 = *reinterpret_cast<const int2*>(sn1 + i);
    o.r0e = fma(-SVB_TWO_PI, int_to_double(o.a0.x), pu.x - pc.x);
    o.r0o = fma(-SVB_TWO_PI, int_to_double(o.a0.y), pu.y - pc.y);
    o.r1e = fma(-SVB_TWO_PI, int_to_double(o.a1.x), pc.y - pc.x);
    o.r1o = fma(-SVB_TWO_PI, int_to_double(o.a1.y), pr - pc.y);
    return o;
}

template <int NT, int MINB, int STAGES>
__global__ void __launch_bounds__(4 * NT, MINB) villain_smem_filtered_kernel(const __grid_constant__ VillainArgs a,
                                                                             const __grid_constant__ FilterConsts fc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, VH = V / 2, T = 4 * NT;
    constexpr int PER = VH / T;                                  // sites per thread per colour (rows x0 + 8 q)
    static_assert(PER >= 2 && PER % 2 == 0, "villain_smem_filtered_kernel: unsupported geometry");
    constexpr uint32_t bytes_phi = V * sizeof(double);
    constexpr uint32_t bytes_n = 2 * V * sizeof(int32_t);
    constexpr uint32_t stage_bytes = bytes_phi + bytes_n;
    const int tid = threadIdx.x;
    float* rc0 = reinterpret_cast<float*>(smem_raw + STAGES * stage_bytes);      // [colour][VH]: residual of link (0, x)
    float* rc1 = rc0 + V;                                                         // [colour][VH]: residual of link (1, x)
    double* scratch = reinterpret_cast<double*>(rc1 + V);                         // 6 * 32 doubles
    uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + 6 * 32);

    if (tid == 0) {
        for (int b = 0; b < STAGES; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const bool want_obs = a.obs != nullptr;
    const uint32_t K = (uint32_t)(2 * a.interval_n + 1);
    const int W = a.W, interval_n = a.interval_n;

    // per-thread geometry: rows row8 + 8 q of the compact column k
    const int row8 = tid / HN, k = tid - row8 * HN;
    const int cc = row8 & 1;                                      // colour of the even-column site of this thread's pairs
    const int wrap0 = (row8 == 0) ? VH : 0;                       // backward-0 neighbour of row 0 is row N - 1

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
        const float hk = (float)half_kappa;
        const float hkA = 1.0001f * hk * fc.bA, hkB = 1.0001f * hk * fc.bB + 2e-5f;

        mbar_wait(&bar[b], (uint32_t)((STAGES == 2 ? (it >> 1) : it) & 1));

        int n_acc = 0;
        double sum_A_all = 0.0;
        for (int s = 0; s < a.n_sweeps; ++s) {
            float sum_A = 0.0f;
            // ---- r = d(phi) - 2 pi n   (neighborhood.py:91) in fp64, stored rounded to fp32 ----
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int jj = tid + T * q;
                const PairResiduals<N> pr = villain_pair_residuals<N>(sphi, sn0, sn1, jj);
                rc0[cc * VH + jj] = (float)pr.r0e;
                rc1[cc * VH + jj] = (float)pr.r1e;
                rc0[(cc ^ 1) * VH + jj] = (float)pr.r0o;
                rc1[(cc ^ 1) * VH + jj] = (float)pr.r1o;
            }
            __syncthreads();

            const unsigned long long gc = a.chain0 + (unsigned long long)chain, gs = a.sweep0 + (unsigned long long)s;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int par = (row8 + c) & 1;                    // column parity of this thread's sites of colour c
                const int x1 = 2 * k + par;
                float* R0own = rc0 + c * VH;
                float* R1own = rc1 + c * VH;
                float* R0oth = rc0 + (c ^ 1) * VH;
                float* R1oth = rc1 + (c ^ 1) * VH;
                const int ob1 = par ? 0 : ((k == 0) ? (1 - HN) : 1);          // compact index of x - e1 is j - ob1
                const int wrap1 = (x1 == 0) ? N : 0;
#pragma unroll
                for (int p = 0; p < PER / 2; ++p) {
                    const uint32_t c0 = (uint32_t)((row8 + 16 * p) * N + x1);                 // villain_pair_counter
                    const Philox4 bits = philox_site_keys(a, gc, gs, c0);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * p + h;
                        const int j = tid + T * q;
                        const int jb0 = j - HN + ((q == 0) ? wrap0 : 0);
                        const int jb1 = j - ob1;
                        const uint32_t wA = h ? bits.z : bits.x, wB = h ? bits.w : bits.y;
                        // proposal: four base-K digits, then the leading 32 bits of the uniform
                        uint32_t f = wB;
                        int dig[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint64_t prod = (uint64_t)f * K;
                            f = (uint32_t)prod;
                            dig[i] = (int)(prod >> 32);
                        }
                        const float U = __uint_as_float(0x3F800000u | (wA >> 9)) - 0.99999994f;       // in (0, 1), 23 bits, centred
                        const float dphi = fmaf(fc.two_I, U, -fc.I);
                        const float g0 = __int_as_float(0x4B400000 + dig[0]) - fc.g_bias;
                        const float g1 = __int_as_float(0x4B400000 + dig[1]) - fc.g_bias;
                        const float g2 = __int_as_float(0x4B400000 + dig[2]) - fc.g_bias;
                        const float g3 = __int_as_float(0x4B400000 + dig[3]) - fc.g_bias;
                        const float r_f0 = R0own[j], r_f1 = R1own[j], r_b0 = R0oth[jb0], r_b1 = R1oth[jb1];
                        // dr = d(dphi) - 2 pi dn   (neighborhood.py:110)
                        const float dr_f0 = fmaf(-fc.c, g0, -dphi), dr_b0 = fmaf(-fc.c, g1, dphi);
                        const float dr_f1 = fmaf(-fc.c, g2, -dphi), dr_b1 = fmaf(-fc.c, g3, dphi);
                        float acc2 = dr_f0 * fmaf(2.0f, r_f0, dr_f0);
                        acc2 = fmaf(dr_b0, fmaf(2.0f, r_b0, dr_b0), acc2);
                        acc2 = fmaf(dr_f1, fmaf(2.0f, r_f1, dr_f1), acc2);
                        acc2 = fmaf(dr_b1, fmaf(2.0f, r_b1, dr_b1), acc2);
                        const float dS = hk * acc2;
                        const float L = fmaf(__log2f((float)f + 0.5f), -0.6931471805599453f, 22.18070977791825f);   // -ln((f + 1/2) 2^-32)
                        const float Rmax = fmaxf(fmaxf(fabsf(r_f0), fabsf(r_b0)), fmaxf(fabsf(r_f1), fabsf(r_b1)));
                        const float band = fmaf(hkA, Rmax, fmaf(4e-6f, L, hkB));
                        const float diff = dS - L;
                        bool ok = diff < 0.0f;
                        sum_A += fminf(exp2f(-1.4426950408889634f * dS), 1.0f);
                        const int i_c = 2 * j + par;
                        const int i_b0 = i_c - N + ((q == 0) ? 2 * wrap0 : 0);
                        const int i_b1 = i_c - 1 + wrap1;
                        if (!(fabsf(diff) > band) || f < 65536u) {
                            ExactProposal ep;
                            ep.phi = sphi; ep.n0 = sn0; ep.n1 = sn1;
                            const int x0 = row8 + 8 * q;
                            ep.i_c = i_c; ep.i_b0 = i_b0; ep.i_b1 = i_b1;
                            ep.i_f0 = ((x0 + 1) & (N - 1)) * N + x1;
                            ep.i_f1 = x0 * N + ((x1 + 1) & (N - 1));
                            ep.half_kappa = half_kappa;
                            ep.c = SVB_TWO_PI * (double)W;
                            ep.dphi = villain_dphi_from_word(wA, a.interval_phi);
#pragma unroll
                            for (int i = 0; i < 4; ++i) ep.g[i] = dig[i] - interval_n;
                            ep.d.f = f; ep.d.c0 = c0; ep.d.half = (uint32_t)h;
                            ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs;
                            ok = villain_exact_decision(ep);
                        }
                        if (ok) {                                               // (:121-129)
                            sphi[i_c] = __dadd_rn(sphi[i_c], villain_dphi_from_word(wA, a.interval_phi));
                            sn0[i_c] += W * (dig[0] - interval_n);
                            sn0[i_b0] += W * (dig[1] - interval_n);
                            sn1[i_c] += W * (dig[2] - interval_n);
                            sn1[i_b1] += W * (dig[3] - interval_n);
                            R0own[j] = r_f0 + dr_f0;
                            R0oth[jb0] = r_b0 + dr_b0;
                            R1own[j] = r_f1 + dr_f1;
                            R1oth[jb1] = r_b1 + dr_b1;
                            ++n_acc;
                        }
                    }
                }
                __syncthreads();
                if (STAGES == 2 && s == 0 && c == 0 && tid == 0 && next < a.chains) {
                    bulk_wait_read0();
                    issue_load(next, b ^ 1);
                }
            }
            sum_A_all += (double)sum_A;
        }

        if (want_obs) {
            // one fp64 pass over site pairs: sum r^2, sum n, sum (dn)^2 of the final state
            ChainSums cs;
            cs.sumA = sum_A_all; cs.accepted = n_acc; cs.action = 0.0; cs.w0 = 0; cs.w1 = 0;
            long long dn2 = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int jj = tid + T * q;
                const PairResiduals<N> pr = villain_pair_residuals<N>(sphi, sn0, sn1, jj);
                cs.action = fma(pr.r0e, pr.r0e, cs.action);
                cs.action = fma(pr.r0o, pr.r0o, cs.action);
                cs.action = fma(pr.r1e, pr.r1e, cs.action);
                cs.action = fma(pr.r1o, pr.r1o, cs.action);
                const int x0 = row8 + 8 * q;
                const int iup = ((x0 + 1) & (N - 1)) * N + 2 * k;
                const int hr = sn0[x0 * N + ((2 * k + 2) & (N - 1))];                // n0[x + 2 e1]
                const int2 up = *reinterpret_cast<const int2*>(sn1 + iup);           // n1[x + e0]
                // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x])      (compact.py d,1 rows)
                const int d0 = (up.x - pr.a1.x) - (pr.a0.y - pr.a0.x), d1 = (up.y - pr.a1.y) - (hr - pr.a0.y);
                dn2 += (long long)d0 * d0 + (long long)d1 * d1;
                cs.w0 += pr.a0.x + pr.a0.y;
                cs.w1 += pr.a1.x + pr.a1.y;
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

template <int NT, int MINB, int STAGES>
static int launch_villain_filtered(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    auto kern = villain_smem_filtered_kernel<NT, MINB, STAGES>;
    const size_t V = (size_t)NT * NT;
    const size_t smem = STAGES * V * 16 + 2 * V * sizeof(float) + 6 * 32 * sizeof(double) + 16;
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 4 * NT, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "filtered villain kernel does not fit an SM at N=%d", NT);
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > a.chains) grid = a.chains;
    const FilterConsts fc = make_filter_consts(a.interval_phi, a.W, a.interval_n);
    kern<<<(unsigned)grid, 4 * NT, smem, stream>>>(a, fc);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

