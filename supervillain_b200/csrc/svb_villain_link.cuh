// svb_villain_link.cuh -- LinkUpdate (supervillain/generator/villain/link.py:53-101) for chains that fit shared memory
// (included by svb_villain.cu inside namespace svb, after svb_villain_filtered.cuh).
//
// A LinkUpdate proposal changes n on ONE link by c (a nonzero multiple of W) and its dS depends on nothing but that link's
// own residual, so all 2 N^2 proposals of a sweep are independent: no colouring, no barriers, phi constant.  One CTA per
// chain: phi and n arrive by TMA (16 B per site), every thread owns the forward links of its site pairs for all fused
// sweeps, the observables of the final state are summed from shared memory, and only n goes back (8 B per site): 24 B per
// site and sweep instead of the three launches (zero, update out of global memory with per-block atomics, observables) of
// villain_link_kernel.  Arithmetic: dS in the reference's operation order (link.py:83-86), bit for bit
// villain_link_kernel's; the Metropolis test through the fp32 log filter with the exact test inside its guard band, so
// every decision is the exact one; the acceptance statistic in fp32 (1e-5, as in the production sweep).
#pragma once

struct LinkProposal {
    double dS;
    LazyUniform lu;
    RefineCtx rc;
};
static __device__ __noinline__ bool villain_link_exact_decision(const LinkProposal& p) {
    return decide_lazy(exp_clipped(-p.dS), p.lu, STREAM_VILLAIN_LINK_REFINE, p.rc);
}

template <int NT, int MINB>
__global__ void __launch_bounds__(4 * NT, MINB) villain_link_smem_kernel(const __grid_constant__ LinkArgs a, int n_sweeps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, T = 4 * NT, NW = T / 32, PER = (V / 2) / T;
    constexpr uint32_t bytes_phi = V * sizeof(double), bytes_n = 2 * V * sizeof(int32_t);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* sphi = reinterpret_cast<double*>(smem_raw);
    int32_t* sn0 = reinterpret_cast<int32_t*>(smem_raw + bytes_phi);
    int32_t* sn1 = sn0 + V;
    double* red_state = reinterpret_cast<double*>(smem_raw + bytes_phi + bytes_n);
    double* red_count = red_state + 4 * NW;
    uint64_t* bar = reinterpret_cast<uint64_t*>(red_count + 2 * NW);
    constexpr int kWriter = 32;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int row8 = tid / HN, k = tid - row8 * HN;
    const int up_off = (row8 == 7) ? (N - V) : N;                 // the row below the last of this thread's rows wraps
    const uint32_t K = (uint32_t)(2 * a.interval);
    auto issue_load = [&](long long chain) {
        mbar_expect_tx(bar, bytes_phi + bytes_n);
        bulk_g2s(sphi, a.phi + chain * V, bytes_phi, bar);
        bulk_g2s(sn0, a.n + chain * 2 * V, bytes_n, bar);
    };
    long long chain = blockIdx.x;
    if (tid == 0 && chain < a.chains) issue_load(chain);

    for (int it = 0; chain < a.chains; chain += gridDim.x, ++it) {
        const long long next = chain + gridDim.x;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double m2pik = __dmul_rn(-SVB_TWO_PI, kappa);
        const double* pp = sphi + row8 * N + 2 * k;
        const double* pp_r = sphi + row8 * N + ((2 * k + 2) & (N - 1));
        int32_t* pn0 = sn0 + row8 * N + 2 * k;
        int32_t* pn1 = sn1 + row8 * N + 2 * k;
        mbar_wait(bar, (uint32_t)(it & 1));

        int n_acc = 0;
        double sum_A_all = 0.0;
        RefineCtx rc;
        rc.seed = a.seed; rc.chain = a.chain0 + (unsigned long long)chain;
#pragma unroll 1
        for (int s = 0; s < n_sweeps; ++s) {
            rc.sweep = a.sweep + (unsigned long long)s;
            float sum_A = 0.0f;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int o = 8 * N * q;
                const int uo = (q == PER - 1) ? up_off : N;
                const double2 pc = *reinterpret_cast<const double2*>(pp + o);
                const double2 pu = *reinterpret_cast<const double2*>(pp + o + uo);
                const double pr = pp_r[o];
                // d(phi) on the four forward links of the pair   (link.py:74)
                const double dphi[4] = {__dsub_rn(pu.x, pc.x), __dsub_rn(pc.y, pc.x), __dsub_rn(pu.y, pc.y), __dsub_rn(pr, pc.y)};
                int2 m0 = *reinterpret_cast<const int2*>(pn0 + o), m1 = *reinterpret_cast<const int2*>(pn1 + o);
                int nl[4] = {m0.x, m1.x, m0.y, m1.y};             // (site e: mu 0, mu 1), (site o: mu 0, mu 1)
                const int site_e = (row8 + 8 * q) * N + 2 * k;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t site = (uint32_t)(site_e + h);
                    const Philox4 p = philox_site(a.seed, rc.chain, rc.sweep, site, STREAM_VILLAIN_LINK);
#pragma unroll
                    for (int mu = 0; mu < 2; ++mu) {
                        const int j = 2 * h + mu;
                        const uint64_t pz = (uint64_t)(mu ? p.y : p.x) * (uint64_t)K;
                        const int idx = (int)(pz >> 32);
                        const uint32_t f = (uint32_t)pz;
                        const int c = a.W * ((idx < a.interval) ? idx - a.interval : idx - a.interval + 1);
                        // dS = -2 pi kappa change (dphi - 2 pi n - pi change), numpy's left-to-right order   (link.py:83-86)
                        const double t1 = __dmul_rn(m2pik, (double)c);
                        const double t2 = __dsub_rn(__dsub_rn(dphi[j], __dmul_rn(SVB_TWO_PI, (double)nl[j])),
                                                    __dmul_rn(3.141592653589793116, (double)c));
                        const double dS = __dmul_rn(t1, t2);
                        double prob;
                        const double u_mid = (__hiloint2double(0x43300000, (int)f) - 4503599627370495.5) * 2.3283064365386963e-10;
                        int r = (f >= 65536u) ? metropolis_log_filter(dS, u_mid, 7.62939453125e-06f, prob) : -1;
                        if (r < 0) {
                            LinkProposal lp;
                            lp.dS = dS; lp.lu.f = f; lp.lu.c0 = site; lp.lu.word = (uint32_t)mu; lp.rc = rc;
                            prob = exp_clipped(-dS);
                            r = villain_link_exact_decision(lp) ? 1 : 0;
                        }
                        sum_A += (float)prob;
                        n_acc += r;
                        nl[j] += r ? c : 0;
                    }
                }
                *reinterpret_cast<int2*>(pn0 + o) = make_int2(nl[0], nl[2]);
                *reinterpret_cast<int2*>(pn1 + o) = make_int2(nl[1], nl[3]);
            }
            sum_A_all += (double)sum_A;
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(a.n + chain * 2 * V, sn0, bytes_n);           // phi did not change
            bulk_commit();
        }
        if (a.obs) {
            double action = 0.0;
            int w0 = 0, w1 = 0;
            long long dn2 = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int o = 8 * N * q;
                const int uo = (q == PER - 1) ? up_off : N;
                const PairResiduals pr = villain_pair_residuals(pp + o, pp + o + uo, pp_r + o, pn0 + o, pn1 + o);
                action = fma(pr.r0e, pr.r0e, action);
                action = fma(pr.r0o, pr.r0o, action);
                action = fma(pr.r1e, pr.r1e, action);
                action = fma(pr.r1o, pr.r1o, action);
                const int hr = sn0[(row8 + 8 * q) * N + ((2 * k + 2) & (N - 1))];
                const int2 up = *reinterpret_cast<const int2*>(pn1 + o + uo);
                const int d0 = (up.x - pr.a1.x) - (pr.a0.y - pr.a0.x), d1 = (up.y - pr.a1.y) - (hr - pr.a0.y);
                dn2 += (long long)d0 * d0 + (long long)d1 * d1;
                w0 += pr.a0.x + pr.a0.y;
                w1 += pr.a1.x + pr.a1.y;
            }
            chain_partials<true, true>(red_state, red_count, lane, warp, action, dn2, w0, w1, sum_A_all, n_acc);
        }
        __syncthreads();
        if (tid == 0) {
            bulk_wait_read0();
            if (next < a.chains) issue_load(next);
        }
        if (tid == kWriter && a.obs) {
            double* row = a.obs + chain * SVB_VOBS_COUNT;
            chain_finish<NW, true, true>(red_state, red_count, kappa / 2, row, row);
        }
    }
    if (tid == 0) bulk_wait0();
}

template <int NT, int MINB>
static int launch_villain_link_smem(const LinkArgs& a, int n_sweeps, cudaStream_t stream, const DeviceInfo& info) {
    auto kern = villain_link_smem_kernel<NT, MINB>;
    const size_t V = (size_t)NT * NT;
    const size_t smem = V * 16 + 6 * (4 * NT / 32) * sizeof(double) + 16;
    static int per_sm_cache[64];
    int per_sm = (info.device < 64) ? per_sm_cache[info.device] : 0;
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 4 * NT, smem));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "link kernel does not fit an SM at N=%d", NT);
        if (info.device < 64) per_sm_cache[info.device] = per_sm;
    }
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > a.chains) grid = a.chains;
    kern<<<(unsigned)grid, 4 * NT, smem, stream>>>(a, n_sweeps);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Chains beyond one SM's shared memory (L = 128): the same update on STRIPS of ROWS rows.  Link proposals are independent and
// phi is read only, so a strip needs nothing from its neighbours but the phi row below it (loaded as row ROWS of the strip);
// work item = (chain, strip).  The per-chain records cannot be finished inside one CTA any more: the acceptance counters
// are added atomically to a zeroed record, the state columns come from svb_villain_observables afterwards.
// ------------------------------------------------------------------------------------------
template <int NT, int ROWS, int MINB>
__global__ void __launch_bounds__(4 * NT, MINB) villain_link_strip_kernel(const __grid_constant__ LinkArgs a, int n_sweeps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, T = 4 * NT, VS = ROWS * N, PER = ROWS / 8, STRIPS = N / ROWS;
    static_assert(ROWS % 8 == 0 && N % ROWS == 0 && ROWS < N, "villain_link_strip_kernel: strips of whole row groups");
    constexpr uint32_t bytes_phi = VS * sizeof(double), bytes_row = N * sizeof(double), bytes_n = VS * sizeof(int32_t);
    const int tid = threadIdx.x, lane = tid & 31;
    double* sphi = reinterpret_cast<double*>(smem_raw);           // ROWS + 1 rows
    int32_t* sn0 = reinterpret_cast<int32_t*>(smem_raw + bytes_phi + bytes_row);
    int32_t* sn1 = sn0 + VS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sn1 + VS);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int row8 = tid / HN, k = tid - row8 * HN;
    const uint32_t K = (uint32_t)(2 * a.interval);
    const long long items = a.chains * STRIPS;
    auto issue_load = [&](long long item) {
        const long long chain = item / STRIPS;
        const int r0 = (int)(item - chain * STRIPS) * ROWS;
        mbar_expect_tx(bar, bytes_phi + bytes_row + 2 * bytes_n);
        bulk_g2s(sphi, a.phi + chain * V + (long long)r0 * N, bytes_phi, bar);
        bulk_g2s(sphi + VS, a.phi + chain * V + (long long)((r0 + ROWS) % N) * N, bytes_row, bar);      // the row below the strip
        bulk_g2s(sn0, a.n + chain * 2 * V + (long long)r0 * N, bytes_n, bar);
        bulk_g2s(sn1, a.n + chain * 2 * V + V + (long long)r0 * N, bytes_n, bar);
    };
    long long item = blockIdx.x;
    if (tid == 0 && item < items) issue_load(item);

    for (int it = 0; item < items; item += gridDim.x, ++it) {
        const long long next = item + gridDim.x;
        const long long chain = item / STRIPS;
        const int r0 = (int)(item - chain * STRIPS) * ROWS;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double m2pik = __dmul_rn(-SVB_TWO_PI, kappa);
        const double* pp = sphi + row8 * N + 2 * k;
        const double* pp_r = sphi + row8 * N + ((2 * k + 2) & (N - 1));
        int32_t* pn0 = sn0 + row8 * N + 2 * k;
        int32_t* pn1 = sn1 + row8 * N + 2 * k;
        mbar_wait(bar, (uint32_t)(it & 1));

        int n_acc = 0;
        double sum_A_all = 0.0;
        RefineCtx rc;
        rc.seed = a.seed; rc.chain = a.chain0 + (unsigned long long)chain;
#pragma unroll 1
        for (int s = 0; s < n_sweeps; ++s) {
            rc.sweep = a.sweep + (unsigned long long)s;
            float sum_A = 0.0f;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int o = 8 * N * q;
                const double2 pc = *reinterpret_cast<const double2*>(pp + o);
                const double2 pu = *reinterpret_cast<const double2*>(pp + o + N);
                const double pr = pp_r[o];
                const double dphi[4] = {__dsub_rn(pu.x, pc.x), __dsub_rn(pc.y, pc.x), __dsub_rn(pu.y, pc.y), __dsub_rn(pr, pc.y)};
                int2 m0 = *reinterpret_cast<const int2*>(pn0 + o), m1 = *reinterpret_cast<const int2*>(pn1 + o);
                int nl[4] = {m0.x, m1.x, m0.y, m1.y};
                const int site_e = (r0 + row8 + 8 * q) * N + 2 * k;                   // global site: the draws do not know about strips
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t site = (uint32_t)(site_e + h);
                    const Philox4 p = philox_site(a.seed, rc.chain, rc.sweep, site, STREAM_VILLAIN_LINK);
#pragma unroll
                    for (int mu = 0; mu < 2; ++mu) {
                        const int j = 2 * h + mu;
                        const uint64_t pz = (uint64_t)(mu ? p.y : p.x) * (uint64_t)K;
                        const int idx = (int)(pz >> 32);
                        const uint32_t f = (uint32_t)pz;
                        const int c = a.W * ((idx < a.interval) ? idx - a.interval : idx - a.interval + 1);
                        const double t1 = __dmul_rn(m2pik, (double)c);                // (link.py:83-86), numpy's order
                        const double t2 = __dsub_rn(__dsub_rn(dphi[j], __dmul_rn(SVB_TWO_PI, (double)nl[j])),
                                                    __dmul_rn(3.141592653589793116, (double)c));
                        const double dS = __dmul_rn(t1, t2);
                        double prob;
                        const double u_mid = (__hiloint2double(0x43300000, (int)f) - 4503599627370495.5) * 2.3283064365386963e-10;
                        int r = (f >= 65536u) ? metropolis_log_filter(dS, u_mid, 7.62939453125e-06f, prob) : -1;
                        if (r < 0) {
                            LinkProposal lp;
                            lp.dS = dS; lp.lu.f = f; lp.lu.c0 = site; lp.lu.word = (uint32_t)mu; lp.rc = rc;
                            prob = exp_clipped(-dS);
                            r = villain_link_exact_decision(lp) ? 1 : 0;
                        }
                        sum_A += (float)prob;
                        n_acc += r;
                        nl[j] += r ? c : 0;
                    }
                }
                *reinterpret_cast<int2*>(pn0 + o) = make_int2(nl[0], nl[2]);
                *reinterpret_cast<int2*>(pn1 + o) = make_int2(nl[1], nl[3]);
            }
            sum_A_all += (double)sum_A;
        }
        if (a.obs) {                                                   // one atomic pair per warp
            const double wa = warp_sum(sum_A_all);
            const int wn = __reduce_add_sync(0xffffffffu, n_acc);
            if (lane == 0) {
                atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, (double)wn);
                atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, wa);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(a.n + chain * 2 * V + (long long)r0 * N, sn0, bytes_n);
            bulk_s2g(a.n + chain * 2 * V + V + (long long)r0 * N, sn1, bytes_n);
            bulk_commit();
            bulk_wait_read0();
            if (next < items) issue_load(next);
        }
    }
    if (tid == 0) bulk_wait0();
}

template <int NT, int ROWS, int MINB>
static int launch_villain_link_strip(const LinkArgs& a, int n_sweeps, cudaStream_t stream, const DeviceInfo& info) {
    auto kern = villain_link_strip_kernel<NT, ROWS, MINB>;
    const size_t smem = (size_t)(ROWS + 1) * NT * sizeof(double) + (size_t)2 * ROWS * NT * sizeof(int32_t) + 16;
    static int per_sm_cache[64];
    int per_sm = (info.device < 64) ? per_sm_cache[info.device] : 0;
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 4 * NT, smem));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "link strip kernel does not fit an SM at N=%d", NT);
        if (info.device < 64) per_sm_cache[info.device] = per_sm;
    }
    const long long items = a.chains * (NT / ROWS);
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > items) grid = items;
    kern<<<(unsigned)grid, 4 * NT, smem, stream>>>(a, n_sweeps);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}
