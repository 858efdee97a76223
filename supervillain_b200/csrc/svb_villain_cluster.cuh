// svb_villain_cluster.cuh -- the fp32-filtered Villain sweep for L = 128 (config 4), one chain per thread-block CLUSTER
// (included by svb_villain.cu inside namespace svb, after svb_villain_filtered.cuh).
//
// A 128 x 128 chain (256 KiB of phi and n, plus 128 KiB of fp32 residuals) does not fit one SM, but it fits the shared
// memory of four: a cluster of CL = 4 CTAs holds one chain, CTA `rank` owning the 32 rows [32 rank, 32 rank + 32) -- its
// strip of phi and n arrives by three 1-D TMA bulk copies and never leaves shared memory during the launch's sweeps.  The
// arithmetic, the draw mapping and the thread geometry are those of villain_smem_filtered_kernel (a thread owns rows
// r, r + 8, r + 16, r + 24 of one column slot; pairs of rows share a Philox block); the only new ingredient is the strip
// boundary, 1/32 of the sites:
//   * the residual of the backward link (0, x - e0) of a site in a strip's first row lives in the previous CTA's last row,
//     and an accepted proposal there patches that CTA's residual and n_0: distributed shared memory (DSMEM) loads / stores;
//   * building the residuals of a strip's last row (and the exact path of its sites) reads phi of the next CTA's first row.
// Colour passes are separated by cluster barriers instead of block barriers.  No ghost zones, no redundant proposals, no
// ping-pong workspace: HBM traffic is the algorithmic 32 B per site-update (the tiled kernel moves 38 B and repeats 13 % of
// the proposals).
#pragma once

// Cluster primitives as PTX (this header sits inside namespace svb; <cooperative_groups.h> cannot be included here).
__device__ __forceinline__ int cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return (int)r;
}
// all threads of all CTAs of the cluster; release / acquire at cluster scope (also a block barrier)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The two halves separately: work that touches nothing another thread writes may sit between them.
// arrive: a release at cluster scope costs a gpu-scope MEMBAR (SASS: MEMBAR.ALL.GPU, ~0.3 us, and 16 warps of them queue
// up) -- the top stall of the first version of this kernel.  One warp releases on behalf of the CTA instead: the block
// barrier orders every thread's writes before warp 0's release fence, and fences are cumulative, so the other warps arrive
// relaxed (the warps that store into a NEIGHBOUR's shared memory release as well, see cluster_arrive).  203 -> 190 us per config-4 shard sweep.  (No release at all,
// -DSVB_CLUSTER_RELAXED: 179 us, passes every test, and is a data race by the PTX memory model -- not used.)
#ifdef SVB_CLUSTER_RELAXED   /* EXPERIMENT ONLY */
__device__ __forceinline__ void cluster_arrive(int = 32) { asm volatile("fence.acq_rel.cta;\n\tbarrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
#else
// `releasers`: the first so many threads (whole warps) release.  Warp 0 always does, on behalf of every write into the CTA's
// own shared memory; the warps whose threads store into a NEIGHBOUR's shared memory release their own stores themselves.
__device__ __forceinline__ void cluster_arrive(int releasers = 32) {
    __syncthreads();
    if ((int)threadIdx.x < releasers) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
#endif
// the generic address of `p` (a shared-memory address of this CTA) in the CTA of rank `rank`
template <typename T>
__device__ __forceinline__ T* cluster_map(T* p, int rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t)p), "r"(rank));
    return reinterpret_cast<T*>(out);
}

// Everything the exact (cold) path needs about one proposal, by address: neighbours may live in another CTA of the cluster.
struct ExactProposalPtr {
    const double *p_c, *p_f0, *p_b0, *p_f1, *p_b1;
    const int32_t *n_f0, *n_b0, *n_f1, *n_b1;
    double half_kappa, c, dphi;
    int g[4];
    VillainDraw d;
    RefineCtx rc;
};

__device__ __noinline__ bool villain_exact_decision_ptr(const ExactProposalPtr& p) {
    const double pc = *p.p_c;
    const double r_f0 = fma(-SVB_TWO_PI, (double)*p.n_f0, *p.p_f0 - pc);
    const double r_b0 = fma(-SVB_TWO_PI, (double)*p.n_b0, pc - *p.p_b0);
    const double r_f1 = fma(-SVB_TWO_PI, (double)*p.n_f1, *p.p_f1 - pc);
    const double r_b1 = fma(-SVB_TWO_PI, (double)*p.n_b1, pc - *p.p_b1);
    const double dr_f0 = fma(-p.c, (double)p.g[0], -p.dphi), dr_b0 = fma(-p.c, (double)p.g[1], p.dphi);
    const double dr_f1 = fma(-p.c, (double)p.g[2], -p.dphi), dr_b1 = fma(-p.c, (double)p.g[3], p.dphi);
    double acc2 = dr_f0 * fma(2.0, r_f0, dr_f0);
    acc2 = fma(dr_b0, fma(2.0, r_b0, dr_b0), acc2);
    acc2 = fma(dr_f1, fma(2.0, r_f1, dr_f1), acc2);
    acc2 = fma(dr_b1, fma(2.0, r_b1, dr_b1), acc2);
    const double dS = p.half_kappa * acc2;
    return villain_decide_lazy(exp_clipped(-dS), p.d, p.rc);
}

// OVERLAP: the overlapped-launch protocol of svb_villain_sweep_overlapped (see villain_smem_filtered_kernel): thread 0 of
// every CTA acquires the chain's epoch before loading its strip; rank 0 publishes the epochs of all the cluster's chains at
// the end, behind a cluster barrier that follows every CTA's completed bulk stores.
// TPB threads per CTA (8 or 16 row groups of N/2 column slots); STAGES = 2: phi and n double-buffered, so that the next
// chain's strip arrives and the previous one's leaves while this one is swept (one CTA of 1024 threads per SM).
// ... and in the reference's operation order (villain_exact_decision_strict): p.c = 2 pi, p.g = dn in units of 1.
__device__ __noinline__ bool villain_exact_decision_strict_ptr(const ExactProposalPtr& p) {
    const double pc = *p.p_c;
    const double r_f0 = __dsub_rn(__dsub_rn(*p.p_f0, pc), __dmul_rn(SVB_TWO_PI, (double)*p.n_f0));
    const double r_b0 = __dsub_rn(__dsub_rn(pc, *p.p_b0), __dmul_rn(SVB_TWO_PI, (double)*p.n_b0));
    const double r_f1 = __dsub_rn(__dsub_rn(*p.p_f1, pc), __dmul_rn(SVB_TWO_PI, (double)*p.n_f1));
    const double r_b1 = __dsub_rn(__dsub_rn(pc, *p.p_b1), __dmul_rn(SVB_TWO_PI, (double)*p.n_b1));
    const double dr_f0 = __dsub_rn(-p.dphi, __dmul_rn(p.c, (double)p.g[0])), dr_b0 = __dsub_rn(p.dphi, __dmul_rn(p.c, (double)p.g[1]));
    const double dr_f1 = __dsub_rn(-p.dphi, __dmul_rn(p.c, (double)p.g[2])), dr_b1 = __dsub_rn(p.dphi, __dmul_rn(p.c, (double)p.g[3]));
    const double hk = p.half_kappa;
    double dS = __dmul_rn(__dmul_rn(hk, dr_f0), __dadd_rn(__dmul_rn(2.0, r_f0), dr_f0));
    dS = __dadd_rn(dS, __dmul_rn(__dmul_rn(hk, dr_b0), __dadd_rn(__dmul_rn(2.0, r_b0), dr_b0)));
    dS = __dadd_rn(dS, __dmul_rn(__dmul_rn(hk, dr_f1), __dadd_rn(__dmul_rn(2.0, r_f1), dr_f1)));
    dS = __dadd_rn(dS, __dmul_rn(__dmul_rn(hk, dr_b1), __dadd_rn(__dmul_rn(2.0, r_b1), dr_b1)));
    return villain_decide_lazy(exp_clipped(-dS), p.d, p.rc);
}

// The cold half of the epoch wait (the producer launch has not stored the chain yet), out of line: it must not cost the
// sweep loop registers.
static __device__ __noinline__ void cluster_epoch_spin(const uint32_t* epoch, uint32_t want) {
    unsigned ns = 32, naps = 0;
    while (true) {
        uint32_t e;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(epoch) : "memory");
        if (e == want) return;
        __nanosleep(ns);
        if (ns < 1024) ns *= 2;
        if (++naps > (1u << 21)) __trap();          // > 2 s: a producer that never comes is a caller error
    }
}

constexpr int cluster_min_blocks(int NT, int CL, int TPB, int STAGES) {
    return (TPB >= 1024 || (STAGES * 16 + 8) * (NT / CL) * NT > 112 * 1024) ? 1 : 2;          // two CTAs per SM where they fit
}
// MODE: SVB_FILT_FAST (the NeighborhoodUpdate sweep), SVB_FILT_STRICT / SVB_FILT_SITE / SVB_FILT_EXACT (the decoupled updates with
// a STRICT cold path), exactly as in villain_smem_filtered_kernel.
// SPARSE (one sweep per launch, no record of the state after it), as in villain_smem_filtered_kernel: the staged phi and n are
// dead once the residuals of EVERY strip are built (cluster barrier S2), accepted proposals go to global memory as
// fire-and-forget reductions, the exact path reads global memory, nothing is stored back, and the next chain's strip is
// loaded while this one is swept -- the store -> load -> barrier phase that bounded this kernel is gone.
template <int NT, int CL, int TPB, int STAGES, bool OVERLAP, int MODE, bool SPARSE = false>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(TPB, cluster_min_blocks(NT, CL, TPB, STAGES))
    villain_cluster_kernel(const __grid_constant__ VillainArgs a, const __grid_constant__ FilterConsts fc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, ROWS = N / CL, VL = ROWS * N, VHL = ROWS * HN, T = TPB, NW = T / 32;
    constexpr int RG = T / HN;                                    // row groups
    constexpr int PER = ROWS / RG;                                // sites per thread per colour (rows r0 + 8 q)
    constexpr int Q = 8 * HN;                                     // compact-index distance between a thread's rows
    static_assert(RG % 8 == 0 && RG * HN == T && PER * RG == ROWS && PER >= 2 && PER % 2 == 0 && (STAGES == 1 || STAGES == 2),
                  "villain_cluster_kernel: unsupported geometry");
    constexpr uint32_t bytes_phi = VL * sizeof(double), bytes_n = VL * sizeof(int32_t), stage_bytes = bytes_phi + 2 * bytes_n;
    const int rank = cluster_cta_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    double* sphi = reinterpret_cast<double*>(smem_raw);           // the stage being swept (set per chain)
    int32_t* sn0 = reinterpret_cast<int32_t*>(smem_raw + bytes_phi);
    int32_t* sn1 = sn0 + VL;
    float* rc0 = reinterpret_cast<float*>(smem_raw + STAGES * stage_bytes);   // [colour][VHL]: residual of link (0, x)
    float* rc1 = rc0 + 2 * VHL;                                   // [colour][VHL]: residual of link (1, x)
    double* red_state = reinterpret_cast<double*>(rc1 + 2 * VHL); // [NW][4] per-warp partial sums
    double* red_count = red_state + 4 * NW;                       // [NW][2]
    double* cta_sums = red_count + 2 * NW;                        // [6]: this CTA's share of the chain's record
    uint64_t* bar = reinterpret_cast<uint64_t*>(cta_sums + 6);
    // interval_n == 1 (K = 3): the 81 proposals of the four dn as residual changes -(2 pi W) digit, looked up by the code the
    // multiply by 81 leaves (svb_villain_filtered.cuh); 0, c and 2 c are exact in fp32, so base + lut is the fma it replaces
    float4* dn_lut = reinterpret_cast<float4*>(bar + 2 * STAGES);
    constexpr int kWriter = 32;

    // the neighbouring strips (DSMEM)
    const int next_rank = (rank + 1) % CL, prev_rank = (rank + CL - 1) % CL;
    // (mapped where they are used -- one `mapa` each -- rather than held in registers)
#define NEXT_PHI cluster_map(sphi, next_rank)                              /* its row 0 is my row ROWS */
#define NEXT_N1 cluster_map(sn1, next_rank)
#define PREV_PHI cluster_map(sphi, prev_rank)                              /* its row ROWS - 1 is my row -1 */
#define PREV_N0 cluster_map(sn0, prev_rank)
#define PREV_RC0 cluster_map(rc0, prev_rank)

    if (tid == 0) {
        for (int b = 0; b < STAGES; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
    }
    if (OVERLAP) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (a.grid_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    for (int i = tid; i < 81; i += T)
        dn_lut[i] = make_float4(-fc.c * (float)(i / 27), -fc.c * (float)((i / 9) % 3), -fc.c * (float)((i / 3) % 3), -fc.c * (float)(i % 3));
    cluster_sync_all();
    const bool obs_of_input = a.obs_in != nullptr;
    const bool want_obs = a.obs != nullptr && !obs_of_input;
    static_assert(!(OVERLAP && MODE != SVB_FILT_FAST), "overlapped launches serve the NeighborhoodUpdate sweep");
    static_assert(!SPARSE || (MODE == SVB_FILT_FAST && STAGES == 1), "SPARSE serves the NeighborhoodUpdate sweep, one stage");
    const int interval_n = MODE == SVB_FILT_SITE ? 0 : a.interval_n;
    const uint32_t K = (MODE == SVB_FILT_EXACT) ? (uint32_t)(2 * interval_n) : (uint32_t)(2 * interval_n + 1);
    const int W = MODE == SVB_FILT_EXACT ? 1 : a.W, mWI = -W * interval_n;
    const bool lut3 = (MODE == SVB_FILT_FAST || MODE == SVB_FILT_STRICT) && K == 3u;     // block-uniform
    const float cIn = fc.c * (float)interval_n;
    const float2 cIn2 = make_float2(cIn, cIn), negc2 = make_float2(-fc.c, -fc.c), two2 = make_float2(2.0f, 2.0f);
    const double two_I_scaled = (2.0 * a.interval_phi) * 2.3283064365386963e-10;

    // a thread's rows are r0 + 8 q, q < PER; rows r0 + 16 p and r0 + 16 p + 8 share a Philox block
    const int rg = tid / HN, k = tid - rg * HN;
    const int r0 = (rg & 7) + 8 * PER * (rg >> 3);
    const int jb = r0 * HN + k;                                   // compact index of the thread's first site
    const int cc = r0 & 1;
    const bool first_row_thread = (r0 == 0), last_row_thread = (r0 + 8 * (PER - 1) == ROWS - 1);

    const long long n_clusters = gridDim.x / CL, cluster_id = blockIdx.x / CL;
    auto peek_epoch = [&](long long chain) -> uint32_t {
        uint32_t e = a.wait_epoch;
        if (OVERLAP && !a.grid_wait)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(a.epochs + chain) : "memory");
        return e;
    };
    auto issue_load = [&](long long chain, int b, uint32_t seen) {
        if (OVERLAP && !a.grid_wait) {
            if (seen != a.wait_epoch) cluster_epoch_spin(a.epochs + chain, a.wait_epoch);
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        mbar_expect_tx(&bar[b], stage_bytes);
        bulk_g2s(stage, reinterpret_cast<const double*>(a.phi) + chain * V + (long long)rank * VL, bytes_phi, &bar[b]);
        bulk_g2s(stage + bytes_phi, a.n + chain * 2 * V + (long long)rank * VL, bytes_n, &bar[b]);
        bulk_g2s(stage + bytes_phi + bytes_n, a.n + chain * 2 * V + V + (long long)rank * VL, bytes_n, &bar[b]);
    };
    // the next chain's strip on its way into L2 while this one is being swept: the load phase then runs at L2 speed
    auto prefetch_l2 = [&](long long chain) {
#ifndef SVB_CLUSTER_NO_PREFETCH
        const double* gp = reinterpret_cast<const double*>(a.phi) + chain * V + (long long)rank * VL;
        const int32_t* g0 = a.n + chain * 2 * V + (long long)rank * VL;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gp), "r"(bytes_phi) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g0), "r"(bytes_n) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g0 + V), "r"(bytes_n) : "memory");
#endif
    };
    // rank 0 gathers the four CTAs' shares of a finished chain's record (call behind a cluster barrier that follows them)
    // Records: ONE thread (lane 0 of warp 1; thread 0 is busy with the bulk copies) adds the slots in a fixed order.  A
    // shuffle-tree version over warp 1 measured 10 % SLOWER per launch -- with or without records: it costs registers in
    // the 64-register budget of the sweep loop (88 vs 44 bytes spilled).
    auto gather_record = [&](long long chain, double kappa) {
        if (rank == 0 && tid == kWriter && (a.obs || a.obs_in)) {
            double t[6] = {0, 0, 0, 0, 0, 0};
            for (int r = 0; r < CL; ++r) {
                const double* s = cluster_map(cta_sums, r);
#pragma unroll
                for (int i = 0; i < 6; ++i) t[i] += s[i];
            }
            double* state_row = (obs_of_input ? a.obs_in : a.obs) + chain * SVB_VOBS_COUNT;
            state_row[SVB_VOBS_ACTION] = (kappa / 2) * t[0];
            state_row[SVB_VOBS_SUM_DN2] = t[1];
            state_row[SVB_VOBS_WRAP0] = t[2];
            state_row[SVB_VOBS_WRAP1] = t[3];
            if (a.obs) {
                double* row = a.obs + chain * SVB_VOBS_COUNT;
                row[SVB_VOBS_ACCEPTANCE] = t[4];
                row[SVB_VOBS_ACCEPTED] = t[5];
            }
        }
    };
    auto cta_share = [&](bool state, bool counters) {
        if (tid == kWriter) {
            if (state) {
                double t[4] = {0, 0, 0, 0};
                for (int w = 0; w < NW; ++w)
#pragma unroll
                    for (int i = 0; i < 4; ++i) t[i] += red_state[4 * w + i];
#pragma unroll
                for (int i = 0; i < 4; ++i) cta_sums[i] = t[i];
            }
            if (counters) {
                double t0 = 0, t1 = 0;
                for (int w = 0; w < NW; ++w) { t0 += red_count[2 * w]; t1 += red_count[2 * w + 1]; }
                cta_sums[4] = t0; cta_sums[5] = t1;
            }
        }
    };
    // sums of the strip's forward links and plaquettes from the current phi and n (fp64), as in the smem kernel
    // part 0: the rows that need nothing from another strip (q < PER - 1); part 1: q = PER - 1; part 2: all
    auto state_sums = [&](double& action, long long& dn2, int& w0, int& w1, bool store_r, bool sums, int part) {
        float* w0e = rc0 + cc * VHL + jb;
        float* w1e = rc1 + cc * VHL + jb;
        float* w0o = rc0 + (cc ^ 1) * VHL + jb;
        float* w1o = rc1 + (cc ^ 1) * VHL + jb;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            if ((part == 0 && q == PER - 1) || (part == 1 && q != PER - 1)) continue;
            const int lx0 = r0 + 8 * q;
            const double* p_c = sphi + lx0 * N + 2 * k;
            const bool below_remote = (q == PER - 1) && last_row_thread;
            const double* p_u = below_remote ? NEXT_PHI + 2 * k : p_c + N;
            const PairResiduals pr = villain_pair_residuals(p_c, p_u, sphi + lx0 * N + ((2 * k + 2) & (N - 1)), sn0 + lx0 * N + 2 * k,
                                                            sn1 + lx0 * N + 2 * k);
            if (store_r) {
                w0e[Q * q] = (float)pr.r0e;
                w1e[Q * q] = (float)pr.r1e;
                w0o[Q * q] = (float)pr.r0o;
                w1o[Q * q] = (float)pr.r1o;
            }
            if (sums) {
                action = fma(pr.r0e, pr.r0e, action);
                action = fma(pr.r0o, pr.r0o, action);
                action = fma(pr.r1e, pr.r1e, action);
                action = fma(pr.r1o, pr.r1o, action);
                const int hr = sn0[lx0 * N + ((2 * k + 2) & (N - 1))];                            // n0[x + 2 e1]
                const int32_t* up_p = below_remote ? NEXT_N1 + 2 * k : sn1 + (lx0 + 1) * N + 2 * k;
                const int2 up = *reinterpret_cast<const int2*>(up_p);                               // n1[x + e0]
                const int d0 = (up.x - pr.a1.x) - (pr.a0.y - pr.a0.x), d1 = (up.y - pr.a1.y) - (hr - pr.a0.y);
                dn2 += (long long)d0 * d0 + (long long)d1 * d1;
                w0 += pr.a0.x + pr.a0.y;
                w1 += pr.a1.x + pr.a1.y;
            }
        }
    };

    // (plain strided chain assignment: the ChainMap of the other persistent kernels costs this one registers it does not
    // have -- 213 instead of 191 us per launch)
    long long chain = cluster_id;
    if (tid == 0 && chain < a.chains) issue_load(chain, 0, peek_epoch(chain));
    long long pending_chain = -1;
    double pending_kappa = 0.0;

    for (int it = 0; chain < a.chains; chain += n_clusters, ++it) {
        const long long next = chain + n_clusters;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double half_kappa = kappa / 2;
        const float hk2 = (float)(half_kappa * 1.4426950408889634);
        const float hkA = 1.0001f * hk2 * fc.bA, hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
        const float2 hk22 = make_float2(hk2, hk2), hkA2 = make_float2(hkA, hkA), hkB2 = make_float2(hkB, hkB);

        const int b = (STAGES == 2) ? (it & 1) : 0;
        sphi = reinterpret_cast<double*>(smem_raw + (size_t)b * stage_bytes);
        sn0 = reinterpret_cast<int32_t*>(smem_raw + (size_t)b * stage_bytes + bytes_phi);
        sn1 = sn0 + VL;
        // SPARSE: the chain in global memory (accepted proposals and the exact path go there)
        double* gphi = reinterpret_cast<double*>(a.phi) + chain * V;
        int32_t* gn0 = a.n + chain * 2 * V;
        uint32_t seen_sparse = 0;
        if (SPARSE && tid == 0 && next < a.chains) seen_sparse = peek_epoch(next);
        mbar_wait(&bar[b], (uint32_t)((STAGES == 2 ? (it >> 1) : it) & 1));
        if (STAGES == 1 && tid == 0 && next < a.chains) prefetch_l2(next);
        cluster_arrive();                                          // S1: every strip of the chain has landed (waited for below)

        int n_acc = 0;
        double sum_A_all = 0.0;
        for (int s = 0; s < a.n_sweeps; ++s) {
            float sum_A = 0.0f;
            // ---- r = d(phi) - 2 pi n   (neighborhood.py:91) in fp64, stored rounded to fp32 ----
            {
                const bool sums = obs_of_input && s == 0;
                double action = 0.0;
                int w0 = 0, w1 = 0;
                long long dn2 = 0;
                if (s == 0) {
                    // the strip's own rows while the other strips are still landing; its last row needs the next strip
                    state_sums(action, dn2, w0, w1, true, sums, 0);
                    cluster_wait();                                // S1
                    // the other stage is free in every strip now (its last readers have arrived here): fetch the next chain
                    if (STAGES == 2 && tid == 0 && next < a.chains) {
                        bulk_wait_read0();                         // (the previous chain's store has long read it)
                        issue_load(next, b ^ 1, peek_epoch(next));
                    }
                    if (pending_chain >= 0) gather_record(pending_chain, pending_kappa);
                    state_sums(action, dn2, w0, w1, true, sums, 1);
                } else {
                    state_sums(action, dn2, w0, w1, true, false, 2);
                }
                if (sums) chain_partials<true, false>(red_state, red_count, lane, warp, action, dn2, w0, w1, 0.0, 0);
            }
            cluster_arrive();                                      // S2: the residuals (also the neighbours') are built

            const unsigned long long gc = a.chain0 + (unsigned long long)chain, gs = a.sweep0 + (unsigned long long)s;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int par = (r0 + c) & 1;
                const int x1 = 2 * k + par;
                const int ob1 = par ? 0 : ((k == 0) ? (1 - HN) : 1);
                const int wrap1 = (x1 == 0) ? N : 0;
                float* R0own = rc0 + c * VHL + jb;
                float* R1own = rc1 + c * VHL + jb;
                float* R0b = rc0 + (c ^ 1) * VHL + jb - HN;        // backward link (0, x - e0): row above, same compact column
                float* R0b_q0 = first_row_thread ? PREV_RC0 + (c ^ 1) * VHL + (ROWS - 1) * HN + k : R0b;   // ... in the previous strip
                float* R1b = rc1 + (c ^ 1) * VHL + jb - ob1;
                double* Pc = sphi + 2 * jb + par;
                int32_t* N0c = sn0 + 2 * jb + par;
                int32_t* N1c = sn1 + 2 * jb + par;
                int32_t* N0b = N0c - N;
                int32_t* N0b_q0 = first_row_thread ? PREV_N0 + (ROWS - 1) * N + x1 : N0b;
                int32_t* N1b = N1c - 1 + wrap1;
                // the draws depend on nothing in memory: they fill the wait for the slowest strip
#ifndef SVB_CLUSTER_PRE
#define SVB_CLUSTER_PRE 1                  /* how many of the pass's PER / 2 Philox blocks are drawn before the wait: both
                                              cost registers the sweep needs (200 us), none leaves the wait empty (203 us) */
#endif
                Philox4 bits_p[PER / 2];
#pragma unroll
                for (int p = 0; p < PER / 2; ++p)
                    if (p < SVB_CLUSTER_PRE) bits_p[p] = philox_site_keys(a, gc, gs, (uint32_t)((rank * ROWS + r0 + 16 * p) * N + x1));
                cluster_wait();                                    // S2 / S3: the previous pass is complete in every strip
                // SPARSE: every strip's residuals are built -- nobody reads any strip's staged phi and n again
                if (SPARSE && c == 0 && tid == 0 && next < a.chains) issue_load(next, 0, seen_sparse);
                if (c == 0 && obs_of_input && s == 0) cta_share(true, false);
#pragma unroll
                for (int p = 0; p < PER / 2; ++p) {
                    const int gx0 = rank * ROWS + r0 + 16 * p;                              // global row of the pair's first site
                    const uint32_t c0 = (uint32_t)(gx0 * N + x1);                             // villain_pair_counter (bit 3 clear)
                    const Philox4 bits = (p < SVB_CLUSTER_PRE) ? bits_p[p] : philox_site_keys(a, gc, gs, c0);
                    const int qA = 2 * p, qB = 2 * p + 1;
                    float* r0bA = (p == 0) ? R0b_q0 : R0b;
                    int32_t* n0bA = (p == 0) ? N0b_q0 : N0b;
                    uint32_t fA = bits.y, fB = bits.w;
                    uint32_t codeA = 0, codeB = 0;
                    int digA[4] = {0, 0, 0, 0}, digB[4] = {0, 0, 0, 0};
                    if (MODE == SVB_FILT_EXACT) {                 // z from word B; as "digits": I - z forward, I + z backward
                        const uint64_t pa = (uint64_t)fA * K, pb = (uint64_t)fB * K;
                        fA = (uint32_t)pa; fB = (uint32_t)pb;
                        const int ia = (int)(pa >> 32), ib = (int)(pb >> 32);
                        const int za = (ia < interval_n) ? ia - interval_n : ia - interval_n + 1;
                        const int zb = (ib < interval_n) ? ib - interval_n : ib - interval_n + 1;
                        digA[0] = digA[2] = interval_n - za; digA[1] = digA[3] = interval_n + za;
                        digB[0] = digB[2] = interval_n - zb; digB[1] = digB[3] = interval_n + zb;
                    } else if (MODE == SVB_FILT_SITE) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) digA[i] = digB[i] = 0;
                    } else if (lut3) {
                        // one multiply by 81 = four successive multiply-highs by 3: hi = the code 27 d0 + 9 d1 + 3 d2 + d3, lo = the
                        // remainder; the digits are decoded only on the rare paths
                        const uint64_t pa = (uint64_t)fA * 81u, pb = (uint64_t)fB * 81u;
                        fA = (uint32_t)pa; fB = (uint32_t)pb;
                        codeA = (uint32_t)(pa >> 32); codeB = (uint32_t)(pb >> 32);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint64_t pa = (uint64_t)fA * K, pb = (uint64_t)fB * K;
                            fA = (uint32_t)pa; digA[i] = (int)(pa >> 32);
                            fB = (uint32_t)pb; digB[i] = (int)(pb >> 32);
                        }
                    }
                    float2 U = make_float2(__uint_as_float(0x3F800000u | (bits.x >> 9)), __uint_as_float(0x3F800000u | (bits.z >> 9)));
                    U = __fadd2_rn(U, make_float2(-0.99999994f, -0.99999994f));
                    const float2 dphi = __ffma2_rn(make_float2(fc.two_I, fc.two_I), U, make_float2(-fc.I, -fc.I));
                    const float2 base_f = __ffma2_rn(dphi, make_float2(-1.0f, -1.0f), cIn2), base_b = __fadd2_rn(cIn2, dphi);
                    const float2 r_f0 = make_float2(R0own[Q * qA], R0own[Q * qB]), r_f1 = make_float2(R1own[Q * qA], R1own[Q * qB]);
                    const float2 r_b0 = make_float2(r0bA[Q * qA], R0b[Q * qB]);
                    const float2 r_b1 = make_float2(R1b[Q * qA], R1b[Q * qB]);
                    float2 dr_f0, dr_b0, dr_f1, dr_b1;
                    if (lut3) {
                        const float4 tA = dn_lut[codeA], tB = dn_lut[codeB];
                        dr_f0 = __fadd2_rn(base_f, make_float2(tA.x, tB.x));
                        dr_b0 = __fadd2_rn(base_b, make_float2(tA.y, tB.y));
                        dr_f1 = __fadd2_rn(base_f, make_float2(tA.z, tB.z));
                        dr_b1 = __fadd2_rn(base_b, make_float2(tA.w, tB.w));
                    } else {
                        dr_f0 = MODE == SVB_FILT_SITE ? base_f : __ffma2_rn(negc2, make_float2((float)digA[0], (float)digB[0]), base_f);
                        dr_b0 = MODE == SVB_FILT_SITE ? base_b : __ffma2_rn(negc2, make_float2((float)digA[1], (float)digB[1]), base_b);
                        dr_f1 = MODE == SVB_FILT_SITE ? base_f : __ffma2_rn(negc2, make_float2((float)digA[2], (float)digB[2]), base_f);
                        dr_b1 = MODE == SVB_FILT_SITE ? base_b : __ffma2_rn(negc2, make_float2((float)digA[3], (float)digB[3]), base_b);
                    }
                    float2 acc2 = __fmul2_rn(dr_f0, __ffma2_rn(two2, r_f0, dr_f0));
                    acc2 = __ffma2_rn(dr_b0, __ffma2_rn(two2, r_b0, dr_b0), acc2);
                    acc2 = __ffma2_rn(dr_f1, __ffma2_rn(two2, r_f1, dr_f1), acc2);
                    acc2 = __ffma2_rn(dr_b1, __ffma2_rn(two2, r_b1, dr_b1), acc2);
                    const float2 dS2 = __fmul2_rn(hk22, acc2);
                    const float2 L2 = __ffma2_rn(make_float2(fast_lg2((float)fA), fast_lg2((float)fB)), make_float2(-1.0f, -1.0f),
                                                 make_float2(32.0f, 32.0f));
                    const float2 Rmax = make_float2(fmaxf(fmaxf(fabsf(r_f0.x), fabsf(r_b0.x)), fmaxf(fabsf(r_f1.x), fabsf(r_b1.x))),
                                                    fmaxf(fmaxf(fabsf(r_f0.y), fabsf(r_b0.y)), fmaxf(fabsf(r_f1.y), fabsf(r_b1.y))));
                    const float2 band = __ffma2_rn(hkA2, Rmax, __ffma2_rn(make_float2(4e-6f, 4e-6f), L2, hkB2));
                    const float2 diff = __ffma2_rn(L2, make_float2(-1.0f, -1.0f), dS2);
                    sum_A += fminf(fast_ex2(-dS2.x), 1.0f) + fminf(fast_ex2(-dS2.y), 1.0f);
                    const float2 n_f0 = __fadd2_rn(r_f0, dr_f0), n_b0 = __fadd2_rn(r_b0, dr_b0);
                    const float2 n_f1 = __fadd2_rn(r_f1, dr_f1), n_b1 = __fadd2_rn(r_b1, dr_b1);
                    // certainly rejected (the overwhelming majority): nothing more to do; everything else behind ONE branch per pair
                    const bool candA = !(diff.x > band.x) || fA < 65536u, candB = !(diff.y > band.y) || fB < 65536u;
                    if (candA || candB) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (!(h ? candB : candA)) continue;
                        const int q = 2 * p + h;
                        float* r0b_site = (h ? R0b : r0bA) + Q * q;                   // q == 0 of row 0: in the previous strip
                        int32_t* n0b_site = (h ? N0b : n0bA) + 2 * Q * q;
                        const uint32_t wA = h ? bits.z : bits.x;
                        const uint32_t f = h ? fB : fA;
                        int dig[4];
                        if (lut3) {
                            uint32_t code = h ? codeB : codeA;
                            dig[0] = (int)((code * 2428u) >> 16); code -= 27u * (uint32_t)dig[0];
                            dig[1] = (int)((code * 7282u) >> 16); code -= 9u * (uint32_t)dig[1];
                            dig[2] = (int)((code * 21846u) >> 16); dig[3] = (int)(code - 3u * (uint32_t)dig[2]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) dig[i] = h ? digB[i] : digA[i];
                        }
                        const float dif = h ? diff.y : diff.x, bnd = h ? band.y : band.x;
                        bool ok = dif < 0.0f;
                        if (SPARSE && (!(fabsf(dif) > bnd) || f < 65536u)) {
                            const int gx = rank * ROWS + r0 + 8 * q;
                            ExactProposal ep;
                            ep.phi = gphi; ep.n0 = gn0; ep.n1 = gn0 + V;
                            ep.i_c = gx * N + x1;
                            ep.i_b0 = ((gx - 1) & (N - 1)) * N + x1;
                            ep.i_b1 = gx * N + ((x1 - 1) & (N - 1));
                            ep.i_f0 = ((gx + 1) & (N - 1)) * N + x1;
                            ep.i_f1 = gx * N + ((x1 + 1) & (N - 1));
                            ep.half_kappa = half_kappa;
                            ep.dphi = villain_dphi_from_word(wA, a.interval_phi);
                            ep.d.f = f; ep.d.c0 = c0; ep.d.half = (uint32_t)h;
                            ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs; ep.rc.stream = a.refine_stream; ep.rc.wide = 0;
                            ep.c = SVB_TWO_PI * (double)W;
#pragma unroll
                            for (int i = 0; i < 4; ++i) ep.g[i] = dig[i] - interval_n;
                            ok = villain_exact_decision(ep);
                        }
                        if (!SPARSE && (!(fabsf(dif) > bnd) || f < 65536u)) {
                            const int lx0 = r0 + 8 * q;
                            ExactProposalPtr ep;
                            ep.p_c = sphi + lx0 * N + x1;
                            ep.p_f0 = (lx0 + 1 < ROWS) ? ep.p_c + N : NEXT_PHI + x1;
                            ep.p_b0 = (lx0 >= 1) ? ep.p_c - N : PREV_PHI + (ROWS - 1) * N + x1;
                            ep.p_f1 = sphi + lx0 * N + ((x1 + 1) & (N - 1));
                            ep.p_b1 = sphi + lx0 * N + ((x1 - 1) & (N - 1));
                            ep.n_f0 = sn0 + lx0 * N + x1;
                            ep.n_b0 = n0b_site;
                            ep.n_f1 = sn1 + lx0 * N + x1;
                            ep.n_b1 = sn1 + lx0 * N + ((x1 - 1) & (N - 1));
                            ep.half_kappa = half_kappa;
                            ep.dphi = (MODE == SVB_FILT_EXACT) ? 0.0 : villain_dphi_from_word(wA, a.interval_phi);
                            ep.d.f = f; ep.d.c0 = c0; ep.d.half = (uint32_t)h;
                            ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs; ep.rc.stream = a.refine_stream; ep.rc.wide = 0;
                            if (MODE == SVB_FILT_FAST) {
                                ep.c = SVB_TWO_PI * (double)W;
#pragma unroll
                                for (int i = 0; i < 4; ++i) ep.g[i] = dig[i] - interval_n;
                                ok = villain_exact_decision_ptr(ep);
                            } else {
                                ep.c = SVB_TWO_PI;
#pragma unroll
                                for (int i = 0; i < 4; ++i) ep.g[i] = W * (dig[i] - interval_n);
                                ok = villain_exact_decision_strict_ptr(ep);
                            }
                        }
                        n_acc += ok ? 1 : 0;
                        if (ok) {
                            const double Ah = __hiloint2double(0x43300000, (int)wA) - 4503599627370495.5;
                            if (SPARSE) {
                                // straight to global memory: one fp64 reduction (rounds once, to nearest, like phi + dphi) and four
                                // integer ones; nothing waits for them
                                const int gx = rank * ROWS + r0 + 8 * q, ic = gx * N + x1;
                                atomicAdd(gphi + ic, __dadd_rn(-a.interval_phi, __dmul_rn(two_I_scaled, Ah)));
                                atomicAdd(gn0 + ic, W * dig[0] + mWI);
                                atomicAdd(gn0 + ((gx - 1) & (N - 1)) * N + x1, W * dig[1] + mWI);
                                atomicAdd(gn0 + V + ic, W * dig[2] + mWI);
                                atomicAdd(gn0 + V + gx * N + ((x1 - 1) & (N - 1)), W * dig[3] + mWI);
                            } else {
                            Pc[2 * Q * q] = __dadd_rn(Pc[2 * Q * q], __dadd_rn(-a.interval_phi, __dmul_rn(two_I_scaled, Ah)));
                            if (MODE != SVB_FILT_SITE) {
                                atomicAdd(N0c + 2 * Q * q, W * dig[0] + mWI);  // only this thread touches these links in this pass
                                atomicAdd(n0b_site, W * dig[1] + mWI);
                                atomicAdd(N1c + 2 * Q * q, W * dig[2] + mWI);
                                atomicAdd(N1b + 2 * Q * q, W * dig[3] + mWI);
                            }
                            }
                            R0own[Q * q] = h ? n_f0.y : n_f0.x;
                            *r0b_site = h ? n_b0.y : n_b0.x;
                            R1own[Q * q] = h ? n_f1.y : n_f1.x;
                            R1b[Q * q] = h ? n_b1.y : n_b1.x;
                        }
                    }
                    }
                }
                if (s == a.n_sweeps - 1 && c == 1) {
                    if (a.obs) chain_partials<false, true>(red_state, red_count, lane, warp, 0.0, 0, 0, 0, sum_A_all + (double)sum_A, n_acc);
                    if (!SPARSE) asm volatile("fence.proxy.async;" ::: "memory");   // phi / n writes (also into the previous strip) -> bulk store
                }
                cluster_arrive(HN > 32 ? HN : 32);                 // S3 / S4: this strip's share of the colour pass is complete
                                                                   // (threads < HN own row 0: they wrote into the previous strip)
            }
            cluster_wait();                                        // S4
            sum_A_all += (double)sum_A;
        }
        if (a.obs) cta_share(false, true);

        if (want_obs) {
            // observables of the final state: one more fp64 pass over the strip (reads the next strip's first row)
            double action = 0.0;
            int w0 = 0, w1 = 0;
            long long dn2 = 0;
            state_sums(action, dn2, w0, w1, false, true, 2);
            chain_partials<true, false>(red_state, red_count, lane, warp, action, dn2, w0, w1, 0.0, 0);
            if (STAGES == 1) { cluster_arrive(); cluster_wait(); }     // S5: nobody reads this strip any more, it may be refilled
            else __syncthreads();                                      // (two stages: the refill waits behind the next chain's S1)
            cta_share(true, false);
        }

        // ---- store the strip; fetch the next chain's (SPARSE: nothing to store, the next strip is already on its way) ----
        if (!SPARSE && tid == 0) {
            bulk_s2g(reinterpret_cast<double*>(a.phi) + chain * V + (long long)rank * VL, sphi, bytes_phi);
            bulk_s2g(a.n + chain * 2 * V + (long long)rank * VL, sn0, bytes_n);
            bulk_s2g(a.n + chain * 2 * V + V + (long long)rank * VL, sn1, bytes_n);
            bulk_commit();
            const uint32_t seen_next = (STAGES == 1 && next < a.chains) ? peek_epoch(next) : 0;
            if (STAGES == 1) {
                bulk_wait_read0();
                if (next < a.chains) issue_load(next, 0, seen_next);
            }
        }
        pending_chain = chain;
        pending_kappa = kappa;
    }
    if (!SPARSE && tid == 0) {
        bulk_wait0();                                              // this CTA's bulk stores are complete ...
        if (OVERLAP) asm volatile("fence.proxy.async;" ::: "memory");
    }
    cluster_sync_all();                                                // every CTA's share of the last record is in place
    if (pending_chain >= 0) gather_record(pending_chain, pending_kappa);
    cluster_sync_all();                                                // no CTA leaves while its shared memory may still be read
    if (OVERLAP && rank == 0 && warp == 0) {
        // ... and so are every other CTA's (cluster barrier) and the records (block barrier): release all the cluster's chains
        const int count = (int)((a.chains - cluster_id + n_clusters - 1) / n_clusters);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        for (int i = lane; i < count; i += 32)
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.epochs + cluster_id + (long long)i * n_clusters), "r"(a.signal_epoch)
                         : "memory");
    }
#undef NEXT_PHI
#undef NEXT_N1
#undef PREV_PHI
#undef PREV_N0
#undef PREV_RC0
}

template <int NT, int CL, int TPB, int STAGES>
static int launch_villain_cluster(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    const bool overlap = a.epochs != nullptr;
    const int mode = a.exact_mode ? SVB_FILT_EXACT : (a.filtered_strict ? (a.interval_n == 0 ? SVB_FILT_SITE : SVB_FILT_STRICT) : SVB_FILT_FAST);
    if (mode != SVB_FILT_FAST && overlap) return fail(SVB_E_UNSUPPORTED, "overlapped launches serve the NeighborhoodUpdate sweep only");
    const char* env_sparse = getenv("SVB_VILLAIN_SPARSE");
    const bool sparse = STAGES == 1 && mode == SVB_FILT_FAST && a.n_sweeps == 1 && (a.obs == nullptr || a.obs_in != nullptr) &&
                        !(env_sparse && env_sparse[0] == '0');
    auto kern = mode == SVB_FILT_EXACT    ? villain_cluster_kernel<NT, CL, TPB, STAGES, false, SVB_FILT_EXACT>
                : mode == SVB_FILT_SITE   ? villain_cluster_kernel<NT, CL, TPB, STAGES, false, SVB_FILT_SITE>
                : mode == SVB_FILT_STRICT ? villain_cluster_kernel<NT, CL, TPB, STAGES, false, SVB_FILT_STRICT>
                : sparse ? (overlap ? villain_cluster_kernel<NT, CL, TPB, 1, true, SVB_FILT_FAST, true>
                                    : villain_cluster_kernel<NT, CL, TPB, 1, false, SVB_FILT_FAST, true>)
                : overlap ? villain_cluster_kernel<NT, CL, TPB, STAGES, true, SVB_FILT_FAST>
                          : villain_cluster_kernel<NT, CL, TPB, STAGES, false, SVB_FILT_FAST>;
    constexpr int ROWS = NT / CL, VL = ROWS * NT, VHL = ROWS * NT / 2, NW = TPB / 32;
    const size_t smem = (size_t)STAGES * VL * 16 + (size_t)4 * VHL * sizeof(float) + (size_t)(6 * NW + 6) * sizeof(double) + 16 * STAGES +
                        81 * sizeof(float4);
    static int clusters_cache[7][64];
    const int variant = mode != SVB_FILT_FAST ? 1 + mode : sparse ? 5 + (overlap ? 1 : 0) : (overlap ? 1 : 0);
    int clusters = (info.device < 64) ? clusters_cache[variant][info.device] : 0;
    if (clusters == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CL * info.sm_count); cfg.blockDim = dim3(TPB); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SVB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg));
        if (clusters < 1) return fail(SVB_E_UNSUPPORTED, "cluster villain kernel does not fit the device at N=%d", NT);
        if (info.device < 64) clusters_cache[variant][info.device] = clusters;
    }
    long long n_clusters = clusters;
    if (n_clusters > a.chains) n_clusters = a.chains;
    const FilterConsts fc = make_filter_consts(a.interval_phi, mode == SVB_FILT_EXACT ? 1 : a.W, a.interval_n);
    if (overlap) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_clusters * CL)); cfg.blockDim = dim3(TPB); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SVB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a, fc));
        return 0;
    }
    kern<<<(unsigned)(n_clusters * CL), TPB, smem, stream>>>(a, fc);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}
