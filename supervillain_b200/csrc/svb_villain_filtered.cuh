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
// (included inside namespace svb; <type_traits> comes with svb_villain.cu)

// Band of the fp32 decision (everything in units of ln 2 inside the kernel; natural units here).  With
// D = interval_phi + 2 pi W interval_n >= |dr|, R = max |r| over the four links and eps = 2^-24, worst cases:
//   delta_dr <= 2.9e-7 D    dphi32 from 23 centred bits (I 2^-23) + fp32 constants (2 pi, I) + the roundings of
//                           c I_n -+ dphi and of the fused multiply-add
//   delta_r  <= 3 eps R + delta_dr     rounding of r to fp32 + at most one fp32 patch by an accepted neighbour
//   one term dr (2r + dr):  (|2r + dr| + |dr|) delta_dr + 2 |dr| delta_r + 4 eps |dr| |2r + dr|
//   four terms, their accumulation and the final multiply:
//       |dS32 - dS| <= (kappa/2) (7.6e-6 D R + 6.6e-6 D^2)
// The kernel uses bA = 1.0e-5 D and bB = 9.0e-6 D^2 (1.3 x the worst case; typical errors are ~30 x smaller).  On the
// other side, -log2 u is known to 2.2e-5 (bracket [f, f + 1] 2^-32 once f >= 2^16; smaller f always take the exact path)
// + 2.4e-7 (1 + |lg2|) (MUFU.LG2) + 2e-6 (roundings): the kernel allows 3.7e-5 + 4e-6 L.
struct FilterConsts {
    float I, two_I;          // interval_phi, 2 interval_phi
    float c;                 // 2 pi W
    float bA, bB;            // band coefficients
    float g_bias;            // 12582912 + interval_n: subtracting it from the planted digit gives dg as a float
};

static FilterConsts make_filter_consts(double interval_phi, int W, int interval_n) {
    FilterConsts fc;
    const double D = interval_phi + SVB_TWO_PI * W * interval_n;
    fc.I = (float)interval_phi;
    fc.two_I = (float)(2.0 * interval_phi);
    fc.c = (float)(SVB_TWO_PI * W);
    fc.bA = (float)(1.0e-5 * D);
    fc.bB = (float)(9.0e-6 * D * D);
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

// The same decision in the reference's operation order (STRICT arithmetic of villain_site_update: one rounding per numpy
// operation, neighborhood.py:91,110-112): p.c = 2 pi and p.g = dn in units of 1.  With this fallback the filtered kernel is
// decision for decision the STRICT kernel -- the filter only ever answers when the answer does not depend on the last bits.
__device__ __noinline__ bool villain_exact_decision_strict(const ExactProposal& p) {
    const double pc = p.phi[p.i_c];
    const double r_f0 = __dsub_rn(__dsub_rn(p.phi[p.i_f0], pc), __dmul_rn(SVB_TWO_PI, (double)p.n0[p.i_c]));
    const double r_b0 = __dsub_rn(__dsub_rn(pc, p.phi[p.i_b0]), __dmul_rn(SVB_TWO_PI, (double)p.n0[p.i_b0]));
    const double r_f1 = __dsub_rn(__dsub_rn(p.phi[p.i_f1], pc), __dmul_rn(SVB_TWO_PI, (double)p.n1[p.i_c]));
    const double r_b1 = __dsub_rn(__dsub_rn(pc, p.phi[p.i_b1]), __dmul_rn(SVB_TWO_PI, (double)p.n1[p.i_b1]));
    const double dr_f0 = __dsub_rn(-p.dphi, __dmul_rn(p.c, (double)p.g[0])), dr_b0 = __dsub_rn(p.dphi, __dmul_rn(p.c, (double)p.g[1]));
    const double dr_f1 = __dsub_rn(-p.dphi, __dmul_rn(p.c, (double)p.g[2])), dr_b1 = __dsub_rn(p.dphi, __dmul_rn(p.c, (double)p.g[3]));
    const double hk = p.half_kappa;
    double dS = __dmul_rn(__dmul_rn(hk, dr_f0), __dadd_rn(__dmul_rn(2.0, r_f0), dr_f0));
    dS = __dadd_rn(dS, __dmul_rn(__dmul_rn(hk, dr_b0), __dadd_rn(__dmul_rn(2.0, r_b0), dr_b0)));
    dS = __dadd_rn(dS, __dmul_rn(__dmul_rn(hk, dr_f1), __dadd_rn(__dmul_rn(2.0, r_f1), dr_f1)));
    dS = __dadd_rn(dS, __dmul_rn(__dmul_rn(hk, dr_b1), __dadd_rn(__dmul_rn(2.0, r_b1), dr_b1)));
    return villain_decide_lazy(exp_clipped(-dS), p.d, p.rc);
}

__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifdef SVB_FILT_CVT_I2F
#define SVB_FILT_CVT(n) ((double)(n))
#else
#define SVB_FILT_CVT(n) int_to_double(n)
#endif

// fp64 residuals of the forward links of the sites (x0, 2k) and (x0, 2k + 1) and the integers they were built from.
struct PairResiduals {
    double r0e, r0o, r1e, r1o;
    int2 a0, a1;
};
// p_c: phi at (x0, 2k); p_u: phi at (x0 + 1, 2k); p_r: phi at (x0, 2k + 2); n0c / n1c: n at (x0, 2k)
__device__ __forceinline__ PairResiduals villain_pair_residuals(const double* p_c, const double* p_u, const double* p_r,
                                                                 const int32_t* n0c, const int32_t* n1c) {
    const double2 pc = *reinterpret_cast<const double2*>(p_c);
    const double2 pu = *reinterpret_cast<const double2*>(p_u);
    const double pr = *p_r;
    PairResiduals o;
    o.a0 = *reinterpret_cast<const int2*>(n0c);
    o.a1 = *reinterpret_cast<const int2*>(n1c);
    o.r0e = fma(-SVB_TWO_PI, SVB_FILT_CVT(o.a0.x), pu.x - pc.x);
    o.r0o = fma(-SVB_TWO_PI, SVB_FILT_CVT(o.a0.y), pu.y - pc.y);
    o.r1e = fma(-SVB_TWO_PI, SVB_FILT_CVT(o.a1.x), pc.y - pc.x);
    o.r1o = fma(-SVB_TWO_PI, SVB_FILT_CVT(o.a1.y), pr - pc.y);
    return o;
}

// Per-chain record in two halves.  chain_partials: every warp reduces its sums and lane 0 parks them in the warp's shared
// slot.  chain_finish (ONE thread, behind the next block barrier): adds the slots in warp order -- deterministic -- and writes
// the record.  No atomics and no fences: a fence in the thread that has just issued the chain's bulk store would wait for it.
// STATE slots (4 doubles per warp) and COUNTER slots (2 per warp) are separate, so the two kinds never collide.
template <bool STATE, bool COUNTERS>
__device__ __forceinline__ void chain_partials(double* red_state, double* red_count, int lane, int warp, double action,
                                               long long dn2, int w0, int w1, double sum_A, int n_acc) {
    unsigned lo = 0, hi = 0;
    if (STATE) {
        action = warp_sum(action);
        lo = __reduce_add_sync(0xffffffffu, (unsigned)(dn2 & 0xFFFFFF));
        hi = __reduce_add_sync(0xffffffffu, (unsigned)((unsigned long long)dn2 >> 24));
        w0 = __reduce_add_sync(0xffffffffu, w0);
        w1 = __reduce_add_sync(0xffffffffu, w1);
    }
    if (COUNTERS) {
        sum_A = warp_sum(sum_A);
        n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    }
    if (lane == 0) {
        if (STATE) {
            double* slot = red_state + 4 * warp;
            slot[0] = action;
            slot[1] = (double)((long long)lo + ((long long)hi << 24));
            slot[2] = (double)w0; slot[3] = (double)w1;
        }
        if (COUNTERS) {
            double* slot = red_count + 2 * warp;
            slot[0] = sum_A; slot[1] = (double)n_acc;
        }
    }
}
template <int NW, bool STATE, bool COUNTERS>
__device__ __forceinline__ void chain_finish(const double* red_state, const double* red_count, double half_kappa,
                                             double* state_row, double* counter_row) {
    if (STATE) {
        double t[4] = {0, 0, 0, 0};
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int i = 0; i < 4; ++i) t[i] += red_state[4 * w + i];
        state_row[SVB_VOBS_ACTION] = half_kappa * t[0];
        state_row[SVB_VOBS_SUM_DN2] = t[1];
        state_row[SVB_VOBS_WRAP0] = t[2];
        state_row[SVB_VOBS_WRAP1] = t[3];
    }
    if (COUNTERS) {
        double t[2] = {0, 0};
        for (int w = 0; w < NW; ++w) { t[0] += red_count[2 * w]; t[1] += red_count[2 * w + 1]; }
        counter_row[SVB_VOBS_ACCEPTED] = t[1];
        counter_row[SVB_VOBS_ACCEPTANCE] = t[0];
    }
}

// OVERLAP: the launch takes part in the overlapped-launch protocol of svb_villain_sweep_overlapped -- it may begin
// while its predecessor in the stream is still running (programmatic dependent launch), and every chain is ordered
// individually through `a.epochs`: a chain is loaded only once its epoch reads a.wait_epoch (written by the launch that
// last stored it), and a CTA sets the epochs of its chains to a.signal_epoch once all its stores and records are complete.
// UNIT: W == 1 and interval_n == 1 (the reference's defaults) as compile-time constants.
// MODE: SVB_FILT_FAST       NeighborhoodUpdate, the cold path decides in FAST fp64 arithmetic (the production sweep);
//       SVB_FILT_STRICT     the same proposals, the cold path decides in the reference's operation order: identical to
//                           the STRICT kernels;  SVB_FILT_SITE: the same with interval_n == 0 -- SiteUpdate (no dn
//                           proposals, site.py:43-120);
//       SVB_FILT_EXACT      ExactUpdate (exact.py:50-129): dphi = 0, dn = d z with z one of the 2 a.interval_n nonzero values
//                           drawn from word B (villain_get_draw), STRICT cold path.
#ifdef SVB_TRACE
// Evidence build only (tools/overlap_trace.py, -DSVB_TRACE): every CTA of an overlapped launch records %globaltimer when it starts
// and when it has published its chains, indexed by the launch's signal epoch.  Not part of the product library.
constexpr int kTraceLaunches = 128, kTraceCtas = 1536;
__device__ unsigned long long g_svb_trace[kTraceLaunches][kTraceCtas][2];
__device__ __forceinline__ unsigned long long svb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif
#define SVB_FILT_FAST 0
#define SVB_FILT_STRICT 1
#define SVB_FILT_EXACT 2
#define SVB_FILT_SITE 3          /* SVB_FILT_STRICT with interval_n == 0 at compile time: no digits, no n updates */
// SPARSE (one sweep per launch, no record of the state AFTER the sweep): phi and n leave shared memory for good once the
//       residuals are built.  An accepted proposal is applied to GLOBAL memory (five fire-and-forget reductions: phi += dphi
//       as one fp64 RED -- the reference's one rounding -- and the four n), the exact path reads the current phi and n from
//       global memory (L2), and NOTHING is stored back: an unchanged value is not rewritten.  The staging buffer is free as
//       soon as the residuals exist, so the next chain's bulk load is issued right behind the build barrier and lands during
//       the two colour passes -- the load latency that a one-stage CTA otherwise waits out per chain is gone, and so is the
//       store phase.  DRAM traffic: one read of the state + the dirty sectors (acceptance is 0.5 - 5 %).
template <int NT, int MINB, int STAGES, bool OVERLAP, bool UNIT, int MODE, bool SPARSE = false>
__global__ void __launch_bounds__(4 * NT, MINB) villain_smem_filtered_kernel(const __grid_constant__ VillainArgs a,
                                                                             const __grid_constant__ FilterConsts fc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, VH = V / 2, T = 4 * NT, NW = T / 32;
    constexpr int PER = VH / T;                                  // sites per thread per colour (rows x0 + 8 q)
    static_assert(STAGES == 1, "one shared-memory stage per CTA: occupancy, not double buffering, hides the copies");
    static_assert(PER >= 2 && PER % 2 == 0, "villain_smem_filtered_kernel: unsupported geometry");
    constexpr uint32_t bytes_phi = V * sizeof(double);
    constexpr uint32_t bytes_n = 2 * V * sizeof(int32_t);
    constexpr uint32_t stage_bytes = bytes_phi + bytes_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef SVB_TRACE
    if (OVERLAP && tid == 0 && blockIdx.x < kTraceCtas) g_svb_trace[a.signal_epoch % kTraceLaunches][blockIdx.x][0] = svb_globaltimer();
#endif
    float* rc0 = reinterpret_cast<float*>(smem_raw + STAGES * stage_bytes);      // [colour][VH]: residual of link (0, x)
    float* rc1 = rc0 + V;                                                         // [colour][VH]: residual of link (1, x)
    double* red_state = reinterpret_cast<double*>(rc1 + V);                       // [NW][4] per-warp partial sums
    double* red_count = red_state + 4 * NW;                                       // [NW][2]
    uint64_t* bar = reinterpret_cast<uint64_t*>(red_count + 2 * NW);
    // UNIT: the 81 proposals of (dn_f0, dn_b0, dn_f1, dn_b1) as the residual changes they make, -2 pi (digit): the code that the
    // multiply by 81 leaves in the high word looks up all four at once (one LDS.128 instead of nine integer operations and four
    // conversions per site).  0, 2 pi and 4 pi are exact in fp32, so base + lut is the fused multiply-add it replaces, bit for bit.
    float4* dn_lut = reinterpret_cast<float4*>(bar + 2);
    constexpr int kWriter = 32;            // finishes the records: lane 0 of warp 1 (thread 0 is busy with the bulk copies)

    if (tid == 0) {
        for (int b = 0; b < STAGES; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
    }
    if (UNIT)
        for (int i = tid; i < 81; i += T)
            dn_lut[i] = make_float4(-fc.c * (float)(i / 27), -fc.c * (float)((i / 9) % 3), -fc.c * (float)((i / 3) % 3), -fc.c * (float)(i % 3));
    if (OVERLAP) {
        // let the next launch in the stream start as soon as every CTA of this one is resident ...
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        // ... and, unless the caller vouches for the predecessor, wait for everything before this launch
        if (a.grid_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    __syncthreads();
    const bool obs_of_input = a.obs_in != nullptr;                // state columns describe the chain as it ARRIVES
    const bool want_obs = a.obs != nullptr && !obs_of_input;     // ... or as it leaves (one more fp64 pass)
    static_assert(!(UNIT && MODE != SVB_FILT_FAST), "UNIT is the production sweep");
    static_assert(!SPARSE || MODE == SVB_FILT_FAST, "SPARSE serves the NeighborhoodUpdate sweep");
    const int interval_n = UNIT ? 1 : (MODE == SVB_FILT_SITE ? 0 : a.interval_n);
    const uint32_t K = (MODE == SVB_FILT_EXACT) ? (uint32_t)(2 * interval_n) : (uint32_t)(2 * interval_n + 1);
    const int W = (UNIT || MODE == SVB_FILT_EXACT) ? 1 : a.W, mWI = -W * interval_n;
    const float cIn = fc.c * (float)interval_n;
    const float2 cIn2 = make_float2(cIn, cIn), negc2 = make_float2(-fc.c, -fc.c), two2 = make_float2(2.0f, 2.0f);
    const double two_I_scaled = (2.0 * a.interval_phi) * 2.3283064365386963e-10;          // (2 I) 2^-32, exact scaling

    // per-thread geometry: rows row8 + 8 q of the compact column k
    const int row8 = tid / HN, k = tid - row8 * HN;
    const int cc = row8 & 1;                                      // colour of the even-column site of this thread's pairs
    const int wrap0 = (row8 == 0) ? VH : 0;                       // backward-0 neighbour of row 0 is row N - 1
    const int up_off = (row8 == 7) ? (N - V) : N;                 // row below the LAST of this thread's rows wraps to row 0

    // The epochs of ALL of this CTA's chains are released together when the CTA is done: one gpu-scope release fence
    // (a memory barrier that costs ~0.7 us on a busy SM -- per chain it would eat most of what overlapping wins)
    // followed by one relaxed store per chain.  Call behind a block barrier before which thread 0 -- the thread whose
    // bulk stores they were -- has seen the stores complete and the last warp has written the records.
    auto publish_all = [&](int count) {
#ifdef SVB_OV_NOPUB
        return;
#endif
        if (OVERLAP) {
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (int i = lane; i < count; i += 32)
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.epochs + blockIdx.x + (long long)i * gridDim.x),
                             "r"(a.signal_epoch)
                             : "memory");
        }
    };
    // `seen`: a value of the chain's epoch read earlier (peek_epoch), so that the common case -- the producer finished
    // long ago -- costs no round trip to L2 at the point where the load is issued
    auto peek_epoch = [&](long long chain) -> uint32_t {
        uint32_t e = a.wait_epoch;
#ifdef SVB_OV_NOWAIT
        return e;
#endif
        if (OVERLAP && !a.grid_wait)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(a.epochs + chain) : "memory");
        return e;
    };
    auto issue_load = [&](long long chain, int b, uint32_t seen) {
        if (OVERLAP && !a.grid_wait) {
            uint32_t e = seen;
            unsigned ns = 32, naps = 0;
            while (true) {
                if (e == a.wait_epoch) break;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(a.epochs + chain) : "memory");
                if (e == a.wait_epoch) break;
                __nanosleep(ns);
                if (ns < 1024) ns *= 2;
                // a producer that never comes is a caller error (wrong epochs): fail the launch instead of hanging the GPU
                if (++naps > (1u << 21)) __trap();          // > 2 s
            }
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        mbar_expect_tx(&bar[b], stage_bytes);
        bulk_g2s(stage, reinterpret_cast<const double*>(a.phi) + chain * V, bytes_phi, &bar[b]);
        bulk_g2s(stage + bytes_phi, a.n + chain * 2 * V, bytes_n, &bar[b]);
    };

    long long chain = blockIdx.x;
    if (tid == 0 && chain < a.chains) issue_load(chain, 0, peek_epoch(chain));

    int it = 0;
    for (; chain < a.chains; chain += gridDim.x, ++it) {
        constexpr int b = 0;
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        double* sphi = reinterpret_cast<double*>(stage);
        int32_t* sn0 = reinterpret_cast<int32_t*>(stage + bytes_phi);
        int32_t* sn1 = sn0 + V;
        const long long next = chain + gridDim.x;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double half_kappa = kappa / 2;
        const float hk2 = (float)(half_kappa * 1.4426950408889634);                 // decisions are taken in units of ln 2
        const float hkA = 1.0001f * hk2 * fc.bA, hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
        const float2 hk22 = make_float2(hk2, hk2), hkA2 = make_float2(hkA, hkA), hkB2 = make_float2(hkB, hkB);
        // pair pointers: phi / n at (row8, 2k); rows advance by 8 N per q
        const double* pp = sphi + row8 * N + 2 * k;
        const double* pp_r = sphi + row8 * N + ((2 * k + 2) & (N - 1));
        const int32_t* pn0 = sn0 + row8 * N + 2 * k;
        const int32_t* pn1 = sn1 + row8 * N + 2 * k;

        // SPARSE: the fields of this chain in global memory (accepted proposals and the exact path go there)
        double* gphi = reinterpret_cast<double*>(a.phi) + chain * V;
        int32_t* gn0 = a.n + chain * 2 * V;
        int32_t* gn1 = gn0 + V;
        uint32_t seen_sparse = 0;
        if (SPARSE && tid == 0 && next < a.chains) seen_sparse = peek_epoch(next);      // lands during the residual build

        mbar_wait(&bar[b], (uint32_t)(it & 1));

        int n_acc = 0;
        double sum_A_all = 0.0;        // per sweep in fp32 (<= 2 PER terms per thread), across sweeps in fp64
        for (int s = 0; s < a.n_sweeps; ++s) {
            float sum_A = 0.0f;
            // ---- r = d(phi) - 2 pi n   (neighborhood.py:91) in fp64, stored rounded to fp32 ----
            {
                float* w0e = rc0 + cc * VH + tid;
                float* w1e = rc1 + cc * VH + tid;
                float* w0o = rc0 + (cc ^ 1) * VH + tid;
                float* w1o = rc1 + (cc ^ 1) * VH + tid;
                const bool sums = obs_of_input && s == 0;            // the observables of the arriving state ride along
                double action = 0.0;
                int w0 = 0, w1 = 0;
                long long dn2 = 0;
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const int o = 8 * N * q;
                    const int uo = (q == PER - 1) ? up_off : N;
                    const PairResiduals pr = villain_pair_residuals(pp + o, pp + o + uo, pp_r + o, pn0 + o, pn1 + o);
                    w0e[T * q] = (float)pr.r0e;
                    w1e[T * q] = (float)pr.r1e;
                    w0o[T * q] = (float)pr.r0o;
                    w1o[T * q] = (float)pr.r1o;
                    if (sums) {
                        action = fma(pr.r0e, pr.r0e, action);
                        action = fma(pr.r0o, pr.r0o, action);
                        action = fma(pr.r1e, pr.r1e, action);
                        action = fma(pr.r1o, pr.r1o, action);
                        const int hr = sn0[(row8 + 8 * q) * N + ((2 * k + 2) & (N - 1))];            // n0[x + 2 e1]
                        const int2 up = *reinterpret_cast<const int2*>(pn1 + o + uo);                 // n1[x + e0]
                        const int d0 = (up.x - pr.a1.x) - (pr.a0.y - pr.a0.x), d1 = (up.y - pr.a1.y) - (hr - pr.a0.y);
                        dn2 += (long long)d0 * d0 + (long long)d1 * d1;
                        w0 += pr.a0.x + pr.a0.y;
                        w1 += pr.a1.x + pr.a1.y;
                    }
                }
                if (sums) chain_partials<true, false>(red_state, red_count, lane, warp, action, dn2, w0, w1, 0.0, 0);
            }
            __syncthreads();
            // SPARSE: nobody reads the staged phi and n again -- the next chain may land on them while this one is swept
            if (SPARSE && tid == 0 && next < a.chains) issue_load(next, 0, seen_sparse);
            if (obs_of_input && s == 0 && tid == kWriter)
                chain_finish<NW, true, false>(red_state, red_count, kappa / 2, a.obs_in + chain * SVB_VOBS_COUNT, nullptr);

            const unsigned long long gc = a.chain0 + (unsigned long long)chain, gs = a.sweep0 + (unsigned long long)s;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int par = (row8 + c) & 1;                    // column parity of this thread's sites of colour c
                const int x1 = 2 * k + par;
                const int ob1 = par ? 0 : ((k == 0) ? (1 - HN) : 1);          // compact index of x - e1 is j - ob1
                const int wrap1 = (x1 == 0) ? N : 0;
                // q-independent bases; per q the offsets are compile-time immediates
                float* R0own = rc0 + c * VH + tid;
                float* R1own = rc1 + c * VH + tid;
                float* R0b = rc0 + (c ^ 1) * VH + tid - HN;        // backward link (0, x - e0): row above, same compact column
                float* R0b_q0 = R0b + wrap0;
                float* R1b = rc1 + (c ^ 1) * VH + tid - ob1;       // backward link (1, x - e1)
                double* Pc = sphi + 2 * tid + par;
                int32_t* N0c = sn0 + 2 * tid + par;
                int32_t* N1c = sn1 + 2 * tid + par;
                int32_t* N0b = N0c - N;
                int32_t* N0b_q0 = N0b + 2 * wrap0;
                int32_t* N1b = N1c - 1 + wrap1;
#pragma unroll
                for (int p = 0; p < PER / 2; ++p) {
                    const uint32_t c0 = (uint32_t)((row8 + 16 * p) * N + x1);                 // villain_pair_counter
                    const Philox4 bits = philox_site_keys(a, gc, gs, c0);
                    const int qA = 2 * p, qB = 2 * p + 1;
                    float* r0bA = (p == 0) ? R0b_q0 : R0b;             // q == 0 is the only row whose e0-neighbour wraps
                    int32_t* n0bA = (p == 0) ? N0b_q0 : N0b;
                    // proposals: four base-K digits each, then the leading 32 bits of the uniform
                    uint32_t fA = bits.y, fB = bits.w;
                    uint32_t codeA = 0, codeB = 0;
                    int digA[4], digB[4];
                    if (MODE == SVB_FILT_EXACT) {
                        // z = the idx-th of the 2 I nonzero values in [-I, I]; n += d z: forward links -z, backward links +z.
                        // As "digits" (dn = digit - I): I - z forward, I + z backward.
                        const uint64_t pa = (uint64_t)fA * K, pb = (uint64_t)fB * K;
                        fA = (uint32_t)pa; fB = (uint32_t)pb;
                        const int ia = (int)(pa >> 32), ib = (int)(pb >> 32);
                        const int za = (ia < interval_n) ? ia - interval_n : ia - interval_n + 1;
                        const int zb = (ib < interval_n) ? ib - interval_n : ib - interval_n + 1;
                        digA[0] = digA[2] = interval_n - za; digA[1] = digA[3] = interval_n + za;
                        digB[0] = digB[2] = interval_n - zb; digB[1] = digB[3] = interval_n + zb;
                    } else if (MODE == SVB_FILT_SITE) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) digA[i] = digB[i] = 0;            // word B is the uniform's leading bits as it stands
                    } else if (UNIT) {
                        // four successive multiply-highs by 3 == one by 81: hi = 27 d0 + 9 d1 + 3 d2 + d3 (the code), lo = the remainder
                        // (one quarter-rate IMAD.WIDE instead of a chain of four); the digits themselves are needed only by the
                        // rare paths (unit_digits below) -- the residual changes come from the look-up table
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
                    // dphi from 23 centred bits
                    float2 U = make_float2(__uint_as_float(0x3F800000u | (bits.x >> 9)), __uint_as_float(0x3F800000u | (bits.z >> 9)));
                    U = __fadd2_rn(U, make_float2(-0.99999994f, -0.99999994f));
                    const float2 dphi = __ffma2_rn(make_float2(fc.two_I, fc.two_I), U, make_float2(-fc.I, -fc.I));
                    const float2 base_f = __ffma2_rn(dphi, make_float2(-1.0f, -1.0f), cIn2), base_b = __fadd2_rn(cIn2, dphi);
                    const float2 r_f0 = make_float2(R0own[T * qA], R0own[T * qB]), r_f1 = make_float2(R1own[T * qA], R1own[T * qB]);
                    const float2 r_b0 = make_float2(r0bA[T * qA], R0b[T * qB]), r_b1 = make_float2(R1b[T * qA], R1b[T * qB]);
                    // dr = d(dphi) - 2 pi dn   (neighborhood.py:110), dn = W (digit - interval_n)
                    float2 dr_f0, dr_b0, dr_f1, dr_b1;
                    if (UNIT) {
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
                    const float2 dS2 = __fmul2_rn(hk22, acc2);                          // dS / ln 2
                    // -log2(f 2^-32); u lies in [f, f + 1] 2^-32
                    const float2 L2 = __ffma2_rn(make_float2(fast_lg2((float)fA), fast_lg2((float)fB)), make_float2(-1.0f, -1.0f),
                                                 make_float2(32.0f, 32.0f));
                    const float2 Rmax = make_float2(fmaxf(fmaxf(fabsf(r_f0.x), fabsf(r_b0.x)), fmaxf(fabsf(r_f1.x), fabsf(r_b1.x))),
                                                    fmaxf(fmaxf(fabsf(r_f0.y), fabsf(r_b0.y)), fmaxf(fabsf(r_f1.y), fabsf(r_b1.y))));
                    const float2 band = __ffma2_rn(hkA2, Rmax, __ffma2_rn(make_float2(4e-6f, 4e-6f), L2, hkB2));
                    const float2 diff = __ffma2_rn(L2, make_float2(-1.0f, -1.0f), dS2);
                    sum_A += fminf(fast_ex2(-dS2.x), 1.0f) + fminf(fast_ex2(-dS2.y), 1.0f);
                    // the residuals an accepted proposal leaves behind
                    const float2 n_f0 = __fadd2_rn(r_f0, dr_f0), n_b0 = __fadd2_rn(r_b0, dr_b0);
                    const float2 n_f1 = __fadd2_rn(r_f1, dr_f1), n_b1 = __fadd2_rn(r_b1, dr_b1);
                    // certainly rejected (the overwhelming majority of proposals): nothing more to do.  Everything else -- accepted, or
                    // inside the error band of the fp32 comparison -- is handled behind ONE branch per pair of sites.
                    const bool candA = !(diff.x > band.x) || fA < 65536u, candB = !(diff.y > band.y) || fB < 65536u;
                    if (candA || candB) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (!(h ? candB : candA)) continue;
                        const int q = 2 * p + h;
                        float* r0b = h ? R0b : r0bA;
                        int32_t* n0b = h ? N0b : n0bA;
                        const uint32_t wA = h ? bits.z : bits.x;
                        const uint32_t f = h ? fB : fA;
                        int dig[4];
                        if (UNIT) {
                            uint32_t code = h ? codeB : codeA;
                            dig[0] = (int)((code * 2428u) >> 16); code -= 27u * (uint32_t)dig[0];
                            dig[1] = (int)((code * 7282u) >> 16); code -= 9u * (uint32_t)dig[1];
                            dig[2] = (int)((code * 21846u) >> 16); dig[3] = (int)(code - 3u * (uint32_t)dig[2]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) dig[i] = h ? digB[i] : digA[i];
                        }
                        const float dif = h ? diff.y : diff.x, bnd = h ? band.y : band.x;
                        const int x0 = row8 + 8 * q;
                        bool ok = dif < 0.0f;
                        if (!(fabsf(dif) > bnd) || f < 65536u) {
                            ExactProposal ep;
                            if (SPARSE) { ep.phi = gphi; ep.n0 = gn0; ep.n1 = gn1; }
                            else { ep.phi = sphi; ep.n0 = sn0; ep.n1 = sn1; }
                            ep.i_c = x0 * N + x1;
                            ep.i_b0 = ((x0 - 1) & (N - 1)) * N + x1;
                            ep.i_b1 = x0 * N + ((x1 - 1) & (N - 1));
                            ep.i_f0 = ((x0 + 1) & (N - 1)) * N + x1;
                            ep.i_f1 = x0 * N + ((x1 + 1) & (N - 1));
                            ep.half_kappa = half_kappa;
                            ep.dphi = (MODE == SVB_FILT_EXACT) ? 0.0 : villain_dphi_from_word(wA, a.interval_phi);
                            ep.d.f = f; ep.d.c0 = c0; ep.d.half = (uint32_t)h;
                            ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs; ep.rc.stream = a.refine_stream; ep.rc.wide = 0;
                            if (MODE == SVB_FILT_FAST) {
                                ep.c = SVB_TWO_PI * (double)W;
#pragma unroll
                                for (int i = 0; i < 4; ++i) ep.g[i] = dig[i] - interval_n;
                                ok = villain_exact_decision(ep);
                            } else {
                                ep.c = SVB_TWO_PI;
#pragma unroll
                                for (int i = 0; i < 4; ++i) ep.g[i] = W * (dig[i] - interval_n);
                                ok = villain_exact_decision_strict(ep);
                            }
                        }
                        n_acc += ok ? 1 : 0;
                        if (ok) {                                               // (:121-129)
                            // phi += dphi with dphi = -I + fl((2 I 2^-32) (A + 1/2)): the scaling by 2^-32 is exact, so this
                            // is villain_dphi_from_word bit for bit with one multiply less
                            const double Ah = __hiloint2double(0x43300000, (int)wA) - 4503599627370495.5;
                            if (SPARSE) {
                                // straight to global memory, nothing waits for it: the fp64 reduction rounds once, to nearest,
                                // like the reference's phi + dphi
                                const int ic = x0 * N + x1;
                                atomicAdd(gphi + ic, __dadd_rn(-a.interval_phi, __dmul_rn(two_I_scaled, Ah)));
                                atomicAdd(gn0 + ic, W * dig[0] + mWI);
                                atomicAdd(gn0 + ((x0 - 1) & (N - 1)) * N + x1, W * dig[1] + mWI);
                                atomicAdd(gn1 + ic, W * dig[2] + mWI);
                                atomicAdd(gn1 + x0 * N + ((x1 - 1) & (N - 1)), W * dig[3] + mWI);
                            } else {
                                Pc[2 * T * q] = __dadd_rn(Pc[2 * T * q], __dadd_rn(-a.interval_phi, __dmul_rn(two_I_scaled, Ah)));
                                if (MODE != SVB_FILT_SITE) {
                                    atomicAdd(N0c + 2 * T * q, W * dig[0] + mWI);  // only this thread touches these links in this pass
                                    atomicAdd(n0b + 2 * T * q, W * dig[1] + mWI);
                                    atomicAdd(N1c + 2 * T * q, W * dig[2] + mWI);
                                    atomicAdd(N1b + 2 * T * q, W * dig[3] + mWI);
                                }
                            }
                            R0own[T * q] = h ? n_f0.y : n_f0.x;
                            r0b[T * q] = h ? n_b0.y : n_b0.x;
                            R1own[T * q] = h ? n_f1.y : n_f1.x;
                            R1b[T * q] = h ? n_b1.y : n_b1.x;
                        }
                    }
                    }
                }
                if (!SPARSE && s == a.n_sweeps - 1 && c == 1) fence_proxy_async();      // this thread's phi / n writes -> visible to the bulk store
                // SPARSE (one sweep): this launch's counters ride on the barrier that ends the last pass
                if (SPARSE && c == 1 && a.obs != nullptr)
                    chain_partials<false, true>(red_state, red_count, lane, warp, 0.0, 0, 0, 0, sum_A_all + (double)sum_A, n_acc);
                __syncthreads();
            }
            sum_A_all += (double)sum_A;
        }

        if (SPARSE) {
            // nothing to store; the state columns went to obs_in with the build, the counters' slots are complete
            if (tid == kWriter && a.obs != nullptr)
                chain_finish<NW, false, true>(red_state, red_count, kappa / 2, nullptr, a.obs + chain * SVB_VOBS_COUNT);
            continue;
        }
        // ---- store the chain (nothing below writes phi or n) ----
        uint32_t seen_next = 0;
        if (tid == 0) {
            bulk_s2g(reinterpret_cast<double*>(a.phi) + chain * V, sphi, bytes_phi);
            bulk_s2g(a.n + chain * 2 * V, sn0, bytes_n);
            bulk_commit();
            if (next < a.chains) seen_next = peek_epoch(next);     // lands during the observable pass
        }

        if (want_obs) {
            // one fp64 pass over site pairs: sum r^2, sum n, sum (dn)^2 of the final state
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
                const int hr = sn0[(row8 + 8 * q) * N + ((2 * k + 2) & (N - 1))];            // n0[x + 2 e1]
                const int2 up = *reinterpret_cast<const int2*>(pn1 + o + uo);                 // n1[x + e0]
                // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x])      (compact.py d,1 rows)
                const int d0 = (up.x - pr.a1.x) - (pr.a0.y - pr.a0.x), d1 = (up.y - pr.a1.y) - (hr - pr.a0.y);
                dn2 += (long long)d0 * d0 + (long long)d1 * d1;
                w0 += pr.a0.x + pr.a0.y;
                w1 += pr.a1.x + pr.a1.y;
            }
            chain_partials<true, true>(red_state, red_count, lane, warp, action, dn2, w0, w1, sum_A_all, n_acc);
        } else if (a.obs != nullptr) {
            // the state columns went to obs_in at the start; only this launch's counters remain
            chain_partials<false, true>(red_state, red_count, lane, warp, 0.0, 0, 0, 0, sum_A_all, n_acc);
        }
        __syncthreads();               // every warp has read the final state before the buffer is refilled
        if (tid == 0) {
            bulk_wait_read0();
            if (next < a.chains) issue_load(next, 0, seen_next);
        }
        if (tid == kWriter && a.obs != nullptr) {
            double* row = a.obs + chain * SVB_VOBS_COUNT;
            if (want_obs) chain_finish<NW, true, true>(red_state, red_count, kappa / 2, row, row);
            else chain_finish<NW, false, true>(red_state, red_count, kappa / 2, nullptr, row);
        }
    }
    if (!SPARSE && tid == 0) bulk_wait0();
    if (OVERLAP) {
        __syncthreads();                                   // every store has completed, every record is written
        if (warp == 0) publish_all(it);
    }
#ifdef SVB_TRACE
    if (OVERLAP && tid == 0 && blockIdx.x < kTraceCtas) g_svb_trace[a.signal_epoch % kTraceLaunches][blockIdx.x][1] = svb_globaltimer();
#endif
}

template <int NT, int MINB, int STAGES>
static int launch_villain_filtered(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    const bool overlap = a.epochs != nullptr;
    const int mode = a.exact_mode ? SVB_FILT_EXACT : (a.filtered_strict ? (a.interval_n == 0 ? SVB_FILT_SITE : SVB_FILT_STRICT) : SVB_FILT_FAST);
    const bool unit = mode == SVB_FILT_FAST && a.W == 1 && a.interval_n == 1;
    if (mode != SVB_FILT_FAST && overlap) return fail(SVB_E_UNSUPPORTED, "overlapped launches serve the NeighborhoodUpdate sweep only");
    // one sweep per launch and no record of the state after it: nothing is stored back (SPARSE, see the kernel)
    const char* env_sparse = getenv("SVB_VILLAIN_SPARSE");
    const bool sparse = mode == SVB_FILT_FAST && a.n_sweeps == 1 && (a.obs == nullptr || a.obs_in != nullptr) &&
                        !(env_sparse && env_sparse[0] == '0');
    auto kern = mode == SVB_FILT_EXACT    ? villain_smem_filtered_kernel<NT, MINB, STAGES, false, false, SVB_FILT_EXACT>
                : mode == SVB_FILT_SITE   ? villain_smem_filtered_kernel<NT, MINB, STAGES, false, false, SVB_FILT_SITE>
                : mode == SVB_FILT_STRICT ? villain_smem_filtered_kernel<NT, MINB, STAGES, false, false, SVB_FILT_STRICT>
                : sparse ? (overlap ? (unit ? villain_smem_filtered_kernel<NT, MINB, STAGES, true, true, SVB_FILT_FAST, true>
                                            : villain_smem_filtered_kernel<NT, MINB, STAGES, true, false, SVB_FILT_FAST, true>)
                                    : (unit ? villain_smem_filtered_kernel<NT, MINB, STAGES, false, true, SVB_FILT_FAST, true>
                                            : villain_smem_filtered_kernel<NT, MINB, STAGES, false, false, SVB_FILT_FAST, true>))
                : overlap ? (unit ? villain_smem_filtered_kernel<NT, MINB, STAGES, true, true, SVB_FILT_FAST>
                                  : villain_smem_filtered_kernel<NT, MINB, STAGES, true, false, SVB_FILT_FAST>)
                          : (unit ? villain_smem_filtered_kernel<NT, MINB, STAGES, false, true, SVB_FILT_FAST>
                                  : villain_smem_filtered_kernel<NT, MINB, STAGES, false, false, SVB_FILT_FAST>);
    const size_t V = (size_t)NT * NT;
    const size_t smem = STAGES * V * 16 + 2 * V * sizeof(float) + 6 * (4 * NT / 32) * sizeof(double) + 32 + 81 * sizeof(float4);
    // kernel attributes and occupancy are set / queried once per (instantiation, device)
    static int per_sm_cache[11][64];
    const int variant = mode != SVB_FILT_FAST ? 3 + mode : sparse ? 7 + (overlap ? 1 : 0) + (unit ? 2 : 0) : (overlap ? 1 : 0) + (unit ? 2 : 0);
    int per_sm = (info.device < 64) ? per_sm_cache[variant][info.device] : 0;
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 4 * NT, smem));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "filtered villain kernel does not fit an SM at N=%d", NT);
        if (info.device < 64) per_sm_cache[variant][info.device] = per_sm;
    }
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > a.chains) grid = a.chains;
    const FilterConsts fc = make_filter_consts(a.interval_phi, mode == SVB_FILT_EXACT ? 1 : a.W, a.interval_n);
    if (overlap) {
        // programmatic dependent launch: this grid may start once every CTA of the previous kernel in the stream has
        // executed griddepcontrol.launch_dependents (or exited); the per-chain epochs order the data
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(4 * NT); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SVB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a, fc));
        return 0;
    }
    kern<<<(unsigned)grid, 4 * NT, smem, stream>>>(a, fc);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// TILED path with fp32-filtered decisions (configs 4 and 5: lattices beyond shared memory; FAST arithmetic).
//
// The geometry, ghost zones and ping-pong of villain_tiled_kernel (svb_villain.cu) with the arithmetic of the kernel
// above: the residuals of every link inside the 37 x 38 region are built once in fp64 and kept in fp32, proposals are
// decided in fp32 with the exact fp64 + lazy-uniform test inside the error band, phi / n are touched on acceptance.
// Philox is keyed by the GLOBAL site, so redundant ghost updates are bit-identical in every tile that computes them --
// including the (rare) exact-path decisions, which depend only on the global state around the site.
//
// Table-driven: what does not depend on the tile -- the list of work items of each colour pass -- is tabulated once per
// process.  An item is the PAIR of region rows (r, r + 8) of one column that share a Philox block (rows 0..7 and 16..23 of
// the tile, bit 3 of the global row clear), or a single site for the ghost rows whose partner lies outside the region;
// per tile only the wrapped global coordinates of the region's rows and columns are tabulated.  A proposal then costs a
// table entry and two coordinate look-ups instead of divisions and wraps, and half a Philox block.
// ------------------------------------------------------------------------------------------
constexpr int kItemsPerPass = 384;                 // 3 rounds of 128 threads
constexpr unsigned kNoSite = 0xFFFFu;

// per colour: local index of site A | local index of site B << 16 (kNoSite: none).  The same for every tile of every
// lattice: filled once per process by villain_tiled_items_kernel (like a kernel attribute, not state of any call).
__device__ uint32_t g_tiled_items[2][kItemsPerPass];

struct TiledShared {
    double phi[kRegSize];
    int32_t n0[kRegSize];
    int32_t n1[kRegSize];
    float r0[kRegSize];            // residual of link (0, x): x -> x + e0 (valid for local rows < 36)
    float r1[kRegSize];            // residual of link (1, x): x -> x + e1 (valid for local cols < 36)
    int x0row[kRegRows + 3];       // wrapped global row of local row i
    int x1col[kRegCols + 2];       // wrapped global column of local column j
    double red[2 * 32];
};

// One proposal at local index l of the region (global site (x0, x1)); returns the acceptance estimate, sets ok.
__device__ __forceinline__ float villain_tiled_site(TiledShared& sh, const VillainArgs& a, const FilterConsts& fc, int l, uint32_t wA,
                                                    uint32_t wB, uint32_t c0, uint32_t half, double half_kappa, float hk2, float hkA,
                                                    float hkB, uint32_t K, int W, int mWI, float cIn, unsigned long long gc,
                                                    unsigned long long gs, bool& ok) {
    uint32_t f = wB;
    int dig[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint64_t prod = (uint64_t)f * K;
        f = (uint32_t)prod;
        dig[q] = (int)(prod >> 32);
    }
    const float U = __uint_as_float(0x3F800000u | (wA >> 9)) - 0.99999994f;
    const float dphi = fmaf(fc.two_I, U, -fc.I);
    const float base_f = cIn - dphi, base_b = cIn + dphi;
    const float r_f0 = sh.r0[l], r_b0 = sh.r0[l - kRegCols], r_f1 = sh.r1[l], r_b1 = sh.r1[l - 1];
    const float dr_f0 = fmaf(-fc.c, (float)dig[0], base_f), dr_b0 = fmaf(-fc.c, (float)dig[1], base_b);
    const float dr_f1 = fmaf(-fc.c, (float)dig[2], base_f), dr_b1 = fmaf(-fc.c, (float)dig[3], base_b);
    float acc2 = dr_f0 * fmaf(2.0f, r_f0, dr_f0);
    acc2 = fmaf(dr_b0, fmaf(2.0f, r_b0, dr_b0), acc2);
    acc2 = fmaf(dr_f1, fmaf(2.0f, r_f1, dr_f1), acc2);
    acc2 = fmaf(dr_b1, fmaf(2.0f, r_b1, dr_b1), acc2);
    const float dS2 = hk2 * acc2;
    const float L2 = 32.0f - fast_lg2((float)f);
    const float Rmax = fmaxf(fmaxf(fabsf(r_f0), fabsf(r_b0)), fmaxf(fabsf(r_f1), fabsf(r_b1)));
    const float band = fmaf(hkA, Rmax, fmaf(4e-6f, L2, hkB));
    const float diff = dS2 - L2;
    ok = diff < 0.0f;
    const float Aest = fminf(fast_ex2(-dS2), 1.0f);
    if (!(fabsf(diff) > band) || f < 65536u) {
        ExactProposal ep;
        ep.phi = sh.phi; ep.n0 = sh.n0; ep.n1 = sh.n1;
        ep.i_c = l; ep.i_b0 = l - kRegCols; ep.i_b1 = l - 1; ep.i_f0 = l + kRegCols; ep.i_f1 = l + 1;
        ep.half_kappa = half_kappa;
        ep.c = SVB_TWO_PI * (double)W;
        ep.dphi = villain_dphi_from_word(wA, a.interval_phi);
#pragma unroll
        for (int q = 0; q < 4; ++q) ep.g[q] = dig[q] - a.interval_n;
        ep.d.f = f; ep.d.c0 = c0; ep.d.half = half;
        ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs; ep.rc.stream = a.refine_stream; ep.rc.wide = 0;
        ok = villain_exact_decision(ep);
    }
    if (ok) {
        sh.phi[l] = __dadd_rn(sh.phi[l], villain_dphi_from_word(wA, a.interval_phi));
        sh.n0[l] += W * dig[0] + mWI;
        sh.n0[l - kRegCols] += W * dig[1] + mWI;
        sh.n1[l] += W * dig[2] + mWI;
        sh.n1[l - 1] += W * dig[3] + mWI;
        sh.r0[l] = r_f0 + dr_f0;
        sh.r0[l - kRegCols] = r_b0 + dr_b0;
        sh.r1[l] = r_f1 + dr_f1;
        sh.r1[l - 1] = r_b1 + dr_b1;
    }
    return Aest;
}

// The work items of the two colour passes.  Colour c works on local [lo, lo + side)^2, lo = 1 / 2, side = 35 / 33; tile row
// r = i - 2.  Rows r in {0..7, 16..23} pair with r + 8 (same column, same colour); the remaining rows of the pass (c = 0:
// r = -1, 32, 33; c = 1: r = 32) are single.
__global__ void villain_tiled_items_kernel() {
    for (int c = 0; c < 2; ++c) {
        const int lo = (c == 0) ? 1 : 2, side = (c == 0) ? kTile + 3 : kTile + 1, per_row = (side + 1) / 2;
        const int n_single = (c == 0) ? 3 : 1;
        for (int e = threadIdx.x; e < kItemsPerPass; e += blockDim.x) {
            uint32_t item = kNoSite | (kNoSite << 16);
            const int row_slot = e / per_row, kcol = e - row_slot * per_row;
            if (row_slot < 16 + n_single) {
                int rA, rB = -100;
                if (row_slot < 16) {
                    rA = (row_slot < 8) ? row_slot : row_slot + 8;
                    rB = rA + 8;
                } else {
                    rA = (c == 0) ? ((row_slot == 16) ? -1 : 32 + (row_slot - 17)) : 32;
                }
                const int iA = rA + 2;
                const int j = lo + 2 * kcol + ((iA + lo + c) & 1);          // (i + j) & 1 == c (tile origins are even)
                if (j < lo + side) {
                    const uint32_t lA = (uint32_t)(iA * kRegCols + j);
                    const uint32_t lB = (rB > -100) ? (uint32_t)((rB + 2) * kRegCols + j) : kNoSite;
                    item = lA | (lB << 16);
                }
            }
            g_tiled_items[c][e] = item;
        }
    }
}

__global__ void __launch_bounds__(128, 6) villain_tiled_filtered_kernel(const __grid_constant__ VillainArgs a,
                                                                        const __grid_constant__ FilterConsts fc,
                                                                        const double* __restrict__ phi_in,
                                                                        const int32_t* __restrict__ n_in, double* __restrict__ phi_out,
                                                                        int32_t* __restrict__ n_out, int sweep, int tiles_per_side,
                                                                        int fuse_obs) {
    extern __shared__ __align__(16) unsigned char tiled_smem_raw[];
    TiledShared& sh = *reinterpret_cast<TiledShared*>(tiled_smem_raw);
    const int N = a.N;
    const unsigned V = (unsigned)N * (unsigned)N;
    const int tiles = tiles_per_side * tiles_per_side;
    const int tid = threadIdx.x;

    const uint32_t K = (uint32_t)(2 * a.interval_n + 1);
    const int W = a.W, mWI = -a.W * a.interval_n;
    const float cIn = fc.c * (float)a.interval_n;
    const unsigned long long gs = a.sweep0 + (unsigned long long)sweep;
    // this thread's work items of both passes: fetched now, needed after the region has been loaded
    uint32_t my_items[2][kItemsPerPass / 128];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int u = 0; u < kItemsPerPass / 128; ++u) my_items[c][u] = g_tiled_items[c][tid + 128 * u];

    {
        const long long t = blockIdx.x;                             // one tile per CTA: CTA scheduling staggers the loads
        const long long chain = t / tiles;
        const int tile = (int)(t - chain * tiles);
        const int a0 = (tile / tiles_per_side) * kTile, a1 = (tile % tiles_per_side) * kTile;
        const double* gphi = phi_in + chain * V;
        const int32_t* gn0 = n_in + chain * 2 * V;
        const int32_t* gn1 = gn0 + V;
        // wrapped global coordinates of the region's rows and columns, for the colour passes (first read behind two barriers)
        if (tid < kRegRows) {
            int x0 = a0 - 2 + tid;  x0 += (x0 < 0) ? N : 0;  x0 -= (x0 >= N) ? N : 0;
            sh.x0row[tid] = x0;
        } else if (tid >= 64 && tid < 64 + kRegCols) {
            int x1 = a1 - 2 + (tid - 64);  x1 += (x1 < 0) ? N : 0;  x1 -= (x1 >= N) ? N : 0;
            sh.x1col[tid - 64] = x1;
        }

        // ---- load the region, two sites at a time (N and the origins are even: a pair never straddles the wrap); all of a
        //      thread's loads are issued before its first shared-memory store ----
        {
            constexpr int kPairs = kRegRows * (kRegCols / 2), kIter = (kPairs + 127) / 128;
            double2 vp[kIter];
            int2 v0[kIter], v1[kIter];
#pragma unroll
            for (int u = 0; u < kIter; ++u) {
                const int p = tid + 128 * u;
                if (p < kPairs) {
                    const int i = p / (kRegCols / 2), jj = 2 * (p - i * (kRegCols / 2));
                    int x0 = a0 - 2 + i;  x0 += (x0 < 0) ? N : 0;  x0 -= (x0 >= N) ? N : 0;
                    int x1 = a1 - 2 + jj; x1 += (x1 < 0) ? N : 0;  x1 -= (x1 >= N) ? N : 0;
                    const unsigned g = (unsigned)x0 * (unsigned)N + (unsigned)x1;
                    vp[u] = *reinterpret_cast<const double2*>(gphi + g);
                    v0[u] = *reinterpret_cast<const int2*>(gn0 + g);
                    v1[u] = *reinterpret_cast<const int2*>(gn1 + g);
                }
            }
#pragma unroll
            for (int u = 0; u < kIter; ++u) {
                const int p = tid + 128 * u;
                if (p < kPairs) {
                    *reinterpret_cast<double2*>(sh.phi + 2 * p) = vp[u];        // 2 p = i * kRegCols + jj (kRegCols is even)
                    *reinterpret_cast<int2*>(sh.n0 + 2 * p) = v0[u];
                    *reinterpret_cast<int2*>(sh.n1 + 2 * p) = v1[u];
                }
            }
        }
        __syncthreads();
        // ---- r = d(phi) - 2 pi n   (neighborhood.py:91) for the links inside the region, fp64 rounded to fp32, two sites
        //      per step (rows 0..35; the last pair of a row produces a link past the region that nothing reads) ----
        for (int p = tid; p < (kRegRows - 1) * (kRegCols / 2); p += 128) {
            const int l = 2 * p;
            const double2 pc = *reinterpret_cast<const double2*>(sh.phi + l);
            const double2 pu = *reinterpret_cast<const double2*>(sh.phi + l + kRegCols);
            const double pr = sh.phi[l + 2];
            const int2 b0 = *reinterpret_cast<const int2*>(sh.n0 + l), b1 = *reinterpret_cast<const int2*>(sh.n1 + l);
            float2 o0, o1;
            o0.x = (float)fma(-SVB_TWO_PI, int_to_double(b0.x), pu.x - pc.x);
            o0.y = (float)fma(-SVB_TWO_PI, int_to_double(b0.y), pu.y - pc.y);
            o1.x = (float)fma(-SVB_TWO_PI, int_to_double(b1.x), pc.y - pc.x);
            o1.y = (float)fma(-SVB_TWO_PI, int_to_double(b1.y), pr - pc.y);
            *reinterpret_cast<float2*>(sh.r0 + l) = o0;
            *reinterpret_cast<float2*>(sh.r1 + l) = o1;
        }
        __syncthreads();

        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double half_kappa = kappa / 2;
        const float hk2 = (float)(half_kappa * 1.4426950408889634);
        const float hkA = 1.0001f * hk2 * fc.bA, hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
        const unsigned long long gc = a.chain0 + (unsigned long long)chain;
        float n_acc = 0.0f, sum_A = 0.0f;

#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll 1
            for (int u = 0; u < kItemsPerPass / 128; ++u) {
                const uint32_t item = (u == 0) ? my_items[c][0] : (u == 1) ? my_items[c][1] : my_items[c][2];
                const uint32_t lA = item & 0xFFFFu, lB = item >> 16;
                if (lA == kNoSite) continue;
                const int iA = (int)lA / kRegCols, j = (int)lA - iA * kRegCols;
                const int x0 = sh.x0row[iA], x1 = sh.x1col[j];
                const uint32_t c0 = villain_pair_counter(x0, x1, N);
                const Philox4 bits = philox_site_keys(a, gc, gs, c0);
                const uint32_t halfA = villain_pair_half(x0);                 // 0 for the first of a pair
                bool ok;
                const float AA = villain_tiled_site(sh, a, fc, (int)lA, halfA ? bits.z : bits.x, halfA ? bits.w : bits.y, c0, halfA,
                                                    half_kappa, hk2, hkA, hkB, K, W, mWI, cIn, gc, gs, ok);
                // counters: owned sites only (local [2, 34)^2)
                if (iA >= 2 && iA < 2 + kTile && j >= 2 && j < 2 + kTile) { n_acc += ok ? 1.0f : 0.0f; sum_A += AA; }
                if (lB != kNoSite) {
                    const float AB = villain_tiled_site(sh, a, fc, (int)lB, bits.z, bits.w, c0, 1u, half_kappa, hk2, hkA, hkB, K, W, mWI,
                                                        cIn, gc, gs, ok);
                    if (iA + 8 < 2 + kTile && j >= 2 && j < 2 + kTile) { n_acc += ok ? 1.0f : 0.0f; sum_A += AB; }
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
            const unsigned g = (unsigned)(a0 + i) * (unsigned)N + (unsigned)(a1 + jj);
            const int l = (i + 2) * kRegCols + (jj + 2);
            *reinterpret_cast<double2*>(ophi + g) = *reinterpret_cast<const double2*>(sh.phi + l);
            *reinterpret_cast<int2*>(on0 + g) = *reinterpret_cast<const int2*>(sh.n0 + l);
            *reinterpret_cast<int2*>(on1 + g) = *reinterpret_cast<const int2*>(sh.n1 + l);
        }
        if (a.obs) {
            double sred[2] = {(double)n_acc, (double)sum_A};
            block_sum<2>(sred, sh.red);
            if (tid == 0) {
                atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, sred[0]);
                atomicAdd(a.obs + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, sred[1]);
            }
            if (fuse_obs) {
                __syncthreads();
                ChainSums cs;
                cs.action = 0.0; cs.sumA = 0.0; cs.dn2 = 0; cs.w0 = 0; cs.w1 = 0; cs.accepted = 0;
                for (int p = tid; p < kTile * kTile; p += 128) {
                    const int i = 2 + p / kTile, j = 2 + (p % kTile);
                    villain_obs_site<double, 0>(sh.phi, sh.n0, sh.n1, kRegCols, i, j, cs.action, cs.dn2, cs.w0, cs.w1);
                }
                double* scratch = reinterpret_cast<double*>(sh.r0);        // the residuals are no longer needed: 6 * 32 doubles
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
}

static int launch_villain_tiled_filtered(const VillainArgs& a, const FilterConsts& fc, const double* phi_in, const int32_t* n_in,
                                         double* phi_out, int32_t* n_out, int sweep, int tps, int fuse, long long blocks,
                                         cudaStream_t st) {
    static bool ready[64];
    const size_t smem = sizeof(TiledShared);
    int dev = 0;
    SVB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !ready[dev]) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(villain_tiled_filtered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(villain_tiled_filtered_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        villain_tiled_items_kernel<<<1, 128, 0, st>>>();           // idempotent; ordered before the sweep on this stream
        SVB_CUDA_TRY(cudaGetLastError());
        SVB_CUDA_TRY(cudaStreamSynchronize(st));                    // once per process and device: other streams may follow
        if (dev >= 0 && dev < 64) ready[dev] = true;
    }
    villain_tiled_filtered_kernel<<<(unsigned)blocks, 128, smem, st>>>(a, fc, phi_in, n_in, phi_out, n_out, sweep, tps, fuse);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// svb_villain_observables for N in {16, 32, 64}, fp64 phi: one CTA per chain at a time, the chain staged by a 1-D TMA bulk
// copy, one conflict-free fp64 pass over site pairs (the pass of the sweep kernel above), records without atomics.
// HBM-bound: one read of the state.
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(4 * NT) villain_obs_smem_kernel(const double* __restrict__ phi, const int32_t* __restrict__ n,
                                                                  long long chains, double kappa_scalar,
                                                                  const double* __restrict__ kappa_chain, double* __restrict__ obs,
                                                                  int keep_counters) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, T = 4 * NT, NW = T / 32, PER = V / 2 / T;
    constexpr uint32_t bytes_phi = V * sizeof(double), bytes_n = 2 * V * sizeof(int32_t);
    double* sphi = reinterpret_cast<double*>(smem_raw);
    int32_t* sn0 = reinterpret_cast<int32_t*>(smem_raw + bytes_phi);
    int32_t* sn1 = sn0 + V;
    double* red_state = reinterpret_cast<double*>(smem_raw + bytes_phi + bytes_n);
    uint64_t* bar = reinterpret_cast<uint64_t*>(red_state + 4 * NW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue_load = [&](long long chain) {
        mbar_expect_tx(bar, bytes_phi + bytes_n);
        bulk_g2s(sphi, phi + chain * V, bytes_phi, bar);
        bulk_g2s(sn0, n + chain * 2 * V, bytes_n, bar);
    };
    const int row8 = tid / HN, k = tid - row8 * HN;
    const int up_off = (row8 == 7) ? (N - V) : N;
    const double* pp = sphi + row8 * N + 2 * k;
    const double* pp_r = sphi + row8 * N + ((2 * k + 2) & (N - 1));
    const int32_t* pn0 = sn0 + row8 * N + 2 * k;
    const int32_t* pn1 = sn1 + row8 * N + 2 * k;
    long long chain = blockIdx.x;
    if (tid == 0 && chain < chains) issue_load(chain);
    for (int it = 0; chain < chains; chain += gridDim.x, ++it) {
        mbar_wait(bar, (uint32_t)(it & 1));
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
        chain_partials<true, false>(red_state, nullptr, lane, warp, action, dn2, w0, w1, 0.0, 0);
        __syncthreads();                       // every warp has read the chain; the slots are written
        const long long next = chain + gridDim.x;
        if (tid == 0 && next < chains) issue_load(next);
        if (tid == 32) {
            const double kappa = kappa_chain ? kappa_chain[chain] : kappa_scalar;
            double* row = obs + chain * SVB_VOBS_COUNT;
            chain_finish<NW, true, false>(red_state, nullptr, kappa / 2, row, nullptr);
            if (!keep_counters) { row[SVB_VOBS_ACCEPTED] = 0.0; row[SVB_VOBS_ACCEPTANCE] = 0.0; }
        }
        __syncthreads();                       // the slots are free again (the next chain's partials follow its load)
    }
}

template <int NT>
static int launch_villain_obs_smem(const double* phi, const int32_t* n, long long chains, double kappa, const double* kappa_chain,
                                   double* obs, int keep_counters, cudaStream_t stream) {
    auto kern = villain_obs_smem_kernel<NT>;
    const size_t smem = (size_t)NT * NT * 16 + 4 * (4 * NT / 32) * sizeof(double) + 16;
    static int grid_cap = 0;
    if (grid_cap == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        SVB_CUDA_TRY(cudaGetDevice(&dev));
        SVB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 4 * NT, smem));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "villain observable kernel does not fit an SM at N=%d", NT);
        grid_cap = per_sm * sms;
    }
    const long long grid = chains < grid_cap ? chains : grid_cap;
    kern<<<(unsigned)grid, 4 * NT, smem, stream>>>(phi, n, chains, kappa, kappa_chain, obs, keep_counters);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}
