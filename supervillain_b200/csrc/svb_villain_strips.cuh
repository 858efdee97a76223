// svb_villain_strips.cuh -- the fp32-filtered Villain sweep for L = 128 (config 4) with ONE CHAIN PER CTA
// (included by svb_villain.cu inside namespace svb, after svb_villain_filtered.cuh).
//
// NeighborhoodUpdate.step (supervillain/generator/villain/neighborhood.py:59-137), Philox draws, FAST arithmetic, one sweep
// per launch with the record protocol of the production kernels (SPARSE of villain_smem_filtered_kernel: accepted proposals go
// to global memory as reductions, the exact test reads global memory, nothing is stored back).
//
// In a sparse sweep phi and n are needed in shared memory for ONE thing only: building the fp32 residuals.  So a 128 x 128
// chain -- 256 KiB of phi and n, which is why it was spread over a cluster of four CTAs (svb_villain_cluster.cuh) -- does fit
// one SM after all: the 128 KiB of residuals stay resident, and phi and n stream through a ring of four 8-row strips
// (17.5 KiB each: 9 rows of phi, 8 of n0, 9 of n1, by 1-D TMA bulk copies) while the residuals are built.  What that buys:
// no cluster barriers (the top stall of the cluster kernel) and no distributed shared memory, the arithmetic and the thread
// geometry of villain_smem_filtered_kernel unchanged (a thread owns rows r, r + 8, ... of one column slot; pairs of rows share a
// Philox block), block barriers only.  The strips of the NEXT chain that fit the ring land during the colour passes.
//
// Measured (config-4 shard of bench.py, 8192 chains, one sweep + record per launch, overlapped launches): 892 - 911 us against
// 1283 - 1347 us for the cluster kernel (0.72 - 0.74 of the HBM roofline against 0.49 - 0.51); without a record 847 us (0.78).
// 143 thread-instructions per site-update (the cluster kernel: 227), issue slots 47 % busy with 16 warps per SM; a tenth of the
// instructions are warps spinning on a strip's mbarrier: the ring keeps 70 KiB in flight per SM, and the build -- memory --
// and the passes -- arithmetic -- of a chain do not overlap.  Tried on top: 1024 threads per CTA at 64 registers (HALVES = 2:
// 1079 against 1048 us), an L2 prefetch of the rest of the next chain during the passes (975 against 961 us).  What the knobs
// below are worth: all eight pairs of a pass in flight per thread 1048 -> 1029 us; the build unrolled 4, 8, 16 strips deep 1008 /
// 961, 931, 911 us.
#pragma once

#ifndef SVB_STRIPS_RING
#define SVB_STRIPS_RING 4
#endif
constexpr int kStripRows = 8, kStripRing = SVB_STRIPS_RING;
#ifndef SVB_STRIPS_HALVES
#define SVB_STRIPS_HALVES 1
#endif
#ifndef SVB_STRIPS_UNROLL_P
#define SVB_STRIPS_UNROLL_P 8      /* pairs of rows of a colour pass in flight per thread */
#endif
#ifndef SVB_STRIPS_UNROLL_B
#define SVB_STRIPS_UNROLL_B 16      /* strips of the residual build in flight per thread */
#endif
constexpr int kStripsUnrollP = SVB_STRIPS_UNROLL_P, kStripsUnrollB = SVB_STRIPS_UNROLL_B;

// HALVES = 2: twice the threads -- the strips with an odd place in the load order and the upper half of the pairs of rows belong
// to threads 4 N .. 8 N - 1 -- for 32 warps per SM at 64 registers instead of 16 at up to 128.
template <int NT, bool OVERLAP, bool UNIT, int HALVES>
__global__ void __launch_bounds__(4 * NT * HALVES, 1) villain_strips_kernel(const __grid_constant__ VillainArgs a, const __grid_constant__ FilterConsts fc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, VH = V / 2, TQ = 4 * NT, T = TQ * HALVES, NW = T / 32;
    constexpr int PER = VH / TQ;                                 // rows row8 + 8 q of a column slot = strips per chain
    static_assert(PER == N / kStripRows && kStripRing <= PER && PER % (2 * HALVES) == 0 && (HALVES == 1 || HALVES == 2),
                  "villain_strips_kernel: unsupported geometry");
    constexpr uint32_t bytes_phi = (kStripRows + 1) * N * sizeof(double);        // rows 8 q .. 8 q + 8
    constexpr uint32_t bytes_n0 = kStripRows * N * sizeof(int32_t);              // rows 8 q .. 8 q + 7
    constexpr uint32_t bytes_n1 = (kStripRows + 1) * N * sizeof(int32_t);        // rows 8 q .. 8 q + 8
    constexpr uint32_t strip_bytes = bytes_phi + bytes_n0 + bytes_n1;
    static_assert(strip_bytes % 128 == 0, "strips are 128-byte aligned");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tq = tid % TQ, half = tid / TQ;
    float* rc0 = reinterpret_cast<float*>(smem_raw + kStripRing * strip_bytes);   // [colour][VH]: residual of link (0, x)
    float* rc1 = rc0 + V;                                                         // [colour][VH]: residual of link (1, x)
    double* red_state = reinterpret_cast<double*>(rc1 + V);                       // [NW][4] per-warp partial sums
    double* red_count = red_state + 4 * NW;                                       // [NW][2]
    uint64_t* full = reinterpret_cast<uint64_t*>(red_count + 2 * NW);             // [ring]: the strip has landed
    uint64_t* empty = full + kStripRing;                                          // [ring]: every warp has read it
    float4* dn_lut = reinterpret_cast<float4*>(empty + kStripRing);
    constexpr int kWriter = 32;

    if (tid == 0) {
        for (int b = 0; b < kStripRing; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&empty[b], NW / HALVES);
        }
        fence_mbar_init();
    }
    if (UNIT)
        for (int i = tid; i < 81; i += T)
            dn_lut[i] = make_float4(-fc.c * (float)(i / 27), -fc.c * (float)((i / 9) % 3), -fc.c * (float)((i / 3) % 3), -fc.c * (float)(i % 3));
    if (OVERLAP) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (a.grid_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    __syncthreads();
    const bool obs_of_input = a.obs_in != nullptr;
    const int interval_n = UNIT ? 1 : a.interval_n;
    const uint32_t K = (uint32_t)(2 * interval_n + 1);
    const int W = UNIT ? 1 : a.W, mWI = -W * interval_n;
    const float cIn = fc.c * (float)interval_n;
    const float2 cIn2 = make_float2(cIn, cIn), negc2 = make_float2(-fc.c, -fc.c), two2 = make_float2(2.0f, 2.0f);
    const double two_I_scaled = (2.0 * a.interval_phi) * 2.3283064365386963e-10;          // (2 I) 2^-32, exact scaling

    // per-thread geometry: rows row8 + 8 q of the compact column k
    const int row8 = tq / HN, k = tq - row8 * HN;
    const int cc = row8 & 1;
    const int wrap0 = (row8 == 0) ? VH : 0;                       // backward-0 neighbour of row 0 is row N - 1

    auto peek_epoch = [&](long long chain) -> uint32_t {
        uint32_t e = a.wait_epoch;
        if (OVERLAP && !a.grid_wait)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(a.epochs + chain) : "memory");
        return e;
    };
    // (thread 0) a chain may be read once its epoch says that the launch before this one has finished with it
    auto wait_chain = [&](long long chain, uint32_t seen) {
        if (OVERLAP && !a.grid_wait) {
            uint32_t e = seen;
            unsigned ns = 32, naps = 0;
            while (true) {
                if (e == a.wait_epoch) break;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(a.epochs + chain) : "memory");
                if (e == a.wait_epoch) break;
                __nanosleep(ns);
                if (ns < 1024) ns *= 2;
                if (++naps > (1u << 21)) __trap();          // > 2 s: a producer that never comes is a caller error
            }
            asm volatile("fence.proxy.async;" ::: "memory");
        }
    };
    // strips are loaded in the order s = 0, 1, ...: strip s holds rows 8 q(s) .. 8 q(s) + 8 and is read by the threads of half
    // s % HALVES
    auto strip_q = [&](int s_) { return HALVES == 1 ? s_ : (s_ >> 1) + (PER / 2) * (s_ & 1); };
    // (one thread) the loads of strip s of `chain` into ring slot s % ring (the row after the last one is row 0)
    auto issue_strip = [&](long long chain, int s_, int slot) {
        const int q = strip_q(s_);
        const int b = slot, r0 = kStripRows * q;
        unsigned char* st = smem_raw + (size_t)b * strip_bytes;
        const double* gp = reinterpret_cast<const double*>(a.phi) + chain * V;
        const int32_t* g0 = a.n + chain * 2 * V;
        const int32_t* g1 = g0 + V;
        mbar_expect_tx(&full[b], strip_bytes);
        if (r0 + kStripRows < N) {
            bulk_g2s(st, gp + r0 * N, bytes_phi, &full[b]);
            bulk_g2s(st + bytes_phi + bytes_n0, g1 + r0 * N, bytes_n1, &full[b]);
        } else {
            bulk_g2s(st, gp + r0 * N, bytes_phi - N * sizeof(double), &full[b]);
            bulk_g2s(st + bytes_phi - N * sizeof(double), gp, N * sizeof(double), &full[b]);
            bulk_g2s(st + bytes_phi + bytes_n0, g1 + r0 * N, bytes_n1 - N * sizeof(int32_t), &full[b]);
            bulk_g2s(st + bytes_phi + bytes_n0 + bytes_n1 - N * sizeof(int32_t), g1, N * sizeof(int32_t), &full[b]);
        }
        bulk_g2s(st + bytes_phi, g0 + r0 * N, bytes_n0, &full[b]);
    };

    long long chain = blockIdx.x;
    if (tid == 0 && chain < a.chains) {
        wait_chain(chain, peek_epoch(chain));
        for (int s_ = 0; s_ < kStripRing; ++s_) issue_strip(chain, s_, s_);
    }

    int it = 0;
    for (; chain < a.chains; chain += gridDim.x, ++it) {
        const long long next = chain + gridDim.x;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double half_kappa = kappa / 2;
        const float hk2 = (float)(half_kappa * 1.4426950408889634);                 // decisions are taken in units of ln 2
        const float hkA = 1.0001f * hk2 * fc.bA, hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
        const float2 hk22 = make_float2(hk2, hk2), hkA2 = make_float2(hkA, hkA), hkB2 = make_float2(hkB, hkB);
        double* gphi = reinterpret_cast<double*>(a.phi) + chain * V;
        int32_t* gn0 = a.n + chain * 2 * V;
        int32_t* gn1 = gn0 + V;
        uint32_t seen_next = 0;
        if (tq == 0 && next < a.chains) seen_next = peek_epoch(next);                // lands during the residual build

        int n_acc = 0;
        float sum_A = 0.0f;
        // ---- r = d(phi) - 2 pi n   (neighborhood.py:91) in fp64, stored rounded to fp32, strip by strip ----
        {
            float* w0e = rc0 + cc * VH + tq;
            float* w1e = rc1 + cc * VH + tq;
            float* w0o = rc0 + (cc ^ 1) * VH + tq;
            float* w1o = rc1 + (cc ^ 1) * VH + tq;
            double action = 0.0;
            int w0 = 0, w1 = 0;
            long long dn2 = 0;
#pragma unroll kStripsUnrollB
            for (int s_ = half; s_ < PER; s_ += HALVES) {
                const int q = strip_q(s_);
                // the g-th strip this CTA reads lives in slot g % ring, in the slot's (g / ring)-th use
                const unsigned g = (unsigned)it * PER + (unsigned)s_;
                const int b = (int)(g % kStripRing);
                const uint32_t parity = (g / kStripRing) & 1u;
                const unsigned char* st = smem_raw + (size_t)b * strip_bytes;
                const double* sp = reinterpret_cast<const double*>(st) + row8 * N;
                const int32_t* s0 = reinterpret_cast<const int32_t*>(st + bytes_phi) + row8 * N;
                const int32_t* s1 = reinterpret_cast<const int32_t*>(st + bytes_phi + bytes_n0) + row8 * N;
                mbar_wait(&full[b], parity);
                const PairResiduals pr = villain_pair_residuals(sp + 2 * k, sp + N + 2 * k, sp + ((2 * k + 2) & (N - 1)), s0 + 2 * k, s1 + 2 * k);
                w0e[TQ * q] = (float)pr.r0e;
                w1e[TQ * q] = (float)pr.r1e;
                w0o[TQ * q] = (float)pr.r0o;
                w1o[TQ * q] = (float)pr.r1o;
                if (obs_of_input) {                                  // the observables of the arriving state ride along
                    action = fma(pr.r0e, pr.r0e, action);
                    action = fma(pr.r0o, pr.r0o, action);
                    action = fma(pr.r1e, pr.r1e, action);
                    action = fma(pr.r1o, pr.r1o, action);
                    const int hr = s0[(2 * k + 2) & (N - 1)];                                    // n0[x + 2 e1]
                    const int2 up = *reinterpret_cast<const int2*>(s1 + N + 2 * k);              // n1[x + e0]
                    const int d0 = (up.x - pr.a1.x) - (pr.a0.y - pr.a0.x), d1 = (up.y - pr.a1.y) - (hr - pr.a0.y);
                    dn2 += (long long)d0 * d0 + (long long)d1 * d1;
                    w0 += pr.a0.x + pr.a0.y;
                    w1 += pr.a1.x + pr.a1.y;
                }
                // the warp has read the strip; when every warp of the half has, its first thread refills the slot -- with a later
                // strip of this chain or, towards the end, with the first strips of the next one (they land during the colour passes)
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[b]);
                if (tq == 0) {
                    mbar_wait(&empty[b], parity);
                    const int sn = s_ + kStripRing;
                    if (sn < PER) {
                        issue_strip(chain, sn, b);
                    } else if (next < a.chains) {
                        if (sn < PER + HALVES) wait_chain(next, seen_next);          // the first strip of the next chain this thread loads
                        issue_strip(next, sn - PER, b);
                    }
                }
            }
            if (obs_of_input) chain_partials<true, false>(red_state, red_count, lane, warp, action, dn2, w0, w1, 0.0, 0);
        }
        __syncthreads();
        if (obs_of_input && tid == kWriter)
            chain_finish<NW, true, false>(red_state, red_count, kappa / 2, a.obs_in + chain * SVB_VOBS_COUNT, nullptr);

        const unsigned long long gc = a.chain0 + (unsigned long long)chain, gs = a.sweep0;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            const int par = (row8 + c) & 1;                    // column parity of this thread's sites of colour c
            const int x1 = 2 * k + par;
            const int ob1 = par ? 0 : ((k == 0) ? (1 - HN) : 1);          // compact index of x - e1 is j - ob1
            float* R0own = rc0 + c * VH + tq;
            float* R1own = rc1 + c * VH + tq;
            float* R0b = rc0 + (c ^ 1) * VH + tq - HN;         // backward link (0, x - e0): row above, same compact column
            float* R0b_q0 = R0b + wrap0;
            float* R1b = rc1 + (c ^ 1) * VH + tq - ob1;        // backward link (1, x - e1)
#pragma unroll kStripsUnrollP
            for (int pl = 0; pl < PER / 2 / HALVES; ++pl) {
                const int p = pl + half * (PER / 2 / HALVES);      // the pair of rows (row8 + 16 p, row8 + 16 p + 8)
                const uint32_t c0 = (uint32_t)((row8 + 16 * p) * N + x1);                 // villain_pair_counter
                const Philox4 bits = philox_site_keys(a, gc, gs, c0);
                const int qA = 2 * p, qB = 2 * p + 1;
                float* r0bA = (p == 0) ? R0b_q0 : R0b;             // q == 0 is the only row whose e0-neighbour wraps
                // proposals: four base-K digits each, then the leading 32 bits of the uniform
                uint32_t fA = bits.y, fB = bits.w;
                uint32_t codeA = 0, codeB = 0;
                int digA[4], digB[4];
                if (UNIT) {
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
                const float2 r_f0 = make_float2(R0own[TQ * qA], R0own[TQ * qB]), r_f1 = make_float2(R1own[TQ * qA], R1own[TQ * qB]);
                const float2 r_b0 = make_float2(r0bA[TQ * qA], R0b[TQ * qB]), r_b1 = make_float2(R1b[TQ * qA], R1b[TQ * qB]);
                // dr = d(dphi) - 2 pi dn   (neighborhood.py:110), dn = W (digit - interval_n)
                float2 dr_f0, dr_b0, dr_f1, dr_b1;
                if (UNIT) {
                    const float4 tA = dn_lut[codeA], tB = dn_lut[codeB];
                    dr_f0 = __fadd2_rn(base_f, make_float2(tA.x, tB.x));
                    dr_b0 = __fadd2_rn(base_b, make_float2(tA.y, tB.y));
                    dr_f1 = __fadd2_rn(base_f, make_float2(tA.z, tB.z));
                    dr_b1 = __fadd2_rn(base_b, make_float2(tA.w, tB.w));
                } else {
                    dr_f0 = __ffma2_rn(negc2, make_float2((float)digA[0], (float)digB[0]), base_f);
                    dr_b0 = __ffma2_rn(negc2, make_float2((float)digA[1], (float)digB[1]), base_b);
                    dr_f1 = __ffma2_rn(negc2, make_float2((float)digA[2], (float)digB[2]), base_f);
                    dr_b1 = __ffma2_rn(negc2, make_float2((float)digA[3], (float)digB[3]), base_b);
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
                            // the exact test reads the current phi and n from global memory (L2: what this CTA's earlier
                            // reductions made of them is visible behind the block barrier between the passes; never an L1 line)
                            ExactProposalPtr ep;
                            const int ic = x0 * N + x1, ib0 = ((x0 - 1) & (N - 1)) * N + x1, ib1 = x0 * N + ((x1 - 1) & (N - 1));
                            const int if0 = ((x0 + 1) & (N - 1)) * N + x1, if1 = x0 * N + ((x1 + 1) & (N - 1));
                            // (ExactProposalPtr dereferences generic pointers: copy the nine values out of L2 first)
                            const double v_c = __ldcg(gphi + ic), v_f0 = __ldcg(gphi + if0), v_b0 = __ldcg(gphi + ib0);
                            const double v_f1 = __ldcg(gphi + if1), v_b1 = __ldcg(gphi + ib1);
                            const int32_t m_f0 = __ldcg(gn0 + ic), m_b0 = __ldcg(gn0 + ib0), m_f1 = __ldcg(gn1 + ic), m_b1 = __ldcg(gn1 + ib1);
                            ep.p_c = &v_c; ep.p_f0 = &v_f0; ep.p_b0 = &v_b0; ep.p_f1 = &v_f1; ep.p_b1 = &v_b1;
                            ep.n_f0 = &m_f0; ep.n_b0 = &m_b0; ep.n_f1 = &m_f1; ep.n_b1 = &m_b1;
                            ep.half_kappa = half_kappa;
                            ep.c = SVB_TWO_PI * (double)W;
                            ep.dphi = villain_dphi_from_word(wA, a.interval_phi);
#pragma unroll
                            for (int i = 0; i < 4; ++i) ep.g[i] = dig[i] - interval_n;
                            ep.d.f = f; ep.d.c0 = c0; ep.d.half = (uint32_t)h;
                            ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs; ep.rc.stream = a.refine_stream; ep.rc.wide = 0;
                            ok = villain_exact_decision_ptr(ep);
                        }
                        n_acc += ok ? 1 : 0;
                        if (ok) {                                               // (:121-129)
                            // straight to global memory, nothing waits for it: the fp64 reduction rounds once, to nearest, like
                            // the reference's phi + dphi
                            const double Ah = __hiloint2double(0x43300000, (int)wA) - 4503599627370495.5;
                            const int ic = x0 * N + x1;
                            atomicAdd(gphi + ic, __dadd_rn(-a.interval_phi, __dmul_rn(two_I_scaled, Ah)));
                            atomicAdd(gn0 + ic, W * dig[0] + mWI);
                            atomicAdd(gn0 + ((x0 - 1) & (N - 1)) * N + x1, W * dig[1] + mWI);
                            atomicAdd(gn1 + ic, W * dig[2] + mWI);
                            atomicAdd(gn1 + x0 * N + ((x1 - 1) & (N - 1)), W * dig[3] + mWI);
                            R0own[TQ * q] = h ? n_f0.y : n_f0.x;
                            r0b[TQ * q] = h ? n_b0.y : n_b0.x;
                            R1own[TQ * q] = h ? n_f1.y : n_f1.x;
                            R1b[TQ * q] = h ? n_b1.y : n_b1.x;
                        }
                    }
                }
            }
            // this launch's counters ride on the barrier that ends the last pass
            if (c == 1 && a.obs != nullptr) chain_partials<false, true>(red_state, red_count, lane, warp, 0.0, 0, 0, 0, (double)sum_A, n_acc);
            __syncthreads();
        }
        if (tid == kWriter && a.obs != nullptr)
            chain_finish<NW, false, true>(red_state, red_count, kappa / 2, nullptr, a.obs + chain * SVB_VOBS_COUNT);
    }
    if (OVERLAP) {
        __syncthreads();                                   // every reduction is issued, every record is written
        if (warp == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (int i = lane; i < it; i += 32)
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.epochs + blockIdx.x + (long long)i * gridDim.x), "r"(a.signal_epoch)
                             : "memory");
        }
    }
}

// one sweep, no record of the state after it: the launches the strips kernel serves
static bool villain_strips_serves(const VillainArgs& a) {
    const char* e = getenv("SVB_VILLAIN_KERNEL128");              // "cluster": the cluster kernel, for an A/B
    if (e && e[0] == 'c') return false;
    const char* env_sparse = getenv("SVB_VILLAIN_SPARSE");
    return a.N == 128 && !a.exact_mode && !a.filtered_strict && !a.accept_mask && !a.dS_out && !a.wide && a.n_sweeps == 1 &&
           (a.obs == nullptr || a.obs_in != nullptr) && !(env_sparse && env_sparse[0] == '0');
}

static int launch_villain_strips(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    constexpr int NT = 128;
    const bool overlap = a.epochs != nullptr, unit = a.W == 1 && a.interval_n == 1;
    int halves = SVB_STRIPS_HALVES;
    if (const char* e = getenv("SVB_STRIPS_HALVES")) halves = atoi(e) == 2 ? 2 : 1;
    auto kern1 = overlap ? (unit ? villain_strips_kernel<NT, true, true, 1> : villain_strips_kernel<NT, true, false, 1>)
                         : (unit ? villain_strips_kernel<NT, false, true, 1> : villain_strips_kernel<NT, false, false, 1>);
    auto kern2 = overlap ? (unit ? villain_strips_kernel<NT, true, true, 2> : villain_strips_kernel<NT, true, false, 2>)
                         : (unit ? villain_strips_kernel<NT, false, true, 2> : villain_strips_kernel<NT, false, false, 2>);
    auto kern = halves == 2 ? kern2 : kern1;
    const int threads = 4 * NT * halves;
    constexpr size_t strip = (size_t)(kStripRows + 1) * NT * 8 + (size_t)kStripRows * NT * 4 + (size_t)(kStripRows + 1) * NT * 4;
    const size_t smem = kStripRing * strip + 2 * (size_t)NT * NT * sizeof(float) + 6 * (threads / 32) * sizeof(double) + 2 * kStripRing * 8 +
                        81 * sizeof(float4);
    static int ready[8][64];
    const int variant = (overlap ? 1 : 0) + (unit ? 2 : 0) + (halves == 2 ? 4 : 0);
    if (info.device >= 64 || !ready[variant][info.device]) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int per_sm = 0;
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "the strips kernel does not fit an SM");
        if (info.device < 64) ready[variant][info.device] = 1;
    }
    long long grid = info.sm_count;
    if (grid > a.chains) grid = a.chains;
    const FilterConsts fc = make_filter_consts(a.interval_phi, a.W, a.interval_n);
    if (overlap) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SVB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a, fc));
        return 0;
    }
    kern<<<(unsigned)grid, threads, smem, stream>>>(a, fc);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}
