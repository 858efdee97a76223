// svb_worldline_table.cuh -- the production worldline sweep kernel (included by svb_worldline.cu, inside namespace svb).
//
// PlaquetteUpdate's move (supervillain/generator/worldline/plaquette.py:79-101) in the red/black order of
// VortexUpdate / CoexactUpdate (worldline/vortex.py:86-128), for W = 1, Philox draws and N in {16, 32, 64}:
//
//  * With W = 1 every f = m - delta v (plaquette.py:53) is an integer, and so is everything dS is made of:
//      delta_f = dm - dv in {-2..2},   s = f1 + f2 - f3 - f4 + 2 delta_f,   dS = (delta_f / kappa) s   (plaquette.py:84-85).
//    The acceptance probability takes a few hundred distinct values per chain; they are TABULATED per (delta_f, s) -- each
//    entry computed in fp64 exactly as the general path computes it, fl(fl(delta_f (1/kappa)) s) then the fp64 exponential --
//    as 32-bit integer thresholds.  A proposal is decided by ONE integer comparison of the leading 32 bits of its uniform
//    with the threshold; if they are within one unit (probability 2^-31) or s is outside the table, the exact lazy test of
//    the general path decides.  No floating-point instruction is left on the hot path except the acceptance statistic.
//  * The m fields are transformed IN PLACE to f when a chain arrives and back to m = f + delta v when it leaves, so a
//    proposal reads four integers (v is touched only on acceptance) and the action, winding and wrapping sums are plain
//    integer sums over f (sum_x m_mu = sum_x f_mu on a torus).
//  * One Philox4x32-10 block serves the four plaquettes (x0 + {0, 8, 16, 24}, x1) a thread owns (draw mapping version 2).
//
// Geometry: T = 4 N threads per CTA, one chain at a time, grid-stride over chains, 1-D TMA bulk copies, three block
// barriers per chain (five with observables).
#pragma once

__device__ __forceinline__ float fast_ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct WlTableEntry {
    uint32_t thr;      // floor(A 2^32) for A < 1
    float A;           // the acceptance probability, for the generator's statistic
};
constexpr int kWlTableS = 32;                      // s in [-32, 31]
constexpr int kWlTableSize = 5 * 2 * kWlTableS;    // delta_f in [-2, 2]

// MODE: SVB_WL_JOINT (PlaquetteUpdate's move), SVB_WL_VORTEX (dv = a alone: delta_f = -a, vortex.py:108-112) or SVB_WL_COEXACT
// (dm = a alone: delta_f = +a, coexact.py:102-106), the latter two for interval <= 2.  Their reference formula sums four
// separately rounded link terms; the table holds fl(fl(delta_f / kappa) s), equal to a few ulp, so the integer comparison is
// trusted only at distance >= 2 from the threshold and the exact path evaluates the reference's own expression.
// OVERLAP: the launch takes part in the overlapped-launch protocol (svb_common.cuh, svb_worldline_sweep_overlapped).
// TM, STAGES: TM N threads per CTA and STAGES chains of shared memory.  (4, 1): a chain per CTA, the next one loaded when this
// one has been stored.  (8, 2) (N = 64: config 3): twice the threads on a chain -- a thread owns ONE Philox block of four
// plaquettes per colour -- and the next chain already in the other stage when this one is finished, so a CTA never waits for
// a load: two CTAs of 16 warps per SM instead of four of 8, the same warps and the same shared memory.
template <int MODE, int NT, int MINB, bool OVERLAP, int TM = 4, int STAGES = 1>
__global__ void __launch_bounds__(TM * NT, MINB) worldline_smem_table_kernel(const __grid_constant__ WorldlineArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, HN = N / 2, VH = V / 2, T = TM * NT, NW = T / 32;
    constexpr int GSTEP = TM / 4;                                // threads that share a column slot and a row8 split the Philox blocks
    constexpr int PER = VH / (4 * NT);                           // plaquettes per (row8, column slot) per colour (rows row8 + 8 q)
    constexpr int QUADS = (PER + 3) / 4, QW = PER < 4 ? PER : 4; // Philox blocks per (row8, column slot) per colour, words used of each
    static_assert(QUADS % GSTEP == 0, "every thread owns the same number of Philox blocks");
    constexpr uint32_t bytes_m = 2 * V * sizeof(int32_t);
    constexpr uint32_t bytes_v = V * sizeof(int32_t);
    constexpr uint32_t stage_bytes = bytes_m + bytes_v;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WlTableEntry* table = reinterpret_cast<WlTableEntry*>(smem_raw + STAGES * stage_bytes);
    long long* red = reinterpret_cast<long long*>(table + kWlTableSize);          // [NW][4] integer partial sums
    float* redA = reinterpret_cast<float*>(red + 4 * 32);                         // [NW]
    uint64_t* bar = reinterpret_cast<uint64_t*>(redA + 32);
    constexpr int kWriter = 32;

    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < STAGES; ++st) mbar_init(&bar[st], 1);
        fence_mbar_init();
    }
    if (OVERLAP) overlap_prologue(a.ov);
    __syncthreads();
    const bool want_obs = a.obs != nullptr;

    auto issue_load = [&](long long chain, uint32_t seen, int st) {
        if (OVERLAP) overlap_wait(a.ov, chain, seen);
        int32_t* dst = reinterpret_cast<int32_t*>(smem_raw + st * stage_bytes);
        mbar_expect_tx(&bar[st], stage_bytes);
        bulk_g2s(dst, a.m + chain * 2 * V, bytes_m, &bar[st]);
        bulk_g2s(dst + 2 * V, a.v + chain * V, bytes_v, &bar[st]);
    };
    // acceptance table of one coupling: entry (delta_f, s) from the same fp64 expressions as the general path
    auto build_table = [&](double kappa) {
        const double inv_kappa = __ddiv_rn(1.0, kappa);
        for (int i = tid; i < kWlTableSize; i += T) {
            const int df = i / (2 * kWlTableS) - 2, s = i % (2 * kWlTableS) - kWlTableS;
            const double dS = __dmul_rn(__dmul_rn((double)df, inv_kappa), (double)s);
            const double A = exp_clipped(-dS);
            WlTableEntry e;
            const double scaled = A * 4294967296.0;
            e.thr = (A >= 1.0) ? 0xFFFFFFFFu : (uint32_t)__double2ull_rd(scaled);
            e.A = (float)A;
            table[i] = e;
        }
    };

    // thread -> (row8, k): it owns the plaquettes (row8 + 8 q, 2 k + parity).  A warp takes 16 column slots of TWO adjacent rows
    // (their column parities are opposite): its 32 accesses to a row-major int32 array then fall on 32 different banks, where
    // 32 slots of ONE row (stride 2) collide pairwise -- a third of this kernel's shared-memory wavefronts were such replays.
    // (Any bijection gives the same chain: the draws are keyed by the site.)
    int row8, k;
    const int tq = tid % (4 * NT), g_first = tid / (4 * NT);     // (row8, column slot) and the first Philox block of this thread
    if (HN >= 16) {
        const int rest = tq >> 5;
        row8 = 2 * (rest / (HN / 16)) + ((tq >> 4) & 1);
        k = 16 * (rest % (HN / 16)) + (tq & 15);
    } else {
        row8 = tq / HN;
        k = tq - row8 * HN;
    }
    long long chain = blockIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < STAGES; ++st) {
            const long long c = chain + (long long)st * gridDim.x;
            if (c < a.chains) issue_load(c, OVERLAP ? overlap_peek(a.ov, c) : 0u, st);
        }
    }
    if (!a.kappa_chain) build_table(a.kappa);

    int it = 0;
    for (; chain < a.chains; chain += gridDim.x, ++it) {
        const int stage = it % STAGES;
        int32_t* F0 = reinterpret_cast<int32_t*>(smem_raw + stage * stage_bytes);     // m_0 on arrival, f_0 during the sweeps
        int32_t* F1 = F0 + V;
        int32_t* sv = F1 + V;
        const long long next = chain + (long long)STAGES * gridDim.x;                 // the chain that takes this stage next
        uint32_t seen_next = 0;
        if (OVERLAP && tid == 0 && next < a.chains) seen_next = overlap_peek(a.ov, next);      // lands during the sweep
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        if (a.kappa_chain) build_table(kappa);          // the previous chain's last reader is behind two barriers
        mbar_wait(&bar[stage], (uint32_t)((it / STAGES) & 1));

        // ---- m -> f = m - delta v, in place, four consecutive sites per step   (plaquette.py:53; compact.py delta,2 rows:
        //      (delta v)_0[x] = v[x] - v[x - e1],  (delta v)_1[x] = -(v[x] - v[x - e0]))
#pragma unroll
        for (int i4 = tid; i4 < V / 4; i4 += T) {
            const int i = 4 * i4, x0 = i / N, x1 = i - x0 * N;
            const int4 vc = *reinterpret_cast<const int4*>(sv + i);
            const int4 vu = *reinterpret_cast<const int4*>(sv + ((x0 - 1) & (N - 1)) * N + x1);        // v[x - e0]
            const int vl = sv[x0 * N + ((x1 - 1) & (N - 1))];                                           // v[x - e1] of the first
            int4 m0 = *reinterpret_cast<int4*>(F0 + i), m1 = *reinterpret_cast<int4*>(F1 + i);
            m0.x -= vc.x - vl;   m0.y -= vc.y - vc.x; m0.z -= vc.z - vc.y; m0.w -= vc.w - vc.z;
            m1.x -= vu.x - vc.x; m1.y -= vu.y - vc.y; m1.z -= vu.z - vc.z; m1.w -= vu.w - vc.w;
            *reinterpret_cast<int4*>(F0 + i) = m0;
            *reinterpret_cast<int4*>(F1 + i) = m1;
        }
        __syncthreads();

        int n_acc = 0;
        float sum_A = 0.0f;
        const double inv_kappa = __ddiv_rn(1.0, kappa), half_inv_kappa = __ddiv_rn(0.5, kappa);
        const float inv_kappa_f = (float)inv_kappa;
        for (int s = 0; s < a.n_sweeps; ++s) {
            const unsigned long long gc = a.chain0 + (unsigned long long)chain, gs = a.sweep0 + (unsigned long long)s;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int par = (row8 + c) & 1;
                const int x1 = 2 * k + par;
                const int xp1 = (x1 + 1) & (N - 1);
                // the four links of plaquette x: (0,x) F0[x], (1,x+e0) F1[x+e0], (0,x+e1) F0[x+e1], (1,x) F1[x]
                int32_t* pF0c = F0 + row8 * N + x1;
                int32_t* pF1c = F1 + row8 * N + x1;
                int32_t* pF0r = F0 + row8 * N + xp1;
                int32_t* pv = sv + row8 * N + x1;
#pragma unroll
                for (int gi = 0; gi < QUADS / GSTEP; ++gi) {
                    const int g = GSTEP == 1 ? gi : g_first + GSTEP * gi;
                    const uint32_t c0 = (uint32_t)((row8 + 32 * g) * N + x1);                 // worldline_quad_counter
                    const Philox4 bits = philox_plaquette_keys(a, gc, gs, c0);
                    // The plaquettes of a block lie eight rows apart: they share no link and no site, so their loads, decisions
                    // and stores are independent.  Written as three phases over the block -- all loads, all decisions, all
                    // stores -- the shared-memory latency of the four overlaps instead of adding up (the compiler cannot prove
                    // by itself that a store of one does not alias a load of the next).
                    int f0c[QW], f1d[QW], f0r[QW], f1c[QW];
#pragma unroll
                    for (int wd = 0; wd < QW; ++wd) {
                        const int q = 4 * g + wd;
                        const int o = 8 * N * q;                                               // row row8 + 8 q
                        const int od = (q == PER - 1 && row8 == 7) ? (o + N - V) : (o + N);    // row below (wraps after the last)
                        f0c[wd] = pF0c[o]; f1d[wd] = pF1c[od]; f0r[wd] = pF0r[o]; f1c[wd] = pF1c[o];
                    }
                    int dfs[QW], dvs[QW];
                    unsigned okm = 0;
#pragma unroll
                    for (int wd = 0; wd < QW; ++wd) {
                        const uint32_t w = (wd == 0) ? bits.x : (wd == 1) ? bits.y : (wd == 2) ? bits.z : bits.w;
                        int df, dv;                                           // delta_f on the (+) links; change of v[x]
                        uint32_t f;
                        if (MODE == SVB_WL_JOINT) {
                            // dm = +-1 from bit 31, dv in {-1, 0, 1} from the next bits, the remainder leads the uniform
                            const uint64_t p = (uint64_t)(w << 1) * 3ull;
                            const int hi = (int)(p >> 32);
                            f = (uint32_t)p;
                            dv = hi - 1;
                            df = 2 * (int)(w >> 31) - hi;                     // delta_f = dm - dv, dm = 2 (w >> 31) - 1
                        } else {
                            const uint64_t p = (uint64_t)w * (uint64_t)(2 * a.interval);
                            const int idx = (int)(p >> 32);
                            f = (uint32_t)p;
                            const int ch = (idx < a.interval) ? idx - a.interval : idx - a.interval + 1;
                            dv = (MODE == SVB_WL_VORTEX) ? ch : 0;
                            df = (MODE == SVB_WL_VORTEX) ? -ch : ch;
                        }
                        const int sI = (f0c[wd] + f1d[wd]) - f0r[wd] - f1c[wd] + 2 * df;      // f1 + f2 - f3 - f4 + 2 delta_f
                        const int kk = df * sI;
                        // dS > 0 (kk > 0) is a real Metropolis test: ONE integer comparison with the tabulated threshold.  Everything
                        // else about it -- s beyond the table, a uniform within two units of the threshold -- is the cold branch.
                        const int sC = min(max(sI, -kWlTableS), kWlTableS - 1);
                        const WlTableEntry e = table[(df + 2) * (2 * kWlTableS) + sC + kWlTableS];
                        const bool test = kk > 0, in = sC == sI;
                        bool ok = !test || (in && f < e.thr);
                        float A = test ? e.A : 1.0f;
                        if (test && (!in || (f - (e.thr - 2u)) <= 3u)) {
                            bool exact = true;                                // f within 2 of thr: the bracket of u may touch A
                            if (!in) {
                                // beyond the table A is below the edge entry of the same delta_f (dS grows with |s|): a uniform at
                                // least 2 units above the edge threshold rejects for sure; otherwise (probability ~ A_edge) exact
                                A = fast_ex2f(-1.4426950408889634f * inv_kappa_f * (float)kk);
                                ok = false;
                                exact = f < e.thr + 2u || e.thr >= 0xFFFFFFFEu;
                            }
                            if (exact) {
                                double dS;
                                if (MODE == SVB_WL_JOINT) {
                                    dS = __dmul_rn(__dmul_rn((double)df, inv_kappa), (double)sI);      // plaquette.py:84-85
                                } else {
                                    // coface_sum order: T(1,x) + T(1,x+e0) + T(0,x) + T(0,x+e1), T = ((0.5/kappa) c)((2 f) + c)
                                    const double Pp = __dmul_rn(half_inv_kappa, (double)df), Pm = __dmul_rn(half_inv_kappa, (double)(-df));
                                    dS = __dadd_rn(__dmul_rn(Pm, (double)(2 * f1c[wd] - df)), __dmul_rn(Pp, (double)(2 * f1d[wd] + df)));
                                    dS = __dadd_rn(dS, __dmul_rn(Pp, (double)(2 * f0c[wd] + df)));
                                    dS = __dadd_rn(dS, __dmul_rn(Pm, (double)(2 * f0r[wd] - df)));
                                }
                                const double Ad = exp_clipped(-dS);
                                LazyUniform lu;
                                lu.f = f; lu.c0 = c0; lu.word = (uint32_t)wd;
                                RefineCtx rc;
                                rc.seed = a.seed; rc.chain = gc; rc.sweep = gs; rc.stream = a.refine_stream; rc.wide = 0;
                                A = (float)Ad;
                                ok = decide_lazy(Ad, lu, rc.stream, rc);
                            }
                        }
                        sum_A += A;
                        n_acc += ok ? 1 : 0;
                        okm |= ok ? (1u << wd) : 0u;
                        dfs[wd] = df; dvs[wd] = dv;
                    }
#pragma unroll
                    for (int wd = 0; wd < QW; ++wd) {
                        const int q = 4 * g + wd;
                        const int o = 8 * N * q;
                        const int od = (q == PER - 1 && row8 == 7) ? (o + N - V) : (o + N);
                        if (okm & (1u << wd)) {                               // plaquette.py:91-101
                            const int df = dfs[wd];
                            pF0c[o] = f0c[wd] + df;
                            pF1c[od] = f1d[wd] + df;
                            pF0r[o] = f0r[wd] - df;
                            pF1c[o] = f1c[wd] - df;
                            if (MODE != SVB_WL_COEXACT && dvs[wd] != 0) atomicAdd(pv + o, dvs[wd]);    // only this thread touches x in this pass
                        }
                    }
                }
                __syncthreads();
            }
        }

        if (want_obs) {
            // integer sums over f: sum f^2 (action), sum (d f)^2 (winding), sum f_mu = sum m_mu (wrapping)
            long long f2 = 0, df2 = 0;
            int w0 = 0, w1 = 0;
#pragma unroll
            for (int i4 = tid; i4 < V / 4; i4 += T) {
                const int i = 4 * i4, x0 = i / N, x1 = i - x0 * N;
                const int4 a0 = *reinterpret_cast<const int4*>(F0 + i), a1 = *reinterpret_cast<const int4*>(F1 + i);
                const int4 d1 = *reinterpret_cast<const int4*>(F1 + ((x0 + 1) & (N - 1)) * N + x1);     // f_1[x + e0]
                const int r0 = F0[x0 * N + ((x1 + 4) & (N - 1))];                                        // f_0[x + e1] of the last
                f2 += (long long)a0.x * a0.x + (long long)a0.y * a0.y + (long long)a0.z * a0.z + (long long)a0.w * a0.w;
                f2 += (long long)a1.x * a1.x + (long long)a1.y * a1.y + (long long)a1.z * a1.z + (long long)a1.w * a1.w;
                // (d f)[x] = (f_1[x+e0] - f_1[x]) - (f_0[x+e1] - f_0[x])      (compact.py d,1 rows)
                const int c0 = (d1.x - a1.x) - (a0.y - a0.x), c1 = (d1.y - a1.y) - (a0.z - a0.y);
                const int c2 = (d1.z - a1.z) - (a0.w - a0.z), c3 = (d1.w - a1.w) - (r0 - a0.w);
                df2 += (long long)c0 * c0 + (long long)c1 * c1 + (long long)c2 * c2 + (long long)c3 * c3;
                w0 += a0.x + a0.y + a0.z + a0.w;
                w1 += a1.x + a1.y + a1.z + a1.w;
            }
            f2 = warp_sum(f2);
            df2 = warp_sum(df2);
            w0 = __reduce_add_sync(0xffffffffu, w0);
            w1 = __reduce_add_sync(0xffffffffu, w1);
            n_acc = __reduce_add_sync(0xffffffffu, n_acc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum_A += __shfl_xor_sync(0xffffffffu, sum_A, off);
            if (lane == 0) {
                long long* slot = red + 4 * warp;
                slot[0] = f2; slot[1] = df2; slot[2] = ((long long)w0 << 32) | (unsigned)w1; slot[3] = n_acc;
                redA[warp] = sum_A;
            }
            __syncthreads();           // every warp has read f; the slots are written
            if (tid == kWriter) {
                long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
                double tA = 0.0;
                for (int w = 0; w < NW; ++w) {
                    t0 += red[4 * w]; t1 += red[4 * w + 1];
                    t2 += red[4 * w + 2] >> 32; t3 += (int)(red[4 * w + 2] & 0xFFFFFFFFLL);
                    t4 += red[4 * w + 3];
                    tA += (double)redA[w];
                }
                double* o = a.obs + chain * SVB_WOBS_COUNT;
                o[SVB_WOBS_SUM_F2] = (double)t0;
                o[SVB_WOBS_SUM_DF2] = (double)t1;
                o[SVB_WOBS_WRAP0] = (double)t2;
                o[SVB_WOBS_WRAP1] = (double)t3;
                o[SVB_WOBS_ACCEPTED] = (double)t4;
                o[SVB_WOBS_ACCEPTANCE] = tA;
                o[SVB_WOBS_DELTA_M_ABS] = -1.0;       // not evaluated by the sweep (the move preserves delta m identically)
            }
        }

        // ---- f -> m = f + delta v with the final v, in place (VORTEX leaves m as it was: nothing to write back) ----
        if (MODE != SVB_WL_VORTEX)
#pragma unroll
        for (int i4 = tid; i4 < V / 4; i4 += T) {
            const int i = 4 * i4, x0 = i / N, x1 = i - x0 * N;
            const int4 vc = *reinterpret_cast<const int4*>(sv + i);
            const int4 vu = *reinterpret_cast<const int4*>(sv + ((x0 - 1) & (N - 1)) * N + x1);
            const int vl = sv[x0 * N + ((x1 - 1) & (N - 1))];
            int4 m0 = *reinterpret_cast<int4*>(F0 + i), m1 = *reinterpret_cast<int4*>(F1 + i);
            m0.x += vc.x - vl;   m0.y += vc.y - vc.x; m0.z += vc.z - vc.y; m0.w += vc.w - vc.z;
            m1.x += vu.x - vc.x; m1.y += vu.y - vc.y; m1.z += vu.z - vc.z; m1.w += vu.w - vc.w;
            *reinterpret_cast<int4*>(F0 + i) = m0;
            *reinterpret_cast<int4*>(F1 + i) = m1;
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (MODE != SVB_WL_VORTEX) bulk_s2g(a.m + chain * 2 * V, F0, bytes_m);
            if (MODE != SVB_WL_COEXACT) bulk_s2g(a.v + chain * V, sv, bytes_v);
            bulk_commit();
            bulk_wait_read0();
            if (next < a.chains) issue_load(next, seen_next, stage);
        }
    }
    if (tid == 0) bulk_wait0();
    if (OVERLAP) {
        __syncthreads();                                   // every store has completed, every record is written
        if (warp == 0) overlap_publish_all(a.ov, lane, it, blockIdx.x, gridDim.x);
    }
}

template <int MODE, int NT, int MINB, int TM = 4, int STAGES = 1>
static int launch_worldline_table(const WorldlineArgs& a, cudaStream_t stream, int sm_count) {
    const bool overlap = a.ov.epochs != nullptr;
    auto kern = overlap ? worldline_smem_table_kernel<MODE, NT, MINB, true, TM, STAGES> : worldline_smem_table_kernel<MODE, NT, MINB, false, TM, STAGES>;
    const size_t smem = (size_t)STAGES * NT * NT * 3 * sizeof(int32_t) + kWlTableSize * sizeof(WlTableEntry) + 4 * 32 * sizeof(long long) +
                        32 * sizeof(float) + 8 * STAGES + 8;
    static int per_sm_cache[2] = {0, 0};
    int per_sm = per_sm_cache[overlap ? 1 : 0];
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TM * NT, smem));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "worldline table kernel does not fit an SM at N=%d", NT);
        per_sm_cache[overlap ? 1 : 0] = per_sm;
    }
    long long grid = (long long)per_sm * sm_count;
    if (grid > a.chains) grid = a.chains;
    if (overlap) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TM * NT); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SVB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
        return 0;
    }
    kern<<<(unsigned)grid, TM * NT, smem, stream>>>(a);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// svb_worldline_observables for W = 1 and N in {16, 32, 64}: the chain staged by a 1-D TMA bulk copy, f = m - delta v built
// in place as in the sweep kernel, then integer sums over f (action, winding, wrapping) and |delta m| from the staged m.
// HBM-bound: one read of the state (12 B per site).
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(4 * NT) worldline_obs_smem_kernel(const int32_t* __restrict__ m, const int32_t* __restrict__ v,
                                                                    long long chains, double* __restrict__ obs, int keep_counters) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = N * N, T = 4 * NT, NW = T / 32;
    constexpr uint32_t bytes_m = 2 * V * sizeof(int32_t), bytes_v = V * sizeof(int32_t);
    int32_t* F0 = reinterpret_cast<int32_t*>(smem_raw);
    int32_t* F1 = F0 + V;
    int32_t* sv = F1 + V;
    long long* red = reinterpret_cast<long long*>(smem_raw + bytes_m + bytes_v);          // [NW][4]
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 4 * NW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue_load = [&](long long chain) {
        mbar_expect_tx(bar, bytes_m + bytes_v);
        bulk_g2s(F0, m + chain * 2 * V, bytes_m, bar);
        bulk_g2s(sv, v + chain * V, bytes_v, bar);
    };
    long long chain = blockIdx.x;
    if (tid == 0 && chain < chains) issue_load(chain);
    for (int it = 0; chain < chains; chain += gridDim.x, ++it) {
        mbar_wait(bar, (uint32_t)(it & 1));
        // |delta m| from the staged m:  (delta m)[x] = -(m0[x] - m0[x-e0]) - (m1[x] - m1[x-e1])   (compact.py delta,1 rows)
        long long dm_abs = 0;
#pragma unroll
        for (int i4 = tid; i4 < V / 4; i4 += T) {
            const int i = 4 * i4, x0 = i / N, x1 = i - x0 * N;
            const int4 a0 = *reinterpret_cast<const int4*>(F0 + i), a1 = *reinterpret_cast<const int4*>(F1 + i);
            const int4 u0 = *reinterpret_cast<const int4*>(F0 + ((x0 - 1) & (N - 1)) * N + x1);
            const int l1 = F1[x0 * N + ((x1 - 1) & (N - 1))];
            const int d0 = -(a0.x - u0.x) - (a1.x - l1), d1 = -(a0.y - u0.y) - (a1.y - a1.x);
            const int d2 = -(a0.z - u0.z) - (a1.z - a1.y), d3 = -(a0.w - u0.w) - (a1.w - a1.z);
            dm_abs += abs(d0) + abs(d1) + abs(d2) + abs(d3);
        }
        __syncthreads();
        // m -> f = m - delta v, in place (as in the sweep kernel)
#pragma unroll
        for (int i4 = tid; i4 < V / 4; i4 += T) {
            const int i = 4 * i4, x0 = i / N, x1 = i - x0 * N;
            const int4 vc = *reinterpret_cast<const int4*>(sv + i);
            const int4 vu = *reinterpret_cast<const int4*>(sv + ((x0 - 1) & (N - 1)) * N + x1);
            const int vl = sv[x0 * N + ((x1 - 1) & (N - 1))];
            int4 m0 = *reinterpret_cast<int4*>(F0 + i), m1 = *reinterpret_cast<int4*>(F1 + i);
            m0.x -= vc.x - vl;   m0.y -= vc.y - vc.x; m0.z -= vc.z - vc.y; m0.w -= vc.w - vc.z;
            m1.x -= vu.x - vc.x; m1.y -= vu.y - vc.y; m1.z -= vu.z - vc.z; m1.w -= vu.w - vc.w;
            *reinterpret_cast<int4*>(F0 + i) = m0;
            *reinterpret_cast<int4*>(F1 + i) = m1;
        }
        __syncthreads();
        long long f2 = 0, df2 = 0;
        int w0 = 0, w1 = 0;
#pragma unroll
        for (int i4 = tid; i4 < V / 4; i4 += T) {
            const int i = 4 * i4, x0 = i / N, x1 = i - x0 * N;
            const int4 a0 = *reinterpret_cast<const int4*>(F0 + i), a1 = *reinterpret_cast<const int4*>(F1 + i);
            const int4 d1 = *reinterpret_cast<const int4*>(F1 + ((x0 + 1) & (N - 1)) * N + x1);
            const int r0 = F0[x0 * N + ((x1 + 4) & (N - 1))];
            f2 += (long long)a0.x * a0.x + (long long)a0.y * a0.y + (long long)a0.z * a0.z + (long long)a0.w * a0.w;
            f2 += (long long)a1.x * a1.x + (long long)a1.y * a1.y + (long long)a1.z * a1.z + (long long)a1.w * a1.w;
            const int c0 = (d1.x - a1.x) - (a0.y - a0.x), c1 = (d1.y - a1.y) - (a0.z - a0.y);
            const int c2 = (d1.z - a1.z) - (a0.w - a0.z), c3 = (d1.w - a1.w) - (r0 - a0.w);
            df2 += (long long)c0 * c0 + (long long)c1 * c1 + (long long)c2 * c2 + (long long)c3 * c3;
            w0 += a0.x + a0.y + a0.z + a0.w;
            w1 += a1.x + a1.y + a1.z + a1.w;
        }
        f2 = warp_sum(f2);
        df2 = warp_sum(df2);
        dm_abs = warp_sum(dm_abs);
        w0 = __reduce_add_sync(0xffffffffu, w0);
        w1 = __reduce_add_sync(0xffffffffu, w1);
        if (lane == 0) {
            long long* slot = red + 4 * warp;
            slot[0] = f2; slot[1] = df2; slot[2] = ((long long)w0 << 32) | (unsigned)w1; slot[3] = dm_abs;
        }
        __syncthreads();                       // every warp has read the chain; the slots are written
        const long long next = chain + gridDim.x;
        if (tid == 0 && next < chains) issue_load(next);
        if (tid == 32) {
            long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
            for (int w = 0; w < NW; ++w) {
                t0 += red[4 * w]; t1 += red[4 * w + 1];
                t2 += red[4 * w + 2] >> 32; t3 += (int)(red[4 * w + 2] & 0xFFFFFFFFLL);
                t4 += red[4 * w + 3];
            }
            double* o = obs + chain * SVB_WOBS_COUNT;
            o[SVB_WOBS_SUM_F2] = (double)t0;
            o[SVB_WOBS_SUM_DF2] = (double)t1;
            o[SVB_WOBS_WRAP0] = (double)t2;
            o[SVB_WOBS_WRAP1] = (double)t3;
            o[SVB_WOBS_DELTA_M_ABS] = (double)t4;
            if (!keep_counters) { o[SVB_WOBS_ACCEPTED] = 0.0; o[SVB_WOBS_ACCEPTANCE] = 0.0; }
        }
        __syncthreads();
    }
}

template <int NT>
static int launch_worldline_obs_smem(const int32_t* m, const int32_t* v, long long chains, double* obs, int keep_counters,
                                     cudaStream_t stream) {
    auto kern = worldline_obs_smem_kernel<NT>;
    const size_t smem = (size_t)NT * NT * 12 + 4 * (4 * NT / 32) * sizeof(long long) + 16;
    static int grid_cap = 0;
    if (grid_cap == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        SVB_CUDA_TRY(cudaGetDevice(&dev));
        SVB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 4 * NT, smem));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "worldline observable kernel does not fit an SM at N=%d", NT);
        grid_cap = per_sm * sms;
    }
    const long long grid = chains < grid_cap ? chains : grid_cap;
    kern<<<(unsigned)grid, 4 * NT, smem, stream>>>(m, v, chains, obs, keep_counters);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}
