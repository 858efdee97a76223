// svb_villain_stream.cuh -- the STREAMING Villain sweep kernels (included by svb_villain.cu).
//
// NeighborhoodUpdate.step (supervillain/generator/villain/neighborhood.py:59-137) for Philox draws, fp64 phi, FAST
// arithmetic and even N that is a multiple of 16, with NO shared-memory copy of the fields at all:
//
//  * A colour pass reads what it needs straight from global memory (L1 / L2): a thread owns the site pair (r, x1),
//    (r + 8, x1) that shares a Philox block (draw mapping version 2), loads the five phi and four n around each site
//    (one 16-byte and three 8-byte loads of phi, one 8-byte and three 4-byte loads of n; a warp's loads are contiguous
//    along the row), builds the four residuals r = d(phi) - 2 pi n (neighborhood.py:91) in fp64 and rounds them to fp32.
//    Same-colour sites touch disjoint links and read only phi of the OTHER colour, so within a pass no
//    thread reads anything another thread writes; the two passes of a sweep are separated by a block barrier (a chain per
//    CTA) or by a launch boundary (one big lattice).
//  * The decision is the fp32-filtered one of svb_villain_filtered.cuh -- dS32 against -log2 u with the error band, the
//    exact fp64 test (from the fp64 residuals the thread already holds) and the lazily refined uniform inside it -- so every
//    decision equals the exact fp64 decision.  The residuals are rebuilt from the current fields in every pass, never
//    patched: their fp32 error is one rounding, inside the band derived for the patched copies.
//  * phi and n are WRITTEN ONLY WHERE A PROPOSAL IS ACCEPTED (five stores).  An unchanged value is not rewritten: the DRAM
//    traffic of a sweep is one read of the state plus the dirty sectors (acceptance is 0.5 - 5 %), against the algorithmic
//    read + write of 32 B per site-update, and there is no store phase, no staging buffer to wait for and no ghost zone.
//  * The next chain is prefetched into L2 (cp.async.bulk.prefetch.L2) while the current one is swept, so a pass starts from
//    L2, not from DRAM.
//  * A thread keeps its column and walks down the rows (r, r + 16, r + 32, ...): everything about the column -- parity, the
//    wrapped neighbour columns -- is computed once, a step costs a handful of offset additions.
//
// Two kernels share the pair routine:
//   villain_stream_chain_kernel   N in {16, 32, 64, 128}: one chain per CTA at a time (4 N threads), both colours (and all fused
//                                 sweeps) in one launch, persistent CTAs striding over the chains, overlapped launches with
//                                 per-chain epochs as svb_villain_filtered.cuh, records without atomics.
//   villain_stream_pass_kernel    any N that is a multiple of 16 (config 5: L = 4096): one launch per colour pass; a CTA (8 warps)
//                                 owns a band of 64 columns and a run of row groups; in place -- no workspace, no ping-pong,
//                                 no ghost zones.
#pragma once
// (included inside namespace svb, after svb_villain_filtered.cuh)

// Programmatic dependent launch for the kernels of an in-place step (records -> sum dn^2 -> colour 0 -> colour 1): every kernel
// lets its successor's CTAs be scheduled as soon as its own are resident (pdl_prologue) and waits for its predecessors to have
// completed before it touches memory, so the launch latency and ramp-up of one kernel hide under the tail of the one before.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// A thread's column slot: constant while it walks down the rows.
struct StreamCol {
    int k2;          // 2 k: the aligned column pair [2k, 2k + 1] of every row holds the site and one in-row neighbour
    int par;         // the site is column 2k + par
    int x1, xm1, xp1;
    int xo;          // the in-row neighbour outside the aligned pair: xp1 if par else xm1
};
__device__ __forceinline__ StreamCol stream_col(int k, int par, int N) {
    StreamCol c;
    c.k2 = 2 * k; c.par = par; c.x1 = 2 * k + par;
    c.xm1 = (c.x1 == 0) ? N - 1 : c.x1 - 1;
    c.xp1 = (c.x1 + 1 == N) ? 0 : c.x1 + 1;
    c.xo = par ? c.xp1 : c.xm1;
    return c;
}

// The four residuals of the site at row offset o (rows above / below at om / op), rounded to fp32.  SUMS: the observables of
// the state as it is now -- every link has exactly one endpoint of this colour, and a site's own four links are written by
// nobody else during the pass.  (sum (dn)^2 is NOT collected here: a plaquette's other links belong to same-colour sites that
// may be writing them right now -- it has a pass of its own, stream_dn2_rows.)
template <bool SUMS>
__device__ __forceinline__ float4 stream_site_residuals(const double* gphi, const int32_t* gn0, const int32_t* gn1, int o, int om,
                                                        int op, const StreamCol& col, double& action, int& w0, int& w1) {
    const double2 pc2 = *reinterpret_cast<const double2*>(gphi + o + col.k2);
    const double po = gphi[o + col.xo];
    const double pu = gphi[op + col.x1], pd = gphi[om + col.x1];
    const int2 n1p = *reinterpret_cast<const int2*>(gn1 + o + col.k2);
    const int n1e = gn1[o + col.xo];                              // column 2k - 1 when par == 0 (unused otherwise)
    const int n0c = gn0[o + col.x1], n0b = gn0[om + col.x1];
    const double pc = col.par ? pc2.y : pc2.x;
    const double pl = col.par ? pc2.x : po, pr = col.par ? po : pc2.y;
    const int n1c = col.par ? n1p.y : n1p.x, n1b = col.par ? n1p.x : n1e;
    const double rf0 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n0c), pu - pc);
    const double rb0 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n0b), pc - pd);
    const double rf1 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n1c), pr - pc);
    const double rb1 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n1b), pc - pl);
    if (SUMS) {
        action = fma(rf0, rf0, action);
        action = fma(rb0, rb0, action);
        action = fma(rf1, rf1, action);
        action = fma(rb1, rb1, action);
        w0 += n0c + n0b;
        w1 += n1c + n1b;
    }
    return make_float4((float)rf0, (float)rb0, (float)rf1, (float)rb1);
}

// Geometry of a TMA-staged tile (villain_tile_pass_kernel, villain_tile_wave_kernel below).
constexpr int kTileRows = 16, kTileCols = 128;
constexpr int kTilePhiRows = kTileRows + 2, kTilePhiCols = kTileCols + 4;     // rows R-1 .. R+16, columns C-2 .. C+129
constexpr int kTileNRows = kTileRows + 1, kTileNCols = kTileCols + 4;        // rows R-1 .. R+15, columns C-4 .. C+127
constexpr int kTilePhiBytes = kTilePhiRows * kTilePhiCols * 8;                 // 19008
constexpr int kTileNBytes = kTileNRows * kTileNCols * 4;                       // 8976
constexpr int kTilePhiSlot = (kTilePhiBytes + 127) / 128 * 128, kTileNSlot = (kTileNBytes + 127) / 128 * 128;
constexpr int kTileStageBytes = kTilePhiSlot + 2 * kTileNSlot;
constexpr int kTileLutBytes = 81 * 16;                                         // the 81 proposals of (dn) as residual changes (stream_fill_lut)
constexpr int kTileSmemBytes = 2 * kTileStageBytes + 64 + kTileLutBytes;
static_assert(kTilePhiCols == kTileNCols, "the exact test of a staged tile indexes phi and n with one stride");

// The exact test of a proposal whose fp32 comparison is inside its error band (a few 1e-5 of the proposals): dS in fp64 from
// the current phi and n (L1 / L2), the fp64 exponential, the lazily refined uniform.  Out of line on purpose: the hot loop
// should not pay registers for it.
__device__ __noinline__ bool stream_exact(const VillainArgs& a, const double* gphi, const int32_t* gn0, int V, int o, int om, int op, int x1,
                                          int xm1, int xp1, uint32_t wA, uint32_t f, uint32_t c0, uint32_t half, int g0, int g1, int g2,
                                          int g3, int W, double half_kappa, unsigned long long gc, unsigned long long gs) {
    ExactProposal ep;
    ep.phi = gphi; ep.n0 = gn0; ep.n1 = gn0 + V;
    ep.i_c = o + x1; ep.i_b0 = om + x1; ep.i_b1 = o + xm1; ep.i_f0 = op + x1; ep.i_f1 = o + xp1;
    ep.half_kappa = half_kappa; ep.c = SVB_TWO_PI * (double)W;
    ep.dphi = villain_dphi_from_word(wA, a.interval_phi);
    ep.g[0] = g0; ep.g[1] = g1; ep.g[2] = g2; ep.g[3] = g3;
    ep.d.f = f; ep.d.c0 = c0; ep.d.half = half;
    ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs; ep.rc.stream = a.refine_stream; ep.rc.wide = 0;
    return villain_exact_decision(ep);
}

// The same test from a tile staged in shared memory (villain_tile_wave_kernel): idx = row * kTilePhiCols + column slot + 4 in the n
// boxes, the phi box is shifted by two columns.  Nothing a site's test reads is written by anybody else during its colour phase,
// so the staged copy is as current as global memory -- and, unlike a plain global load, it cannot come from a stale L1 line.
__device__ __noinline__ bool tile_exact(const VillainArgs& a, const double* P, const int32_t* N0, const int32_t* N1, int idx, uint32_t wA,
                                        uint32_t f, uint32_t c0, uint32_t half, int g0, int g1, int g2, int g3, int W, double half_kappa,
                                        unsigned long long gc, unsigned long long gs) {
    ExactProposal ep;
    ep.phi = P - 2; ep.n0 = N0; ep.n1 = N1;
    ep.i_c = idx; ep.i_b0 = idx - kTilePhiCols; ep.i_b1 = idx - 1; ep.i_f0 = idx + kTilePhiCols; ep.i_f1 = idx + 1;
    ep.half_kappa = half_kappa; ep.c = SVB_TWO_PI * (double)W;
    ep.dphi = villain_dphi_from_word(wA, a.interval_phi);
    ep.g[0] = g0; ep.g[1] = g1; ep.g[2] = g2; ep.g[3] = g3;
    ep.d.f = f; ep.d.c0 = c0; ep.d.half = half;
    ep.rc.seed = a.seed; ep.rc.chain = gc; ep.rc.sweep = gs; ep.rc.stream = a.refine_stream; ep.rc.wide = 0;
    return villain_exact_decision(ep);
}

// An accepted proposal (:121-129), applied to global memory as fire-and-forget reductions: the fp64 one rounds once, to nearest,
// like the reference's phi + dphi.  (Acceptance is 0.5 - 5 %, so some lane of most warps comes through here: it is inline.)
__device__ __forceinline__ void stream_accept(double* gphi, int32_t* gn0, int V, int o, int om, int x1, int xm1, uint32_t wA,
                                              double interval_phi, double two_I_scaled, int d0, int d1, int d2, int d3) {
    const double Ah = __hiloint2double(0x43300000, (int)wA) - 4503599627370495.5;          // villain_dphi_from_word, one multiply less
    atomicAdd(gphi + o + x1, __dadd_rn(-interval_phi, __dmul_rn(two_I_scaled, Ah)));
    atomicAdd(gn0 + o + x1, d0);
    atomicAdd(gn0 + om + x1, d1);
    atomicAdd(gn0 + V + o + x1, d2);
    atomicAdd(gn0 + V + o + xm1, d3);
}

// The decision half of a colour pass over a pair of sites (row offsets oA and oB = oA + 8 N, column x1) that share the Philox
// block `bits` (counter word 0 = oA + x1, draw mapping version 2), from their fp32 residuals rA, rB = (f0, b0, f1, b1).
template <bool UNIT, bool LUT = true, bool TILE = false>
__device__ __forceinline__ void stream_pair_decide(const VillainArgs& a, const FilterConsts& fc, double* gphi, int32_t* gn0, int V, int oA,
                                                   int oAm, int oAp, int oB, int oBm, int oBp, int x1, int xm1, int xp1, const Philox4& bits,
                                                   uint32_t c0, const float4& rA, const float4& rB, unsigned long long gc,
                                                   unsigned long long gs, double half_kappa, float hk2, float hkA, float hkB, int& n_acc,
                                                   float& sum_A, const float4* dn_lut, const double* tP = nullptr,
                                                   const int32_t* tN0 = nullptr, const int32_t* tN1 = nullptr, int tIdx = 0) {
    const int interval_n = UNIT ? 1 : a.interval_n;
    const uint32_t K = (uint32_t)(2 * interval_n + 1);
    const int W = UNIT ? 1 : a.W;
    // proposals: four base-K digits each, then the leading 32 bits of the uniform (svb_villain_filtered.cuh)
    uint32_t fA = bits.y, fB = bits.w;
    uint32_t codeA = 0, codeB = 0;
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    if (UNIT) {
        // one multiply by 81: hi = 27 d0 + 9 d1 + 3 d2 + d3 (the code, which looks the four residual changes up in dn_lut), lo = the
        // remainder that leads the uniform; the digits themselves are decoded only on the rare paths
        const uint64_t pa = (uint64_t)fA * 81u, pb = (uint64_t)fB * 81u;
        fA = (uint32_t)pa; fB = (uint32_t)pb;
        codeA = (uint32_t)(pa >> 32); codeB = (uint32_t)(pb >> 32);
        if (!LUT) {
            // (the tile kernel keeps its shared-memory bandwidth for the staged fields: there the digits are decoded arithmetically)
            uint32_t c = codeA;
            a0 = (int)((c * 2428u) >> 16); c -= 27u * (uint32_t)a0;
            a1 = (int)((c * 7282u) >> 16); c -= 9u * (uint32_t)a1;
            a2 = (int)((c * 21846u) >> 16); a3 = (int)(c - 3u * (uint32_t)a2);
            c = codeB;
            b0 = (int)((c * 2428u) >> 16); c -= 27u * (uint32_t)b0;
            b1 = (int)((c * 7282u) >> 16); c -= 9u * (uint32_t)b1;
            b2 = (int)((c * 21846u) >> 16); b3 = (int)(c - 3u * (uint32_t)b2);
        }
    } else {
        uint64_t pa, pb;
        pa = (uint64_t)fA * K; fA = (uint32_t)pa; a0 = (int)(pa >> 32);  pb = (uint64_t)fB * K; fB = (uint32_t)pb; b0 = (int)(pb >> 32);
        pa = (uint64_t)fA * K; fA = (uint32_t)pa; a1 = (int)(pa >> 32);  pb = (uint64_t)fB * K; fB = (uint32_t)pb; b1 = (int)(pb >> 32);
        pa = (uint64_t)fA * K; fA = (uint32_t)pa; a2 = (int)(pa >> 32);  pb = (uint64_t)fB * K; fB = (uint32_t)pb; b2 = (int)(pb >> 32);
        pa = (uint64_t)fA * K; fA = (uint32_t)pa; a3 = (int)(pa >> 32);  pb = (uint64_t)fB * K; fB = (uint32_t)pb; b3 = (int)(pb >> 32);
    }
    const float cIn = fc.c * (float)interval_n;
    const float2 cIn2 = make_float2(cIn, cIn), negc2 = make_float2(-fc.c, -fc.c), two2 = make_float2(2.0f, 2.0f);
    // dphi from 23 centred bits
    float2 U = make_float2(__uint_as_float(0x3F800000u | (bits.x >> 9)), __uint_as_float(0x3F800000u | (bits.z >> 9)));
    U = __fadd2_rn(U, make_float2(-0.99999994f, -0.99999994f));
    const float2 dphi = __ffma2_rn(make_float2(fc.two_I, fc.two_I), U, make_float2(-fc.I, -fc.I));
    const float2 base_f = __ffma2_rn(dphi, make_float2(-1.0f, -1.0f), cIn2), base_b = __fadd2_rn(cIn2, dphi);
    const float2 r_f0 = make_float2(rA.x, rB.x), r_b0 = make_float2(rA.y, rB.y);
    const float2 r_f1 = make_float2(rA.z, rB.z), r_b1 = make_float2(rA.w, rB.w);
    // dr = d(dphi) - 2 pi dn   (neighborhood.py:110), dn = W (digit - interval_n)
    float2 dr_f0, dr_b0, dr_f1, dr_b1;
    if (UNIT && LUT) {
        const float4 tA = dn_lut[codeA], tB = dn_lut[codeB];          // -2 pi digit: exact in fp32, so the add is the fma below
        dr_f0 = __fadd2_rn(base_f, make_float2(tA.x, tB.x));
        dr_b0 = __fadd2_rn(base_b, make_float2(tA.y, tB.y));
        dr_f1 = __fadd2_rn(base_f, make_float2(tA.z, tB.z));
        dr_b1 = __fadd2_rn(base_b, make_float2(tA.w, tB.w));
    } else {
        dr_f0 = __ffma2_rn(negc2, make_float2((float)a0, (float)b0), base_f);
        dr_b0 = __ffma2_rn(negc2, make_float2((float)a1, (float)b1), base_b);
        dr_f1 = __ffma2_rn(negc2, make_float2((float)a2, (float)b2), base_f);
        dr_b1 = __ffma2_rn(negc2, make_float2((float)a3, (float)b3), base_b);
    }
    float2 acc2 = __fmul2_rn(dr_f0, __ffma2_rn(two2, r_f0, dr_f0));
    acc2 = __ffma2_rn(dr_b0, __ffma2_rn(two2, r_b0, dr_b0), acc2);
    acc2 = __ffma2_rn(dr_f1, __ffma2_rn(two2, r_f1, dr_f1), acc2);
    acc2 = __ffma2_rn(dr_b1, __ffma2_rn(two2, r_b1, dr_b1), acc2);
    const float2 dS2 = __fmul2_rn(make_float2(hk2, hk2), acc2);                          // dS / ln 2
    // -log2(f 2^-32); u lies in [f, f + 1] 2^-32
    const float2 L2 = __ffma2_rn(make_float2(fast_lg2((float)fA), fast_lg2((float)fB)), make_float2(-1.0f, -1.0f),
                                 make_float2(32.0f, 32.0f));
    const float2 Rmax = make_float2(fmaxf(fmaxf(fabsf(r_f0.x), fabsf(r_b0.x)), fmaxf(fabsf(r_f1.x), fabsf(r_b1.x))),
                                    fmaxf(fmaxf(fabsf(r_f0.y), fabsf(r_b0.y)), fmaxf(fabsf(r_f1.y), fabsf(r_b1.y))));
    const float2 band = __ffma2_rn(make_float2(hkA, hkA), Rmax, __ffma2_rn(make_float2(4e-6f, 4e-6f), L2, make_float2(hkB, hkB)));
    const float2 diff = __ffma2_rn(L2, make_float2(-1.0f, -1.0f), dS2);
    sum_A += fminf(fast_ex2(-dS2.x), 1.0f) + fminf(fast_ex2(-dS2.y), 1.0f);
    // certainly rejected (the overwhelming majority): nothing more to do.  Everything else -- accepted, or inside the error band of
    // the fp32 comparison -- is handled behind ONE branch per pair of sites.
    const bool candA = !(diff.x > band.x) || fA < 65536u, candB = !(diff.y > band.y) || fB < 65536u;
    if (candA || candB) {
        const int mWI = -W * interval_n;
        const double two_I_scaled = (2.0 * a.interval_phi) * 2.3283064365386963e-10;          // (2 I) 2^-32, exact scaling
        if (UNIT && LUT) {
            uint32_t c = codeA;
            a0 = (int)((c * 2428u) >> 16); c -= 27u * (uint32_t)a0;
            a1 = (int)((c * 7282u) >> 16); c -= 9u * (uint32_t)a1;
            a2 = (int)((c * 21846u) >> 16); a3 = (int)(c - 3u * (uint32_t)a2);
            c = codeB;
            b0 = (int)((c * 2428u) >> 16); c -= 27u * (uint32_t)b0;
            b1 = (int)((c * 7282u) >> 16); c -= 9u * (uint32_t)b1;
            b2 = (int)((c * 21846u) >> 16); b3 = (int)(c - 3u * (uint32_t)b2);
        }
        if (candA) {
            bool ok = diff.x < 0.0f;
            if (!(fabsf(diff.x) > band.x) || fA < 65536u)
                ok = TILE ? tile_exact(a, tP, tN0, tN1, tIdx, bits.x, fA, c0, 0u, a0 - interval_n, a1 - interval_n, a2 - interval_n,
                                       a3 - interval_n, W, half_kappa, gc, gs)
                          : stream_exact(a, gphi, gn0, V, oA, oAm, oAp, x1, xm1, xp1, bits.x, fA, c0, 0u, a0 - interval_n, a1 - interval_n,
                                         a2 - interval_n, a3 - interval_n, W, half_kappa, gc, gs);
            if (ok) {
                n_acc += 1;
                stream_accept(gphi, gn0, V, oA, oAm, x1, xm1, bits.x, a.interval_phi, two_I_scaled, W * a0 + mWI, W * a1 + mWI, W * a2 + mWI,
                              W * a3 + mWI);
            }
        }
        if (candB) {
            bool ok = diff.y < 0.0f;
            if (!(fabsf(diff.y) > band.y) || fB < 65536u)
                ok = TILE ? tile_exact(a, tP, tN0, tN1, tIdx + 8 * kTilePhiCols, bits.z, fB, c0, 1u, b0 - interval_n, b1 - interval_n,
                                       b2 - interval_n, b3 - interval_n, W, half_kappa, gc, gs)
                          : stream_exact(a, gphi, gn0, V, oB, oBm, oBp, x1, xm1, xp1, bits.z, fB, c0, 1u, b0 - interval_n, b1 - interval_n,
                                         b2 - interval_n, b3 - interval_n, W, half_kappa, gc, gs);
            if (ok) {
                n_acc += 1;
                stream_accept(gphi, gn0, V, oB, oBm, x1, xm1, bits.z, a.interval_phi, two_I_scaled, W * b0 + mWI, W * b1 + mWI, W * b2 + mWI,
                              W * b3 + mWI);
            }
        }
    }
}

// One colour pass over the pair of sites at row offsets oA and oB = oA + 8 N of the column slot `col`, straight from global memory.
template <bool UNIT, bool SUMS>
__device__ __forceinline__ void stream_pair(const VillainArgs& a, const FilterConsts& fc, double* gphi, int32_t* gn0, int V, int oA,
                                            int oAm, int oAp, int oB, int oBm, int oBp, const StreamCol& col, unsigned long long gc,
                                            unsigned long long gs, double half_kappa, float hk2, float hkA, float hkB, int& n_acc,
                                            float& sum_A, double& action, int& w0, int& w1, const float4* dn_lut) {
    // the Philox block does not depend on memory: issue it first so that it overlaps the loads
    const uint32_t c0 = (uint32_t)(oA + col.x1);                          // villain_pair_counter: (x0 & ~8) N + x1
    const Philox4 bits = philox_site_keys(a, gc, gs, c0);
    const float4 rA = stream_site_residuals<SUMS>(gphi, gn0, gn0 + V, oA, oAm, oAp, col, action, w0, w1);
    const float4 rB = stream_site_residuals<SUMS>(gphi, gn0, gn0 + V, oB, oBm, oBp, col, action, w0, w1);
    stream_pair_decide<UNIT>(a, fc, gphi, gn0, V, oA, oAm, oAp, oB, oBm, oBp, col.x1, col.xm1, col.xp1, bits, c0, rA, rB, gc, gs, half_kappa,
                             hk2, hkA, hkB, n_acc, sum_A, dn_lut);
}

// The 81 proposals of (dn_f0, dn_b0, dn_f1, dn_b1) at interval_n = 1 as the residual changes they make, -2 pi W digit (exact in
// fp32 for digits 0, 1, 2), indexed by the code 27 d0 + 9 d1 + 3 d2 + d3 (svb_villain_filtered.cuh).  Call before a block barrier.
__device__ __forceinline__ void stream_fill_lut(float4* dn_lut, float c, int tid, int threads) {
    for (int i = tid; i < 81; i += threads)
        dn_lut[i] = make_float4(-c * (float)(i / 27), -c * (float)((i / 9) % 3), -c * (float)((i / 3) % 3), -c * (float)(i % 3));
}

// sum (dn)^2 of the plaquettes at the four sites (r, 2k), (r, 2k + 1), (r + 8, 2k), (r + 8, 2k + 1); reads n only.  Must not run
// concurrently with a colour pass on the same lattice.
__device__ __forceinline__ long long stream_dn2_rows(const int32_t* gn0, const int32_t* gn1, int N, int r, int k) {
    const int x2 = (2 * k + 2 == N) ? 0 : 2 * k + 2;
    long long t = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int rr = r + 8 * h, o = rr * N, op = ((rr + 1 == N) ? 0 : rr + 1) * N;
        const int2 m0 = *reinterpret_cast<const int2*>(gn0 + o + 2 * k), m1 = *reinterpret_cast<const int2*>(gn1 + o + 2 * k);
        const int2 up = *reinterpret_cast<const int2*>(gn1 + op + 2 * k);
        const int hr = gn0[o + x2];
        // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x])      (compact.py d,1 rows)
        const int d0 = (up.x - m1.x) - (m0.y - m0.x), d1 = (up.y - m1.y) - (hr - m0.y);
        t += (long long)d0 * d0 + (long long)d1 * d1;
    }
    return t;
}

// Observables of the CURRENT state from the four sites (r, 2k), (r, 2k + 1), (r + 8, 2k), (r + 8, 2k + 1): forward links and the
// plaquette of each (the full record of a launch that does not use obs_in).
__device__ __forceinline__ void stream_obs_rows(const double* gphi, const int32_t* gn0, const int32_t* gn1, int N, int r, int k,
                                                double& action, long long& dn2, int& w0, int& w1) {
    const int x2 = (2 * k + 2 == N) ? 0 : 2 * k + 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int rr = r + 8 * h, o = rr * N, op = ((rr + 1 == N) ? 0 : rr + 1) * N;
        const double2 pc = *reinterpret_cast<const double2*>(gphi + o + 2 * k);
        const double2 pu = *reinterpret_cast<const double2*>(gphi + op + 2 * k);
        const double pr = gphi[o + x2];
        const int2 m0 = *reinterpret_cast<const int2*>(gn0 + o + 2 * k), m1 = *reinterpret_cast<const int2*>(gn1 + o + 2 * k);
        const int2 up = *reinterpret_cast<const int2*>(gn1 + op + 2 * k);
        const int hr = gn0[o + x2];
        const double r0e = fma(-SVB_TWO_PI, SVB_FILT_CVT(m0.x), pu.x - pc.x), r0o = fma(-SVB_TWO_PI, SVB_FILT_CVT(m0.y), pu.y - pc.y);
        const double r1e = fma(-SVB_TWO_PI, SVB_FILT_CVT(m1.x), pc.y - pc.x), r1o = fma(-SVB_TWO_PI, SVB_FILT_CVT(m1.y), pr - pc.y);
        action = fma(r0e, r0e, action);
        action = fma(r0o, r0o, action);
        action = fma(r1e, r1e, action);
        action = fma(r1o, r1o, action);
        const int d0 = (up.x - m1.x) - (m0.y - m0.x), d1 = (up.y - m1.y) - (hr - m0.y);
        dn2 += (long long)d0 * d0 + (long long)d1 * d1;
        w0 += m0.x + m0.y;
        w1 += m1.x + m1.y;
    }
}

// ------------------------------------------------------------------------------------------
// One chain per CTA at a time (N in {16, 32, 64, 128}), persistent CTAs striding over the chains.  4 N threads: thread
// (w8, k) = (tid / (N/2), tid % (N/2)) owns column slot k and walks the rows 16 g + w8 (and + 8), g < N / 16.
// OVERLAP: the overlapped-launch protocol of svb_villain_sweep_overlapped (svb_villain_filtered.cuh): programmatic dependent
// launch, a chain is swept only once its epoch reads a.wait_epoch, all of a CTA's chains are released together at its end.
// ------------------------------------------------------------------------------------------
template <int NT, int MINB, bool OVERLAP, bool UNIT>
__global__ void __launch_bounds__(4 * NT, MINB) villain_stream_chain_kernel(const __grid_constant__ VillainArgs a,
                                                                            const __grid_constant__ FilterConsts fc) {
    constexpr int N = NT, V = N * N, HN = N / 2, THREADS = 4 * NT, NW = THREADS / 32, GROUPS = N / 16;
    static_assert(N % 16 == 0, "villain_stream_chain_kernel: unsupported geometry");
    __shared__ double red_state[4 * NW];
    __shared__ double red_count[2 * NW];
    __shared__ float4 dn_lut[81];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWriter = (NW > 1) ? 32 : 0;
    const int w8 = tid / HN, k = tid - w8 * HN;
    if (UNIT) stream_fill_lut(dn_lut, fc.c, tid, THREADS);          // (the barrier below orders it)

    if (OVERLAP) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (a.grid_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const bool obs_of_input = a.obs_in != nullptr;
    const bool want_obs = a.obs != nullptr && !obs_of_input;

    auto wait_epoch = [&](long long chain) {
        if (OVERLAP && !a.grid_wait) {
            uint32_t e;
            unsigned ns = 32, naps = 0;
            while (true) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(a.epochs + chain) : "memory");
                if (e == a.wait_epoch) break;
                __nanosleep(ns);
                if (ns < 1024) ns *= 2;
                if (++naps > (1u << 21)) __trap();          // a producer that never comes is a caller error: fail, do not hang
            }
        }
    };
    auto prefetch = [&](long long chain) {
        // the whole chain into L2 (it is the coherence point: whatever is prefetched there is current)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const double*>(a.phi) + chain * V),
                     "r"((uint32_t)(V * sizeof(double)))
                     : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.n + chain * 2 * V), "r"((uint32_t)(2 * V * sizeof(int32_t))) : "memory");
    };

    long long chain = blockIdx.x;
    if (tid == 0 && chain < a.chains) wait_epoch(chain);
    __syncthreads();

    int it = 0;
    for (; chain < a.chains; chain += gridDim.x, ++it) {
        const long long next = chain + gridDim.x;
        if (tid == 0 && next < a.chains) prefetch(next);
        double* gphi = reinterpret_cast<double*>(a.phi) + chain * V;
        int32_t* gn0 = a.n + chain * 2 * V;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double half_kappa = kappa / 2;
        const float hk2 = (float)(half_kappa * 1.4426950408889634);                 // decisions are taken in units of ln 2
        const float hkA = 1.0001f * hk2 * fc.bA, hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
        const unsigned long long gc = a.chain0 + (unsigned long long)chain;

        double action = 0.0, sum_A_all = 0.0;
        long long dn2 = 0;
        int w0 = 0, w1 = 0, n_acc = 0;
        float sum_A = 0.0f;
        if (obs_of_input) {
            // sum (dn)^2 of the arriving state: every thread reads n here, nobody writes it before the barrier
#pragma unroll 1
            for (int g = 0; g < GROUPS; ++g) dn2 += stream_dn2_rows(gn0, gn0 + V, N, 16 * g + w8, k);
            __syncthreads();
        }
        for (int sw = 0; sw < a.n_sweeps; ++sw) {
            const unsigned long long gs = a.sweep0 + (unsigned long long)sw;
            const bool sums = obs_of_input && sw == 0;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const StreamCol col = stream_col(k, (w8 + c) & 1, N);
#pragma unroll 1
                for (int g = 0; g < GROUPS; ++g) {
                    const int r = 16 * g + w8;
                    const int oA = r * N, oAm = ((r == 0) ? N - 1 : r - 1) * N, oB = oA + 8 * N, oBp = ((r + 9 == N) ? 0 : r + 9) * N;
                    if (c == 0 && sums)
                        stream_pair<UNIT, true>(a, fc, gphi, gn0, V, oA, oAm, oA + N, oB, oB - N, oBp, col, gc, gs, half_kappa, hk2, hkA, hkB,
                                                n_acc, sum_A, action, w0, w1, dn_lut);
                    else
                        stream_pair<UNIT, false>(a, fc, gphi, gn0, V, oA, oAm, oA + N, oB, oB - N, oBp, col, gc, gs, half_kappa, hk2, hkA, hkB,
                                                 n_acc, sum_A, action, w0, w1, dn_lut);
                }
                if (c == 0 && sums) chain_partials<true, false>(red_state, red_count, lane, warp, action, dn2, w0, w1, 0.0, 0);
                const bool last = (sw == a.n_sweeps - 1) && c == 1;
                if (last && !want_obs) {
                    if (a.obs != nullptr)
                        chain_partials<false, true>(red_state, red_count, lane, warp, 0.0, 0, 0, 0, sum_A_all + (double)sum_A, n_acc);
                    // the next chain may be swept once its epoch has been seen: one thread looks before the barrier
                    if (tid == 0 && next < a.chains) wait_epoch(next);
                }
                __syncthreads();               // this colour's writes are visible to the whole CTA
                if (c == 0 && sums && tid == kWriter)
                    chain_finish<NW, true, false>(red_state, red_count, half_kappa, a.obs_in + chain * SVB_VOBS_COUNT, nullptr);
            }
            sum_A_all += (double)sum_A;
            sum_A = 0.0f;
        }
        if (want_obs) {
            // the full record of the state after the sweeps: one more pass over the chain (L1 / L2)
            action = 0.0; dn2 = 0; w0 = 0; w1 = 0;
#pragma unroll 1
            for (int g = 0; g < GROUPS; ++g) stream_obs_rows(gphi, gn0, gn0 + V, N, 16 * g + w8, k, action, dn2, w0, w1);
            chain_partials<true, true>(red_state, red_count, lane, warp, action, dn2, w0, w1, sum_A_all, n_acc);
            if (tid == 0 && next < a.chains) wait_epoch(next);
            __syncthreads();
        }
        if (tid == kWriter && a.obs != nullptr) {
            double* row = a.obs + chain * SVB_VOBS_COUNT;
            if (want_obs) chain_finish<NW, true, true>(red_state, red_count, half_kappa, row, row);
            else chain_finish<NW, false, true>(red_state, red_count, half_kappa, nullptr, row);
        }
    }
    if (OVERLAP) {
        __syncthreads();                                   // every write of every thread, every record
        if (warp == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (int i = lane; i < it; i += 32)
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.epochs + blockIdx.x + (long long)i * gridDim.x), "r"(a.signal_epoch)
                             : "memory");
        }
    }
}

template <int NT, int MINB>
static int launch_villain_stream_chain(const VillainArgs& a, cudaStream_t stream, const DeviceInfo& info) {
    const bool overlap = a.epochs != nullptr;
    const bool unit = a.W == 1 && a.interval_n == 1;
    constexpr int THREADS = 4 * NT;
    auto kern = overlap ? (unit ? villain_stream_chain_kernel<NT, MINB, true, true> : villain_stream_chain_kernel<NT, MINB, true, false>)
                        : (unit ? villain_stream_chain_kernel<NT, MINB, false, true> : villain_stream_chain_kernel<NT, MINB, false, false>);
    static int per_sm_cache[4][64];
    const int variant = (overlap ? 1 : 0) + (unit ? 2 : 0);
    int per_sm = (info.device < 64) ? per_sm_cache[variant][info.device] : 0;
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "streaming villain kernel does not fit an SM at N=%d", NT);
        if (info.device < 64) per_sm_cache[variant][info.device] = per_sm;
    }
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > a.chains) grid = a.chains;
    const FilterConsts fc = make_filter_consts(a.interval_phi, a.W, a.interval_n);
    if (overlap) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SVB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a, fc));
        return 0;
    }
    kern<<<(unsigned)grid, THREADS, 0, stream>>>(a, fc);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// One colour pass per launch for lattices beyond a CTA (config 5, L = 4096), IN PLACE: a pass writes only accepted proposals
// and nothing it writes is read by another thread of the same pass.  A CTA of 8 warps owns a band of 64 columns (warp w8 = row
// 16 g + w8 and its partner + 8, lane = column slot) and `run` consecutive row groups g; one CTA per (chain, band, run), so the
// hardware scheduler balances the load.  Records: per-CTA partial sums added atomically (a handful of atomics per CTA).
//   pass 0 with state_out: the state columns ACTION, WRAP0, WRAP1 of the lattice AS IT ARRIVES; both passes: counters.
//   sum (dn)^2 of the arriving state: villain_stream_dn2_kernel (reads n only), launched before pass 0.
// ------------------------------------------------------------------------------------------
template <bool UNIT>
__global__ void __launch_bounds__(256, 3) villain_stream_pass_kernel(const __grid_constant__ VillainArgs a, const __grid_constant__ FilterConsts fc,
                                                                     int colour, int sweep, int bands, int runs, int run,
                                                                     double* __restrict__ state_out, double* __restrict__ counter_out) {
    __shared__ double scratch[5 * 32];
    __shared__ float4 dn_lut[81];
    const int N = a.N, V = N * N, groups = N / 16;
    const int tid = threadIdx.x, w8 = tid >> 5, lane = tid & 31;
    if (UNIT) {
        stream_fill_lut(dn_lut, fc.c, tid, 256);
        __syncthreads();
    }
    pdl_prologue();
    // blockIdx.x = (chain * runs + run index) * bands + band: neighbouring CTAs work on neighbouring bands of the same rows
    const unsigned per_chain = (unsigned)(bands * runs);
    const long long chain = blockIdx.x / per_chain;
    const int rem = (int)(blockIdx.x - chain * per_chain);
    const int ri = rem / bands, band = rem - ri * bands;
    const int k = 32 * band + lane;
    const bool sums = state_out != nullptr;
    const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
    const double half_kappa = kappa / 2;
    const float hk2 = (float)(half_kappa * 1.4426950408889634);
    const float hkA = 1.0001f * hk2 * fc.bA, hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
    const unsigned long long gc = a.chain0 + (unsigned long long)chain, gs = a.sweep0 + (unsigned long long)sweep;
    double* gphi = reinterpret_cast<double*>(a.phi) + chain * (long long)V;
    int32_t* gn0 = a.n + chain * 2 * (long long)V;
    double action = 0.0;
    int w0 = 0, w1 = 0, n_acc = 0;
    float sum_A = 0.0f;
    if (k < N / 2) {                                              // N not a multiple of 64: the last band is ragged
        const StreamCol col = stream_col(k, (w8 + colour) & 1, N);
        const int g_end = min(groups, (ri + 1) * run);
#pragma unroll 1
        for (int g = ri * run; g < g_end; ++g) {
            const int r = 16 * g + w8;
            const int oA = r * N, oAm = ((r == 0) ? N - 1 : r - 1) * N, oB = oA + 8 * N, oBp = ((r + 9 == N) ? 0 : r + 9) * N;
            if (sums)
                stream_pair<UNIT, true>(a, fc, gphi, gn0, V, oA, oAm, oA + N, oB, oB - N, oBp, col, gc, gs, half_kappa, hk2, hkA, hkB, n_acc,
                                        sum_A, action, w0, w1, dn_lut);
            else
                stream_pair<UNIT, false>(a, fc, gphi, gn0, V, oA, oAm, oA + N, oB, oB - N, oBp, col, gc, gs, half_kappa, hk2, hkA, hkB, n_acc,
                                         sum_A, action, w0, w1, dn_lut);
        }
    }
    if (sums || counter_out) {
        double v[5] = {action, (double)w0, (double)w1, (double)n_acc, (double)sum_A};
        block_sum<5>(v, scratch);
        if (tid == 0) {
            if (sums) {
                double* o = state_out + chain * SVB_VOBS_COUNT;
                atomicAdd(o + SVB_VOBS_ACTION, half_kappa * v[0]);
                atomicAdd(o + SVB_VOBS_WRAP0, v[1]);
                atomicAdd(o + SVB_VOBS_WRAP1, v[2]);
            }
            if (counter_out) {
                atomicAdd(counter_out + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, v[3]);
                atomicAdd(counter_out + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, v[4]);
            }
        }
    }
}

// sum (dn)^2 of every chain, reading n once: a warp owns 128 columns (a lane four: 16-byte loads, 512 contiguous bytes per warp and
// row) and walks down `rows` consecutive rows; a lane keeps n1[x] of the current row and needs only the next row's n1 (which it
// keeps for the next step) and its right-hand neighbour's first n0 (a shuffle; the last lane of a band loads it).
// (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x]).  Grid: chains x bands x row runs; per-warp partial sums added atomically (the
// caller zeroes SVB_VOBS_SUM_DN2).
__global__ void __launch_bounds__(256) villain_stream_dn2_kernel(const int32_t* __restrict__ n, long long chains, int N, int bands,
                                                                 int runs, int rows, double* __restrict__ state_out) {
    const int tid = threadIdx.x, lane = tid & 31;
    pdl_prologue();
    const long long w = (long long)blockIdx.x * 8 + (tid >> 5);                  // global warp index = (chain, run, band)
    const long long per_chain = (long long)bands * runs;
    if (w >= chains * per_chain) return;
    const long long chain = w / per_chain;
    const int rem = (int)(w - chain * per_chain);
    const int run = rem / bands, band = rem - run * bands;
    const int q = 32 * band + lane;                                               // column quad [4 q, 4 q + 3]
    const long long V = (long long)N * N;
    const int32_t* gn0 = n + chain * 2 * V;
    const int32_t* gn1 = gn0 + V;
    long long dn2 = 0;
    const bool active = q < N / 4;
    const int c = active ? 4 * q : 0;
    const int cr = (c + 4 == N) ? 0 : c + 4;                                       // the column right of the quad, wrapped
    const bool edge = lane == 31 || q + 1 >= N / 4;
    const int r_lo = run * rows, r_hi = min(N, r_lo + rows);
    int4 m1 = *reinterpret_cast<const int4*>(gn1 + (long long)r_lo * N + c);
    int r = r_lo;
    // four rows at a time: their eight (nine) loads are in flight together -- a warp has little else to hide DRAM latency with
    for (; r + 4 <= r_hi; r += 4) {
        int4 m0[4], up[4];
        int hrq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rn = (r + i + 1 == N) ? 0 : r + i + 1;
            m0[i] = *reinterpret_cast<const int4*>(gn0 + (long long)(r + i) * N + c);
            up[i] = *reinterpret_cast<const int4*>(gn1 + (long long)rn * N + c);
            hrq[i] = edge ? gn0[(long long)(r + i) * N + cr] : 0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int sh = __shfl_down_sync(0xffffffffu, m0[i].x, 1);
            const int hr = edge ? hrq[i] : sh;
            const int d0 = (up[i].x - m1.x) - (m0[i].y - m0[i].x), d1 = (up[i].y - m1.y) - (m0[i].z - m0[i].y);
            const int d2 = (up[i].z - m1.z) - (m0[i].w - m0[i].z), d3 = (up[i].w - m1.w) - (hr - m0[i].w);
            if (active) dn2 += (long long)d0 * d0 + (long long)d1 * d1 + (long long)d2 * d2 + (long long)d3 * d3;
            m1 = up[i];
        }
    }
    for (; r < r_hi; ++r) {
        const int rn = (r + 1 == N) ? 0 : r + 1;
        const int4 m0 = *reinterpret_cast<const int4*>(gn0 + (long long)r * N + c);
        const int4 up = *reinterpret_cast<const int4*>(gn1 + (long long)rn * N + c);
        int hr = __shfl_down_sync(0xffffffffu, m0.x, 1);
        if (edge) hr = gn0[(long long)r * N + cr];
        const int d0 = (up.x - m1.x) - (m0.y - m0.x), d1 = (up.y - m1.y) - (m0.z - m0.y);
        const int d2 = (up.z - m1.z) - (m0.w - m0.z), d3 = (up.w - m1.w) - (hr - m0.w);
        if (active) dn2 += (long long)d0 * d0 + (long long)d1 * d1 + (long long)d2 * d2 + (long long)d3 * d3;
        m1 = up;
    }
    dn2 = warp_sum(dn2);
    if (lane == 0) atomicAdd(state_out + chain * SVB_VOBS_COUNT + SVB_VOBS_SUM_DN2, (double)dn2);
}

static int launch_villain_stream_dn2(const int32_t* n, long long chains, int N, double* state_out, cudaStream_t stream, const DeviceInfo& info) {
    const int bands = (N / 4 + 31) / 32;
    // rows per warp: long runs amortise the set-up and the one redundant row, short runs fill the machine
    int rows = N;
    while (rows > 16 && chains * bands * ((N + rows - 1) / rows) < 4LL * 64 * info.sm_count) rows = (rows + 1) / 2;
    const int runs = (N + rows - 1) / rows;
    const long long warps = chains * bands * runs;
    SVB_CUDA_TRY(launch_pdl(villain_stream_dn2_kernel, (unsigned)((warps + 7) / 8), 256, 0, stream, n, chains, N, bands, runs, rows, state_out));
    return 0;
}

// n_sweeps sweeps in place; obs_in (optional): the state columns of the arriving lattice (zeroed by the caller), filled by a pass
// over n and the first colour pass; counters (optional): ACCEPTED / ACCEPTANCE of this call are ADDED to it.
static int launch_villain_stream_passes(const VillainArgs& a, double* obs_in, double* counters, cudaStream_t stream, const DeviceInfo& info) {
    const bool unit = a.W == 1 && a.interval_n == 1;
    auto kern = unit ? villain_stream_pass_kernel<true> : villain_stream_pass_kernel<false>;
    static bool ready[2][64];
    if (info.device >= 64 || !ready[unit ? 1 : 0][info.device]) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
        if (info.device < 64) ready[unit ? 1 : 0][info.device] = true;
    }
    const int bands = (a.N / 2 + 31) / 32, groups = a.N / 16;
    // rows per CTA: long runs amortise a thread's column set-up, short runs balance the load -- aim at >= 8 CTAs per SM slot
    int run = groups;
    while (run > 2 && (long long)a.chains * bands * ((groups + run - 1) / run) < 24LL * info.sm_count) run = (run + 1) / 2;
    const int runs = (groups + run - 1) / run;
    const long long grid = (long long)a.chains * bands * runs;
    if (grid > 0x7fffffffLL) return fail(SVB_E_SHAPE, "streaming villain passes: too many CTAs (%lld)", grid);
    const FilterConsts fc = make_filter_consts(a.interval_phi, a.W, a.interval_n);
    if (obs_in) {
        const int rc_dn2 = launch_villain_stream_dn2(a.n, a.chains, a.N, obs_in, stream, info);
        if (rc_dn2) return rc_dn2;
    }
    for (int sw = 0; sw < a.n_sweeps; ++sw)
        for (int c = 0; c < 2; ++c) {
            SVB_CUDA_TRY(launch_pdl(kern, (unsigned)grid, 256, 0, stream, a, fc, c, sw, bands, runs, run, (sw == 0 && c == 0) ? obs_in : nullptr,
                                    counters));
        }
    return 0;
}

// Observables of the current state for lattices that do not fit a CTA (N a multiple of 16, fp64): a thread takes the four
// sites (r, 2k), (r, 2k + 1), (r + 8, 2k), (r + 8, 2k + 1) with 16-byte loads (stream_obs_rows), CTAs stride over chunks of 256
// such quads, block partials are added atomically (the caller zeroes the state columns).  One read of the state.
__global__ void __launch_bounds__(256) villain_stream_obs_kernel(const double* __restrict__ phi, const int32_t* __restrict__ n,
                                                                 long long chains, int N, double kappa_scalar,
                                                                 const double* __restrict__ kappa_chain, double* __restrict__ obs) {
    __shared__ double scratch[4 * 32];
    const long long V = (long long)N * N;
    const int HN = N / 2;
    const long long quads_per_chain = (long long)(N / 2) * HN;
    const int tid = threadIdx.x;
    const int chunks = (int)((quads_per_chain + 255) / 256);
    long long chain_of_sum = -1;
    double action = 0.0;
    long long dn2 = 0;
    int w0 = 0, w1 = 0;
    auto flush = [&](long long chain) {
        double v[4] = {action, (double)dn2, (double)w0, (double)w1};
        block_sum<4>(v, scratch);
        if (tid == 0) {
            const double kappa = kappa_chain ? kappa_chain[chain] : kappa_scalar;
            double* o = obs + chain * SVB_VOBS_COUNT;
            atomicAdd(o + SVB_VOBS_ACTION, (kappa / 2) * v[0]);
            atomicAdd(o + SVB_VOBS_SUM_DN2, v[1]);
            atomicAdd(o + SVB_VOBS_WRAP0, v[2]);
            atomicAdd(o + SVB_VOBS_WRAP1, v[3]);
        }
        __syncthreads();
        action = 0.0; dn2 = 0; w0 = 0; w1 = 0;
    };
    for (long long t = blockIdx.x; t < chains * chunks; t += gridDim.x) {
        const long long chain = t / chunks;
        if (chain != chain_of_sum) {
            if (chain_of_sum >= 0) flush(chain_of_sum);
            chain_of_sum = chain;
        }
        const long long q = (t - chain * chunks) * 256 + tid;
        if (q < quads_per_chain) {
            const int rp = (int)(q / HN), k = (int)(q - (long long)rp * HN);
            stream_obs_rows(phi + chain * V, n + chain * 2 * V, n + chain * 2 * V + V, N, (rp & 7) | ((rp >> 3) << 4), k, action, dn2, w0, w1);
        }
    }
    if (chain_of_sum >= 0) flush(chain_of_sum);
}

// ------------------------------------------------------------------------------------------
// The same colour pass with the tile STAGED IN SHARED MEMORY BY TMA TENSOR LOADS (config 5: L = 4096; N a multiple of 128).
//
// A pass over a big lattice out of global memory is latency bound: a thread's nine loads per site come from DRAM, and the
// registers that would hold a second iteration's loads are what limits occupancy.  Here the copy engine does the waiting:
//   * a tile is 16 rows x 128 columns (1024 sites of the colour); three 3-D tensor boxes (cp.async.bulk.tensor, SASS UTMALDG)
//     bring phi (18 rows x 132 columns: one halo row above and below, the neighbour columns left and right), n0 and n1
//     (17 x 132: the row above for the backward links, the column to the left) into one shared-memory stage and signal an
//     mbarrier; the tensor maps describe the fields as (column, row, chain [x component]), so a halo row or column beyond the
//     edge of the lattice is out of bounds and zero-filled, never a neighbouring chain's data;
//   * two stages per CTA: the boxes of the tile after next are issued as soon as the CTA has finished reading a stage, and land
//     while the next tile is being swept -- the load latency is hidden whatever the occupancy;
//   * the periodic wrap touches only the tiles on the edge of the lattice (2 of 32 columns of tiles, 2 of 256 rows at L = 4096):
//     their halo row / column is patched with ordinary loads behind the mbarrier;
//   * persistent CTAs (3 per SM) stride over the tiles; a thread owns two column slots of a tile and the row pair (w8, w8 + 8);
//   * everything after the residuals is stream_pair_decide: accepted proposals go to GLOBAL memory as reductions, in place --
//     no workspace, no ping-pong, no ghost-zone recomputation, nothing stored that did not change.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void tensor_box_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// residuals of the site at local row i (phi / n row index: 0 = the halo row R - 1), column slot base jc (even), parity par
template <bool SUMS>
__device__ __forceinline__ float4 tile_site_residuals(const double* P, const int32_t* N0, const int32_t* N1, int i, int jc, int par,
                                                      double& action, int& w0, int& w1) {
    const double* prow = P + i * kTilePhiCols + jc + 2;             // local column 0 is global C - 2
    const int32_t* n0row = N0 + i * kTileNCols + jc + 4;            // local column 0 is global C - 4
    const int32_t* n1row = N1 + i * kTileNCols + jc + 4;
    const double2 pc2 = *reinterpret_cast<const double2*>(prow);
    const double po = prow[par ? 2 : -1];
    const double pu = prow[kTilePhiCols + par], pd = prow[par - kTilePhiCols];
    const int2 n1p = *reinterpret_cast<const int2*>(n1row);
    const int n1e = n1row[-1];
    const int n0c = n0row[par], n0b = n0row[par - kTileNCols];
    const double pc = par ? pc2.y : pc2.x;
    const double pl = par ? pc2.x : po, pr = par ? po : pc2.y;
    const int n1c = par ? n1p.y : n1p.x, n1b = par ? n1p.x : n1e;
    const double rf0 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n0c), pu - pc);
    const double rb0 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n0b), pc - pd);
    const double rf1 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n1c), pr - pc);
    const double rb1 = fma(-SVB_TWO_PI, SVB_FILT_CVT(n1b), pc - pl);
    if (SUMS) {
        action = fma(rf0, rf0, action);
        action = fma(rb0, rb0, action);
        action = fma(rf1, rf1, action);
        action = fma(rb1, rb1, action);
        w0 += n0c + n0b;
        w1 += n1c + n1b;
    }
    return make_float4((float)rf0, (float)rb0, (float)rf1, (float)rb1);
}

template <bool UNIT>
__global__ void __launch_bounds__(256, 3) villain_tile_pass_kernel(const __grid_constant__ VillainArgs a, const __grid_constant__ FilterConsts fc,
                                                                   const __grid_constant__ CUtensorMap map_phi,
                                                                   const __grid_constant__ CUtensorMap map_n, int colour, int sweep,
                                                                   double* __restrict__ state_out, double* __restrict__ counter_out) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ double scratch[8 * 5];         // per-warp partial sums of the records (shared memory is what limits this kernel to 3 CTAs per SM)
    __shared__ int tile_coord[2][4];          // per stage: chain, R, C of the tile that was loaded into it (written by thread 0)
    const int N = a.N, V = N * N;
    const int tiles_x = N / kTileCols, tiles_y = N / kTileRows, tiles_per_chain = tiles_x * tiles_y;
    const long long tiles = (long long)tiles_per_chain * a.chains;
    const int tid = threadIdx.x, w8 = tid >> 5, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile_smem + 2 * kTileStageBytes);
    const bool sums = state_out != nullptr;
    const unsigned long long gs = a.sweep0 + (unsigned long long)sweep;
    const int par = (w8 + colour) & 1;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    // (three CTAs of 75.6 KiB still fit an SM: the table costs no occupancy, and the look-up saves the arithmetic decode of the
    // digits on the hot path -- 161.2 -> 160.3 us per config-5 step)
    float4* dn_lut = reinterpret_cast<float4*>(tile_smem + 2 * kTileStageBytes + 64);
    if (UNIT) stream_fill_lut(dn_lut, fc.c, tid, 256);
    __syncthreads();
    pdl_prologue();
    auto issue = [&](long long t, int stage) {
        // (32-bit arithmetic: the tiles of a call are fewer than 2^32 -- a 64-bit division by one thread cost this kernel 6 % of
        // its issue slots, every other thread of the warp sitting it out)
        const unsigned tu = (unsigned)t, tpc = (unsigned)tiles_per_chain;
        const unsigned chain = tu / tpc, tile = tu - chain * tpc;
        const unsigned ty = tile / (unsigned)tiles_x, tx = tile - ty * (unsigned)tiles_x;
        const int R = (int)ty * kTileRows, C = (int)tx * kTileCols;
        tile_coord[stage][0] = (int)chain; tile_coord[stage][1] = R; tile_coord[stage][2] = C;
        unsigned char* st = tile_smem + stage * kTileStageBytes;
        mbar_expect_tx(&bar[stage], (uint32_t)(kTilePhiBytes + 2 * kTileNBytes));
        tensor_box_3d(st, &map_phi, C - 2, R - 1, (int)chain, &bar[stage]);
        tensor_box_3d(st + kTilePhiSlot, &map_n, C - 4, R - 1, (int)(2 * chain), &bar[stage]);
        tensor_box_3d(st + kTilePhiSlot + kTileNSlot, &map_n, C - 4, R - 1, (int)(2 * chain + 1), &bar[stage]);
    };
    long long t = blockIdx.x;
    if (tid == 0) {
        if (t < tiles) issue(t, 0);
        if (t + gridDim.x < tiles) issue(t + gridDim.x, 1);
    }
    __syncthreads();                                               // the coordinates of the first two tiles are published

    double action = 0.0, half_kappa = 0.0;
    int w0 = 0, w1 = 0, n_acc = 0;
    float sum_A = 0.0f, hk2 = 0.0f, hkA = 0.0f, hkB = 0.0f;
    long long chain_of_sums = -1;
    auto flush = [&](long long chain) {
        double v[5] = {action, (double)w0, (double)w1, (double)n_acc, (double)sum_A};
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = warp_sum(v[i]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) scratch[5 * w8 + i] = v[i];
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                v[i] = 0.0;
                for (int w = 0; w < 8; ++w) v[i] += scratch[5 * w + i];
            }
            if (sums) {
                double* o = state_out + chain * SVB_VOBS_COUNT;
                atomicAdd(o + SVB_VOBS_ACTION, half_kappa * v[0]);
                atomicAdd(o + SVB_VOBS_WRAP0, v[1]);
                atomicAdd(o + SVB_VOBS_WRAP1, v[2]);
            }
            if (counter_out) {
                atomicAdd(counter_out + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, v[3]);
                atomicAdd(counter_out + chain * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, v[4]);
            }
        }
        __syncthreads();
        action = 0.0; w0 = 0; w1 = 0; n_acc = 0; sum_A = 0.0f;
    };

    for (int it = 0; t < tiles; t += gridDim.x, ++it) {
        const int stage = it & 1;
        // (one thread did the divisions when it issued the loads; everybody else reads the result)
        const long long chain = tile_coord[stage][0];
        const int R = tile_coord[stage][1], C = tile_coord[stage][2];
        if (chain != chain_of_sums) {
            if (chain_of_sums >= 0 && (sums || counter_out)) flush(chain_of_sums);
            chain_of_sums = chain;
            const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
            half_kappa = kappa / 2;
            hk2 = (float)(half_kappa * 1.4426950408889634);
            hkA = 1.0001f * hk2 * fc.bA; hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
        }
        double* gphi = reinterpret_cast<double*>(a.phi) + chain * (long long)V;
        int32_t* gn0 = a.n + chain * 2 * (long long)V;
        unsigned char* st = tile_smem + stage * kTileStageBytes;
        double* P = reinterpret_cast<double*>(st);
        int32_t* N0 = reinterpret_cast<int32_t*>(st + kTilePhiSlot);
        int32_t* N1 = reinterpret_cast<int32_t*>(st + kTilePhiSlot + kTileNSlot);

        mbar_wait(&bar[stage], (uint32_t)((it >> 1) & 1));
        const bool top = R == 0, bottom = R + kTileRows == N, left = C == 0, right = C + kTileCols == N;
        if (top || bottom || left || right) {
            // the periodic wrap: what lies beyond the edge of the lattice arrived as zeros; patch it from the other side
            if (top && tid < kTileCols) {
                P[2 + tid] = gphi[(N - 1) * N + C + tid];                                  // halo row R - 1 = row N - 1
                N0[4 + tid] = gn0[(N - 1) * N + C + tid];
            }
            if (bottom && tid >= 128 && tid < 128 + kTileCols) P[(kTilePhiRows - 1) * kTilePhiCols + 2 + (tid - 128)] = gphi[C + (tid - 128)];
            if (left && tid < kTileRows) {
                P[(tid + 1) * kTilePhiCols + 1] = gphi[(R + tid) * N + N - 1];            // column C - 1 = column N - 1
                N1[(tid + 1) * kTileNCols + 3] = gn0[V + (R + tid) * N + N - 1];
            }
            if (right && tid >= 32 && tid < 32 + kTileRows) P[(tid - 31) * kTilePhiCols + 2 + kTileCols] = gphi[(R + tid - 32) * N];   // column N = 0
            __syncthreads();
        }

        const unsigned long long gc = a.chain0 + (unsigned long long)chain;
        const int r = R + w8;
        const int oA = r * N, oAm = ((r == 0) ? N - 1 : r - 1) * N, oB = oA + 8 * N, oBp = ((r + 9 == N) ? 0 : r + 9) * N;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int jc = 2 * (lane + 32 * h);
            const int x1 = C + jc + par;
            const int xm1 = (x1 == 0) ? N - 1 : x1 - 1, xp1 = (x1 + 1 == N) ? 0 : x1 + 1;
            const uint32_t c0 = (uint32_t)(oA + x1);
            const Philox4 bits = philox_site_keys(a, gc, gs, c0);
            float4 rA, rB;
            if (sums) {
                rA = tile_site_residuals<true>(P, N0, N1, w8 + 1, jc, par, action, w0, w1);
                rB = tile_site_residuals<true>(P, N0, N1, w8 + 9, jc, par, action, w0, w1);
            } else {
                rA = tile_site_residuals<false>(P, N0, N1, w8 + 1, jc, par, action, w0, w1);
                rB = tile_site_residuals<false>(P, N0, N1, w8 + 9, jc, par, action, w0, w1);
            }
            stream_pair_decide<UNIT, true>(a, fc, gphi, gn0, V, oA, oAm, oA + N, oB, oB - N, oBp, x1, xm1, xp1, bits, c0, rA, rB, gc, gs,
                                           half_kappa, hk2, hkA, hkB, n_acc, sum_A, dn_lut);
        }
        __syncthreads();                                           // every thread has read the stage (and its coordinates): refill it
        const long long t2 = t + 2LL * gridDim.x;
        if (tid == 0 && t2 < tiles) issue(t2, stage);
        // (the coordinates thread 0 has just written belong to the iteration after next: a barrier -- the next iteration's --
        // lies between this write and their first read)
    }
    if (chain_of_sums >= 0 && (sums || counter_out)) flush(chain_of_sums);
}

// Tensor maps of the fields for the tile kernel: phi as (column, row, chain), n as (column, row, 2 chain + component).
typedef CUresult (*svb_tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int villain_tile_maps(const VillainArgs& a, CUtensorMap& map_phi, CUtensorMap& map_n) {
    static svb_tensor_map_encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult status;
        SVB_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status));
        if (status != cudaDriverEntryPointSuccess || !fn) return fail(SVB_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<svb_tensor_map_encode_fn>(fn);
    }
    const cuuint64_t N = (cuuint64_t)a.N;
    const cuuint32_t ones[3] = {1, 1, 1};
    {
        const cuuint64_t dims[3] = {N, N, (cuuint64_t)a.chains}, strides[2] = {N * 8, N * N * 8};
        const cuuint32_t box[3] = {(cuuint32_t)kTilePhiCols, (cuuint32_t)kTilePhiRows, 1};
        const CUresult r = encode(&map_phi, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, a.phi, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SVB_E_UNSUPPORTED, "cuTensorMapEncodeTiled(phi) failed: %d", (int)r);
    }
    {
        const cuuint64_t dims[3] = {N, N, 2 * (cuuint64_t)a.chains}, strides[2] = {N * 4, N * N * 4};
        const cuuint32_t box[3] = {(cuuint32_t)kTileNCols, (cuuint32_t)kTileNRows, 1};
        const CUresult r = encode(&map_n, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, a.n, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SVB_E_UNSUPPORTED, "cuTensorMapEncodeTiled(n) failed: %d", (int)r);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// sum (dn)^2 of big lattices (N a multiple of 128) as a TMA-fed stream: the pass over n that precedes the colour passes of a
// step with a record.  villain_stream_dn2_kernel's per-thread loads reach 3.5 TB/s (38 us of a 173 us step at L = 4096); here
// the copy engine does the streaming: a tile is 32 rows x 128 columns of plaquettes, its n0 and n1 arrive as two 3-D tensor
// boxes of 33 rows x 132 columns (the row below for n1, the column to the right for n0; beyond the edge of the lattice they
// are zero-filled and patched from the other side), three stages per CTA, persistent CTAs, a thread sums 16 plaquettes
// from conflict-free 16-byte shared-memory loads.
// ------------------------------------------------------------------------------------------
constexpr int kDn2Rows = 32, kDn2BoxRows = kDn2Rows + 1, kDn2BoxCols = kTileCols + 4;
constexpr int kDn2BoxBytes = kDn2BoxRows * kDn2BoxCols * 4;                            // 17424
constexpr int kDn2BoxSlot = (kDn2BoxBytes + 127) / 128 * 128;
#ifndef SVB_DN2_STAGES
/* two CTAs of 256 threads per SM with three stages each (config-5 step 161.4 us); one CTA per SM with six stages, five tiles
 * in flight: 164.2 us with 256 threads, 164.7 us with 512 */
#define SVB_DN2_STAGES 3
#define SVB_DN2_THREADS 256
#define SVB_DN2_MINB 2
#endif
constexpr int kDn2Stages = SVB_DN2_STAGES, kDn2StageBytes = 2 * kDn2BoxSlot, kDn2Threads = SVB_DN2_THREADS, kDn2Warps = kDn2Threads / 32;
constexpr int kDn2SmemBytes = kDn2Stages * kDn2StageBytes + 64;

__global__ void __launch_bounds__(kDn2Threads, SVB_DN2_MINB) villain_tile_dn2_kernel(const __grid_constant__ CUtensorMap map_n, const int32_t* __restrict__ n,
                                                                  long long chains, int N, double* __restrict__ state_out) {
    extern __shared__ __align__(128) unsigned char dn2_smem[];
    const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(dn2_smem + kDn2Stages * kDn2StageBytes);
    const int tiles_x = N / kTileCols, tiles_y = N / kDn2Rows, tiles_per_chain = tiles_x * tiles_y;
    const long long tiles = (long long)tiles_per_chain * chains;
    const long long V = (long long)N * N;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kDn2Stages; ++s) mbar_init(&bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    pdl_prologue();
    auto issue = [&](long long t, int stage) {
        const long long chain = t / tiles_per_chain;
        const int tile = (int)(t - chain * tiles_per_chain);
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        unsigned char* st = dn2_smem + stage * kDn2StageBytes;
        mbar_expect_tx(&bar[stage], (uint32_t)(2 * kDn2BoxBytes));
        tensor_box_3d(st, &map_n, tx * kTileCols, ty * kDn2Rows, (int)(2 * chain), &bar[stage]);
        tensor_box_3d(st + kDn2BoxSlot, &map_n, tx * kTileCols, ty * kDn2Rows, (int)(2 * chain + 1), &bar[stage]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kDn2Stages; ++s)
            if (blockIdx.x + (long long)s * gridDim.x < tiles) issue(blockIdx.x + (long long)s * gridDim.x, s);
    }
    long long dn2 = 0, chain_of_sum = -1;
    int stage = 0;
    uint32_t parity = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long chain = t / tiles_per_chain;
        const int tile = (int)(t - chain * tiles_per_chain);
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int R = ty * kDn2Rows, C = tx * kTileCols;
        if (chain != chain_of_sum) {
            if (chain_of_sum >= 0) {
                const long long v = warp_sum(dn2);
                if (lane == 0) atomicAdd(state_out + chain_of_sum * SVB_VOBS_COUNT + SVB_VOBS_SUM_DN2, (double)v);
                dn2 = 0;
            }
            chain_of_sum = chain;
        }
        int32_t* N0 = reinterpret_cast<int32_t*>(dn2_smem + stage * kDn2StageBytes);
        int32_t* N1 = reinterpret_cast<int32_t*>(dn2_smem + stage * kDn2StageBytes + kDn2BoxSlot);
        mbar_wait(&bar[stage], parity);
        const bool bottom = R + kDn2Rows == N, right = C + kTileCols == N;
        if (bottom || right) {
            const int32_t* gn0 = n + chain * 2 * V;
            if (bottom && tid < kTileCols) N1[kDn2Rows * kDn2BoxCols + tid] = gn0[V + C + tid];                              // row N = row 0
            if (right && tid >= 128 && tid < 128 + kDn2Rows) N0[(tid - 128) * kDn2BoxCols + kTileCols] = gn0[(long long)(R + tid - 128) * N];   // column N = 0
            __syncthreads();
        }
        // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x]); a thread: rows warp + kDn2Warps h, four columns
#pragma unroll
        for (int h = 0; h < kDn2Rows / kDn2Warps; ++h) {
            const int i = wrp + kDn2Warps * h;
            const int4 m0 = *reinterpret_cast<const int4*>(N0 + i * kDn2BoxCols + 4 * lane);
            const int hr = N0[i * kDn2BoxCols + 4 * lane + 4];
            const int4 m1 = *reinterpret_cast<const int4*>(N1 + i * kDn2BoxCols + 4 * lane);
            const int4 up = *reinterpret_cast<const int4*>(N1 + (i + 1) * kDn2BoxCols + 4 * lane);
            const int d0 = (up.x - m1.x) - (m0.y - m0.x), d1 = (up.y - m1.y) - (m0.z - m0.y);
            const int d2 = (up.z - m1.z) - (m0.w - m0.z), d3 = (up.w - m1.w) - (hr - m0.w);
            dn2 += (long long)d0 * d0 + (long long)d1 * d1 + (long long)d2 * d2 + (long long)d3 * d3;
        }
        __syncthreads();                                           // every thread has read the stage: refill it
        const long long t2 = t + (long long)kDn2Stages * gridDim.x;
        if (tid == 0 && t2 < tiles) issue(t2, stage);
        if (++stage == kDn2Stages) { stage = 0; parity ^= 1u; }
    }
    if (chain_of_sum >= 0) {
        const long long v = warp_sum(dn2);
        if (lane == 0) atomicAdd(state_out + chain_of_sum * SVB_VOBS_COUNT + SVB_VOBS_SUM_DN2, (double)v);
    }
}

static int launch_villain_tile_dn2(const int32_t* n, long long chains, int N, double* state_out, cudaStream_t stream, const DeviceInfo& info) {
    static svb_tensor_map_encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult status;
        SVB_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status));
        if (status != cudaDriverEntryPointSuccess || !fn) return fail(SVB_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<svb_tensor_map_encode_fn>(fn);
    }
    static int per_sm_cache[64];
    int per_sm = (info.device < 64) ? per_sm_cache[info.device] : 0;
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(villain_tile_dn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDn2SmemBytes));
        SVB_CUDA_TRY(cudaFuncSetAttribute(villain_tile_dn2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, villain_tile_dn2_kernel, kDn2Threads, kDn2SmemBytes));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "the (dn)^2 tile kernel does not fit an SM");
        if (info.device < 64) per_sm_cache[info.device] = per_sm;
    }
    CUtensorMap map_n;
    const cuuint64_t NN = (cuuint64_t)N;
    const cuuint32_t ones[3] = {1, 1, 1};
    const cuuint64_t dims[3] = {NN, NN, 2 * (cuuint64_t)chains}, strides[2] = {NN * 4, NN * NN * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kDn2BoxCols, (cuuint32_t)kDn2BoxRows, 1};
    const CUresult r = encode(&map_n, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<int32_t*>(n), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SVB_E_UNSUPPORTED, "cuTensorMapEncodeTiled(n, (dn)^2 boxes) failed: %d", (int)r);
    const long long tiles = (long long)(N / kTileCols) * (N / kDn2Rows) * chains;
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > tiles) grid = tiles;
    SVB_CUDA_TRY(launch_pdl(villain_tile_dn2_kernel, (unsigned)grid, (unsigned)kDn2Threads, (size_t)kDn2SmemBytes, stream, map_n, n, chains, N, state_out));
    return 0;
}

// n_sweeps sweeps in place by TMA-staged colour passes (N a multiple of 128); obs_in / counters as launch_villain_stream_passes.
static int launch_villain_tile_passes(const VillainArgs& a, double* obs_in, double* counters, cudaStream_t stream, const DeviceInfo& info) {
    const bool unit = a.W == 1 && a.interval_n == 1;
    auto kern = unit ? villain_tile_pass_kernel<true> : villain_tile_pass_kernel<false>;
    static int per_sm_cache[2][64];
    int per_sm = (info.device < 64) ? per_sm_cache[unit ? 1 : 0][info.device] : 0;
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, kTileSmemBytes));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "the tile pass kernel does not fit an SM");
        if (info.device < 64) per_sm_cache[unit ? 1 : 0][info.device] = per_sm;
    }
    CUtensorMap map_phi, map_n;
    const int rc = villain_tile_maps(a, map_phi, map_n);
    if (rc) return rc;
    const long long tiles = (long long)(a.N / kTileCols) * (a.N / kTileRows) * a.chains;
    if (tiles + 2LL * per_sm * info.sm_count > 0xffffffffLL) return fail(SVB_E_SHAPE, "the tile pass kernel indexes its tiles with 32 bits (%lld)", tiles);
    long long grid = (long long)per_sm * info.sm_count;
    if (grid > tiles) grid = tiles;
    const FilterConsts fc = make_filter_consts(a.interval_phi, a.W, a.interval_n);
    if (obs_in) {
        const char* e = getenv("SVB_VILLAIN_DN2");                  // "stream": the per-thread-load kernel, for an A/B
        const int rc_dn2 = (e && e[0] == 's') ? launch_villain_stream_dn2(a.n, a.chains, a.N, obs_in, stream, info)
                                              : launch_villain_tile_dn2(a.n, a.chains, a.N, obs_in, stream, info);
        if (rc_dn2) return rc_dn2;
    }
    for (int sw = 0; sw < a.n_sweeps; ++sw)
        for (int c = 0; c < 2; ++c) {
            SVB_CUDA_TRY(launch_pdl(kern, (unsigned)grid, 256, (size_t)kTileSmemBytes, stream, a, fc, map_phi, map_n, c, sw,
                                    (sw == 0 && c == 0) ? obs_in : nullptr, counters));
        }
    return 0;
}


// ------------------------------------------------------------------------------------------
// ONE launch per step for big lattices: the phases of a step as a WAVEFRONT through L2 (svb_villain_sweep_wavefront).
//
// The in-place step above is three launches -- sum (dn)^2 over n, colour 0, colour 1 -- and each of them streams the lattice
// from DRAM: 2.5 reads of the state per sweep where the algorithm needs one (ncu of a pass: 6.0 TB/s of DRAM traffic, 92 % of
// what the part delivers: the passes are DRAM bound at twice the algorithmic bytes).  Nothing in the algorithm asks for that
// order.  Colour 1 of a tile row needs colour 0 of that row and of the rows above and below it, nothing more; so the phases
// can follow one another down the lattice a few tile rows apart, and what phase p + 1 reads is what phase p left in L2 a few
// megabytes ago.  DRAM then sees one read of the state and the eviction of the dirty sectors, the algorithmic 32 B per
// site-update.
//
//   * Work items: (phase p, tile row, tile column).  Phase 0 is the sum (dn)^2 of the arriving state (if a record is asked
//     for), then the colour passes of the sweeps in order.  A chain has R = N / 16 tile rows; the tile rows of all chains are
//     numbered J = chain R + k, and phase p visits them in the order k -> tile row (k + p) mod R: the shift by p makes the
//     dependency the same everywhere, the periodic wrap included -- item (p, J) needs (p - 1, chain R + (k + d) mod R), d = 0, 1, 2.
//   * Order: slot s = (wave, p, tile column), J = wave - p lag (slots whose J falls outside the lattice are skipped).  One CTA
//     per SM; its 24 sweeping warps form three GROUPS of eight, each with its own two shared-memory stages.  Slots are handed
//     out in order through a ticket counter, four at a time, to whichever group asks next: the cheap (dn)^2 tiles and the
//     colour tiles, fast SMs and slow ones, balance themselves.
//   * Each group has a CONTROL WARP (warps 24 - 26, one lane active).  It decodes the group's next slot, checks its dependency,
//     issues its TMA loads (full barrier, complete_tx), and -- when the group's eight warps have arrived on the stage's empty
//     barrier -- publishes the finished tile and reuses the stage.  None of this is on the sweeping warps' path: a gpu-scope
//     release fence and an acquire load cost 0.5 - 0.8 us each, a fifth of a tile's sweep; done by the sweeping threads
//     themselves they made this kernel slower than the three launches it replaces (287 against 173 us at L = 4096), and so
//     did one control warp for the three groups (a warp that sits in a fence serves nobody else: 238 us).
//   * Completion: a counter per (p, J), p >= 1, that counts the finished tiles among the three rows (p, J) depends on.  A
//     finished tile of phase p < P - 1 adds one to the counters of its three dependents (gpu-scope release: the group's
//     arrivals on the empty barrier ordered every thread's reductions before the control lane's fence); the control lane
//     loads a dependent tile once its counter reads 3 tiles_x (one acquire load).  The control lane never blocks: it polls
//     the dependency of the next slot and the empty barrier of the oldest stage in turn, so a tile is always published no
//     matter what its group is waiting for, and slots are loaded in order.  Hence no deadlock, whatever the scheduling: the
//     smallest unfinished slot of the grid has all its dependencies (smaller slots, lag >= 3) finished and published.
//     lag puts every dependency about six grids earlier in slot order, so it is normally complete when it is asked for.
//     The counters live in a caller-owned workspace that is all zero between launches: the last CTA to finish clears them.
//   * Everything a tile reads comes through TMA (L2) or ld.global.cg; the exact test reads the staged tile (tile_exact): no
//     plain global load may see an L1 line from an earlier phase.
//   * The tile itself is villain_tile_pass_kernel's: three tensor boxes per colour tile, two per (dn)^2 tile (n0, n1 at the
//     plaquette origin), a warp = one row pair x 32 column pairs.
//
// MEASURED (B200, L = 4096, one sweep per step; tools/kbench_wave.py, profiles/r2_wavefront_c5.txt).  The protocol does what it
// was built for: with the phases up to 18 waves apart DRAM reads 273 - 287 MB per step (one read of the 268 MB state; the
// three launches read 676 MB), beyond 22 waves the lines are gone from L2 before the next phase arrives (575 MB at 28, 850 MB
// at 36: the reuse distance L2 holds is about half its 126 MB).  It is NOT faster: without a record 129 us per step at its
// best lag against 129 us for the two colour-pass launches, with a record 182 against 164 us.  The colour tiles are bound by
// instruction issue (224 thread-instructions per site-update, issue slots 66 - 77 % busy), not by DRAM, so halving the DRAM
// traffic buys nothing until the tile itself is cheaper; the lag that keeps every group supplied (28 - 36 waves) is larger
// than the lag L2 can hold (about 20); and the (dn)^2 tiles, two per group in flight with nothing to compute, wait on DRAM
// latency.  The three launches therefore remain the default for big lattices; this entry point stays for what it is good
// at (half the DRAM traffic and energy at equal time) and as the base for a cheaper tile.
// ------------------------------------------------------------------------------------------
struct WaveArgs {
    int phases;               // P
    int dn2;                  // 1: phase 0 is the sum (dn)^2 pass
    int lag;                  // waves between consecutive phases of one tile row
    int rows;                 // R = N / kTileRows
    int tiles_x;              // N / kTileCols
    int sweep_first;          // sweep (relative to a.sweep0) of the first colour phase
    int chunk;                // slots drawn per ticket
    unsigned rows_total;      // chains R
    unsigned slots;           // (rows_total + (P - 1) lag) P tiles_x
    int* progress;            // (P - 1) rows_total dependency counters (phases 1 .. P - 1), the count of finished CTAs, the slot ticket
    double* state_out;        // state columns of the arriving state (first colour phase, dn2 phase) or nullptr
    double* counter_out;      // ACCEPTED / ACCEPTANCE are added here, or nullptr
};

constexpr int kWaveGroups = 3, kWaveStages = 2, kWaveChunk = 2;
constexpr int kWaveThreads = 256 * kWaveGroups + 32 * kWaveGroups;       // 24 sweeping warps + one control warp per group
constexpr int kWaveSmemBytes = kWaveGroups * kWaveStages * kTileStageBytes + 2 * kWaveGroups * kWaveStages * 8;

__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool UNIT>
__global__ void __launch_bounds__(kWaveThreads, 1) villain_tile_wave_kernel(const __grid_constant__ VillainArgs a,
                                                                            const __grid_constant__ FilterConsts fc,
                                                                            const __grid_constant__ CUtensorMap map_phi,
                                                                            const __grid_constant__ CUtensorMap map_n,
                                                                            const __grid_constant__ WaveArgs wv) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    // per group and stage: chain (-1: no more work), R, C, phase of the tile loaded into it; k (its row index within the phase)
    __shared__ int tile_coord[kWaveGroups][kWaveStages][4];
    __shared__ int tile_k[kWaveGroups][kWaveStages];
    __shared__ int last_cta;
    const int N = a.N, V = N * N;
    const int tid = threadIdx.x;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tile_smem + kWaveGroups * kWaveStages * kTileStageBytes);
    // full[g][s] = bars[2 (2 g + s)], empty[g][s] = bars[2 (2 g + s) + 1]

    if (tid == 0) {
        for (int i = 0; i < kWaveGroups * kWaveStages; ++i) {
            mbar_init(&bars[2 * i], 1);          // full: the control lane's arrive (+ the bytes of the boxes)
            mbar_init(&bars[2 * i + 1], 8);      // empty: one arrival per sweeping warp
        }
        fence_mbar_init();
        last_cta = 0;
    }
    __syncthreads();
    pdl_prologue();

    if (tid >= 256 * kWaveGroups) {
        // ---------------- the control warps: lane 0 of warp 24 + g serves group g ----------------
        const int g = (tid - 256 * kWaveGroups) >> 5;
        if ((tid & 31) == 0) {
            const unsigned per_wave = (unsigned)(wv.phases * wv.tiles_x);
            const int ready = 3 * wv.tiles_x;
            int* ticket = wv.progress + (size_t)(wv.phases - 1) * wv.rows_total + 1;
            unsigned next = 0, chunk_left = 0;                    // the slots this group has drawn and not yet decoded
            unsigned sat_p = 0xffffffffu, sat_J = 0;              // the last (p, J) whose dependency counter was seen complete
            unsigned loaded = 0, retired = 0;                     // tiles of this group whose loads were issued / whose stage was handed back
            unsigned idle = 0;
            unsigned char* stages = tile_smem + g * kWaveStages * kTileStageBytes;
            uint64_t* gbar = bars + 2 * kWaveStages * g;
            // the slot after the ones in the stages: decoded, and its dependency checked, while the group is busy sweeping -- when
            // a stage comes back its loads go out at once
            bool have = false, dep_ok = false, ended = false;
            int n_chain = 0, n_R = 0, n_C = 0, n_p = 0, n_k = 0;
            // a tile that was handed back and is not published yet (its stage is refilled first)
            bool owe = false;
            int o_chain = 0, o_p = 0, o_k = 0;
            while (true) {
                bool progress = false;
                // (A) a free stage and a slot that may be loaded: issue its boxes
                if (have && dep_ok && loaded - retired < (unsigned)kWaveStages) {
                    const int stage = (int)(loaded & 1u);
                    tile_coord[g][stage][0] = n_chain; tile_coord[g][stage][1] = n_R; tile_coord[g][stage][2] = n_C; tile_coord[g][stage][3] = n_p;
                    tile_k[g][stage] = n_k;
                    unsigned char* st = stages + stage * kTileStageBytes;
                    if (wv.dn2 && n_p == 0) {
                        // plaquettes (R .. R+15, C .. C+127): n0 of columns C .. C+128, n1 of rows R .. R+16
                        mbar_expect_tx(&gbar[2 * stage], (uint32_t)(2 * kTileNBytes));
                        tensor_box_3d(st + kTilePhiSlot, &map_n, n_C, n_R, 2 * n_chain, &gbar[2 * stage]);
                        tensor_box_3d(st + kTilePhiSlot + kTileNSlot, &map_n, n_C, n_R, 2 * n_chain + 1, &gbar[2 * stage]);
                    } else {
                        mbar_expect_tx(&gbar[2 * stage], (uint32_t)(kTilePhiBytes + 2 * kTileNBytes));
                        tensor_box_3d(st, &map_phi, n_C - 2, n_R - 1, n_chain, &gbar[2 * stage]);
                        tensor_box_3d(st + kTilePhiSlot, &map_n, n_C - 4, n_R - 1, 2 * n_chain, &gbar[2 * stage]);
                        tensor_box_3d(st + kTilePhiSlot + kTileNSlot, &map_n, n_C - 4, n_R - 1, 2 * n_chain + 1, &gbar[2 * stage]);
                    }
                    ++loaded; have = false; progress = true;
                }
                // (B) publish the tile that was handed back: the three tile rows of phase p + 1 that read it are k, k - 1, k - 2
                // (mod R) in that phase's order
                if (owe) {
                    int* cnt = wv.progress + (size_t)o_p * wv.rows_total + (size_t)o_chain * wv.rows;
                    // release: the group's arrivals ordered every thread's reductions before this fence.  (A (dn)^2 tile wrote
                    // nothing; what its dependents must not overtake are its READS, and those were complete when its boxes landed.)
                    if (!(wv.dn2 && o_p == 0)) asm volatile("fence.acq_rel.gpu;" ::: "memory");
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        int kk = o_k - d;
                        if (kk < 0) kk += wv.rows;
                        asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(cnt + kk) : "memory");
                    }
                    owe = false; progress = true;
                }
                // (C) the oldest tile in the stages: handed back once the group's eight warps have arrived
                if (retired < loaded) {
                    const int stage = (int)(retired & 1u);
                    if (mbar_try(&gbar[2 * stage + 1], (retired >> 1) & 1u)) {
                        o_p = tile_coord[g][stage][3]; o_chain = tile_coord[g][stage][0]; o_k = tile_k[g][stage];
                        owe = o_p + 1 < wv.phases;
                        ++retired; progress = true;
                        continue;                                              // refill the stage before anything else
                    }
                }
                // (D) draw the next slot -- slots are handed out in order, kWaveChunk at a time, to whichever group asks: a group
                // that gets cheap tiles, or a fast SM, simply comes back sooner -- and decode it
                if (!have && !ended) {
                    if (chunk_left == 0) {
                        next = (unsigned)atomicAdd(ticket, wv.chunk);
                        chunk_left = (unsigned)wv.chunk;
                    }
                    const unsigned sl = next++;
                    --chunk_left;
                    if (sl >= wv.slots) {
                        ended = true;
                    } else {
                        const unsigned wave = sl / per_wave, rem = sl - wave * per_wave;
                        const int pp = (int)(rem / (unsigned)wv.tiles_x), tx = (int)(rem - (unsigned)pp * (unsigned)wv.tiles_x);
                        const long long J = (long long)wave - (long long)pp * wv.lag;
                        if (J >= 0 && J < (long long)wv.rows_total) {            // (else: a slot of the ramp that falls outside the lattice)
                            n_chain = (int)((unsigned)J / (unsigned)wv.rows); n_k = (int)((unsigned)J - (unsigned)n_chain * (unsigned)wv.rows);
                            n_R = ((n_k + pp) % wv.rows) * kTileRows; n_C = tx * kTileCols; n_p = pp;
                            have = true;
                            dep_ok = pp == 0 || ((unsigned)pp == sat_p && (unsigned)J == sat_J);
                        }
                    }
                    progress = true;
                }
                // (E) its dependency: the counter of (p, J) reads 3 tiles_x once the three rows it reads are complete
                if (have && !dep_ok) {
                    int seen;
                    const int* cnt = wv.progress + (size_t)(n_p - 1) * wv.rows_total + (size_t)n_chain * wv.rows + n_k;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                    if (seen >= ready) {
                        asm volatile("fence.proxy.async;" ::: "memory");
                        dep_ok = true; progress = true;
                        sat_p = (unsigned)n_p; sat_J = (unsigned)n_chain * (unsigned)wv.rows + (unsigned)n_k;
                    }
                }
                // (F) no slot left: once both stages are back, wake the group with the end marker
                if (ended && !have && !owe && retired == loaded) {
                    const int stage = (int)(loaded & 1u);
                    tile_coord[g][stage][0] = -1;
                    mbar_arrive(&gbar[2 * stage]);
                    break;
                }
                if (progress) {
                    idle = 0;
                } else {
                    __nanosleep(64);
                    if (++idle > (1u << 25)) __trap();            // > 2 s without progress: fail, do not hang the GPU
                }
            }
        }
    } else {
        // ---------------- the sweeping warps: group g, warp w8 of the group ----------------
        const int g = tid >> 8, t = tid & 255, w8 = t >> 5, lane = t & 31;
        unsigned char* stages = tile_smem + g * kWaveStages * kTileStageBytes;
        uint64_t* gbar = bars + 2 * kWaveStages * g;
        double action = 0.0, hk_of_sums = 0.0;
        long long dn2 = 0;
        int w0 = 0, w1 = 0, n_acc = 0, dirty = 0;               // dirty: 1 = dn2, 2 = action / wrapping, 4 = the counters hold something
        float sum_A = 0.0f;
        int chain_of_sums = -1;
        // the warps add their partial sums to a chain's record when the group moves on to another chain
        auto flush = [&]() {
            if (dirty & 1) {
                const long long v = warp_sum(dn2);
                if (lane == 0) atomicAdd(wv.state_out + (long long)chain_of_sums * SVB_VOBS_COUNT + SVB_VOBS_SUM_DN2, (double)v);
                dn2 = 0;
            }
            if (dirty & 2) {
                const double va = warp_sum(action);
                const int v0 = warp_sum_i32(w0), v1 = warp_sum_i32(w1);
                if (lane == 0) {
                    double* o = wv.state_out + (long long)chain_of_sums * SVB_VOBS_COUNT;
                    atomicAdd(o + SVB_VOBS_ACTION, hk_of_sums * va);
                    atomicAdd(o + SVB_VOBS_WRAP0, (double)v0);
                    atomicAdd(o + SVB_VOBS_WRAP1, (double)v1);
                }
                action = 0.0; w0 = 0; w1 = 0;
            }
            if (dirty & 4) {
                if (wv.counter_out) {
                    const int vn = warp_sum_i32(n_acc);
                    const float vs = warp_sum_f32(sum_A);
                    if (lane == 0) {
                        atomicAdd(wv.counter_out + (long long)chain_of_sums * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTED, (double)vn);
                        atomicAdd(wv.counter_out + (long long)chain_of_sums * SVB_VOBS_COUNT + SVB_VOBS_ACCEPTANCE, (double)vs);
                    }
                }
                n_acc = 0; sum_A = 0.0f;
            }
            dirty = 0;
        };

        for (int it = 0;; ++it) {
            const int stage = it & 1;
            mbar_wait(&gbar[2 * stage], (uint32_t)((it >> 1) & 1));
            const int chain = tile_coord[g][stage][0];
            if (chain < 0) break;
            const int R = tile_coord[g][stage][1], C = tile_coord[g][stage][2], p = tile_coord[g][stage][3];
            const bool is_dn2 = wv.dn2 && p == 0;
            const int q = p - wv.dn2;                                  // colour phase: sweep q / 2, colour q % 2
            const bool sums = wv.state_out != nullptr && q == 0;
            if (chain != chain_of_sums) {
                if (dirty) flush();
                chain_of_sums = chain;
                hk_of_sums = (a.kappa_chain ? a.kappa_chain[chain] : a.kappa) / 2;
            }
            const double half_kappa = hk_of_sums;
            double* gphi = reinterpret_cast<double*>(a.phi) + chain * (long long)V;
            int32_t* gn0 = a.n + chain * 2 * (long long)V;
            unsigned char* st = stages + stage * kTileStageBytes;
            double* P = reinterpret_cast<double*>(st);
            int32_t* N0 = reinterpret_cast<int32_t*>(st + kTilePhiSlot);
            int32_t* N1 = reinterpret_cast<int32_t*>(st + kTilePhiSlot + kTileNSlot);
            const bool bottom = R + kTileRows == N, right = C + kTileCols == N;

            if (is_dn2) {
                if (bottom || right) {
                    if (bottom && t < kTileCols) N1[kTileRows * kTileNCols + t] = __ldcg(gn0 + V + C + t);        // row N = row 0
                    if (right && t >= 128 && t < 128 + kTileRows) N0[(t - 128) * kTileNCols + kTileCols] = __ldcg(gn0 + (R + t - 128) * N);
                    asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory");
                }
                // (dn)[x] = (n1[x+e0] - n1[x]) - (n0[x+e1] - n0[x]); a thread: rows w8 and w8 + 8, four columns
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = w8 + 8 * h;
                    const int4 m0 = *reinterpret_cast<const int4*>(N0 + i * kTileNCols + 4 * lane);
                    const int hr = N0[i * kTileNCols + 4 * lane + 4];
                    const int4 m1 = *reinterpret_cast<const int4*>(N1 + i * kTileNCols + 4 * lane);
                    const int4 up = *reinterpret_cast<const int4*>(N1 + (i + 1) * kTileNCols + 4 * lane);
                    const int d0 = (up.x - m1.x) - (m0.y - m0.x), d1 = (up.y - m1.y) - (m0.z - m0.y);
                    const int d2 = (up.z - m1.z) - (m0.w - m0.z), d3 = (up.w - m1.w) - (hr - m0.w);
                    dn2 += (long long)d0 * d0 + (long long)d1 * d1 + (long long)d2 * d2 + (long long)d3 * d3;
                }
                dirty |= 1;
            } else {
                const bool top = R == 0, left = C == 0;
                if (top || bottom || left || right) {
                    // the periodic wrap: what lies beyond the edge of the lattice arrived as zeros; patch it from the other side (L2)
                    if (top && t < kTileCols) {
                        P[2 + t] = __ldcg(gphi + (N - 1) * N + C + t);                                  // halo row R - 1 = row N - 1
                        N0[4 + t] = __ldcg(gn0 + (N - 1) * N + C + t);
                    }
                    if (bottom && t >= 128 && t < 128 + kTileCols) P[(kTilePhiRows - 1) * kTilePhiCols + 2 + (t - 128)] = __ldcg(gphi + C + (t - 128));
                    if (left && t < kTileRows) {
                        P[(t + 1) * kTilePhiCols + 1] = __ldcg(gphi + (R + t) * N + N - 1);            // column C - 1 = column N - 1
                        N1[(t + 1) * kTileNCols + 3] = __ldcg(gn0 + V + (R + t) * N + N - 1);
                    }
                    if (right && t >= 32 && t < 32 + kTileRows) P[(t - 31) * kTilePhiCols + 2 + kTileCols] = __ldcg(gphi + (R + t - 32) * N);   // column N = 0
                    asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory");
                }
                const float hk2 = (float)(half_kappa * 1.4426950408889634);
                const float hkA = 1.0001f * hk2 * fc.bA, hkB = 1.0001f * hk2 * fc.bB + 3.7e-5f;
                const unsigned long long gc = a.chain0 + (unsigned long long)chain;
                const unsigned long long gs = a.sweep0 + (unsigned long long)(wv.sweep_first + (q >> 1));
                const int par = (w8 + q) & 1;
                const int r = R + w8;
                const int oA = r * N, oAm = ((r == 0) ? N - 1 : r - 1) * N, oB = oA + 8 * N, oBp = ((r + 9 == N) ? 0 : r + 9) * N;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int jc = 2 * (lane + 32 * h);
                    const int x1 = C + jc + par;
                    const int xm1 = (x1 == 0) ? N - 1 : x1 - 1, xp1 = (x1 + 1 == N) ? 0 : x1 + 1;
                    const uint32_t c0 = (uint32_t)(oA + x1);
                    const Philox4 bits = philox_site_keys(a, gc, gs, c0);
                    float4 rA, rB;
                    if (sums) {
                        rA = tile_site_residuals<true>(P, N0, N1, w8 + 1, jc, par, action, w0, w1);
                        rB = tile_site_residuals<true>(P, N0, N1, w8 + 9, jc, par, action, w0, w1);
                    } else {
                        rA = tile_site_residuals<false>(P, N0, N1, w8 + 1, jc, par, action, w0, w1);
                        rB = tile_site_residuals<false>(P, N0, N1, w8 + 9, jc, par, action, w0, w1);
                    }
                    stream_pair_decide<UNIT, false, true>(a, fc, gphi, gn0, V, oA, oAm, oA + N, oB, oB - N, oBp, x1, xm1, xp1, bits, c0, rA, rB, gc,
                                                          gs, half_kappa, hk2, hkA, hkB, n_acc, sum_A, nullptr, P, N0, N1,
                                                          (w8 + 1) * kTilePhiCols + jc + 4 + par);
                }
                dirty |= sums ? 6 : 4;
            }
            // this warp has read the stage and issued its reductions (the arrival releases them at CTA scope to the control lane)
            if (bottom || right || R == 0 || C == 0) fence_proxy_async();      // the patches it wrote precede the stage's next TMA fill
            __syncwarp();
            if (lane == 0) mbar_arrive(&gbar[2 * stage + 1]);
        }
        if (dirty) flush();
    }
    // the workspace is all zero between launches: the last CTA to get here clears it (every other CTA has read its last counter)
    __syncthreads();
    if (tid == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        const size_t done_at = (size_t)(wv.phases - 1) * wv.rows_total;
        const int before = atomicAdd(wv.progress + done_at, 1);
        if (before == (int)gridDim.x - 1) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            last_cta = 1;
        }
    }
    __syncthreads();
    if (last_cta) {
        const size_t total = (size_t)(wv.phases - 1) * wv.rows_total + 2;
        for (size_t i = tid; i < total; i += kWaveThreads) wv.progress[i] = 0;
    }
}

// ints of zeroed workspace svb_villain_sweep_wavefront needs for `sweeps` sweeps in one launch
static long long villain_wave_workspace_ints(long long chains, int N, int sweeps, bool with_dn2) {
    return (long long)((with_dn2 ? 1 : 0) + 2 * sweeps) * chains * (N / kTileRows) + 2;
}

// n_sweeps sweeps in place, one wavefront launch per group of sweeps that the workspace has counters for.
static int launch_villain_tile_wave(const VillainArgs& a, double* obs_in, double* counters, int* progress, long long progress_ints,
                                    cudaStream_t stream, const DeviceInfo& info) {
    const bool unit = a.W == 1 && a.interval_n == 1;
    auto kern = unit ? villain_tile_wave_kernel<true> : villain_tile_wave_kernel<false>;
    static int per_sm_cache[2][64];
    int per_sm = (info.device < 64) ? per_sm_cache[unit ? 1 : 0][info.device] : 0;
    if (per_sm == 0) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWaveSmemBytes));
        SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWaveThreads, kWaveSmemBytes));
        if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "the tile wavefront kernel does not fit an SM");
        if (info.device < 64) per_sm_cache[unit ? 1 : 0][info.device] = per_sm;
    }
    CUtensorMap map_phi, map_n;
    const int rc = villain_tile_maps(a, map_phi, map_n);
    if (rc) return rc;
    const int rows = a.N / kTileRows, tiles_x = a.N / kTileCols;
    const long long rows_total = a.chains * rows;
    const long long tiles = rows_total * tiles_x;
    // every CTA must be resident (a tile may wait for a tile of another CTA): one CTA per SM, three groups of sweeping warps each
    long long grid = info.sm_count;
    if (grid * kWaveGroups > tiles) grid = (tiles + kWaveGroups - 1) / kWaveGroups;
    const long long vgrid = grid * kWaveGroups;
    const FilterConsts fc = make_filter_consts(a.interval_phi, a.W, a.interval_n);
    int done = 0;
    while (done < a.n_sweeps) {
        const int dn2 = (done == 0 && obs_in) ? 1 : 0;
        // as many sweeps per launch as the workspace has counters for (at least one)
        long long fit = ((progress_ints - 2) / rows_total - dn2) / 2;
        if (fit < 1) return fail(SVB_E_SHAPE, "svb_villain_sweep_wavefront: the workspace has %lld ints, one sweep needs %lld", progress_ints,
                                 villain_wave_workspace_ints(a.chains, a.N, 1, dn2 != 0));
        int sweeps = (int)((fit < (long long)(a.n_sweeps - done)) ? fit : (long long)(a.n_sweeps - done));
        if (sweeps > 16) sweeps = 16;
        WaveArgs wv;
        wv.phases = dn2 + 2 * sweeps; wv.dn2 = dn2; wv.rows = rows; wv.tiles_x = tiles_x; wv.sweep_first = done;
        const long long per_wave = (long long)wv.phases * tiles_x;
        // every dependency about six grids (of groups) earlier in slot order: a group holds up to four slots (two stages, the one it
        // has decoded ahead, the rest of its ticket) and a tile is published a microsecond after it is finished -- measured at
        // L = 4096: 210 us per step at lag 14, 188 at 28, 183 at 36 (SVB_WAVE_LAG overrides).  Any lag >= 3 is correct.
        const long long lag_min = 3;
        long long lag = 3 + (6 * vgrid + per_wave - 1) / per_wave;
        if (const char* e = getenv("SVB_WAVE_LAG")) lag = atoll(e);
        if (lag < lag_min) lag = lag_min;
        if (lag > 0x3fffffff) lag = 0x3fffffff;
        wv.lag = (int)lag;
        wv.chunk = kWaveChunk;
        if (const char* e = getenv("SVB_WAVE_CHUNK")) wv.chunk = atoi(e) > 0 ? atoi(e) : 1;
        const long long slots = (rows_total + (long long)(wv.phases - 1) * lag) * per_wave;
        if (rows_total > 0x7fffffffLL || slots > 0xf0000000LL)
            return fail(SVB_E_SHAPE, "svb_villain_sweep_wavefront: too many tiles for one launch (%lld slots)", slots);
        wv.rows_total = (unsigned)rows_total; wv.slots = (unsigned)slots;
        wv.progress = progress; wv.state_out = (done == 0) ? obs_in : nullptr; wv.counter_out = counters;
        SVB_CUDA_TRY(launch_pdl(kern, (unsigned)grid, (unsigned)kWaveThreads, (size_t)kWaveSmemBytes, stream, a, fc, map_phi, map_n, wv));
        done += sweeps;
    }
    return 0;
}
