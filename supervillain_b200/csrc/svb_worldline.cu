// svb_worldline.cu -- batched checkerboard Metropolis sweeps of the worldline action
//   S = 1/(2 kappa) sum_l (m - delta v / W)_l^2 + const,   delta m = 0
// The move is PlaquetteUpdate's (supervillain/generator/worldline/plaquette.py:79-101): the four
// boundary links of the plaquette based at x change by +dm, +dm, -dm, -dm and v[x] by dv.  The
// sweep order is the red/black order of the reference's own checkerboard generators
// VortexUpdate / CoexactUpdate (worldline/vortex.py:86-128, worldline/coexact.py:91-120), whose
// v-only and m-only moves are the VORTEX and COEXACT modes here.
//
// Boundary links of the plaquette at x and the sign of delta(unit 2-form at x) on them:
//   (0,x): +1    (0,x+e1): -1    (1,x): -1    (1,x+e0): +1
// f_l = m_l - (delta v)_l / W with (delta v)_0[x] = v[x] - v[x-e1], (delta v)_1[x] = -(v[x] - v[x-e0])
// (compact.py delta,2 rows).  Plaquettes whose base sites share a colour have disjoint
// boundaries, so a colour is updated concurrently without atomics.

#include "svb_common.cuh"

#ifndef SVB_WL_MINB64
/* The table kernel at N = 64: four CTAs of 256 threads per SM, one chain of shared memory each (3 CTAs: 85 registers, measured
 * slower).  -DSVB_WL_MINB64=2 -DSVB_WL_TM64=8 -DSVB_WL_STAGES64=2 builds the alternative -- two CTAs of 512 threads with two
 * stages each, so that a CTA never waits for a load: 30.8 against 23.2 us per config-3 step; four independent CTAs hide each
 * other's barriers and loads better than two that never wait. */
#define SVB_WL_MINB64 4
#define SVB_WL_TM64 4
#define SVB_WL_STAGES64 1
#endif

namespace svb {

struct WorldlineArgs {
    int32_t* m;
    int32_t* v;
    long long chains;
    int N;
    double kappa;
    const double* kappa_chain;
    int W;
    int interval;
    int n_sweeps;
    unsigned long long seed, sweep0, chain0;
    uint32_t round_key[20];   // Philox key schedule precomputed on the host
    uint32_t stream, refine_stream;   // the Philox streams of the generator kind (JOINT 2 / 5, VORTEX 13 / 14, COEXACT 15 / 16)
    const double* inj_u;
    const int32_t* inj_a;
    const int32_t* inj_b;
    double* obs;
    uint8_t* accept_mask;
    double* dS_out;
    OverlapArgs ov;           // overlapped launches (svb_worldline_sweep_overlapped); ov.epochs == nullptr otherwise
};

struct WlDraw {
    double u;        // INJECTED: the Metropolis uniform.  Philox: the midpoint (f + 1/2) 2^-32 of the known bracket
    int a;           // dm (JOINT), dv (VORTEX), t (COEXACT)
    int b;           // dv (JOINT)
    LazyUniform lu;  // Philox: leading 32 bits of the uniform and where its trailing bits come from (svb_common.cuh)
};

// The Philox draw mapping (version 2): ONE 32-bit word per plaquette per sweep, and more only when a decision needs it.
//   The four plaquettes (x0 with bits 3 and 4 varied, x1) -- same colour -- share one Philox4x32-10 block with counter
//   word 0  c0 = (x0 & ~24) N + x1; plaquette x0 owns word (x0 >> 3) & 3 = w:
//   JOINT    dm = bit 31 of w ? +1 : -1   (rng.choice([-1,+1]), plaquette.py:58);  g = w << 1;  p = 3 g;
//            dv = (p >> 32) - 1           (rng.choice([-1,0,+1]), plaquette.py:59);  f = p mod 2^32
//   VORTEX / COEXACT   p = (2 I) w;  idx = p >> 32 picks from [-I..-1, 1..I] (vortex.py:39, coexact.py:40);  f = p mod 2^32
//   and the remainder f is the leading 32 bits of the Metropolis uniform, refined lazily from the kind's refinement stream
//   (same counter, same word).  Given the choice, f lies on a lattice of step 6 (JOINT) or 2 I (VORTEX / COEXACT): the
//   acceptance probability is quantised in units of at most 256 x 2^-32 = 2^-24 (I <= 128 is enforced), the resolution
//   of an fp32 uniform.  Streams (proposal, refinement) by kind: JOINT (2, 5), VORTEX (13, 14), COEXACT (15, 16), so
//   generators sharing a seed inside Sequentially never share a Philox block.
__host__ __device__ __forceinline__ uint32_t worldline_quad_counter(int x0, int x1, int N) { return (uint32_t)((x0 & ~24) * N + x1); }
__host__ __device__ __forceinline__ uint32_t worldline_quad_word(int x0) { return (uint32_t)((x0 >> 3) & 3); }

template <int MODE>
__device__ __forceinline__ WlDraw worldline_draw_from_word(uint32_t w, int interval) {
    WlDraw d;
    uint32_t f;
    if (MODE == SVB_WL_JOINT) {
        d.a = (w >> 31) ? +1 : -1;
        const uint64_t p = (uint64_t)(w << 1) * 3ull;
        d.b = (int)(p >> 32) - 1;
        f = (uint32_t)p;
    } else {
        const uint64_t p = (uint64_t)w * (uint64_t)(2 * interval);
        const int idx = (int)(p >> 32);
        d.a = (idx < interval) ? idx - interval : idx - interval + 1;
        d.b = 0;
        f = (uint32_t)p;
    }
    d.lu.f = f;
    d.u = (__hiloint2double(0x43300000, (int)f) - 4503599627370495.5) * 2.3283064365386963e-10;    // (f + 1/2) 2^-32
    return d;
}

__device__ __forceinline__ uint32_t philox_word(const Philox4& p, uint32_t word) {
    return (word == 0) ? p.x : (word == 1) ? p.y : (word == 2) ? p.z : p.w;
}

template <int MODE, bool INJECTED>
__device__ __forceinline__ WlDraw worldline_get_draw(const WorldlineArgs& a, long long chain, int sweep, int x0, int x1, int site) {
    if (INJECTED) {
        const long long V = (long long)a.N * a.N;
        const long long base = ((long long)sweep * a.chains + chain) * V + site;
        WlDraw d;
        d.u = a.inj_u[base];
        d.a = a.inj_a[base];
        d.b = (MODE == SVB_WL_JOINT) ? a.inj_b[base] : 0;
        d.lu.f = 0; d.lu.c0 = 0; d.lu.word = 0;
        return d;
    } else {
        const uint32_t c0 = worldline_quad_counter(x0, x1, a.N), word = worldline_quad_word(x0);
        const Philox4 p = philox_site(a.seed, a.chain0 + (unsigned long long)chain, a.sweep0 + (unsigned long long)sweep, c0,
                                      a.stream);
        WlDraw d = worldline_draw_from_word<MODE>(philox_word(p, word), a.interval);
        d.lu.c0 = c0; d.lu.word = word;
        return d;
    }
}

__device__ __forceinline__ RefineCtx worldline_refine_ctx(const WorldlineArgs& a, long long chain, int sweep) {
    RefineCtx rc;
    rc.seed = a.seed; rc.chain = a.chain0 + (unsigned long long)chain; rc.sweep = a.sweep0 + (unsigned long long)sweep;
    rc.stream = a.refine_stream; rc.wide = 0;
    return rc;
}

// u < min(1, e^-dS) for either kind of draw: INJECTED compares the given uniform, Philox decides lazily.
template <bool LAZY>
__device__ __forceinline__ bool worldline_decide(double acc, const WlDraw& d, const RefineCtx& rc) {
    if (LAZY) return decide_lazy(acc, d.lu, rc.stream, rc);
    return d.u < acc;
}

struct PlaqOut {
    double A;
    bool ok;
    double dS;
};

template <int MODE, bool LAZY>
__device__ __forceinline__ PlaqOut worldline_plaquette_update(int32_t* __restrict__ m0, int32_t* __restrict__ m1,
                                                              int32_t* __restrict__ v, int N, int x0, int x1, double kappa,
                                                              double Wd, const WlDraw& d, const RefineCtx& rc) {
    const int xp0 = (x0 + 1 == N) ? 0 : x0 + 1;
    const int xm0 = (x0 == 0) ? N - 1 : x0 - 1;
    const int xp1 = (x1 + 1 == N) ? 0 : x1 + 1;
    const int xm1 = (x1 == 0) ? N - 1 : x1 - 1;
    const int i_c = x0 * N + x1;
    const int i_p0 = xp0 * N + x1, i_m0 = xm0 * N + x1;
    const int i_p1 = x0 * N + xp1, i_m1 = x0 * N + xm1;

    const int vc = v[i_c], vp0 = v[i_p0], vm0 = v[i_m0], vp1 = v[i_p1], vm1 = v[i_m1];
    const int m_0x = m0[i_c];      // link (0, x)
    const int m_0p = m0[i_p1];     // link (0, x + e1)
    const int m_1x = m1[i_c];      // link (1, x)
    const int m_1p = m1[i_p0];     // link (1, x + e0)

    // f = m - delta(v) / W     (plaquette.py:53, vortex.py:111, coexact.py:104)
    const double f_0x = __dsub_rn((double)m_0x, __ddiv_rn((double)(vc - vm1), Wd));
    const double f_0p = __dsub_rn((double)m_0p, __ddiv_rn((double)(vp1 - vc), Wd));
    const double f_1x = __dsub_rn((double)m_1x, __ddiv_rn((double)(vm0 - vc), Wd));
    const double f_1p = __dsub_rn((double)m_1p, __ddiv_rn((double)(vc - vp0), Wd));

    double dS;
    if (MODE == SVB_WL_JOINT) {
        // delta_f = dm - dv / W ;  dS = delta_f / kappa * (f1 + f2 - f3 - f4 + 2 delta_f)   (plaquette.py:84-85)
        const double df = __dsub_rn((double)d.a, __ddiv_rn((double)d.b, Wd));
        double s = __dadd_rn(f_0x, f_1p);            // f1 + f2
        s = __dsub_rn(s, f_0p);                      // - f3
        s = __dsub_rn(s, f_1x);                      // - f4
        s = __dadd_rn(s, __dmul_rn(2.0, df));
        dS = __dmul_rn(__ddiv_rn(df, kappa), s);
    } else {
        const double hk = __ddiv_rn(0.5, kappa);
        double t_1x, t_1p, t_0x, t_0p;
        if (MODE == SVB_WL_VORTEX) {
            // c_l = sign_l a / W ;  T_l = ((0.5/kappa)(-c_l)) ((2 f_l) - c_l)   (vortex.py:108-112)
            const double c_pos = __ddiv_rn((double)d.a, Wd), c_neg = __ddiv_rn((double)(-d.a), Wd);
            t_1x = __dmul_rn(__dmul_rn(hk, -c_neg), __dsub_rn(__dmul_rn(2.0, f_1x), c_neg));
            t_1p = __dmul_rn(__dmul_rn(hk, -c_pos), __dsub_rn(__dmul_rn(2.0, f_1p), c_pos));
            t_0x = __dmul_rn(__dmul_rn(hk, -c_pos), __dsub_rn(__dmul_rn(2.0, f_0x), c_pos));
            t_0p = __dmul_rn(__dmul_rn(hk, -c_neg), __dsub_rn(__dmul_rn(2.0, f_0p), c_neg));
        } else {
            // c_l = sign_l t ;  T_l = ((0.5/kappa) c_l) ((2 f_l) + c_l)           (coexact.py:102-106)
            const double c_pos = (double)d.a, c_neg = (double)(-d.a);
            t_1x = __dmul_rn(__dmul_rn(hk, c_neg), __dadd_rn(__dmul_rn(2.0, f_1x), c_neg));
            t_1p = __dmul_rn(__dmul_rn(hk, c_pos), __dadd_rn(__dmul_rn(2.0, f_1p), c_pos));
            t_0x = __dmul_rn(__dmul_rn(hk, c_pos), __dadd_rn(__dmul_rn(2.0, f_0x), c_pos));
            t_0p = __dmul_rn(__dmul_rn(hk, c_neg), __dadd_rn(__dmul_rn(2.0, f_0p), c_neg));
        }
        // coface_sum order: 0 + T(1,x) + T(1,x+e0) + T(0,x) + T(0,x+e1)   (compact.py coface_sum,1 rows)
        dS = __dadd_rn(t_1x, t_1p);
        dS = __dadd_rn(dS, t_0x);
        dS = __dadd_rn(dS, t_0p);
    }
    const double acc = fmin(exp(-dS), 1.0);
    const bool ok = worldline_decide<LAZY>(acc, d, rc);
    if (ok) {
        if (MODE == SVB_WL_JOINT) {
            m0[i_c] = m_0x + d.a;      // plaquette.py:91-95
            m1[i_p0] = m_1p + d.a;
            m0[i_p1] = m_0p - d.a;
            m1[i_c] = m_1x - d.a;
            v[i_c] = vc + d.b;
        } else if (MODE == SVB_WL_VORTEX) {
            v[i_c] = vc + d.a;         // vortex.py:125-127
        } else {
            m0[i_c] = m_0x + d.a;      // m += delta t, coexact.py:119-120
            m1[i_p0] = m_1p + d.a;
            m0[i_p1] = m_0p - d.a;
            m1[i_c] = m_1x - d.a;
        }
    }
    PlaqOut o;
    o.A = acc;
    o.ok = ok;
    o.dS = dS;
    return o;
}

__device__ __forceinline__ void worldline_obs_partial(const int32_t* __restrict__ m0, const int32_t* __restrict__ m1,
                                                      const int32_t* __restrict__ v, int N, double Wd, int tid, int nthreads,
                                                      double (&s)[5]) {
    const int V = N * N;
    for (int i = tid; i < V; i += nthreads) {
        const int x0 = i / N, x1 = i - x0 * N;
        const int xp0 = (x0 + 1 == N) ? 0 : x0 + 1, xm0 = (x0 == 0) ? N - 1 : x0 - 1;
        const int xp1 = (x1 + 1 == N) ? 0 : x1 + 1, xm1 = (x1 == 0) ? N - 1 : x1 - 1;
        const int i_p0 = xp0 * N + x1, i_m0 = xm0 * N + x1, i_p1 = x0 * N + xp1, i_m1 = x0 * N + xm1;
        const int vc = v[i], a0 = m0[i], a1 = m1[i];
        // f at (0,x), (1,x), (1,x+e0), (0,x+e1)
        const double f0 = (double)a0 - (double)(vc - v[i_m1]) / Wd;
        const double f1 = (double)a1 - (double)(v[i_m0] - vc) / Wd;
        const double f1p = (double)m1[i_p0] - (double)(vc - v[i_p0]) / Wd;
        const double f0p = (double)m0[i_p1] - (double)(v[i_p1] - vc) / Wd;
        s[0] += f0 * f0 + f1 * f1;
        const double df = (f1p - f1) - (f0p - f0);                  // (d f)[x], compact.py d,1 rows
        s[1] += df * df;
        s[2] += (double)a0;
        s[3] += (double)a1;
        // (delta m)[x] = -(m0[x] - m0[x-e0]) - (m1[x] - m1[x-e1])   (compact.py delta,1 rows)
        const int dm = -(a0 - m0[i_m0]) - (a1 - m1[i_m1]);
        s[4] += (double)(dm < 0 ? -dm : dm);
    }
}

template <int MODE, bool INJECTED>
__global__ void __launch_bounds__(256) worldline_smem_kernel(WorldlineArgs a, int use_bulk) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int N = a.N, V = N * N;
    const int tid = threadIdx.x, T = blockDim.x;
    const size_t bytes_m = (size_t)2 * V * sizeof(int32_t);
    const size_t bytes_v = (size_t)V * sizeof(int32_t);
    const size_t off_v = (bytes_m + 15) & ~(size_t)15;
    const size_t off_scr = (off_v + bytes_v + 15) & ~(size_t)15;
    int32_t* sm0 = reinterpret_cast<int32_t*>(smem_raw);
    int32_t* sm1 = sm0 + V;
    int32_t* sv = reinterpret_cast<int32_t*>(smem_raw + off_v);
    double* scratch = reinterpret_cast<double*>(smem_raw + off_scr);   // 7 * 32 doubles
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + off_scr + 7 * 32 * sizeof(double));

    if (use_bulk) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    uint32_t phase = 0;
    const int ncol = n_colours(N);
    const int halfN = N >> 1, nhalf = V >> 1;
    const double Wd = (double)a.W;

    for (long long chain = blockIdx.x; chain < a.chains; chain += gridDim.x) {
        int32_t* gm = a.m + chain * 2 * V;
        int32_t* gv = a.v + chain * V;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;

        if (use_bulk) {
            if (tid == 0) {
                mbar_expect_tx(bar, (uint32_t)(bytes_m + bytes_v));
                bulk_g2s(sm0, gm, (uint32_t)bytes_m, bar);
                bulk_g2s(sv, gv, (uint32_t)bytes_v, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            for (int i = tid; i < 2 * V; i += T) sm0[i] = gm[i];
            for (int i = tid; i < V; i += T) sv[i] = gv[i];
            __syncthreads();
        }

        double n_acc = 0.0, sum_A = 0.0;
        for (int s = 0; s < a.n_sweeps; ++s) {
            const bool last = (s == a.n_sweeps - 1);
            for (int c = 0; c < ncol; ++c) {
                if (ncol == 2) {
                    for (int j = tid; j < nhalf; j += T) {
                        const int x0 = j / halfN;
                        const int x1 = 2 * (j - x0 * halfN) + ((x0 + c) & 1);
                        const int site = x0 * N + x1;
                        const WlDraw d = worldline_get_draw<MODE, INJECTED>(a, chain, s, x0, x1, site);
                        const PlaqOut o = worldline_plaquette_update<MODE, !INJECTED>(sm0, sm1, sv, N, x0, x1, kappa, Wd, d,
                                                                                      worldline_refine_ctx(a, chain, s));
                        n_acc += o.ok ? 1.0 : 0.0;
                        sum_A += o.A;
                        if (last) {
                            if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                            if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
                        }
                    }
                } else {
                    for (int site = tid; site < V; site += T) {
                        const int x0 = site / N, x1 = site - x0 * N;
                        if (site_colour(x0, x1, N) != c) continue;
                        const WlDraw d = worldline_get_draw<MODE, INJECTED>(a, chain, s, x0, x1, site);
                        const PlaqOut o = worldline_plaquette_update<MODE, !INJECTED>(sm0, sm1, sv, N, x0, x1, kappa, Wd, d,
                                                                                      worldline_refine_ctx(a, chain, s));
                        n_acc += o.ok ? 1.0 : 0.0;
                        sum_A += o.A;
                        if (last) {
                            if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                            if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
                        }
                    }
                }
                __syncthreads();
            }
        }

        if (a.obs) {
            double part[5] = {0, 0, 0, 0, 0};
            worldline_obs_partial(sm0, sm1, sv, N, Wd, tid, T, part);
            double s[7] = {part[0], part[1], part[2], part[3], n_acc, sum_A, part[4]};
            block_sum<7>(s, scratch);
            if (tid == 0) {
                double* o = a.obs + chain * SVB_WOBS_COUNT;
                for (int k = 0; k < SVB_WOBS_COUNT; ++k) o[k] = s[k];
            }
        }

        if (use_bulk) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (MODE != SVB_WL_VORTEX) bulk_s2g(gm, sm0, (uint32_t)bytes_m);
                if (MODE != SVB_WL_COEXACT) bulk_s2g(gv, sv, (uint32_t)bytes_v);
                bulk_commit();
                bulk_wait_read0();
            }
            __syncthreads();
        } else {
            if (MODE != SVB_WL_VORTEX)
                for (int i = tid; i < 2 * V; i += T) gm[i] = sm0[i];
            if (MODE != SVB_WL_COEXACT)
                for (int i = tid; i < V; i += T) gv[i] = sv[i];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Production instantiation: W = 1, Philox draws, compile-time geometry, TMA bulk copies with
// STAGES shared-memory stages per CTA.  With W = 1 every f = m - delta v is an integer, so the
// whole neighbourhood arithmetic is int32 and only the final products are fp64 -- with exactly the
// roundings of the reference's expressions (each product below has exact integer factors times one
// rounded constant), so dS is bit-equal to the general path and to the oracle.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Philox4 philox_plaquette_keys(const WorldlineArgs& a, uint64_t chain, uint64_t sweep, uint32_t site) {
    uint32_t c0 = site, c1 = (uint32_t)chain, c2 = (uint32_t)sweep;
    uint32_t c3 = (a.stream << 24) | ((uint32_t)((chain >> 32) & 0xFFu) << 16) |
                  (uint32_t)((sweep >> 32) & 0xFFFFu);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c0, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ a.round_key[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ a.round_key[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ double wl_int_to_double(int n) {   // exact, on the fp64 adder
    return __hiloint2double(0x43300000, n ^ 0x80000000) - 4503601774854144.0;
}

struct WlSums {
    long long f2;      // sum of f^2 over links (an integer when W = 1)
    long long df2;     // sum of (d f)^2 over plaquettes
    int w0, w1;        // sum of m_0, m_1
};

template <int MODE, int NT>
__device__ __forceinline__ PlaqOut worldline_plaquette_update_w1(int32_t* __restrict__ m0, int32_t* __restrict__ m1,
                                                                 int32_t* __restrict__ v, int x0, int x1, double inv_kappa,
                                                                 double half_inv_kappa, const WlDraw& d, bool collect,
                                                                 WlSums& sums, const RefineCtx& rc) {
    const int xp0 = (x0 + 1) & (NT - 1), xm0 = (x0 - 1) & (NT - 1);
    const int xp1 = (x1 + 1) & (NT - 1), xm1 = (x1 - 1) & (NT - 1);
    const int i_c = x0 * NT + x1;
    const int i_p0 = xp0 * NT + x1, i_m0 = xm0 * NT + x1;
    const int i_p1 = x0 * NT + xp1, i_m1 = x0 * NT + xm1;
    const int vc = v[i_c], vp0 = v[i_p0], vm0 = v[i_m0], vp1 = v[i_p1], vm1 = v[i_m1];
    const int m_0x = m0[i_c], m_0p = m0[i_p1], m_1x = m1[i_c], m_1p = m1[i_p0];
    // f = m - delta(v) on the four boundary links (W = 1: exact integers)
    const int f_0x = m_0x - (vc - vm1);
    const int f_0p = m_0p - (vp1 - vc);
    const int f_1x = m_1x - (vm0 - vc);
    const int f_1p = m_1p - (vc - vp0);
    const int curl = (f_0x + f_1p) - f_0p - f_1x;        // f1 + f2 - f3 - f4 = (d f)[x]
    double dS;
    int df_int;                                          // change of f on the (+)-oriented links if accepted
    if (MODE == SVB_WL_JOINT) {
        df_int = d.a - d.b;                              // delta_f = dm - dv / 1
        // (delta_f / kappa) * (f1 + f2 - f3 - f4 + 2 delta_f): |delta_f| <= 2, so delta_f / kappa == delta_f * fl(1/kappa)
        dS = __dmul_rn(__dmul_rn(wl_int_to_double(df_int), inv_kappa), wl_int_to_double(curl + 2 * df_int));
    } else {
        const int a = d.a;
        const double P = __dmul_rn(half_inv_kappa, wl_int_to_double(a));      // (0.5/kappa) * a
        double t_1x, t_1p, t_0x, t_0p;
        if (MODE == SVB_WL_VORTEX) {
            df_int = -a;                                 // f changes by -sign * a
            t_1x = __dmul_rn(P, wl_int_to_double(2 * f_1x + a));
            t_1p = __dmul_rn(-P, wl_int_to_double(2 * f_1p - a));
            t_0x = __dmul_rn(-P, wl_int_to_double(2 * f_0x - a));
            t_0p = __dmul_rn(P, wl_int_to_double(2 * f_0p + a));
        } else {
            df_int = a;
            t_1x = __dmul_rn(-P, wl_int_to_double(2 * f_1x - a));
            t_1p = __dmul_rn(P, wl_int_to_double(2 * f_1p + a));
            t_0x = __dmul_rn(P, wl_int_to_double(2 * f_0x + a));
            t_0p = __dmul_rn(-P, wl_int_to_double(2 * f_0p - a));
        }
        dS = __dadd_rn(t_1x, t_1p);
        dS = __dadd_rn(dS, t_0x);
        dS = __dadd_rn(dS, t_0p);
    }
    double acc;                                          // fp32-accurate statistic; the decision is the exact fp64 one
    bool ok;
    {
        int r = -1;
        if (d.lu.f >= 65536u) r = metropolis_log_filter(dS, d.u, 7.62939453125e-06f, acc);    // bracket half-width <= 2^-17 u
        if (r >= 0) ok = r != 0;
        else {
            acc = exp_clipped(-dS);
            ok = decide_lazy(acc, d.lu, rc.stream, rc);
        }
    }
    if (ok) {
        if (MODE != SVB_WL_VORTEX) {
            m0[i_c] = m_0x + d.a;
            m1[i_p0] = m_1p + d.a;
            m0[i_p1] = m_0p - d.a;
            m1[i_c] = m_1x - d.a;
        }
        if (MODE == SVB_WL_JOINT) v[i_c] = vc + d.b;
        if (MODE == SVB_WL_VORTEX) v[i_c] = vc + d.a;
    }
    if (collect) {
        const int s = ok ? df_int : 0;
        const int g0 = f_0x + s, g1 = f_1p + s, g2 = f_0p - s, g3 = f_1x - s;      // final f on the four links
        sums.f2 += (long long)g0 * g0 + (long long)g1 * g1 + (long long)g2 * g2 + (long long)g3 * g3;
        const int c2 = curl + 4 * s;
        sums.df2 += (long long)c2 * c2;
        sums.w0 += m_0x + m_0p;                                                    // dm cancels within a direction
        sums.w1 += m_1x + m_1p;
    }
    PlaqOut o;
    o.A = acc;
    o.ok = ok;
    o.dS = dS;
    return o;
}

// (d f)^2 of the plaquette at (x0, x1) from the final fields (W = 1)
template <int NT>
__device__ __forceinline__ long long worldline_curl2_w1(const int32_t* __restrict__ m0, const int32_t* __restrict__ m1,
                                                        const int32_t* __restrict__ v, int x0, int x1) {
    const int xp0 = (x0 + 1) & (NT - 1), xm0 = (x0 - 1) & (NT - 1);
    const int xp1 = (x1 + 1) & (NT - 1), xm1 = (x1 - 1) & (NT - 1);
    const int i_c = x0 * NT + x1;
    const int vc = v[i_c];
    // curl of m minus curl of delta v:  (d delta v)[x] = 4 v[x] - v[x-e1] - v[x+e1] - v[x-e0] - v[x+e0]
    const int cm = (m0[i_c] + m1[xp0 * NT + x1]) - m0[x0 * NT + xp1] - m1[i_c];
    const int cv = 4 * vc - v[x0 * NT + xm1] - v[x0 * NT + xp1] - v[xm0 * NT + x1] - v[xp0 * NT + x1];
    const int c = cm - cv;
    return (long long)c * c;
}

template <int MODE, int NT, int TT, int MINB, int STAGES>
__global__ void __launch_bounds__(TT, MINB) worldline_smem_fast_kernel(const __grid_constant__ WorldlineArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int N = NT, V = NT * NT, T = TT;
    constexpr int halfN = N / 2, nhalf = V / 2;
    constexpr uint32_t bytes_m = 2 * V * sizeof(int32_t);
    constexpr uint32_t bytes_v = V * sizeof(int32_t);
    constexpr uint32_t stage_bytes = bytes_m + bytes_v;
    const int tid = threadIdx.x;
    double* scratch = reinterpret_cast<double*>(smem_raw + STAGES * stage_bytes);      // 7 * 32 doubles
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + STAGES * stage_bytes + 7 * 32 * sizeof(double));

    if (tid == 0) {
        for (int b = 0; b < STAGES; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const bool want_obs = a.obs != nullptr;

    auto issue_load = [&](long long chain, int b) {
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        mbar_expect_tx(&bar[b], stage_bytes);
        bulk_g2s(stage, a.m + chain * 2 * V, bytes_m, &bar[b]);
        bulk_g2s(stage + bytes_m, a.v + chain * V, bytes_v, &bar[b]);
    };

    long long chain = blockIdx.x;
    if (tid == 0 && chain < a.chains) issue_load(chain, 0);

    for (int it = 0; chain < a.chains; chain += gridDim.x, ++it) {
        const int b = (STAGES == 2) ? (it & 1) : 0;
        unsigned char* stage = smem_raw + (size_t)b * stage_bytes;
        int32_t* sm0 = reinterpret_cast<int32_t*>(stage);
        int32_t* sm1 = sm0 + V;
        int32_t* sv = reinterpret_cast<int32_t*>(stage + bytes_m);
        const long long next = chain + gridDim.x;
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double inv_kappa = __ddiv_rn(1.0, kappa), half_inv_kappa = __ddiv_rn(0.5, kappa);

        mbar_wait(&bar[b], (uint32_t)((STAGES == 2 ? (it >> 1) : it) & 1));

        int n_acc = 0;
        double sum_A = 0.0;
        WlSums ws;
        ws.f2 = 0; ws.df2 = 0; ws.w0 = 0; ws.w1 = 0;
        for (int s = 0; s < a.n_sweeps; ++s) {
            const bool last = (s == a.n_sweeps - 1);
            const bool debug = last && (a.accept_mask != nullptr || a.dS_out != nullptr);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const bool collect = want_obs && last && (c == 1);
#pragma unroll 1
                for (int j = tid; j < nhalf; j += T) {
                    const int x0 = j / halfN;
                    const int x1 = 2 * (j - x0 * halfN) + ((x0 + c) & 1);
                    const int site = x0 * N + x1;
                    const uint32_t qc0 = worldline_quad_counter(x0, x1, N), qword = worldline_quad_word(x0);
                    const Philox4 bits = philox_plaquette_keys(a, a.chain0 + (unsigned long long)chain,
                                                               a.sweep0 + (unsigned long long)s, qc0);
                    WlDraw d = worldline_draw_from_word<MODE>(philox_word(bits, qword), a.interval);
                    d.lu.c0 = qc0; d.lu.word = qword;
                    const PlaqOut o = worldline_plaquette_update_w1<MODE, NT>(sm0, sm1, sv, x0, x1, inv_kappa, half_inv_kappa, d,
                                                                              collect, ws, worldline_refine_ctx(a, chain, s));
                    n_acc += o.ok ? 1 : 0;
                    sum_A += o.A;
                    if (debug) {
                        if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
                        if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
                    }
                }
                __syncthreads();
                if (STAGES == 2 && s == 0 && c == 0 && tid == 0 && next < a.chains) {
                    bulk_wait_read0();
                    issue_load(next, b ^ 1);
                }
            }
        }

        if (want_obs) {
            // links and colour-1 plaquettes were collected in the last pass; the colour-0 plaquettes' (d f)^2 needs the
            // final fields of their neighbours, so it is evaluated here
#pragma unroll 1
            for (int j = tid; j < nhalf; j += T) {
                const int x0 = j / halfN;
                const int x1 = 2 * (j - x0 * halfN) + (x0 & 1);
                ws.df2 += worldline_curl2_w1<NT>(sm0, sm1, sv, x0, x1);
            }
            double sred[7] = {0, 0, 0, 0, 0, sum_A, 0};
            // integer sums: exact in double as long as they stay below 2^53
            sred[0] = (double)ws.f2; sred[1] = (double)ws.df2; sred[2] = (double)ws.w0; sred[3] = (double)ws.w1;
            sred[4] = (double)n_acc;
            block_sum<7>(sred, scratch);
            if (tid == 0) {
                double* o = a.obs + chain * SVB_WOBS_COUNT;
                o[SVB_WOBS_SUM_F2] = sred[0];
                o[SVB_WOBS_SUM_DF2] = sred[1];
                o[SVB_WOBS_WRAP0] = sred[2];
                o[SVB_WOBS_WRAP1] = sred[3];
                o[SVB_WOBS_ACCEPTED] = sred[4];
                o[SVB_WOBS_ACCEPTANCE] = sred[5];
                o[SVB_WOBS_DELTA_M_ABS] = -1.0;       // not evaluated by the sweep (the move preserves delta m identically)
            }
        }

        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (MODE != SVB_WL_VORTEX) bulk_s2g(a.m + chain * 2 * V, sm0, bytes_m);
            if (MODE != SVB_WL_COEXACT) bulk_s2g(a.v + chain * V, sv, bytes_v);
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

template <int MODE, bool INJECTED>
__global__ void __launch_bounds__(256) worldline_colour_pass_kernel(WorldlineArgs a, int sweep, int colour, int blocks_per_chain,
                                                                    int write_debug) {
    const int N = a.N, V = N * N;
    const long long chain = blockIdx.x / blocks_per_chain;
    const int blk = blockIdx.x - (int)(chain * blocks_per_chain);
    int32_t* gm0 = a.m + chain * 2 * V;
    int32_t* gm1 = gm0 + V;
    int32_t* gv = a.v + chain * V;
    const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
    double n_acc = 0.0, sum_A = 0.0;

    int site = -1, x0 = 0, x1 = 0;
    if ((N & 1) == 0) {
        const int j = blk * blockDim.x + threadIdx.x;
        if (j < (V >> 1)) {
            const int halfN = N >> 1;
            x0 = j / halfN;
            x1 = 2 * (j - x0 * halfN) + ((x0 + colour) & 1);
            site = x0 * N + x1;
        }
    } else {
        const int i = blk * blockDim.x + threadIdx.x;
        if (i < V) {
            x0 = i / N;
            x1 = i - x0 * N;
            if (site_colour(x0, x1, N) == colour) site = i;
        }
    }
    if (site >= 0) {
        const WlDraw d = worldline_get_draw<MODE, INJECTED>(a, chain, sweep, x0, x1, site);
        const PlaqOut o = worldline_plaquette_update<MODE, !INJECTED>(gm0, gm1, gv, N, x0, x1, kappa, (double)a.W, d,
                                                                      worldline_refine_ctx(a, chain, sweep));
        n_acc = o.ok ? 1.0 : 0.0;
        sum_A = o.A;
        if (write_debug) {
            if (a.accept_mask) a.accept_mask[chain * V + site] = o.ok ? 1 : 0;
            if (a.dS_out) a.dS_out[chain * V + site] = o.dS;
        }
    }
    if (a.obs) {
        __shared__ double scratch[2 * 32];
        double s[2] = {n_acc, sum_A};
        block_sum<2>(s, scratch);
        if (threadIdx.x == 0) {
            atomicAdd(a.obs + chain * SVB_WOBS_COUNT + SVB_WOBS_ACCEPTED, s[0]);
            atomicAdd(a.obs + chain * SVB_WOBS_COUNT + SVB_WOBS_ACCEPTANCE, s[1]);
        }
    }
}

__global__ void __launch_bounds__(256) worldline_obs_kernel(const int32_t* __restrict__ m, const int32_t* __restrict__ v,
                                                            long long chains, int N, int W, double* __restrict__ obs,
                                                            int keep_counters) {
    __shared__ double scratch[5 * 32];
    const int V = N * N;
    for (long long chain = blockIdx.x; chain < chains; chain += gridDim.x) {
        const int32_t* gm0 = m + chain * 2 * V;
        double s[5] = {0, 0, 0, 0, 0};
        worldline_obs_partial(gm0, gm0 + V, v + chain * V, N, (double)W, threadIdx.x, blockDim.x, s);
        block_sum<5>(s, scratch);
        if (threadIdx.x == 0) {
            double* o = obs + chain * SVB_WOBS_COUNT;
            o[SVB_WOBS_SUM_F2] = s[0];
            o[SVB_WOBS_SUM_DF2] = s[1];
            o[SVB_WOBS_WRAP0] = s[2];
            o[SVB_WOBS_WRAP1] = s[3];
            o[SVB_WOBS_DELTA_M_ABS] = s[4];
            if (!keep_counters) {
                o[SVB_WOBS_ACCEPTED] = 0.0;
                o[SVB_WOBS_ACCEPTANCE] = 0.0;
            }
        }
        __syncthreads();
    }
}

__global__ void worldline_zero_counters_kernel(double* obs, long long chains) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < chains) {
        obs[c * SVB_WOBS_COUNT + SVB_WOBS_ACCEPTED] = 0.0;
        obs[c * SVB_WOBS_COUNT + SVB_WOBS_ACCEPTANCE] = 0.0;
    }
}

static size_t worldline_smem_bytes(int N) {
    const size_t V = (size_t)N * N;
    const size_t off_v = (2 * V * sizeof(int32_t) + 15) & ~(size_t)15;
    const size_t off_scr = (off_v + V * sizeof(int32_t) + 15) & ~(size_t)15;
    return off_scr + 7 * 32 * sizeof(double) + 16;
}

static int worldline_threads_for(int N) {
    const int V = N * N;
    int t = (V / 8 + 31) / 32 * 32;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
}

template <int MODE, bool INJECTED>
static int launch_worldline_smem(const WorldlineArgs& a, cudaStream_t stream, int sm_count) {
    auto kern = worldline_smem_kernel<MODE, INJECTED>;
    const size_t smem = worldline_smem_bytes(a.N);
    const int threads = worldline_threads_for(a.N);
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "worldline smem kernel does not fit an SM at N=%d", a.N);
    long long grid = (long long)per_sm * sm_count;
    if (grid > a.chains) grid = a.chains;
    const size_t bytes_v = (size_t)a.N * a.N * sizeof(int32_t);
    const int use_bulk = (bytes_v % 16 == 0) && ((uintptr_t)a.m % 16 == 0) && ((uintptr_t)a.v % 16 == 0);
    kern<<<(unsigned)grid, threads, smem, stream>>>(a, use_bulk);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <int MODE, bool INJECTED>
static int launch_worldline_global(const WorldlineArgs& a, cudaStream_t stream) {
    const int N = a.N, V = N * N;
    const int threads = 256;
    const int work = (N & 1) ? V : V / 2;
    const int bpc = (work + threads - 1) / threads;
    const long long blocks = (long long)bpc * a.chains;
    if (blocks > 0x7fffffffLL) return fail(SVB_E_SHAPE, "too many blocks (%lld) for the global path", blocks);
    if (a.obs) {
        worldline_zero_counters_kernel<<<(unsigned)((a.chains + 255) / 256), 256, 0, stream>>>(a.obs, a.chains);
        SVB_CUDA_TRY(cudaGetLastError());
    }
    const int ncol = n_colours(N);
    for (int s = 0; s < a.n_sweeps; ++s) {
        for (int c = 0; c < ncol; ++c) {
            worldline_colour_pass_kernel<MODE, INJECTED>
                <<<(unsigned)blocks, threads, 0, stream>>>(a, s, c, bpc, s == a.n_sweeps - 1);
            SVB_CUDA_TRY(cudaGetLastError());
        }
    }
    if (a.obs) {
        long long grid = a.chains < 148 * 8 ? a.chains : 148 * 8;
        worldline_obs_kernel<<<(unsigned)grid, 256, 0, stream>>>(a.m, a.v, a.chains, N, a.W, a.obs, 1);
        SVB_CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

template <int MODE, int NT, int TT, int MINB, int STAGES>
static int launch_worldline_fast(const WorldlineArgs& a, cudaStream_t stream, int sm_count) {
    auto kern = worldline_smem_fast_kernel<MODE, NT, TT, MINB, STAGES>;
    const size_t smem = (size_t)STAGES * NT * NT * 3 * sizeof(int32_t) + 7 * 32 * sizeof(double) + 16;
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    SVB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TT, smem));
    if (per_sm < 1) return fail(SVB_E_UNSUPPORTED, "fast worldline kernel does not fit an SM at N=%d", NT);
    long long grid = (long long)per_sm * sm_count;
    if (grid > a.chains) grid = a.chains;
    kern<<<(unsigned)grid, TT, smem, stream>>>(a);
    SVB_CUDA_TRY(cudaGetLastError());
    return 0;
}

#include "svb_worldline_table.cuh"

template <int MODE>
static int dispatch_worldline(const WorldlineArgs& a, int rng_mode, int path, cudaStream_t stream) {
    int dev = 0, sm_count = 0, max_smem = 0;
    SVB_CUDA_TRY(cudaGetDevice(&dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (path == SVB_PATH_AUTO) path = (worldline_smem_bytes(a.N) <= (size_t)max_smem) ? SVB_PATH_SMEM : SVB_PATH_GLOBAL;
    if (path == SVB_PATH_SMEM) {
        if (worldline_smem_bytes(a.N) > (size_t)max_smem)
            return fail(SVB_E_UNSUPPORTED, "N=%d does not fit shared memory; use SVB_PATH_GLOBAL", a.N);
        const bool aligned = ((uintptr_t)a.m % 16 == 0) && ((uintptr_t)a.v % 16 == 0);
#ifndef SVB_NO_TABLE_KERNEL
        if ((MODE == SVB_WL_JOINT || a.interval <= 2) && rng_mode == SVB_RNG_PHILOX && a.W == 1 && aligned && !a.accept_mask &&
            !a.dS_out) {
            // production path: tabulated integer acceptance thresholds on resident f (svb_worldline_table.cuh)
            switch (a.N) {
                case 16: return launch_worldline_table<MODE, 16, 16>(a, stream, sm_count);
                case 32: return launch_worldline_table<MODE, 32, 8>(a, stream, sm_count);
                case 64: return launch_worldline_table<MODE, 64, SVB_WL_MINB64, SVB_WL_TM64, SVB_WL_STAGES64>(a, stream, sm_count);
                case 128: return launch_worldline_table<MODE, 128, 1>(a, stream, sm_count);        // 192 KiB: one chain per SM
                default: break;
            }
        }
#endif
        if (rng_mode == SVB_RNG_PHILOX && a.W == 1 && aligned) {
            switch (a.N) {
                case 16: return launch_worldline_fast<MODE, 16, 32, 16, 2>(a, stream, sm_count);
                case 32: return launch_worldline_fast<MODE, 32, 128, 6, 2>(a, stream, sm_count);
                case 64: return launch_worldline_fast<MODE, 64, 256, 4, 1>(a, stream, sm_count);
                default: break;
            }
        }
        return rng_mode == SVB_RNG_INJECTED ? launch_worldline_smem<MODE, true>(a, stream, sm_count)
                                            : launch_worldline_smem<MODE, false>(a, stream, sm_count);
    }
    return rng_mode == SVB_RNG_INJECTED ? launch_worldline_global<MODE, true>(a, stream)
                                        : launch_worldline_global<MODE, false>(a, stream);
}

}  // namespace svb

using namespace svb;

extern "C" int svb_worldline_sweep(int32_t* m, int32_t* v, int64_t chains, int N, double kappa, const double* kappa_chain,
                                   int W, int mode, int interval, int n_sweeps, uint64_t seed, uint64_t sweep0,
                                   uint64_t chain0, int rng_mode, int path, const double* inj_u, const int32_t* inj_a,
                                   const int32_t* inj_b, double* obs, uint8_t* accept_mask, double* dS_out, void* stream) {
    if (!m || !v) return fail(SVB_E_NULL, "svb_worldline_sweep: m and v are required");
    if (chains < 0 || N < 3 || N > 32768) return fail(SVB_E_SHAPE, "svb_worldline_sweep: chains=%lld N=%d", (long long)chains, N);
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_worldline_sweep: kappa must be positive");
    if (W < 1) return fail(SVB_E_PARAM, "svb_worldline_sweep: W must be a finite integer >= 1 (got %d)", W);
    if (mode < SVB_WL_JOINT || mode > SVB_WL_COEXACT) return fail(SVB_E_PARAM, "svb_worldline_sweep: mode %d", mode);
    if (mode != SVB_WL_JOINT && (interval < 1 || interval > 128)) return fail(SVB_E_PARAM, "svb_worldline_sweep: interval must be in [1, 128]");
    if (n_sweeps < 0) return fail(SVB_E_PARAM, "svb_worldline_sweep: n_sweeps < 0");
    if (rng_mode != SVB_RNG_PHILOX && rng_mode != SVB_RNG_INJECTED) return fail(SVB_E_PARAM, "svb_worldline_sweep: rng_mode");
    if (rng_mode == SVB_RNG_INJECTED && (!inj_u || !inj_a || (mode == SVB_WL_JOINT && !inj_b)))
        return fail(SVB_E_NULL, "svb_worldline_sweep: injected mode needs inj_u, inj_a (and inj_b for JOINT)");
    if (path < SVB_PATH_AUTO || path > SVB_PATH_GLOBAL) return fail(SVB_E_PARAM, "svb_worldline_sweep: path");
    if (chains == 0 || n_sweeps == 0) return SVB_OK;

    WorldlineArgs a;
    a.m = m; a.v = v; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W;
    a.interval = interval; a.n_sweeps = n_sweeps; a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    for (int r = 0; r < 10; ++r) {
        a.round_key[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        a.round_key[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    a.stream = (mode == SVB_WL_JOINT) ? STREAM_WORLDLINE_PLAQUETTE : (mode == SVB_WL_VORTEX) ? STREAM_WORLDLINE_VORTEX : STREAM_WORLDLINE_COEXACT;
    a.refine_stream = (mode == SVB_WL_JOINT) ? STREAM_WORLDLINE_REFINE
                                             : (mode == SVB_WL_VORTEX) ? STREAM_WORLDLINE_VORTEX_REFINE : STREAM_WORLDLINE_COEXACT_REFINE;
    a.inj_u = inj_u; a.inj_a = inj_a; a.inj_b = inj_b; a.obs = obs; a.accept_mask = accept_mask; a.dS_out = dS_out;
    a.ov.epochs = nullptr; a.ov.wait_epoch = 0; a.ov.signal_epoch = 0; a.ov.grid_wait = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (mode) {
        case SVB_WL_JOINT: return dispatch_worldline<SVB_WL_JOINT>(a, rng_mode, path, st);
        case SVB_WL_VORTEX: return dispatch_worldline<SVB_WL_VORTEX>(a, rng_mode, path, st);
        default: return dispatch_worldline<SVB_WL_COEXACT>(a, rng_mode, path, st);
    }
}

extern "C" int svb_worldline_sweep_overlapped(int32_t* m, int32_t* v, int64_t chains, int N, double kappa, const double* kappa_chain,
                                              int mode, int interval, int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                                              double* obs,
                                              uint32_t* epochs, uint32_t wait_epoch, uint32_t signal_epoch, int flags,
                                              void* stream) {
    if (!m || !v || !epochs) return fail(SVB_E_NULL, "svb_worldline_sweep_overlapped: m, v and epochs are required");
    if (chains < 0) return fail(SVB_E_SHAPE, "svb_worldline_sweep_overlapped: chains=%lld", (long long)chains);
    if (N != 16 && N != 32 && N != 64 && N != 128)
        return fail(SVB_E_UNSUPPORTED, "svb_worldline_sweep_overlapped: N must be 16, 32, 64 or 128 (got %d)", N);
    if (((uintptr_t)m % 16) || ((uintptr_t)v % 16)) return fail(SVB_E_ALIGN, "svb_worldline_sweep_overlapped: fields must be 16-byte aligned");
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_worldline_sweep_overlapped: kappa must be positive");
    if (mode < SVB_WL_JOINT || mode > SVB_WL_COEXACT) return fail(SVB_E_PARAM, "svb_worldline_sweep_overlapped: mode %d", mode);
    if (mode != SVB_WL_JOINT && (interval < 1 || interval > 2))
        return fail(SVB_E_UNSUPPORTED, "svb_worldline_sweep_overlapped: interval must be 1 or 2 for VORTEX / COEXACT (got %d)", interval);
    if (n_sweeps < 1) return fail(SVB_E_PARAM, "svb_worldline_sweep_overlapped: n_sweeps must be >= 1 (every launch signals its epoch)");
    if (flags & ~SVB_OVERLAP_PREDECESSOR) return fail(SVB_E_PARAM, "svb_worldline_sweep_overlapped: flags");
    if (chains == 0) return SVB_OK;
    WorldlineArgs a;
    a.m = m; a.v = v; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = 1;
    a.interval = (mode == SVB_WL_JOINT) ? 1 : interval; a.n_sweeps = n_sweeps; a.seed = seed; a.sweep0 = sweep0; a.chain0 = chain0;
    for (int r = 0; r < 10; ++r) {
        a.round_key[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        a.round_key[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    a.stream = (mode == SVB_WL_JOINT) ? STREAM_WORLDLINE_PLAQUETTE : (mode == SVB_WL_VORTEX) ? STREAM_WORLDLINE_VORTEX : STREAM_WORLDLINE_COEXACT;
    a.refine_stream = (mode == SVB_WL_JOINT) ? STREAM_WORLDLINE_REFINE
                                             : (mode == SVB_WL_VORTEX) ? STREAM_WORLDLINE_VORTEX_REFINE : STREAM_WORLDLINE_COEXACT_REFINE;
    a.inj_u = nullptr; a.inj_a = nullptr; a.inj_b = nullptr; a.obs = obs; a.accept_mask = nullptr; a.dS_out = nullptr;
    a.ov.epochs = epochs; a.ov.wait_epoch = wait_epoch; a.ov.signal_epoch = signal_epoch;
    a.ov.grid_wait = (flags & SVB_OVERLAP_PREDECESSOR) ? 0 : 1;
    int dev = 0, sm_count = 0;
    SVB_CUDA_TRY(cudaGetDevice(&dev));
    SVB_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define SVB_WL_OV_DISPATCH(M)                                                                  \
    switch (N) {                                                                               \
        case 16: return launch_worldline_table<M, 16, 16>(a, st, sm_count);                    \
        case 32: return launch_worldline_table<M, 32, 8>(a, st, sm_count);                     \
        case 64: return launch_worldline_table<M, 64, SVB_WL_MINB64, SVB_WL_TM64, SVB_WL_STAGES64>(a, st, sm_count);                     \
        default: return launch_worldline_table<M, 128, 1>(a, st, sm_count);                    \
    }
    if (mode == SVB_WL_JOINT) { SVB_WL_OV_DISPATCH(SVB_WL_JOINT) }
    if (mode == SVB_WL_VORTEX) { SVB_WL_OV_DISPATCH(SVB_WL_VORTEX) }
    SVB_WL_OV_DISPATCH(SVB_WL_COEXACT)
#undef SVB_WL_OV_DISPATCH
}

// ------------------------------------------------------------------------------------------
// WrappingUpdate (supervillain/generator/worldline/wrapping.py:43-90): one proposal per torus cycle.
// The mu-direction cycle at perpendicular coordinate k changes m_mu by c on its N links;
//   dS = sum_j ((0.5/kappa) c) ((2 f_j) + c),  f = m - delta v / W  (:64, :71)
// All 2N cycles of a chain read the initial fields and touch disjoint links, so they are decided
// concurrently: one thread per cycle out of shared memory.  2N proposals per chain against N^2 per
// plaquette sweep -- this kernel is never the bottleneck.
// ------------------------------------------------------------------------------------------
namespace svb {

struct WrappingArgs {
    int32_t* m;
    const int32_t* v;
    long long chains;
    int N;
    double kappa;
    const double* kappa_chain;
    int W;
    int interval;
    unsigned long long seed, sweep, chain0;
    const double* inj_u;       // (chains, 2, N)
    const int32_t* inj_c;      // (chains, 2, N)
    double* counters;          // (chains, 2): accepted, sum of min(1, e^-dS)
    double* dS_out;            // (chains, 2, N) optional
};

template <bool INJECTED>
__global__ void __launch_bounds__(128) worldline_wrapping_kernel(WrappingArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, V = N * N;
    int32_t* sm = reinterpret_cast<int32_t*>(smem_raw);       // m: 2V
    int32_t* sv = sm + 2 * V;                                  // v: V
    __shared__ double red[2 * 32];
    const double Wd = (double)a.W;
    for (long long chain = blockIdx.x; chain < a.chains; chain += gridDim.x) {
        int32_t* gm = a.m + chain * 2 * V;
        const int32_t* gv = a.v + chain * V;
        for (int i = threadIdx.x; i < 2 * V; i += blockDim.x) sm[i] = gm[i];
        for (int i = threadIdx.x; i < V; i += blockDim.x) sv[i] = gv[i];
        __syncthreads();
        const double kappa = a.kappa_chain ? a.kappa_chain[chain] : a.kappa;
        const double hk = __ddiv_rn(0.5, kappa);
        double n_acc = 0.0, sum_A = 0.0;
        for (int cyc = threadIdx.x; cyc < 2 * N; cyc += blockDim.x) {
            const int mu = cyc / N, k = cyc - mu * N;
            double u;
            int c;
            if (INJECTED) {
                u = a.inj_u[chain * 2 * N + cyc];
                c = a.inj_c[chain * 2 * N + cyc];
            } else {
                const Philox4 p = philox_site(a.seed, a.chain0 + (unsigned long long)chain, a.sweep, (uint32_t)cyc, STREAM_WORLDLINE_WRAPPING);
                const uint64_t ku = ((uint64_t)p.x << 12) | (uint64_t)(p.y >> 20);
                u = ((double)(long long)ku + 0.5) * 5.6843418860808015e-14;
                const int idx = (int)(((uint64_t)p.w * (uint64_t)(2 * a.interval)) >> 32);
                c = (idx < a.interval) ? idx - a.interval : idx - a.interval + 1;
            }
            const double cd = (double)c;
            const double hc = __dmul_rn(hk, cd);
            double dS = 0.0;
            for (int j = 0; j < N; ++j) {
                // link (mu, x): mu = 0 -> x = (j, k);  mu = 1 -> x = (k, j)
                const int x0 = (mu == 0) ? j : k, x1 = (mu == 0) ? k : j;
                const int i = x0 * N + x1;
                const int dv = (mu == 0) ? (sv[i] - sv[x0 * N + (x1 == 0 ? N - 1 : x1 - 1)])
                                         : (sv[(x0 == 0 ? N - 1 : x0 - 1) * N + x1] - sv[i]);
                const double f = __dsub_rn((double)sm[mu * V + i], __ddiv_rn((double)dv, Wd));
                dS = __dadd_rn(dS, __dmul_rn(hc, __dadd_rn(__dmul_rn(2.0, f), cd)));
            }
            const double A = exp_clipped(-dS);
            const bool ok = u < A;
            sum_A += A;
            if (a.dS_out) a.dS_out[chain * 2 * N + cyc] = dS;
            if (ok) {
                n_acc += 1.0;
                for (int j = 0; j < N; ++j) {
                    const int i = (mu == 0) ? (j * N + k) : (k * N + j);
                    gm[mu * V + i] = sm[mu * V + i] + c;       // cycles touch disjoint links: write straight to HBM
                }
            }
        }
        if (a.counters) {
            double s[2] = {n_acc, sum_A};
            block_sum<2>(s, red);
            if (threadIdx.x == 0) {
                a.counters[chain * 2] = s[0];
                a.counters[chain * 2 + 1] = s[1];
            }
        }
        __syncthreads();
    }
}

}  // namespace svb

extern "C" int svb_worldline_wrapping(int32_t* m, const int32_t* v, int64_t chains, int N, double kappa, const double* kappa_chain,
                                      int W, int interval, uint64_t seed, uint64_t sweep, uint64_t chain0, int rng_mode,
                                      const double* inj_u, const int32_t* inj_c, double* counters, double* dS_out, void* stream) {
    if (!m || !v) return fail(SVB_E_NULL, "svb_worldline_wrapping: m and v are required");
    if (chains < 0 || N < 3 || N > 181) return fail(SVB_E_SHAPE, "svb_worldline_wrapping: chains=%lld N=%d (needs the lattice in shared memory)", (long long)chains, N);
    if (!kappa_chain && !(kappa > 0)) return fail(SVB_E_PARAM, "svb_worldline_wrapping: kappa must be positive");
    if (W < 1 || interval < 1 || interval > 1024) return fail(SVB_E_PARAM, "svb_worldline_wrapping: W / interval");
    if (rng_mode == SVB_RNG_INJECTED && (!inj_u || !inj_c)) return fail(SVB_E_NULL, "svb_worldline_wrapping: injected mode needs inj_u and inj_c");
    if (chains == 0) return SVB_OK;
    WrappingArgs a;
    a.m = m; a.v = v; a.chains = chains; a.N = N; a.kappa = kappa; a.kappa_chain = kappa_chain; a.W = W; a.interval = interval;
    a.seed = seed; a.sweep = sweep; a.chain0 = chain0; a.inj_u = inj_u; a.inj_c = inj_c; a.counters = counters; a.dS_out = dS_out;
    const size_t smem = (size_t)3 * N * N * sizeof(int32_t);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long grid = chains < 148LL * 8 ? chains : 148LL * 8;
    if (rng_mode == SVB_RNG_INJECTED) {
        SVB_CUDA_TRY(cudaFuncSetAttribute(worldline_wrapping_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        worldline_wrapping_kernel<true><<<(unsigned)grid, 128, smem, st>>>(a);
    } else {
        SVB_CUDA_TRY(cudaFuncSetAttribute(worldline_wrapping_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        worldline_wrapping_kernel<false><<<(unsigned)grid, 128, smem, st>>>(a);
    }
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}

extern "C" int svb_worldline_observables(const int32_t* m, const int32_t* v, int64_t chains, int N, int W, double* obs,
                                         void* stream) {
    if (!m || !v || !obs) return fail(SVB_E_NULL, "svb_worldline_observables: m, v, obs are required");
    if (chains < 0 || N < 3 || N > 32768) return fail(SVB_E_SHAPE, "svb_worldline_observables: shape");
    if (W < 1) return fail(SVB_E_PARAM, "svb_worldline_observables: W");
    if (chains == 0) return SVB_OK;
    if (W == 1 && chains >= 64 && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0)) {
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        switch (N) {          // TMA-staged integer pass (svb_worldline_table.cuh)
            case 16: return launch_worldline_obs_smem<16>(m, v, chains, obs, 0, st);
            case 32: return launch_worldline_obs_smem<32>(m, v, chains, obs, 0, st);
            case 64: return launch_worldline_obs_smem<64>(m, v, chains, obs, 0, st);
            default: break;
        }
    }
    long long grid = chains < 148 * 8 ? chains : 148 * 8;
    worldline_obs_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(m, v, chains, N, W, obs, 0);
    SVB_CUDA_TRY(cudaGetLastError());
    return SVB_OK;
}
