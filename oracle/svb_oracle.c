/*
 * svb_oracle.c -- scalar C restatement of the reference's hot-path algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the fast CHECKER for the CUDA path at full benchmark
 * sizes (the numpy restatements in this directory are the readable ones and are what is pinned
 * against the reference's golden vectors; tests/test_oracle_c.py pins this file against them).
 * It is never linked into, imported by, or executed from supervillain_b200/.
 *
 * Restated here, with the reference lines each function follows:
 *   villain_sweep_dense      NeighborhoodUpdate.step, supervillain/generator/villain/neighborhood.py:59-137,
 *                            per site in the reference's colour order with the residual r carried
 *                            incrementally across colours exactly as the reference does (:91, :129)
 *   villain_action           Villain.__call__, supervillain/action/villain.py:51-66
 *   worldline_sweep_dense    the per-plaquette arithmetic of PlaquetteUpdate (worldline/plaquette.py:79-101),
 *                            VortexUpdate (worldline/vortex.py:108-128) and CoexactUpdate
 *                            (worldline/coexact.py:102-120) in red/black order
 *   colour                   Lattice.checkerboarding, supervillain/lattice/compact.py:192-239
 *   philox4x32_10 and the draw mappings: the published Philox algorithm (Salmon et al., SC'11) and
 *                            the bit-to-proposal mapping documented in supervillain_b200/csrc/svb_villain.cu
 *
 * Compile with -ffp-contract=off: every numpy operation of the reference is one IEEE rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TWO_PI 6.283185307179586476925286766559

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10                                                                               */
/* ------------------------------------------------------------------------------------------ */
void svo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox_site(uint64_t seed, uint64_t chain, uint64_t sweep, uint32_t site, uint32_t stream, uint32_t out[4]) {
    uint32_t ctr[4], key[2];
    ctr[0] = site;
    ctr[1] = (uint32_t)chain;
    ctr[2] = (uint32_t)sweep;
    ctr[3] = (stream << 24) | ((uint32_t)((chain >> 32) & 0xFFu) << 16) | (uint32_t)((sweep >> 32) & 0xFFFFu);
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    svo_philox4x32_10(ctr, key, out);
}

/* draw mapping version 2 (see svb_villain.cu): 64 bits per site from the Philox block its pair shares, and the
 * trailing bits of the uniform from the refinement stream (NeighborhoodUpdate: streams 1 and 4).  Wide dn intervals
 * (K^4 > 256): given the digits, the remainder of word B only takes every K^4-th value, so the uniform's leading bits are
 * refinement word 2 half and its trailing bits refinement word 2 half + 1 instead. */
void svo_villain_draw(uint64_t seed, uint64_t chain, uint64_t sweep, uint32_t site, int N, int W, double interval_phi,
                      int interval_n, double* u, double* dphi, int dn[4]) {
    uint32_t w[4], r[4];
    int x0 = (int)(site / (uint32_t)N), x1 = (int)(site % (uint32_t)N);
    uint32_t c0 = (uint32_t)((x0 & ~8) * N + x1);
    int half = (x0 >> 3) & 1;
    philox_site(seed, chain, sweep, c0, 1u, w);
    philox_site(seed, chain, sweep, c0, 4u, r);
    uint32_t A = w[2 * half], B = w[2 * half + 1], e = r[2 * half];
    double Uphi = ((double)A + 0.5) * 0x1p-32;
    double prod = (2.0 * interval_phi) * Uphi;
    *dphi = -interval_phi + prod;
    uint32_t K = (uint32_t)(2 * interval_n + 1);
    uint32_t f = B;
    for (int i = 0; i < 4; ++i) {
        uint64_t p = (uint64_t)f * K;
        f = (uint32_t)p;
        dn[i] = W * ((int)(p >> 32) - interval_n);
    }
    if ((uint64_t)K * K * K * K > 256u) {
        f = r[2 * half];
        e = r[2 * half + 1];
    }
    double frac = ((double)e + 0.5) * 0x1p-32;
    double uu = ((double)f + frac) * 0x1p-32;
    *u = uu < 0x1.fffffffffffffp-1 ? uu : 0x1.fffffffffffffp-1;
}

/* mode 0 joint: a = dm in {-1,+1}, b = dv in {-1,0,+1}; modes 1, 2: a in [-I..-1, 1..I].
 * Draw mapping version 2 (see svb_worldline.cu): one word of the Philox block four plaquettes share. */
void svo_worldline_draw(uint64_t seed, uint64_t chain, uint64_t sweep, uint32_t site, int N, int mode, int interval, double* u,
                        int* a, int* b) {
    uint32_t blk[4], ref[4];
    int x0 = (int)(site / (uint32_t)N), x1 = (int)(site % (uint32_t)N);
    uint32_t c0 = (uint32_t)((x0 & ~24) * N + x1);
    int word = (x0 >> 3) & 3;
    /* streams (proposal, refinement) by generator kind: joint 2 / 5, vortex 13 / 14, coexact 15 / 16 */
    philox_site(seed, chain, sweep, c0, mode == 0 ? 2u : (mode == 1 ? 13u : 15u), blk);
    philox_site(seed, chain, sweep, c0, mode == 0 ? 5u : (mode == 1 ? 14u : 16u), ref);
    uint32_t w = blk[word], e = ref[word];
    uint64_t p;
    if (mode == 0) {
        *a = (w >> 31) ? +1 : -1;
        p = (uint64_t)(uint32_t)(w << 1) * 3ull;
        *b = (int)(p >> 32) - 1;
    } else {
        p = (uint64_t)w * (uint64_t)(2 * interval);
        int idx = (int)(p >> 32);
        *a = (idx < interval) ? idx - interval : idx - interval + 1;
        *b = 0;
    }
    uint32_t f = (uint32_t)p;
    double frac = ((double)e + 0.5) * 0x1p-32;
    double uu = ((double)f + frac) * 0x1p-32;
    *u = uu < 0x1.fffffffffffffp-1 ? uu : 0x1.fffffffffffffp-1;
}

/* ------------------------------------------------------------------------------------------ */
/* colouring                                                                                   */
/* ------------------------------------------------------------------------------------------ */
int svo_n_colours(int N) { return (N % 2 == 0) ? 2 : 4; }

int svo_colour(int x0, int x1, int N) {
    int h = N / 2;
    int c0 = (x0 <= h) ? x0 : x0 - N;     /* lattice/__init__.py:4-9 */
    int c1 = (x1 <= h) ? x1 : x1 - N;
    int parity = ((c0 + c1) % 2 + 2) % 2;
    if (N % 2 == 0) return parity;
    int mixed = ((c0 >= 0) != (c1 >= 0)) ? 1 : 0;
    return 2 * mixed + parity;
}

/* ------------------------------------------------------------------------------------------ */
/* Villain                                                                                     */
/* ------------------------------------------------------------------------------------------ */
double svo_villain_action(const double* phi, const int64_t* n, int N, double kappa) {
    /* plain left-to-right sum: agrees with numpy's pairwise sum to ~1e-15 relative, not bitwise */
    long double s = 0;
    for (int x0 = 0; x0 < N; ++x0)
        for (int x1 = 0; x1 < N; ++x1) {
            int i = x0 * N + x1;
            double r0 = (phi[((x0 + 1) % N) * N + x1] - phi[i]) - TWO_PI * (double)n[i];
            double r1 = (phi[x0 * N + (x1 + 1) % N] - phi[i]) - TWO_PI * (double)n[N * N + i];
            s += (long double)r0 * r0 + (long double)r1 * r1;
        }
    return (kappa / 2) * (double)s;
}

/*
 * One sweep of one chain with dense per-site draws (u, dphi: N*N; dn_fwd, dn_bwd: 2*N*N, indexed
 * by the proposing site).  phi (N*N) and n (2*N*N) are updated in place.  Optional outputs:
 * accept (N*N bytes), dS (N*N).  Returns the number of accepted proposals; *acceptance gets the
 * sum of min(1, e^-dS).  `scratch_r` must hold 2*N*N doubles.
 */
int svo_villain_sweep_dense(double* phi, int64_t* n, int N, double kappa, const double* u, const double* dphi,
                            const int64_t* dn_fwd, const int64_t* dn_bwd, double* scratch_r, uint8_t* accept, double* dS_out,
                            double* acceptance) {
    const int V = N * N;
    double* r0 = scratch_r;
    double* r1 = scratch_r + V;
    int64_t* n0 = n;
    int64_t* n1 = n + V;
    for (int x0 = 0; x0 < N; ++x0)                       /* r = d(phi) - 2 pi n   (:91) */
        for (int x1 = 0; x1 < N; ++x1) {
            int i = x0 * N + x1;
            r0[i] = (phi[((x0 + 1) % N) * N + x1] - phi[i]) - TWO_PI * (double)n0[i];
            r1[i] = (phi[x0 * N + (x1 + 1) % N] - phi[i]) - TWO_PI * (double)n1[i];
        }
    const double hk = kappa / 2;
    int accepted = 0;
    double acc_sum = 0.0;
    const int ncol = svo_n_colours(N);
    for (int c = 0; c < ncol; ++c) {                     /* :93 */
        for (int x0 = 0; x0 < N; ++x0)
            for (int x1 = 0; x1 < N; ++x1) {
                if (svo_colour(x0, x1, N) != c) continue;
                int i = x0 * N + x1;
                int ib0 = ((x0 + N - 1) % N) * N + x1;
                int ib1 = x0 * N + (x1 + N - 1) % N;
                double dp = dphi[i];
                double t;
                t = TWO_PI * (double)dn_fwd[i];      double dr_f0 = (0.0 - dp) - t;    /* :110 */
                t = TWO_PI * (double)dn_bwd[i];      double dr_b0 = (dp - 0.0) - t;
                t = TWO_PI * (double)dn_fwd[V + i];  double dr_f1 = (0.0 - dp) - t;
                t = TWO_PI * (double)dn_bwd[V + i];  double dr_b1 = (dp - 0.0) - t;
                double a, b2;
                a = hk * dr_f0; b2 = 2 * r0[i] + dr_f0;   double s_f0 = a * b2;          /* :111 */
                a = hk * dr_b0; b2 = 2 * r0[ib0] + dr_b0; double s_b0 = a * b2;
                a = hk * dr_f1; b2 = 2 * r1[i] + dr_f1;   double s_f1 = a * b2;
                a = hk * dr_b1; b2 = 2 * r1[ib1] + dr_b1; double s_b1 = a * b2;
                double dS = 0.0 + s_f0;                                                  /* :112 */
                dS = dS + s_b0;
                dS = dS + s_f1;
                dS = dS + s_b1;
                double A = exp(-dS);                                                     /* :115 */
                if (A > 1.0) A = 1.0;
                if (A < 0.0) A = 0.0;
                int ok = u[i] < A;                                                       /* :116 */
                acc_sum += A;
                if (dS_out) dS_out[i] = dS;
                if (accept) accept[i] = (uint8_t)ok;
                if (ok) {
                    ++accepted;
                    phi[i] = phi[i] + dp;                                                /* :127 */
                    n0[i] += dn_fwd[i];                                                  /* :128 */
                    n0[ib0] += dn_bwd[i];
                    n1[i] += dn_fwd[V + i];
                    n1[ib1] += dn_bwd[V + i];
                    t = TWO_PI * (double)dn_fwd[i];      r0[i] = (r0[i] + (0.0 - dp)) - t;       /* :129 */
                    t = TWO_PI * (double)dn_bwd[i];      r0[ib0] = (r0[ib0] + (dp - 0.0)) - t;
                    t = TWO_PI * (double)dn_fwd[V + i];  r1[i] = (r1[i] + (0.0 - dp)) - t;
                    t = TWO_PI * (double)dn_bwd[V + i];  r1[ib1] = (r1[ib1] + (dp - 0.0)) - t;
                }
            }
    }
    if (acceptance) *acceptance = acc_sum;
    return accepted;
}

/*
 * `n_sweeps` sweeps on `chains` chains with the kernels' Philox draws: chain c, sweep s uses
 * counter (site, chain0 + c, sweep0 + s).  phi (chains*N*N), n (chains*2*N*N int64) in place.
 * accepted_out / acceptance_out (per chain, summed over the sweeps) may be NULL.
 */
int svo_villain_sweep_philox(double* phi, int64_t* n, int64_t chains, int N, double kappa, int W, double interval_phi,
                             int interval_n, int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                             int64_t* accepted_out, double* acceptance_out) {
    const int V = N * N;
    double* u = (double*)malloc(sizeof(double) * V);
    double* dphi = (double*)malloc(sizeof(double) * V);
    int64_t* dnf = (int64_t*)malloc(sizeof(int64_t) * 2 * V);
    int64_t* dnb = (int64_t*)malloc(sizeof(int64_t) * 2 * V);
    double* scratch = (double*)malloc(sizeof(double) * 2 * V);
    if (!u || !dphi || !dnf || !dnb || !scratch) return -1;
    for (int64_t c = 0; c < chains; ++c) {
        int64_t acc_total = 0;
        double accp_total = 0.0;
        for (int s = 0; s < n_sweeps; ++s) {
            for (int i = 0; i < V; ++i) {
                int dn[4];
                svo_villain_draw(seed, chain0 + (uint64_t)c, sweep0 + (uint64_t)s, (uint32_t)i, N, W, interval_phi, interval_n,
                                 &u[i], &dphi[i], dn);
                dnf[i] = dn[0]; dnb[i] = dn[1]; dnf[V + i] = dn[2]; dnb[V + i] = dn[3];
            }
            double accp = 0.0;
            acc_total += svo_villain_sweep_dense(phi + c * V, n + c * 2 * V, N, kappa, u, dphi, dnf, dnb, scratch, NULL, NULL, &accp);
            accp_total += accp;
        }
        if (accepted_out) accepted_out[c] = acc_total;
        if (acceptance_out) acceptance_out[c] = accp_total;
    }
    free(u); free(dphi); free(dnf); free(dnb); free(scratch);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Worldline                                                                                   */
/* ------------------------------------------------------------------------------------------ */
/*
 * One red/black sweep of one chain with dense per-plaquette draws.  mode 0 joint (a = dm, b = dv),
 * 1 vortex (a = dv), 2 coexact (a = t).  m (2*N*N), v (N*N) int64 in place.
 */
int svo_worldline_sweep_dense(int64_t* m, int64_t* v, int N, double kappa, int W, int mode, const double* u, const int64_t* a_all,
                              const int64_t* b_all, uint8_t* accept, double* dS_out, double* acceptance) {
    const int V = N * N;
    int64_t* m0 = m;
    int64_t* m1 = m + V;
    const double Wd = (double)W;
    const double hk = 0.5 / kappa;
    int accepted = 0;
    double acc_sum = 0.0;
    const int ncol = svo_n_colours(N);
    for (int c = 0; c < ncol; ++c)
        for (int x0 = 0; x0 < N; ++x0)
            for (int x1 = 0; x1 < N; ++x1) {
                if (svo_colour(x0, x1, N) != c) continue;
                int i = x0 * N + x1;
                int ip0 = ((x0 + 1) % N) * N + x1, im0 = ((x0 + N - 1) % N) * N + x1;
                int ip1 = x0 * N + (x1 + 1) % N, im1 = x0 * N + (x1 + N - 1) % N;
                int64_t vc = v[i];
                double q;
                q = (double)(vc - v[im1]) / Wd;  double f_0x = (double)m0[i] - q;        /* f = m - delta(v)/W */
                q = (double)(v[ip1] - vc) / Wd;  double f_0p = (double)m0[ip1] - q;
                q = (double)(v[im0] - vc) / Wd;  double f_1x = (double)m1[i] - q;
                q = (double)(vc - v[ip0]) / Wd;  double f_1p = (double)m1[ip0] - q;
                int64_t a = a_all[i], b = b_all ? b_all[i] : 0;
                double dS;
                if (mode == 0) {                                                        /* plaquette.py:84-85 */
                    q = (double)b / Wd;
                    double df = (double)a - q;
                    double s = f_0x + f_1p;
                    s = s - f_0p;
                    s = s - f_1x;
                    q = 2 * df;
                    s = s + q;
                    q = df / kappa;
                    dS = q * s;
                } else if (mode == 1) {                                                 /* vortex.py:108-117 */
                    double cp = (double)a / Wd, cn = (double)(-a) / Wd;
                    double t1, t2;
                    t1 = hk * (-cn); t2 = 2 * f_1x; t2 = t2 - cn; dS = t1 * t2;
                    t1 = hk * (-cp); t2 = 2 * f_1p; t2 = t2 - cp; dS = dS + t1 * t2;
                    t1 = hk * (-cp); t2 = 2 * f_0x; t2 = t2 - cp; dS = dS + t1 * t2;
                    t1 = hk * (-cn); t2 = 2 * f_0p; t2 = t2 - cn; dS = dS + t1 * t2;
                } else {                                                                /* coexact.py:102-111 */
                    double cp = (double)a, cn = (double)(-a);
                    double t1, t2;
                    t1 = hk * cn; t2 = 2 * f_1x; t2 = t2 + cn; dS = t1 * t2;
                    t1 = hk * cp; t2 = 2 * f_1p; t2 = t2 + cp; dS = dS + t1 * t2;
                    t1 = hk * cp; t2 = 2 * f_0x; t2 = t2 + cp; dS = dS + t1 * t2;
                    t1 = hk * cn; t2 = 2 * f_0p; t2 = t2 + cn; dS = dS + t1 * t2;
                }
                double A = exp(-dS);
                if (A > 1.0) A = 1.0;
                int ok = u[i] < A;
                acc_sum += A;
                if (dS_out) dS_out[i] = dS;
                if (accept) accept[i] = (uint8_t)ok;
                if (ok) {
                    ++accepted;
                    if (mode == 0 || mode == 2) {
                        m0[i] += a; m1[ip0] += a; m0[ip1] -= a; m1[i] -= a;
                    }
                    if (mode == 0) v[i] += b;
                    if (mode == 1) v[i] += a;
                }
            }
    if (acceptance) *acceptance = acc_sum;
    return accepted;
}

int svo_worldline_sweep_philox(int64_t* m, int64_t* v, int64_t chains, int N, double kappa, int W, int mode, int interval,
                               int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0, int64_t* accepted_out,
                               double* acceptance_out) {
    const int V = N * N;
    double* u = (double*)malloc(sizeof(double) * V);
    int64_t* a = (int64_t*)malloc(sizeof(int64_t) * V);
    int64_t* b = (int64_t*)malloc(sizeof(int64_t) * V);
    if (!u || !a || !b) return -1;
    for (int64_t c = 0; c < chains; ++c) {
        int64_t acc_total = 0;
        double accp_total = 0.0;
        for (int s = 0; s < n_sweeps; ++s) {
            for (int i = 0; i < V; ++i) {
                int ai, bi;
                svo_worldline_draw(seed, chain0 + (uint64_t)c, sweep0 + (uint64_t)s, (uint32_t)i, N, mode, interval, &u[i], &ai, &bi);
                a[i] = ai; b[i] = bi;
            }
            double accp = 0.0;
            acc_total += svo_worldline_sweep_dense(m + c * 2 * V, v + c * V, N, kappa, W, mode, u, a, b, NULL, NULL, &accp);
            accp_total += accp;
        }
        if (accepted_out) accepted_out[c] = acc_total;
        if (acceptance_out) acceptance_out[c] = accp_total;
    }
    free(u); free(a); free(b);
    return 0;
}
