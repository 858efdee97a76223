"""Independent numpy restatement of Philox4x32-10 and of the kernel's draw mappings.

TEST INFRASTRUCTURE ONLY.  The reference draws from numpy's PCG64; the GPU path's production
RNG is a counter-based Philox4x32-10 (Salmon et al., SC'11 -- "Random123"; published algorithm,
known-answer vectors from its kat_vectors file are checked in tests/test_philox.py).  This module
lets the tests regenerate, on the CPU, exactly the proposals a SVB_RNG_PHILOX kernel launch uses,
feed them to the restated reference algorithm, and demand identical fields from the GPU.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_VILLAIN_NEIGHBORHOOD = 1
STREAM_WORLDLINE_PLAQUETTE = 2
STREAM_VILLAIN_REFINE = 4
STREAM_WORLDLINE_REFINE = 5
STREAM_VILLAIN_LINK = 6
STREAM_VILLAIN_LINK_REFINE = 7
STREAM_VILLAIN_COHOMOLOGY = 8
STREAM_VILLAIN_SITE = 9
STREAM_VILLAIN_SITE_REFINE = 10
STREAM_VILLAIN_EXACT = 11
STREAM_VILLAIN_EXACT_REFINE = 12
STREAM_WORLDLINE_VORTEX = 13
STREAM_WORLDLINE_VORTEX_REFINE = 14
STREAM_WORLDLINE_COEXACT = 15
STREAM_WORLDLINE_COEXACT_REFINE = 16
# every generator kind owns a (proposal, refinement) stream pair: generators sharing a seed never share a Philox block
VILLAIN_STREAMS = {'neighborhood': (STREAM_VILLAIN_NEIGHBORHOOD, STREAM_VILLAIN_REFINE),
                   'site': (STREAM_VILLAIN_SITE, STREAM_VILLAIN_SITE_REFINE),
                   'exact': (STREAM_VILLAIN_EXACT, STREAM_VILLAIN_EXACT_REFINE)}
WORLDLINE_STREAMS = {'joint': (STREAM_WORLDLINE_PLAQUETTE, STREAM_WORLDLINE_REFINE),
                     'vortex': (STREAM_WORLDLINE_VORTEX, STREAM_WORLDLINE_VORTEX_REFINE),
                     'coexact': (STREAM_WORLDLINE_COEXACT, STREAM_WORLDLINE_COEXACT_REFINE)}


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counter words are uint64 arrays holding 32-bit values; keys are ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_site(seed, chain, sweep, site, stream_id):
    """The kernels' counter layout: (site, chain[31:0], sweep[31:0], stream<<24 | chain[39:32]<<16 | sweep[47:32])."""
    seed = int(seed) & (2**64 - 1)
    chain = int(chain)
    sweep = int(sweep)
    c3 = (stream_id << 24) | (((chain >> 32) & 0xFF) << 16) | ((sweep >> 32) & 0xFFFF)
    site = np.asarray(site, dtype=np.uint64)
    return philox4x32_10(site, np.full_like(site, chain & 0xFFFFFFFF), np.full_like(site, sweep & 0xFFFFFFFF),
                         np.full_like(site, c3), seed & 0xFFFFFFFF, seed >> 32)


TWO_M44 = 2.0 ** -44
TWO_M52 = 2.0 ** -52
TWO_M32 = 2.0 ** -32


def villain_draws(seed, chain, sweep, N, W=1, interval_phi=np.pi, interval_n=1, kind='neighborhood'):
    """Dense per-site proposals of one chain and sweep, in the layout of villain_np.draw_neighborhood.

    Draw mapping, version 2 (documented in supervillain_b200/csrc/svb_villain.cu): sites (x0, x1) and (x0 ^ 8, x1)
    share the Philox block with counter word 0 = (x0 & ~8) N + x1; the site with bit 3 of x0 clear owns words
    (A, B) = (0, 1), the other one words (2, 3):
      dphi = -I + (2I) * ((A + 1/2) 2^-32)
      dn   = W * (digit_i - interval_n), digit_i = the four leading base-K digits (K = 2 interval_n + 1)
             of the fraction B / 2^32, ordered (fwd 0, bwd 0, fwd 1, bwd 1)
      u    = min(fl(f + (e + 1/2) 2^-32) 2^-32, 1 - 2^-53), f = the remainder of B after the four digits,
             e = word 0 or 2 (by half) of the block with the same counter in the kind's refinement stream
    (the kernels generate e only when u < A is not already decided by f; the decision is the same).
    Given the digits, f only takes every K^4-th value; for K^4 > 256 (interval_n >= 2, "wide") the acceptance probability
    would be visibly quantised, so there f = refinement word 0 or 2 (by half) and e = refinement word 1 or 3.
    kind: 'neighborhood' (streams 1, 4) or 'site' (SiteUpdate = the same mapping with interval_n = 0; streams 9, 10).
    """
    x0, x1 = np.divmod(np.arange(N * N, dtype=np.int64), N)
    c0 = ((x0 & ~8) * N + x1).astype(np.uint64)
    half = ((x0 >> 3) & 1).astype(bool)
    stream, refine = VILLAIN_STREAMS[kind]
    w = philox_site(seed, chain, sweep, c0, stream)
    r = philox_site(seed, chain, sweep, c0, refine)
    A = np.where(half, w[2], w[0])
    B = np.where(half, w[3], w[1])
    e = np.where(half, r[2], r[0])
    Uphi = (A.astype(np.float64) + 0.5) * TWO_M32
    dphi = -interval_phi + (2.0 * interval_phi) * Uphi
    K = np.uint64(2 * interval_n + 1)
    f = B
    digits = []
    for _ in range(4):
        prod = f * K
        digits.append((prod >> np.uint64(32)).astype(np.int64) - interval_n)
        f = prod & MASK32
    if int(K) ** 4 > 256:
        f, e = e, np.where(half, r[3], r[1])
    frac = (e.astype(np.float64) + 0.5) * TWO_M32
    u = np.minimum((f.astype(np.float64) + frac) * TWO_M32, 1.0 - 2.0 ** -53)
    dn_fwd = np.stack([W * digits[0], W * digits[2]]).reshape(2, N, N)
    dn_bwd = np.stack([W * digits[1], W * digits[3]]).reshape(2, N, N)
    return {'u': u.reshape(N, N), 'dphi': dphi.reshape(N, N), 'dn_fwd': dn_fwd, 'dn_bwd': dn_bwd}


def worldline_draws(seed, chain, sweep, N, mode, interval=1):
    """Dense per-plaquette proposals: u, a (dm | dv | t) and, for mode 'joint', b (dv).

    Draw mapping, version 2 (documented in supervillain_b200/csrc/svb_worldline.cu): the plaquettes (x0 with bits 3 and
    4 varied, x1) share the Philox block with counter word 0 = (x0 & ~24) N + x1; plaquette x0 owns word (x0 >> 3) & 3 = w:
      joint             dm = +1 if bit 31 of w else -1;  p = 3 (w << 1 mod 2^32);  dv = (p >> 32) - 1;  f = p mod 2^32
      vortex / coexact  p = (2 I) w;  idx = p >> 32 picks from [-I..-1, 1..I];  f = p mod 2^32
      u = min(fl(f + (e + 1/2) 2^-32) 2^-32, 1 - 2^-53), e = the same word of the block with the same counter in the kind's
      refinement stream (the kernels generate e only when f alone does not decide u < A; same decision).
      Streams (proposal, refinement): joint (2, 5), vortex (13, 14), coexact (15, 16)."""
    x0, x1 = np.divmod(np.arange(N * N, dtype=np.int64), N)
    c0 = ((x0 & ~24) * N + x1).astype(np.uint64)
    word = (x0 >> 3) & 3
    stream, refine = WORLDLINE_STREAMS[mode]
    blk = philox_site(seed, chain, sweep, c0, stream)
    ref = philox_site(seed, chain, sweep, c0, refine)
    w = np.choose(word, blk)
    e = np.choose(word, ref)
    if mode == 'joint':
        a = np.where((w >> np.uint64(31)) == 1, 1, -1).astype(np.int64)
        p = ((w << np.uint64(1)) & MASK32) * np.uint64(3)
        b = (p >> np.uint64(32)).astype(np.int64) - 1
    else:
        p = w * np.uint64(2 * interval)
        idx = (p >> np.uint64(32)).astype(np.int64)
        a = np.where(idx < interval, idx - interval, idx - interval + 1)
        b = np.zeros_like(a)
    f = p & MASK32
    frac = (e.astype(np.float64) + 0.5) * TWO_M32
    u = np.minimum((f.astype(np.float64) + frac) * TWO_M32, 1.0 - 2.0 ** -53)
    return {'u': u.reshape(N, N), 'a': a.reshape(N, N), 'b': b.reshape(N, N)}


def _lazy_uniform(f, e):
    frac = (e.astype(np.float64) + 0.5) * TWO_M32
    return np.minimum((f.astype(np.float64) + frac) * TWO_M32, 1.0 - 2.0 ** -53)


def _nonzero_choice(idx, interval):
    return np.where(idx < interval, idx - interval, idx - interval + 1)


def villain_exact_draws(seed, chain, sweep, N, interval_z=1):
    """ExactUpdate on the GPU: the site's pair block as in villain_draws; word B picks z among the 2 I nonzero values
    (p = (2 I) B, index = p >> 32) and the remainder leads the uniform (streams STREAM_VILLAIN_EXACT / _EXACT_REFINE)."""
    x0, x1 = np.divmod(np.arange(N * N, dtype=np.int64), N)
    c0 = ((x0 & ~8) * N + x1).astype(np.uint64)
    half = ((x0 >> 3) & 1).astype(bool)
    w = philox_site(seed, chain, sweep, c0, STREAM_VILLAIN_EXACT)
    r = philox_site(seed, chain, sweep, c0, STREAM_VILLAIN_EXACT_REFINE)
    B = np.where(half, w[3], w[1])
    e = np.where(half, r[2], r[0])
    p = B * np.uint64(2 * interval_z)
    z = _nonzero_choice((p >> np.uint64(32)).astype(np.int64), interval_z)
    return {'u': _lazy_uniform(p & MASK32, e).reshape(N, N), 'a': z.reshape(N, N)}


def villain_link_draws(seed, chain, sweep, N, W=1, interval_n=1):
    """LinkUpdate on the GPU: block with counter word 0 = site index in STREAM_VILLAIN_LINK, word mu for link (mu, x):
    p = (2 I) w, change = W * nonzero[p >> 32], remainder -> uniform (refined from STREAM_VILLAIN_LINK_REFINE, word mu)."""
    site = np.arange(N * N, dtype=np.uint64)
    w = philox_site(seed, chain, sweep, site, STREAM_VILLAIN_LINK)
    r = philox_site(seed, chain, sweep, site, STREAM_VILLAIN_LINK_REFINE)
    us, cs = [], []
    for mu in range(2):
        p = w[mu] * np.uint64(2 * interval_n)
        cs.append(W * _nonzero_choice((p >> np.uint64(32)).astype(np.int64), interval_n))
        us.append(_lazy_uniform(p & MASK32, r[mu]))
    return {'u': np.stack(us).reshape(2, N, N), 'a': np.stack(cs).reshape(2, N, N)}


def villain_cohomology_draws(seed, chain, sweep, interval_h=1):
    """CohomologyUpdate on the GPU: block with counter word 0 = mu in STREAM_VILLAIN_COHOMOLOGY; word 0 -> h (index =
    (2 I) w >> 32), words 1, 2 -> u = (k52 + 1/2) 2^-52 with k52 = (y & 0xFFFFF) << 32 | z.  Returns (u[2], h[2])."""
    x, y, z, _ = philox_site(seed, chain, sweep, np.arange(2, dtype=np.uint64), STREAM_VILLAIN_COHOMOLOGY)
    h = _nonzero_choice(((x * np.uint64(2 * interval_h)) >> np.uint64(32)).astype(np.int64), interval_h)
    ku = ((y & np.uint64(0xFFFFF)) << np.uint64(32)) | z
    u = (ku.astype(np.float64) + 0.5) * TWO_M52
    return u, h
