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


def villain_draws(seed, chain, sweep, N, W=1, interval_phi=np.pi, interval_n=1):
    """Dense per-site proposals of one chain and sweep, in the layout of villain_np.draw_neighborhood.

    128 Philox bits (x, y, z, w) per site split 44 / 52 / 32:
      dphi = -I + (2I) * ((k44 + 1/2) 2^-44),  k44 = x << 12 | y >> 20
      u    = (k52 + 1/2) 2^-52,                k52 = (y & 0xFFFFF) << 32 | z
      dn   = W * (digit_i - interval_n), digit_i = the four leading base-K digits (K = 2 interval_n + 1)
             of the fraction w / 2^32, ordered (fwd 0, bwd 0, fwd 1, bwd 1)
    """
    site = np.arange(N * N, dtype=np.uint64)
    x, y, z, w = philox_site(seed, chain, sweep, site, STREAM_VILLAIN_NEIGHBORHOOD)
    kphi = (x << np.uint64(12)) | (y >> np.uint64(20))
    ku = ((y & np.uint64(0xFFFFF)) << np.uint64(32)) | z
    Uphi = (kphi.astype(np.float64) + 0.5) * TWO_M44
    u = (ku.astype(np.float64) + 0.5) * TWO_M52
    dphi = -interval_phi + (2.0 * interval_phi) * Uphi
    K = np.uint64(2 * interval_n + 1)
    f = w
    digits = []
    for _ in range(4):
        prod = f * K
        digits.append((prod >> np.uint64(32)).astype(np.int64) - interval_n)
        f = prod & MASK32
    dn_fwd = np.stack([W * digits[0], W * digits[2]]).reshape(2, N, N)
    dn_bwd = np.stack([W * digits[1], W * digits[3]]).reshape(2, N, N)
    return {'u': u.reshape(N, N), 'dphi': dphi.reshape(N, N), 'dn_fwd': dn_fwd, 'dn_bwd': dn_bwd}


def worldline_draws(seed, chain, sweep, N, mode, interval=1):
    """Dense per-plaquette proposals: u, a (dm | dv | t) and, for mode 'joint', b (dv).

    u = (k44 + 1/2) 2^-44 with k44 = x << 12 | y >> 20; dm sign = bit 19 of y; choices = (w * K) >> 32."""
    site = np.arange(N * N, dtype=np.uint64)
    x, y, z, w = philox_site(seed, chain, sweep, site, STREAM_WORLDLINE_PLAQUETTE)
    ku = (x << np.uint64(12)) | (y >> np.uint64(20))
    u = (ku.astype(np.float64) + 0.5) * TWO_M44
    if mode == 'joint':
        a = np.where(((y >> np.uint64(19)) & np.uint64(1)) == 1, 1, -1).astype(np.int64)
        b = ((w * np.uint64(3)) >> np.uint64(32)).astype(np.int64) - 1
    else:
        idx = ((w * np.uint64(2 * interval)) >> np.uint64(32)).astype(np.int64)
        a = np.where(idx < interval, idx - interval, idx - interval + 1)
        b = np.zeros_like(a)
    return {'u': u.reshape(N, N), 'a': a.reshape(N, N), 'b': b.reshape(N, N)}
