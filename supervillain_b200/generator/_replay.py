"""Replay of the reference generators' numpy RNG call sequences into dense per-site arrays.

When a user assigns `generator.rng = np.random.default_rng(seed)` -- the reference's own seam,
test/test_vortex_sparse.py:31-32 -- the GPU generators draw from that numpy Generator with exactly
the calls, in exactly the order, the reference makes, scatter the 1-D draw vectors to the sites
that consume them, and hand the dense arrays to the kernels (SVB_RNG_INJECTED).  The chain
produced is then the reference's chain.  This is host bookkeeping of random numbers, not a
compute path: dS, accept/reject and the field updates all happen in the kernels.
"""
import numpy as np


def villain_neighborhood(rng, lattice, W, interval_phi, interval_n):
    """One sweep of NeighborhoodUpdate draws (neighborhood.py:87, 98, 104-107)."""
    N = lattice.N
    n_changes = np.arange(-interval_n, 1 + interval_n)
    u = rng.uniform(0, 1, (N,) * 2)
    dphi = np.zeros((N, N))
    dn_fwd = np.zeros((2, N, N), dtype=np.int32)
    dn_bwd = np.zeros((2, N, N), dtype=np.int32)
    for color in lattice.checkerboarding:
        count = len(color[0])
        dphi[color] = rng.uniform(-interval_phi, +interval_phi, count)
        for mu in range(2):
            dn_fwd[mu][color] = W * rng.choice(n_changes, count)
            dn_bwd[mu][color] = W * rng.choice(n_changes, count)
    return u, dphi, dn_fwd, dn_bwd


def _nonzero_choices(interval):
    return tuple(range(-interval, 0)) + tuple(range(1, interval + 1))


def worldline_checkerboard(rng, lattice, mode, interval=1):
    """One sweep of VortexUpdate (vortex.py:84, 98-105) or CoexactUpdate (coexact.py:89, 96-99) draws;
    for 'joint' (no reference counterpart in checkerboard order): u, then per colour dm, dv."""
    N = lattice.N
    u = rng.uniform(0, 1, (1, N, N))[0]
    a = np.zeros((N, N), dtype=np.int32)
    b = np.zeros((N, N), dtype=np.int32)
    for color in lattice.checkerboarding:
        count = len(color[0])
        if mode == 'joint':
            a[color] = rng.choice([-1, +1], count)
            b[color] = rng.choice([-1, 0, +1], count)
        else:
            a[color] = rng.choice(_nonzero_choices(interval), count)
    return u, a, b


def worldline_wrapping(rng, lattice, interval=1):
    """One WrappingUpdate step's draws (wrapping.py:59-60, :74): choices for mu = 0, 1, then uniforms for mu = 0, 1."""
    N = lattice.N
    ws = _nonzero_choices(interval)
    c = np.stack([rng.choice(ws, N), rng.choice(ws, N)]).astype(np.int32)
    u = np.stack([rng.uniform(0, 1, N), rng.uniform(0, 1, N)])
    return u, c


def villain_site(rng, lattice, interval_phi):
    """One SiteUpdate sweep's draws (site.py:66, :89): the uniforms, then per colour the dphi."""
    N = lattice.N
    u = rng.uniform(0, 1, (N,) * 2)
    dphi = np.zeros((N, N))
    for color in lattice.checkerboarding:
        dphi[color] = rng.uniform(-interval_phi, +interval_phi, len(color[0]))
    return u, dphi


def villain_link(rng, lattice, W, interval_n):
    """One LinkUpdate sweep's draws (link.py:77, :92): the changes of the whole 1-form, then the uniforms."""
    N = lattice.N
    c = (W * rng.choice(_nonzero_choices(interval_n), size=(2, N, N))).astype(np.int32)
    u = rng.uniform(0, 1, size=(2, N, N))
    return u, c


def villain_exact(rng, lattice, interval_z):
    """One ExactUpdate sweep's draws (exact.py:75, :94): the uniforms, then per colour z."""
    N = lattice.N
    u = rng.uniform(0, 1, (N,) * 2)
    z = np.zeros((N, N), dtype=np.int32)
    for color in lattice.checkerboarding:
        z[color] = rng.choice(_nonzero_choices(interval_z), len(color[0]))
    return u, z


def villain_cohomology(rng, interval_h):
    """One CohomologyUpdate step's draws (cohomology.py:90, :101): per direction h, then the uniform."""
    hs = _nonzero_choices(interval_h)
    h = np.zeros(2, dtype=np.int32); u = np.zeros(2)
    for mu in range(2):
        h[mu] = rng.choice(hs)
        u[mu] = rng.uniform(0, 1)
    return u, h
