"""CPU restatement (numpy) of the reference's worldline hot path, D=2.

TEST INFRASTRUCTURE ONLY -- checker for the CUDA path and timed CPU baseline of bench.py.

Restated, each with the reference lines it follows:
  * `plaquette_step`   -- PlaquetteUpdate.step, supervillain/generator/worldline/plaquette.py:35-104
                          (sequential, random site order from the GLOBAL numpy RNG, :63)
  * `vortex_step`      -- VortexUpdate.step_reference,  worldline/vortex.py:138-198
  * `coexact_step`     -- CoexactUpdate.step_reference, worldline/coexact.py:130-186
  * `draw_checkerboard`-- their RNG call order (vortex.py:84,98-105; coexact.py:89,96-99) replayed
                          and scattered per plaquette
  * `checkerboard_step_dense` -- the per-plaquette arithmetic of SURVEY.md App. B in red/black
                          order for the three modes of the GPU kernel (joint / vortex / coexact)
  * `links`, `action`, observables -- supervillain/action/worldline.py:44-94 and
                          supervillain/observable/{action,energy,winding,wrapping,vortex}.py

W is finite here (integer v); the reference's W=inf (real v, _W = 2 pi) is out of scope.
"""
import numpy as np

from . import lattice_np as lat


def links(m, v, W):
    """f = m - delta(v) / W   (observable/links.py:35-45)."""
    return m - lat.delta2(v) / W


def valid(m):
    """delta m == 0 everywhere (action/worldline.py:54-70)."""
    return bool((lat.delta1(m) == 0).all())


def constant_offset(N, kappa):
    """links/2 * ln(2 pi kappa) - sites * ln(2 pi)   (action/worldline.py:44)."""
    return (2 * N * N) / 2 * np.log(2 * np.pi * kappa) - (N * N) * np.log(2 * np.pi)


def action(m, v, kappa, W):
    """action/worldline.py:72-94.  Raises ValueError when delta m != 0, as the reference does."""
    if not valid(m):
        raise ValueError('The one-form m does not satisfy the constraint δm = 0 everywhere.')
    N = m.shape[-1]
    return 0.5 / kappa * np.sum(links(m, v, W) ** 2, axis=(-3, -2, -1)) + constant_offset(N, kappa)


def action_density(m, v, kappa, W):
    """observable/action.py:35-47."""
    N = m.shape[-1]
    f = links(m, v, W)
    return ((2 * N * N) / 2 - 0.5 / kappa * (f ** 2).sum(axis=(-3, -2, -1))) / (N * N)


def internal_energy_density(m, v, kappa, W):
    """observable/energy.py:68-81."""
    N = m.shape[-1]
    f = links(m, v, W)
    return ((2 * N * N) / 2 - 0.5 / kappa * (f ** 2).sum(axis=(-3, -2, -1))) / (N * N * kappa)


def internal_energy_density_squared(m, v, kappa, W):
    """observable/energy.py:85-102."""
    N = m.shape[-1]
    nl, ns = 2 * N * N, N * N
    f2 = (links(m, v, W) ** 2).sum(axis=(-3, -2, -1))
    p1 = (nl / 2 - 0.5 / kappa * f2) / kappa
    p2 = (f2 / kappa - nl / 2) / kappa ** 2
    return (p1 ** 2 - p2) / ns ** 2


def winding_squared(m, v, kappa, W):
    """observable/winding.py:40-52."""
    f = links(m, v, W)
    return 1 / (np.pi ** 2 * kappa) - np.mean(lat.d1(f) ** 2, axis=(-3, -2, -1)) / (2 * np.pi * kappa) ** 2


def torus_wrapping(m):
    """observable/wrapping.py:28-39."""
    return m.sum(axis=(-2, -1)) / m.shape[-1]


def vortex_vortex(v, W):
    """observable/vortex.py:22-37."""
    vortex = np.exp(2j * np.pi * v / W)
    return lat.correlation(vortex, vortex).mean(axis=-3)


# ---------------------------------------------------------------------------------------------
# generators
# ---------------------------------------------------------------------------------------------

def plaquette_step(m, v, kappa, W, rng, permutation=None, stats=None):
    """PlaquetteUpdate.step (plaquette.py:35-104) on one chain.

    `permutation` supplies the site order; None means the reference's own
    `np.random.permutation(L.coordinates)` from the global numpy RNG (:63).
    """
    N = m.shape[-1]
    m = m.copy()
    v = v.copy()
    f = m - lat.delta2(v) / W                                           # :53
    P = N * N
    dm_all = rng.choice([-1, +1], P)                                    # :58
    dv_all = rng.choice([-1, 0, +1], P)                                 # :59
    u_all = rng.uniform(0, 1, P)                                        # :60
    if permutation is None:
        c = lat.fft_coordinates(N)
        coords = np.stack([a.flatten() for a in np.meshgrid(c, c, indexing='ij')], axis=1)
        permutation = np.random.permutation(coords)                     # :63 (global RNG, FFT coordinates)
    accepted = 0
    acceptance = 0.0
    for idx, here in enumerate(permutation):
        x0, x1 = int(here[0]) % N, int(here[1]) % N                     # negative FFT coordinates index from the end
        p0, p1 = (x0 + 1) % N, (x1 + 1) % N
        f1, f2, f3, f4 = f[0, x0, x1], f[1, p0, x1], f[0, x0, p1], f[1, x0, x1]   # :79-82
        dm, dv = dm_all[idx], dv_all[idx]
        df = dm - dv / W                                                # :84
        dS = df / kappa * (f1 + f2 - f3 - f4 + 2 * df)                  # :85
        A = np.clip(np.exp(-dS), 0, 1)                                  # :87
        acceptance += A
        if u_all[idx] < A:                                              # :90
            m[0, x0, x1] += dm; m[1, p0, x1] += dm; m[0, x0, p1] -= dm; m[1, x0, x1] -= dm
            v[0, x0, x1] += dv
            f[0, x0, x1] += df; f[1, p0, x1] += df; f[0, x0, p1] -= df; f[1, x0, x1] -= df
            accepted += 1
    if stats is not None:
        stats['accepted'] = accepted
        stats['acceptance'] = acceptance
    return m, v


def _choices(interval):
    return tuple(range(-interval, 0)) + tuple(range(1, interval + 1))


def vortex_step(m, v, kappa, W, rng, interval_v=1, stats=None):
    """VortexUpdate.step_reference (vortex.py:138-198): v-only checkerboard sweep, dense."""
    N = m.shape[-1]
    v = v.copy()
    vs = _choices(interval_v)
    u = rng.uniform(0, 1, (1, N, N))                                    # :163
    accepted, acceptance = 0, 0.0
    for (x0, x1) in lat.colour_sites(N):
        dv = lat.delta2(v)                                              # :169
        change = np.zeros((1, N, N), dtype=np.int64)
        change[0, x0, x1] = rng.choice(vs, len(x0))                     # :172-173
        cdv_W = lat.delta2(change) / W                                  # :178-181
        dS_link = (0.5 / kappa) * (-cdv_W) * (2 * (m - dv / W) - cdv_W)  # :182-185
        dS = lat.coface_sum1(dS_link)                                   # :187
        A = np.clip(np.exp(-dS[0, x0, x1]), 0, 1)                       # :189
        acc = u[0, x0, x1] < A
        accepted += int(acc.sum()); acceptance += float(A.sum())
        v[0, x0, x1] += change[0, x0, x1] * acc                         # :195
    if stats is not None:
        stats['accepted'] = accepted; stats['acceptance'] = acceptance
    return m.copy(), v


def coexact_step(m, v, kappa, W, rng, interval_t=1, stats=None):
    """CoexactUpdate.step_reference (coexact.py:130-186): m += delta t checkerboard sweep, dense."""
    N = m.shape[-1]
    m = m.copy()
    ts = _choices(interval_t)
    dvw = lat.delta2(v) / W                                             # :146
    u = rng.uniform(0, 1, (1, N, N))                                    # :153
    accepted, acceptance = 0, 0.0
    for (x0, x1) in lat.colour_sites(N):
        t = np.zeros((1, N, N), dtype=np.int64)
        t[0, x0, x1] = rng.choice(ts, len(x0))                          # :158-159
        cm = lat.delta2(t)                                              # :161
        dS_link = (0.5 / kappa) * cm * (2 * (m - dvw) + cm)             # :165-168
        dS = lat.coface_sum1(dS_link)                                   # :170
        A = np.clip(np.exp(-dS[0, x0, x1]), 0, 1)
        acc = u[0, x0, x1] < A
        accepted += int(acc.sum()); acceptance += float(A.sum())
        t[0, x0, x1] *= acc                                             # :178
        m = m + lat.delta2(t)                                           # :179
    if stats is not None:
        stats['accepted'] = accepted; stats['acceptance'] = acceptance
    return m, v.copy()


def draw_checkerboard(rng, N, mode, interval=1):
    """Replay the checkerboard generators' RNG calls for one sweep, scattered per plaquette.

    'vortex' / 'coexact': u (N,N) first, then one `choice` vector per colour (vortex.py:84,98-105;
    coexact.py:89,96-99).  'joint' has no reference counterpart in checkerboard order; its draw
    order is defined here as u (N,N), then per colour dm = choice([-1,+1]) and dv = choice([-1,0,+1]).
    """
    u = rng.uniform(0, 1, (1, N, N))[0]
    a = np.zeros((N, N), dtype=np.int64)
    b = np.zeros((N, N), dtype=np.int64)
    for (x0, x1) in lat.colour_sites(N):
        if mode == 'joint':
            a[x0, x1] = rng.choice([-1, +1], len(x0))
            b[x0, x1] = rng.choice([-1, 0, +1], len(x0))
        else:
            a[x0, x1] = rng.choice(_choices(interval), len(x0))
    return {'u': u, 'a': a, 'b': b}


def checkerboard_step_dense(m, v, kappa, W, draws, mode, stats=None, accept_mask=None, dS_out=None):
    """One red/black sweep, plaquette by plaquette, f recomputed from the current integers.

    Signs of delta(unit 2-form at x): (0,x) +, (0,x+e1) -, (1,x) -, (1,x+e0) +.
    dS order: T(1,x) + T(1,x+e0) + T(0,x) + T(0,x+e1) (coface_sum,1 rows) for vortex / coexact;
    delta_f / kappa * (f1 + f2 - f3 - f4 + 2 delta_f) for joint (plaquette.py:84-85).
    """
    N = m.shape[-1]
    m = m.copy()
    v = v.copy()
    u, a_all, b_all = draws['u'], draws['a'], draws['b']
    accepted, acceptance = 0, 0.0
    Wd = float(W)
    hk = 0.5 / kappa
    vv = v[0]
    for (xs0, xs1) in lat.colour_sites(N):
        for x0, x1 in zip(xs0.tolist(), xs1.tolist()):
            p0, m0_, p1, m1_ = (x0 + 1) % N, (x0 - 1) % N, (x1 + 1) % N, (x1 - 1) % N
            vc = int(vv[x0, x1])
            f_0x = float(m[0, x0, x1]) - float(vc - int(vv[x0, m1_])) / Wd
            f_0p = float(m[0, x0, p1]) - float(int(vv[x0, p1]) - vc) / Wd
            f_1x = float(m[1, x0, x1]) - float(int(vv[m0_, x1]) - vc) / Wd
            f_1p = float(m[1, p0, x1]) - float(vc - int(vv[p0, x1])) / Wd
            a = int(a_all[x0, x1])
            b = int(b_all[x0, x1])
            if mode == 'joint':
                df = float(a) - float(b) / Wd
                dS = df / kappa * (f_0x + f_1p - f_0p - f_1x + 2 * df)
            elif mode == 'vortex':
                cp, cn = float(a) / Wd, float(-a) / Wd
                dS = (hk * (-cn)) * (2 * f_1x - cn)
                dS = dS + (hk * (-cp)) * (2 * f_1p - cp)
                dS = dS + (hk * (-cp)) * (2 * f_0x - cp)
                dS = dS + (hk * (-cn)) * (2 * f_0p - cn)
            else:
                cp, cn = float(a), float(-a)
                dS = (hk * cn) * (2 * f_1x + cn)
                dS = dS + (hk * cp) * (2 * f_1p + cp)
                dS = dS + (hk * cp) * (2 * f_0x + cp)
                dS = dS + (hk * cn) * (2 * f_0p + cn)
            A = min(max(float(np.exp(-dS)), 0.0), 1.0)
            ok = u[x0, x1] < A
            acceptance += A
            if dS_out is not None:
                dS_out[x0, x1] = dS
            if accept_mask is not None:
                accept_mask[x0, x1] = ok
            if ok:
                accepted += 1
                if mode in ('joint', 'coexact'):
                    m[0, x0, x1] += a; m[1, p0, x1] += a; m[0, x0, p1] -= a; m[1, x0, x1] -= a
                if mode == 'joint':
                    vv[x0, x1] += b
                elif mode == 'vortex':
                    vv[x0, x1] += a
    if stats is not None:
        stats['accepted'] = accepted; stats['acceptance'] = acceptance
    return m, v


def hot_start(rng, N, chains=None):
    """m = delta t with t ~ integers(-2, 3), v ~ integers(-2, 3): test/test_delta_s.py:30-40 (so delta m = 0)."""
    lead = () if chains is None else (chains,)
    t = rng.integers(-2, 3, lead + (1, N, N))
    m = lat.delta2(t)
    v = rng.integers(-2, 3, lead + (1, N, N))
    return m, v


def wrapping_step(m, v, kappa, W, rng, interval_w=1, stats=None):
    """WrappingUpdate.step (supervillain/generator/worldline/wrapping.py:43-90): one proposal per torus cycle.

    A mu-direction cycle changes m_mu by the same amount on all N links along direction mu at fixed
    perpendicular coordinate; dS of a cycle is the sum of the link changes along it (:71).
    """
    N = m.shape[-1]
    ws = _choices(interval_w)
    change = np.zeros((2, N, N), dtype=np.int64)
    change[0] = rng.choice(ws, N).reshape(1, N)                          # :59-60, mu = 0: one value per x1
    change[1] = rng.choice(ws, N).reshape(N, 1)                          #          mu = 1: one value per x0
    dS_link = 0.5 / kappa * change * (2 * (m - lat.delta2(v) / W) + change)    # :64
    accepted, acceptance = 0, 0.0
    for mu in range(2):
        dS = dS_link[mu].sum(axis=mu)                                    # :71
        A = np.clip(np.exp(-dS), 0, 1)
        u = rng.uniform(0, 1, A.shape)                                   # :74
        acc = u < A
        change[mu] *= np.expand_dims(acc, axis=mu)                       # :77
        acceptance += float(A.sum()); accepted += int(acc.sum())
    if stats is not None:
        stats['accepted'] = accepted; stats['acceptance'] = acceptance
    return m + change, v.copy()


def draw_wrapping(rng, N, interval_w=1):
    """The reference's RNG call order for one WrappingUpdate step: choices for mu=0, mu=1, then uniforms for mu=0, mu=1."""
    ws = _choices(interval_w)
    cm = np.stack([rng.choice(ws, N), rng.choice(ws, N)])
    u = np.stack([rng.uniform(0, 1, N), rng.uniform(0, 1, N)])
    return {'cm': cm, 'u': u}


def wrapping_step_dense(m, v, kappa, W, draws, stats=None, dS_out=None):
    """The same step cycle by cycle with dense draws cm, u of shape (2, N): index [mu, k] is the mu-direction cycle
    at perpendicular coordinate k.  Terms are accumulated sequentially along the cycle."""
    N = m.shape[-1]
    m = m.copy()
    f = m - lat.delta2(v) / W
    hk = 0.5 / kappa
    accepted, acceptance = 0, 0.0
    for mu in range(2):
        for k in range(N):
            c = int(draws['cm'][mu, k])
            dS = 0.0
            for j in range(N):
                fl = f[0, j, k] if mu == 0 else f[1, k, j]
                dS = dS + (hk * c) * (2 * fl + c)
            A = min(max(float(np.exp(-dS)), 0.0), 1.0)
            acceptance += A
            if dS_out is not None:
                dS_out[mu, k] = dS
            if draws['u'][mu, k] < A:
                accepted += 1
                if mu == 0:
                    m[0, :, k] += c
                else:
                    m[1, k, :] += c
    if stats is not None:
        stats['accepted'] = accepted; stats['acceptance'] = acceptance
    return m, v.copy()
