#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The reference is imported through oracle/refimport.py (which stubs h5py/matplotlib when those
are absent).  Every fixture stores the inputs, the random draws in dense per-site form, and the
reference's outputs, so the tests that consume it need neither the reference nor numpy's PCG64
stream to be reproducible.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refimport            # noqa: E402
from oracle import villain_np           # noqa: E402
from oracle import worldline_np         # noqa: E402

sv = refimport.import_reference()
Form = sv.lattice.Form


def villain_neighborhood():
    """NeighborhoodUpdate.step chains with rng = default_rng(99) (the reference's own seed in
    test/test_vortex_sparse.py:31) from hot starts built like test/test_delta_s.py:20-27."""
    cases = []
    for (N, kappa, W, sweeps, cfg_seed) in [
        (4, 0.5, 1, 12, 0), (4, 0.1, 2, 12, 1),
        (5, 0.5, 1, 12, 2), (5, 0.1, 1, 12, 3), (5, 0.5, 3, 8, 4),
        (8, 0.5, 1, 12, 5), (8, 0.1, 2, 12, 6), (7, 0.3, 1, 8, 7),
        (6, 1.0, 1, 8, 8), (16, 0.2, 1, 6, 9),
        (32, 0.5, 1, 4, 10), (32, 0.05, 1, 4, 11),
    ]:
        L = sv.lattice.Lattice2D(N)
        S = sv.action.Villain(L, kappa, W=W)
        G = sv.generator.villain.NeighborhoodUpdate(S)
        G.rng = np.random.default_rng(99)
        replay = np.random.default_rng(99)
        phi0, n0 = villain_np.hot_start(np.random.default_rng(cfg_seed), N)
        n0 = n0 * W   # keeps dn = 0 mod W irrelevant here, but makes W>1 starts "valid-like"
        cfg = {'phi': Form(phi0, degree=0, lattice=L), 'n': Form(n0, degree=1, lattice=L)}
        rec = dict(N=N, kappa=kappa, W=W, sweeps=sweeps, phi0=phi0, n0=n0)
        us, dphis, dnf, dnb, phis, ns, acc, accp, act = [], [], [], [], [], [], [], [], []
        for s in range(sweeps):
            draws = villain_np.draw_neighborhood(replay, N, W=W)
            before = (G.accepted, G.acceptance)
            cfg = G.step(cfg)
            us.append(draws['u']); dphis.append(draws['dphi'])
            dnf.append(draws['dn_fwd']); dnb.append(draws['dn_bwd'])
            phis.append(np.asarray(cfg['phi']).copy()); ns.append(np.asarray(cfg['n']).copy())
            acc.append(int(G.accepted - before[0]))
            accp.append(float(G.acceptance - before[1]))      # mean acceptance probability of the sweep
            act.append(float(S(cfg['phi'], cfg['n'])))
        rec.update(u=np.array(us), dphi=np.array(dphis), dn_fwd=np.array(dnf), dn_bwd=np.array(dnb),
                   phi=np.array(phis), n=np.array(ns), accepted=np.array(acc),
                   acceptance=np.array(accp), action=np.array(act))
        cases.append(rec)
    out = {}
    for i, rec in enumerate(cases):
        for k, v in rec.items():
            out[f'case{i}_{k}'] = np.asarray(v)
    out['n_cases'] = np.asarray(len(cases))
    np.savez_compressed(os.path.join(HERE, 'villain_neighborhood.npz'), **out)
    print('villain_neighborhood.npz:', len(cases), 'cases; accepted per case:',
          [int(r['accepted'].sum()) for r in cases])


def villain_observables():
    """Action and the named Villain observables on seeded hot configurations."""
    out = {}
    cases = [(4, 0.7, 0), (5, 0.5, 1), (8, 0.3, 2), (32, 0.5, 3)]
    for i, (N, kappa, seed) in enumerate(cases):
        L = sv.lattice.Lattice2D(N)
        S = sv.action.Villain(L, kappa)
        phi, n = villain_np.hot_start(np.random.default_rng(seed), N)
        fphi, fn = Form(phi, degree=0, lattice=L), Form(n, degree=1, lattice=L)
        O = sv.observable
        out[f'case{i}_N'] = np.asarray(N); out[f'case{i}_kappa'] = np.asarray(kappa)
        out[f'case{i}_phi'] = phi; out[f'case{i}_n'] = n
        out[f'case{i}_action'] = np.asarray(float(S(fphi, fn)))
        out[f'case{i}_links'] = np.asarray(O.Links.Villain(S, fphi, fn))
        out[f'case{i}_ActionDensity'] = np.asarray(float(O.ActionDensity.Villain(S, fphi, fn)))
        out[f'case{i}_InternalEnergyDensity'] = np.asarray(float(O.InternalEnergyDensity.Villain(S, fphi, fn)))
        out[f'case{i}_InternalEnergyDensitySquared'] = np.asarray(float(O.InternalEnergyDensitySquared.Villain(S, fphi, fn)))
        out[f'case{i}_WindingSquared'] = np.asarray(float(O.WindingSquared.Villain(S, fn)))
        tw = np.asarray(O.TorusWrapping.Villain(S, fphi, fn))
        out[f'case{i}_TorusWrapping'] = tw
        out[f'case{i}_WrappingSquared'] = np.asarray(float(O.WrappingSquared.default(S, tw)))
        out[f'case{i}_Spin_Spin'] = np.asarray(O.Spin_Spin.Villain(S, fphi))
        out[f'case{i}_Winding_Winding'] = np.asarray(O.Winding_Winding.Villain(S, fn))
        out[f'case{i}_dn'] = np.asarray(sv.lattice.d(fn))
    out['n_cases'] = np.asarray(len(cases))
    np.savez_compressed(os.path.join(HERE, 'villain_observables.npz'), **out)
    print('villain_observables.npz:', len(cases), 'cases')


def lattice_forms():
    """d, delta, face_sum, coface_sum of random float and int forms, plus the colour maps."""
    out = {}
    i = 0
    for N in (3, 4, 5, 8):
        L = sv.lattice.Lattice2D(N)
        rng = np.random.default_rng(100 + N)
        for p in (0, 1, 2):
            for kind in ('f', 'i'):
                C = (1, 2, 1)[p]
                if kind == 'f':
                    data = rng.uniform(-3, 3, (C, N, N))
                else:
                    data = rng.integers(-5, 6, (C, N, N))
                F = Form(data, degree=p, lattice=L)
                out[f'case{i}_N'] = np.asarray(N); out[f'case{i}_p'] = np.asarray(p)
                out[f'case{i}_in'] = data
                for name, fn in (('d', sv.lattice.d), ('delta', sv.lattice.delta),
                                 ('face_sum', lambda f: f.face_sum()), ('coface_sum', lambda f: f.coface_sum())):
                    res = fn(F)
                    if isinstance(res, np.ndarray):
                        out[f'case{i}_{name}'] = np.asarray(res)
                i += 1
        cmap = np.zeros((N, N), dtype=np.int64)
        for c, color in enumerate(L.checkerboarding):
            cmap[color] = c
        out[f'colour_N{N}'] = cmap
        out[f'colour_order_N{N}'] = np.concatenate([np.stack(color, 0) for color in L.checkerboarding], axis=1)
    for N in (6, 7, 9, 32):
        L = sv.lattice.Lattice2D(N)
        cmap = np.zeros((N, N), dtype=np.int64)
        for c, color in enumerate(L.checkerboarding):
            cmap[color] = c
        out[f'colour_N{N}'] = cmap
        out[f'colour_order_N{N}'] = np.concatenate([np.stack(color, 0) for color in L.checkerboarding], axis=1)
    out['n_cases'] = np.asarray(i)
    np.savez_compressed(os.path.join(HERE, 'lattice_forms.npz'), **out)
    print('lattice_forms.npz:', i, 'cases')


def _pack(cases, name):
    out = {}
    for i, rec in enumerate(cases):
        for k, v in rec.items():
            out[f'case{i}_{k}'] = np.asarray(v)
    out['n_cases'] = np.asarray(len(cases))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(f'{name}.npz:', len(cases), 'cases')


def _wl_random_cfg(L, seed):
    """The construction of test/test_vortex_sparse.py:15-19 (m is NOT constrained there)."""
    rng = np.random.default_rng(seed)
    m = rng.integers(-3, 4, (2,) + L.dims)
    v = rng.integers(-3, 4, (1,) + L.dims)
    return m, v


def worldline_checkerboard():
    """VortexUpdate / CoexactUpdate chains: `step_reference` with rng = default_rng(99), 5 chained
    steps, exactly the pattern of test/test_vortex_sparse.py:22-39 and test/test_coexact_sparse.py:25-39
    (which also assert step == step_reference; re-checked here)."""
    cases = []
    for kind in ('vortex', 'coexact'):
        for (N, W, kappa, interval) in [(4, 1, 0.5, 1), (5, 1, 0.5, 1), (5, 3, 0.5, 1), (8, 1, 0.5, 1), (8, 3, 0.3, 2),
                                        (6, 2, 1.0, 1), (7, 2, 0.5, 1), (16, 1, 0.5, 1)]:
            L = sv.lattice.Lattice2D(N)
            S = sv.action.Worldline(L, kappa, W=W)
            Gen = sv.generator.worldline.VortexUpdate if kind == 'vortex' else sv.generator.worldline.CoexactUpdate
            dense, sparse = Gen(S, interval), Gen(S, interval)
            dense.rng = np.random.default_rng(99); sparse.rng = np.random.default_rng(99)
            replay = np.random.default_rng(99)
            if kind == 'vortex':
                m0, v0 = _wl_random_cfg(L, 10 * 2 + W)
            else:
                m0, v0 = worldline_np.hot_start(np.random.default_rng(N + W), N)
            cfg = {'m': Form(m0, degree=1, lattice=L), 'v': Form(v0, degree=2, lattice=L)}
            cfg_s = cfg
            us, aa, ms, vs, acc, accp = [], [], [], [], [], []
            for s in range(5):
                draws = worldline_np.draw_checkerboard(replay, N, kind, interval)
                before = (dense.accepted, dense.acceptance)
                cfg = cfg | dense.step_reference(cfg)
                cfg_s = cfg_s | sparse.step(cfg_s)
                assert (np.asarray(cfg['m']) == np.asarray(cfg_s['m'])).all() and (np.asarray(cfg['v']) == np.asarray(cfg_s['v'])).all()
                us.append(draws['u']); aa.append(draws['a'])
                ms.append(np.asarray(cfg['m']).copy()); vs.append(np.asarray(cfg['v']).copy())
                acc.append(int(dense.accepted - before[0])); accp.append(float(dense.acceptance - before[1]))
            cases.append(dict(kind=kind, N=N, W=W, kappa=kappa, interval=interval, sweeps=5, m0=m0, v0=v0,
                              u=np.array(us), a=np.array(aa), m=np.array(ms), v=np.array(vs),
                              accepted=np.array(acc), acceptance=np.array(accp)))
    _pack(cases, 'worldline_checkerboard')


def worldline_plaquette():
    """PlaquetteUpdate.step chains: np.random.seed(7) for the site permutation
    (test/test_plaquette_update.py:31) and rng = default_rng(99) for the proposals."""
    cases = []
    for (N, W, kappa) in [(4, 1, 0.5), (5, 1, 0.5), (5, 2, 0.4), (8, 1, 0.5), (8, 3, 0.6)]:
        L = sv.lattice.Lattice2D(N)
        S = sv.action.Worldline(L, kappa, W=W)
        G = sv.generator.worldline.PlaquetteUpdate(S)
        G.rng = np.random.default_rng(99)
        np.random.seed(7)
        m0, v0 = worldline_np.hot_start(np.random.default_rng(N * 10 + W), N)
        cfg = {'m': Form(m0, degree=1, lattice=L), 'v': Form(v0, degree=2, lattice=L)}
        ms, vs, acc, accp, act = [], [], [], [], []
        for s in range(4):
            before = (G.accepted, G.acceptance)
            cfg = G.step(cfg)
            assert S.valid(cfg)
            ms.append(np.asarray(cfg['m']).copy()); vs.append(np.asarray(cfg['v']).copy())
            acc.append(int(G.accepted - before[0])); accp.append(float(G.acceptance - before[1]))
            act.append(float(S(cfg['m'], cfg['v'])))
        cases.append(dict(N=N, W=W, kappa=kappa, sweeps=4, m0=m0, v0=v0, m=np.array(ms), v=np.array(vs),
                          accepted=np.array(acc), acceptance=np.array(accp), action=np.array(act)))
    _pack(cases, 'worldline_plaquette')


def worldline_observables():
    cases = []
    for (N, kappa, W, seed) in [(4, 0.7, 1, 0), (5, 0.5, 2, 1), (8, 0.3, 3, 2), (32, 0.5, 1, 3), (64, 0.5, 1, 4)]:
        L = sv.lattice.Lattice2D(N)
        S = sv.action.Worldline(L, kappa, W=W)
        m, v = worldline_np.hot_start(np.random.default_rng(seed), N)
        fm, fv = Form(m, degree=1, lattice=L), Form(v, degree=2, lattice=L)
        O = sv.observable
        lk = O.Links.Worldline(S, fm, fv)
        cases.append(dict(N=N, kappa=kappa, W=W, m=m, v=v, action=float(S(fm, fv)), links=np.asarray(lk),
                          ActionDensity=float(O.ActionDensity.Worldline(S, lk)),
                          InternalEnergyDensity=float(O.InternalEnergyDensity.Worldline(S, lk)),
                          InternalEnergyDensitySquared=float(O.InternalEnergyDensitySquared.Worldline(S, lk)),
                          WindingSquared=float(O.WindingSquared.Worldline(S, lk)),
                          TorusWrapping=np.asarray(O.TorusWrapping.Worldline(S, fm)),
                          Vortex_Vortex=np.asarray(O.Vortex_Vortex.Worldline(S, fv))))
    _pack(cases, 'worldline_observables')


def worldline_wrapping():
    """WrappingUpdate.step chains (worldline/wrapping.py:43-90) with rng = default_rng(99)."""
    cases = []
    for (N, W, kappa, interval) in [(4, 1, 0.5, 1), (5, 2, 0.4, 1), (8, 1, 0.3, 2), (8, 3, 0.6, 1), (16, 1, 0.25, 1)]:
        L = sv.lattice.Lattice2D(N)
        S = sv.action.Worldline(L, kappa, W=W)
        G = sv.generator.worldline.WrappingUpdate(S, interval)
        G.rng = np.random.default_rng(99)
        replay = np.random.default_rng(99)
        m0, v0 = worldline_np.hot_start(np.random.default_rng(N + W), N)
        cfg = {'m': Form(m0, degree=1, lattice=L), 'v': Form(v0, degree=2, lattice=L)}
        us, cs, ms, acc = [], [], [], []
        for s in range(6):
            d = worldline_np.draw_wrapping(replay, N, interval)
            before = G.accepted
            cfg = G.step(cfg)
            assert S.valid(cfg)
            us.append(d['u']); cs.append(d['cm']); ms.append(np.asarray(cfg['m']).copy()); acc.append(int(G.accepted - before))
        cases.append(dict(N=N, W=W, kappa=kappa, interval=interval, sweeps=6, m0=m0, v0=v0, u=np.array(us), cm=np.array(cs),
                          m=np.array(ms), accepted=np.array(acc)))
    _pack(cases, 'worldline_wrapping')


def villain_decoupled():
    """SiteUpdate / LinkUpdate / ExactUpdate chains (generator/villain/{site,link,exact}.py) with rng = default_rng(99):
    inputs, the dense draws of every sweep, and the reference's fields, accept counts and acceptance after every sweep."""
    cases = []
    gens = sv.generator.villain
    for kind, (N, kappa, W, interval, sweeps, cfg_seed) in [
        ('site', (4, 0.5, 1, np.pi, 8, 0)), ('site', (5, 0.3, 1, 1.0, 8, 1)), ('site', (8, 0.7, 2, np.pi, 6, 2)), ('site', (16, 0.5, 1, np.pi, 4, 3)),
        ('link', (4, 0.5, 1, 1, 8, 4)), ('link', (5, 0.2, 2, 2, 8, 5)), ('link', (8, 0.1, 3, 1, 6, 6)), ('link', (16, 0.5, 1, 1, 4, 7)),
        ('exact', (4, 0.5, 1, 1, 8, 8)), ('exact', (5, 0.1, 1, 2, 8, 9)), ('exact', (8, 0.05, 2, 1, 6, 10)), ('exact', (16, 0.2, 1, 1, 4, 11)),
    ]:
        L = sv.lattice.Lattice2D(N)
        S = sv.action.Villain(L, kappa, W=W)
        G = {'site': lambda: gens.SiteUpdate(S, interval_phi=interval), 'link': lambda: gens.LinkUpdate(S, interval_n=interval),
             'exact': lambda: gens.ExactUpdate(S, interval_z=interval)}[kind]()
        G.rng = np.random.default_rng(99)
        replay = np.random.default_rng(99)
        phi0, n0 = villain_np.hot_start(np.random.default_rng(cfg_seed), N)
        n0 = n0 * W
        cfg = {'phi': Form(phi0, degree=0, lattice=L), 'n': Form(n0, degree=1, lattice=L)}
        us, aa, phis, ns, acc, accp = [], [], [], [], [], []
        for s in range(sweeps):
            if kind == 'site':
                d = villain_np.draw_site(replay, N, interval)
                us.append(d['u']); aa.append(d['dphi'])
            elif kind == 'link':
                d = villain_np.draw_link(replay, N, W=W, interval_n=interval)
                us.append(d['u']); aa.append(d['a'])
            else:
                d = villain_np.draw_exact(replay, N, interval)
                us.append(d['u']); aa.append(d['a'])
            before = (G.accepted, G.acceptance)
            cfg = G.step(cfg)
            phis.append(np.asarray(cfg['phi']).copy()); ns.append(np.asarray(cfg['n']).copy())
            acc.append(int(G.accepted - before[0])); accp.append(float(G.acceptance - before[1]))
        cases.append(dict(kind=np.array({'site': 0, 'link': 1, 'exact': 2}[kind]), N=N, kappa=kappa, W=W, interval=interval,
                          sweeps=sweeps, phi0=phi0, n0=n0, u=np.array(us), a=np.array(aa), phi=np.array(phis), n=np.array(ns),
                          accepted=np.array(acc), acceptance=np.array(accp)))
    _pack(cases, 'villain_decoupled')
    print('villain_decoupled.npz:', len(cases), 'cases; accepted per case:', [int(c['accepted'].sum()) for c in cases])


def villain_cohomology():
    """CohomologyUpdate chains (generator/villain/cohomology.py) with rng = default_rng(99)."""
    cases = []
    for (N, kappa, W, interval, sweeps, cfg_seed) in [(4, 0.05, 1, 1, 12, 0), (5, 0.02, 2, 2, 12, 1), (8, 0.03, 1, 1, 12, 2),
                                                      (16, 0.01, 1, 1, 10, 3), (32, 0.004, 1, 2, 10, 4)]:
        L = sv.lattice.Lattice2D(N)
        S = sv.action.Villain(L, kappa, W=W)
        G = sv.generator.villain.CohomologyUpdate(S, interval_h=interval)
        G.rng = np.random.default_rng(99)
        replay = np.random.default_rng(99)
        phi0, n0 = villain_np.hot_start(np.random.default_rng(cfg_seed), N)
        n0 = n0 * W
        cfg = {'phi': Form(phi0, degree=0, lattice=L), 'n': Form(n0, degree=1, lattice=L)}
        hs = tuple(range(-interval, 0)) + tuple(range(1, interval + 1))
        us, hh, ns, acc, accp = [], [], [], [], []
        for s in range(sweeps):
            u, h = np.zeros(2), np.zeros(2, dtype=np.int64)
            for mu in range(2):
                h[mu] = replay.choice(hs); u[mu] = replay.uniform(0, 1)
            before = (G.accepted, G.acceptance)
            cfg = G.step(cfg)
            us.append(u); hh.append(h); ns.append(np.asarray(cfg['n']).copy())
            acc.append(int(G.accepted - before[0])); accp.append(float(G.acceptance - before[1]))
        cases.append(dict(N=N, kappa=kappa, W=W, interval=interval, sweeps=sweeps, phi0=phi0, n0=n0, u=np.array(us), h=np.array(hh),
                          n=np.array(ns), accepted=np.array(acc), acceptance=np.array(accp)))
    _pack(cases, 'villain_cohomology')
    print('accepted per case:', [int(c['accepted'].sum()) for c in cases])


def autocorrelation():
    """supervillain.analysis.autocorrelation on AR(1) series of various lengths and correlation times, and on an observable
    column of a reference chain."""
    rng = np.random.default_rng(5)
    cases = []
    for T, rho in [(64, 0.5), (257, 0.9), (1000, 0.97), (2048, 0.0), (4096, 0.995)]:
        x = np.zeros(T)
        for t in range(1, T):
            x[t] = rho * x[t - 1] + rng.normal()
        x += 3.0
        C, tau = sv.analysis.autocorrelation(x)
        C2, tau2 = sv.analysis.autocorrelation(x, mean=3.0)
        cases.append(dict(data=x, C=C, tau=tau, C_mean3=C2, tau_mean3=tau2))
    _pack(cases, 'autocorrelation')
    print('tau per case:', [int(c['tau']) for c in cases], [int(c['tau_mean3']) for c in cases])


def resampling():
    """supervillain.analysis.Blocking._block and Bootstrap._resample (analysis/blocking.py:54-66, bootstrap.py:57-67) on
    scalar columns, driven through the reference classes with a minimal ensemble stand-in (they only ask it for its
    length, weight and the column); np.random.seed fixes the reference's own np.random.randint draw of the indices."""
    rng = np.random.default_rng(8)
    cases = []
    for T, width, draws, weighted in [(100, 7, 16, False), (1000, 50, 100, False), (4096, 64, 33, False), (257, 16, 20, True)]:
        class Stub:
            Action = None
            index_stride = 1
            weight = rng.uniform(0.5, 1.5, T) if weighted else np.ones(T)
            def __len__(self):
                return T
        E = Stub()
        E.column = np.cumsum(rng.normal(size=T)) * 0.1 + rng.normal(size=T)
        B = sv.analysis.Blocking(E, width=width)
        blocked = B._block(E.column)
        np.random.seed(1000 + T)
        R = sv.analysis.Bootstrap(E, draws=draws)
        resampled = R._resample(E.column)
        cases.append(dict(data=E.column, weight=np.asarray(E.weight), width=width, drop=B.drop, blocked=blocked,
                          indices=R.indices, resampled=resampled))
    _pack(cases, 'resampling')
    print('resampling cases:', [(len(c['data']), c['width'], c['drop'], c['blocked'].shape, c['resampled'].shape) for c in cases])


def taxicab():
    """Spin_Spin.Worldline (observable/spin.py:50-224) and Vortex_Vortex.Villain (observable/vortex.py:63-189): the taxicab
    reweighting observables on random link fields, even and odd N."""
    rng = np.random.default_rng(12)
    cases = []
    for N, kappa, W in [(4, 0.5, 1), (5, 0.7, 1), (6, 0.3, 2), (8, 1.1, 1)]:
        L = sv.lattice.Lattice2D(N)
        Sw = sv.action.Worldline(L, kappa, W=W)
        m_, v_ = _wl_random_cfg(L, int(rng.integers(1 << 30)))
        cfg = {'m': L.form(1, dtype=int), 'v': L.form(2, dtype=int)}
        cfg['m'][...] = m_; cfg['v'][...] = v_
        links_w = np.asarray(sv.observable.Links.Worldline(Sw, cfg['m'], cfg['v']), dtype=np.float64)
        spin = sv.observable.Spin_Spin.Worldline(Sw, links_w)
        Sv = sv.action.Villain(L, kappa, W=W)
        phi = L.form(0); phi[...] = rng.uniform(-np.pi, np.pi, phi.shape)
        n = L.form(1, dtype=int); n[...] = rng.integers(-2, 3, n.shape)
        links_v = np.asarray(sv.observable.Links.Villain(Sv, phi, n), dtype=np.float64)
        vortex = sv.observable.Vortex_Vortex.Villain(Sv, links_v)
        cases.append(dict(N=N, kappa=kappa, W=W, m=np.asarray(cfg['m']), v=np.asarray(cfg['v']), links_w=links_w, spin_spin=spin,
                          phi=np.asarray(phi), n=np.asarray(n), links_v=links_v, vortex_vortex=vortex))
    _pack(cases, 'taxicab')
    print('taxicab:', [(c['N'], float(c['spin_spin'].min()), float(c['vortex_vortex'].max())) for c in cases])


if __name__ == '__main__':
    which = sys.argv[1:] or ['villain_neighborhood', 'villain_observables', 'lattice_forms',
                             'worldline_checkerboard', 'worldline_plaquette', 'worldline_observables', 'worldline_wrapping',
                             'villain_decoupled', 'villain_cohomology', 'autocorrelation', 'resampling', 'taxicab']
    for name in which:
        globals()[name]()
