"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the RNG replay matches the oracle's, Batch/Configurations keep the reference's contract."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    from supervillain_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, 'include', 'svb200.h')).read()
    declared = set(re.findall(r'\b(svb_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 10
    lib = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(lib, name), f'{name} is declared in svb200.h but missing from libsvb200.so'
    from supervillain_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert _lib.load().svb_version() == 1


def test_constants_mirror_header():
    from supervillain_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'svb200.h')).read()
    defs = dict(re.findall(r'#define\s+SVB_([A-Z0-9_]+)\s+(-?\d+)', header))
    for py, c in (('VOBS_COUNT', 'VOBS_COUNT'), ('WOBS_COUNT', 'WOBS_COUNT'), ('RNG_INJECTED', 'RNG_INJECTED'),
                  ('PATH_GLOBAL', 'PATH_GLOBAL'), ('WL_COEXACT', 'WL_COEXACT'), ('OP_COFACE_SUM', 'OP_COFACE_SUM'),
                  ('I64', 'I64'), ('E_ALIGN', 'E_ALIGN'), ('WOBS_DELTA_M_ABS', 'WOBS_DELTA_M_ABS'),
                  ('VOBS_ACCEPTANCE', 'VOBS_ACCEPTANCE'), ('ARITH_FAST', 'ARITH_FAST')):
        assert getattr(_lib, py) == int(defs[c]), py


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when it cannot reach a GPU."""
    import torch
    from supervillain_b200 import ops
    if torch.cuda.is_available():
        pytest.skip('a device is present')
    phi = torch.zeros((1, 1, 8, 8), dtype=torch.float64)
    n = torch.zeros((1, 2, 8, 8), dtype=torch.int32)
    with pytest.raises(ValueError):
        ops.villain_sweep(phi, n, 0.5)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'supervillain_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f
                assert '/root/reference' not in text, f


def test_replay_matches_oracle_draw_order():
    from oracle import villain_np as V
    from oracle import worldline_np as WL
    import supervillain_b200 as svb
    from supervillain_b200.generator import _replay
    for N in (4, 5, 8):
        L = svb.Lattice2D(N)
        a, b = np.random.default_rng(99), np.random.default_rng(99)
        for _ in range(3):
            u, dphi, dn_fwd, dn_bwd = _replay.villain_neighborhood(a, L, 2, np.pi, 1)
            ref = V.draw_neighborhood(b, N, W=2)
            assert (u == ref['u']).all() and (dphi == ref['dphi']).all()
            assert (dn_fwd == ref['dn_fwd']).all() and (dn_bwd == ref['dn_bwd']).all()
        for mode in ('vortex', 'coexact', 'joint'):
            a, b = np.random.default_rng(7), np.random.default_rng(7)
            u, x, y = _replay.worldline_checkerboard(a, L, mode, 2)
            ref = WL.draw_checkerboard(b, N, mode, 2)
            assert (u == ref['u']).all() and (x == ref['a']).all() and (y == ref['b']).all()


def test_lattice_colouring_matches_reference(golden_lattice_forms):
    import supervillain_b200 as svb
    _, extras = golden_lattice_forms
    for N in (3, 4, 5, 6, 7, 8, 9, 32):
        L = svb.Lattice2D(N)
        assert (L.colour_map == extras[f'colour_N{N}']).all()
        order = np.concatenate([np.stack(c, 0) for c in L.checkerboarding], axis=1)
        assert (order == extras[f'colour_order_N{N}']).all()
    with pytest.raises(NotImplementedError):
        svb.Lattice(3, 4)


def test_batch_contract():
    """test/test_batch.py, test/test_batch_dtype.py: draw-major columns, Form wrapping, lossless casts only."""
    import supervillain_b200 as svb
    L = svb.Lattice2D(4)
    S = svb.Villain(L, 0.5)
    cfgs = S.configurations(3)
    assert cfgs.phi.shape == (3, 1, 4, 4) and cfgs.n.shape == (3, 2, 4, 4)
    assert cfgs.phi.dtype == np.float64 and cfgs.n.dtype == np.int64
    one = cfgs[1]
    assert isinstance(one['phi'], svb.Form) and one['phi'].degree == 0 and one['n'].degree == 1
    cfgs[2] = {'phi': np.ones((1, 4, 4)), 'n': np.full((2, 4, 4), 3, dtype=np.int32)}     # widening is lossless
    assert (cfgs.n[2] == 3).all()
    cfgs[2] = {'n': np.full((2, 4, 4), 2.0)}                                              # integer-valued float is fine
    with pytest.raises(TypeError):
        cfgs[2] = {'n': np.full((2, 4, 4), 2.5)}
    sub = cfgs[1:]
    assert len(sub) == 2 and isinstance(sub.phi, svb.Batch)
    W = svb.Worldline(L, 0.5, W=2).configurations(2)
    assert W.m.shape == (2, 2, 4, 4) and W.v.shape == (2, 1, 4, 4) and W.v.dtype == np.int64


def test_generators_reject_wrong_action_and_report_format():
    import supervillain_b200 as svb
    from supervillain_b200.generator.villain import NeighborhoodUpdate
    from supervillain_b200.generator.worldline import PlaquetteUpdate, VortexUpdate, CoexactUpdate
    L = svb.Lattice2D(4)
    V, Wl = svb.Villain(L, 0.5), svb.Worldline(L, 0.5)
    with pytest.raises(ValueError, match='requires the Villain action'):
        NeighborhoodUpdate(Wl)
    for G in (PlaquetteUpdate, VortexUpdate, CoexactUpdate):
        with pytest.raises(ValueError):
            G(V)
    with pytest.raises(NotImplementedError):
        PlaquetteUpdate(svb.Worldline(L, 0.5, W=float('inf')))
    G = NeighborhoodUpdate(V)
    assert str(G) == 'NeighborhoodUpdate' and G.rng is None and G.accepted == 0 and G.sweeps == 0
    G.accepted, G.proposed, G.acceptance, G.sweeps = 3, 160, 0.2, 10
    assert G.report() == ('There were 3 neighborhood proposals accepted of 160 proposed updates.\n'
                          '    0.018750 acceptance rate\n    0.020000 average Metropolis acceptance probability.')


def test_gpu_generators_accept_the_reference_action_objects():
    """the generators are duck-typed on the action (needs the reference tree: build container only)."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip('reference tree not mounted')
    sv = refimport.import_reference()
    from supervillain_b200.generator.villain import NeighborhoodUpdate
    from supervillain_b200.generator.worldline import PlaquetteUpdate
    L = sv.lattice.Lattice2D(8)
    G = NeighborhoodUpdate(sv.action.Villain(L, 0.5, W=2))
    assert G.kappa == 0.5 and G.Lattice.N == 8 and G.Action.W == 2
    P = PlaquetteUpdate(sv.action.Worldline(L, 0.5))
    assert str(P) == 'PlaquetteUpdate'
    with pytest.raises(ValueError):
        NeighborhoodUpdate(sv.action.Worldline(L, 0.5))
    assert len(G.Lattice.checkerboarding) == 2        # _replay uses the reference lattice's own colour lists


def test_new_generators_keep_the_reference_constructor_contract():
    """SiteUpdate / LinkUpdate / ExactUpdate / CohomologyUpdate: same constructor signatures, attributes and wrong-action
    errors as the reference classes (site.py:23-38, link.py:32-48, exact.py:29-46, cohomology.py:44-60); no GPU needed."""
    import supervillain_b200 as svb
    from supervillain_b200.generator.villain import CohomologyUpdate, ExactUpdate, LinkUpdate, SiteUpdate
    S = svb.Villain(svb.Lattice2D(6), 0.4, W=2)
    Wl = svb.Worldline(svb.Lattice2D(6), 0.4)
    for cls, kw, attr, value in ((SiteUpdate, {'interval_phi': 1.5}, 'interval_phi', 1.5), (LinkUpdate, {'interval_n': 2}, 'n_changes', (-2, -1, 1, 2)),
                                 (ExactUpdate, {'interval_z': 2}, 'zs', (-2, -1, 1, 2)), (CohomologyUpdate, {'interval_h': 1}, 'h', (-1, 1))):
        G = cls(S, **kw)
        assert getattr(G, attr) == value and str(G) == cls.__name__
        assert (G.accepted, G.proposed, G.sweeps, G.acceptance) == (0, 0, 0, 0.) and G.rng is None
        assert G.inline_observables(5) == {}
        with pytest.raises(ValueError):
            cls(Wl)


def test_no_cpu_fallback_for_the_overlapped_and_decoupled_paths():
    import torch
    from supervillain_b200 import ops
    if torch.cuda.is_available():
        pytest.skip('a device is present')
    phi = torch.zeros((2, 1, 32, 32), dtype=torch.float64)
    n = torch.zeros((2, 2, 32, 32), dtype=torch.int32)
    with pytest.raises(ValueError):
        ops.VillainOverlappedSweeps(phi, n, 0.5)
    with pytest.raises(ValueError):
        ops.villain_decoupled('link', phi, n, 0.5)
    with pytest.raises(ValueError):
        ops.villain_cohomology(phi, n, 0.5)
    with pytest.raises(ValueError):
        ops.WorldlineOverlappedSweeps(n, n[:, :1].contiguous(), 0.5)


def test_batched_ensemble_hands_a_chain_to_the_reference_package(tmp_path):
    """`BatchedEnsemble.to_reference`: one chain becomes a reference `Ensemble` (the object `to_h5` writes);
    its field columns are the kept draws, and the inline columns are exactly what the reference measures from those fields
    (so the short-circuit of observable/observable.py:49-54 returns the same numbers).  Build container only."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip('reference tree not mounted')
    sv = refimport.import_reference()
    import supervillain_b200 as svb
    from supervillain_b200._lib import VOBS_ACTION, VOBS_COUNT, VOBS_SUM_DN2, VOBS_WRAP0, VOBS_WRAP1
    from supervillain_b200.generator.villain import NeighborhoodUpdate, villain_inline_values
    from oracle import villain_np as V
    N, chains, steps, keep, kappa = 6, 3, 6, 2, 0.45
    rng = np.random.default_rng(5)
    S = svb.Villain(svb.Lattice2D(N), kappa)
    E = svb.BatchedEnsemble(S, chains, device='cpu')
    phi = rng.uniform(-np.pi, np.pi, (chains, steps, 1, N, N))
    n = rng.integers(-2, 3, (chains, steps, 2, N, N))
    rec = np.zeros((chains, steps, VOBS_COUNT))
    for c in range(chains):
        for t in range(steps):
            rec[c, t, VOBS_ACTION] = V.action(phi[c, t], n[c, t], kappa)
            rec[c, t, VOBS_SUM_DN2] = V.winding_squared(n[c, t]) * N * N
            rec[c, t, VOBS_WRAP0], rec[c, t, VOBS_WRAP1] = V.torus_wrapping(n[c, t])
    kept = keep * (1 + np.arange(steps // keep)) - 1
    E.record, E.steps, E.sweeps_per_step, E.keep_every = rec, steps, 5, keep
    E.index = 5 * (1 + np.arange(steps))
    E.generator = NeighborhoodUpdate(S)
    E.configuration = {'phi': phi[:, kept], 'n': n[:, kept].astype(np.int64)}
    E.observables = villain_inline_values(rec, N, kappa)
    with pytest.raises(IndexError):
        E.to_reference(chains, supervillain=sv)
    R = E.to_reference(1, supervillain=sv)
    assert type(R).__module__.startswith('supervillain') and len(R.configuration) == len(kept)
    assert R.Action.kappa == kappa and R.Action.Lattice.nx == N and R.index_stride == 10
    assert list(sv.batch.Batch.as_array(R.index)) == [10, 20, 30]
    got_phi, got_n = (sv.batch.Batch.as_array(R.configuration.fields[k]) for k in ('phi', 'n'))
    assert got_phi.dtype == np.float64 and (got_phi == phi[1, kept]).all()
    assert got_n.dtype == np.int64 and (got_n == n[1, kept]).all()
    cfg0 = R.configuration[0]
    assert type(cfg0['phi']).__name__ == 'Form' and cfg0['n'].degree == 1
    # the same fields without inline columns: the reference measures for itself
    bare = sv.Ensemble(R.Action).from_configurations(sv.configurations.Configurations(
        {k: R.configuration.fields[k] for k in ('phi', 'n')}))
    for name in ('ActionDensity', 'InternalEnergyDensity', 'WindingSquared', 'TorusWrapping', 'WrappingSquared'):
        inline, measured = np.asarray(getattr(R, name)), np.asarray(getattr(bare, name))
        assert inline.shape == measured.shape, name
        assert np.allclose(inline, measured, rtol=1e-12, atol=1e-12), name
    # a kappa scan hands each chain its own coupling
    import torch
    E.kappa_chain = torch.tensor([0.3, 0.6, 0.9], dtype=torch.float64)
    assert E.to_reference(2, supervillain=sv).Action.kappa == 0.9


def test_batched_worldline_ensemble_hands_a_chain_to_the_reference_package():
    """The worldline side of `to_reference`: integer m, v columns (int64 Forms, delta m = 0 so the reference's action accepts
    them, worldline.py:92-93) and the inline columns equal to the reference's own worldline measurements."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip('reference tree not mounted')
    sv = refimport.import_reference()
    import supervillain_b200 as svb
    from supervillain_b200._lib import WOBS_COUNT, WOBS_SUM_DF2, WOBS_SUM_F2, WOBS_WRAP0, WOBS_WRAP1
    from supervillain_b200.generator.worldline import PlaquetteUpdate, worldline_inline_values
    from oracle import worldline_np as WL
    from oracle import lattice_np as LT
    N, chains, steps, kappa, W = 6, 2, 4, 0.7, 1
    rng = np.random.default_rng(11)
    S = svb.Worldline(svb.Lattice2D(N), kappa)
    E = svb.BatchedEnsemble(S, chains, device='cpu')
    m = np.zeros((chains, steps, 2, N, N), dtype=np.int64)
    v = rng.integers(-2, 3, (chains, steps, 1, N, N))
    rec = np.zeros((chains, steps, WOBS_COUNT))
    for c in range(chains):
        for t in range(steps):
            m[c, t], _ = WL.hot_start(rng, N)
            m[c, t, 0, :, 0] += 1                                   # one unit of wrapping in direction 0 keeps delta m = 0
            assert WL.valid(m[c, t])
            f = WL.links(m[c, t], v[c, t], W)
            rec[c, t, WOBS_SUM_F2] = (f ** 2).sum()
            rec[c, t, WOBS_SUM_DF2] = (LT.d1(f) ** 2).sum()
            rec[c, t, WOBS_WRAP0], rec[c, t, WOBS_WRAP1] = m[c, t, 0].sum(), m[c, t, 1].sum()
    E.record, E.steps, E.sweeps_per_step, E.keep_every = rec, steps, 1, 1
    E.index = 1 + np.arange(steps)
    E.generator = PlaquetteUpdate(S)
    E.configuration = {'m': m, 'v': v.astype(np.int64)}
    E.observables = worldline_inline_values(rec, N, kappa)
    R = E.to_reference(1, supervillain=sv)
    assert type(R.Action).__name__ == 'Worldline' and R.index_stride == 1
    cfg = R.configuration[2]
    assert cfg['m'].degree == 1 and cfg['m'].dtype == np.int64 and (np.asarray(cfg['v']) == v[1, 2]).all()
    bare = sv.Ensemble(R.Action).from_configurations(sv.configurations.Configurations(
        {k: R.configuration.fields[k] for k in ('m', 'v')}))
    for name in ('ActionDensity', 'InternalEnergyDensity', 'InternalEnergyDensitySquared', 'WindingSquared', 'TorusWrapping',
                 'WrappingSquared'):
        inline, measured = np.asarray(getattr(R, name)), np.asarray(getattr(bare, name))
        assert inline.shape == measured.shape, name
        assert np.allclose(inline, measured, rtol=1e-12, atol=1e-12), name


def test_checkpoint_state_is_plain_data_and_round_trips():
    """generator/_state.py: an (action, generator) pair becomes JSON (no pickle, no class paths) and comes back through the
    whitelisted constructors with its Philox (seed, counter), report counters and -- if assigned -- the numpy rng state."""
    import json
    import supervillain_b200 as svb
    from supervillain_b200.generator import _state
    from supervillain_b200.generator.combining import KeepEvery, Sequentially
    from supervillain_b200.generator.villain import CohomologyUpdate, ExactUpdate, LinkUpdate, NeighborhoodUpdate, SiteUpdate
    from supervillain_b200.generator.worldline import PlaquetteUpdate, WrappingUpdate
    S = svb.Villain(svb.Lattice2D(12), 0.37, W=2)
    G = NeighborhoodUpdate(S, 1.5, 2, seed=2**63 + 5, inline=('ActionDensity',), arithmetic='strict')
    G.counter, G.accepted, G.proposed, G.acceptance, G.sweeps = 17, 3, 99, 0.25, 4
    G.rng = np.random.default_rng(99)
    G.rng.uniform(size=7)
    combo = KeepEvery(5, Sequentially((G, SiteUpdate(S, 0.7, seed=1), LinkUpdate(S, 2, seed=2), ExactUpdate(S, 3, seed=3),
                                       CohomologyUpdate(S, 1, seed=4))), blocked_inline=False)
    text = json.dumps({'action': _state.describe_action(S), 'generator': _state.describe_generator(combo)})
    d = json.loads(text)
    S2 = _state.rebuild_action(d['action'])
    assert type(S2).__name__ == 'Villain' and (S2.Lattice.N, S2.kappa, S2.W) == (12, 0.37, 2)
    back = _state.rebuild_generator(d['generator'], S2)
    assert str(back) == str(combo) and back.stride == 5 and back.blocked_inline is False
    G2 = back.generator.generators[0]
    assert (G2.seed, G2.counter, G2.accepted, G2.proposed, G2.acceptance, G2.sweeps) == (2**63 + 5, 17, 3, 99, 0.25, 4)
    assert (G2.interval_phi, G2.interval_n, G2.inline, G2.arithmetic) == (1.5, 2, ('ActionDensity',), 'strict')
    assert (G2.rng.uniform(size=3) == G.rng.uniform(size=3)).all()                  # the numpy stream continues where it was
    assert [g.seed for g in back.generator.generators[1:]] == [1, 2, 3, 4]
    assert back.generator.generators[3].zs == (-3, -2, -1, 1, 2, 3)
    Wl = svb.Worldline(svb.Lattice2D(8), 0.5)
    pair = Sequentially((PlaquetteUpdate(Wl, seed=7, inline=('WindingSquared',)), WrappingUpdate(Wl, 2, seed=8)))
    back = _state.rebuild_generator(json.loads(json.dumps(_state.describe_generator(pair))), Wl)
    assert str(back) == str(pair) and back.generators[1].interval_w == 2 and back.generators[0].inline == ('WindingSquared',)
    with pytest.raises(TypeError):
        _state.describe_generator(object())
    with pytest.raises(ValueError):
        _state.rebuild_generator({'class': 'os.system'}, S)


def test_checkpoints_of_earlier_builds_are_refused_not_unpickled(tmp_path):
    import pickle
    import supervillain_b200 as svb
    path = tmp_path / 'old.npz'
    np.savez(path, state=np.frombuffer(pickle.dumps(('anything',)), dtype=np.uint8), meta=np.zeros(5, dtype=np.int64))
    with pytest.raises(ValueError, match='pickled'):
        svb.BatchedEnsemble.load(path, device='cpu')


def test_chain_pairs_kept_configurations_with_their_own_steps():
    """BatchedEnsemble.chain(): with steps not a multiple of keep_every (11 steps, every 4th kept -> steps 3 and 7) the
    observable columns and the index come from exactly those steps."""
    import supervillain_b200 as svb
    from supervillain_b200._lib import VOBS_ACTION, VOBS_COUNT
    from supervillain_b200.generator.villain import NeighborhoodUpdate, villain_inline_values
    N, chains, steps, keep, sps = 4, 2, 11, 4, 3
    S = svb.Villain(svb.Lattice2D(N), 0.5)
    E = svb.BatchedEnsemble(S, chains, device='cpu')
    rec = np.zeros((chains, steps, VOBS_COUNT))
    rec[..., VOBS_ACTION] = 100 * np.arange(chains)[:, None] + np.arange(steps)[None, :]
    E.record, E.steps, E.sweeps_per_step, E.keep_every = rec, steps, sps, keep
    E.index = sps * (1 + np.arange(steps))
    E.generator = NeighborhoodUpdate(S)
    E.configuration = {'phi': np.zeros((chains, 2, 1, N, N)), 'n': np.zeros((chains, 2, 2, N, N), dtype=np.int64)}
    E.observables = villain_inline_values(rec, N, 0.5)
    c = E.chain(1)
    assert list(np.asarray(c.index)) == [sps * 4, sps * 8] and c.index_stride == keep * sps
    assert list(np.asarray(c.ActionDensity) * N * N) == [103.0, 107.0]
