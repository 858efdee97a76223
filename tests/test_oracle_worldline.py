"""The numpy worldline oracle (oracle/worldline_np.py) against golden vectors produced by the
reference's PlaquetteUpdate / VortexUpdate / CoexactUpdate and observables.  CPU only."""
import numpy as np
import pytest

from oracle import lattice_np as lat
from oracle import worldline_np as WL


def test_vectorised_checkerboard_steps_reproduce_reference(golden_worldline_checkerboard):
    for c in golden_worldline_checkerboard:
        kind, kappa, W, I = str(c['kind']), float(c['kappa']), int(c['W']), int(c['interval'])
        step = WL.vortex_step if kind == 'vortex' else WL.coexact_step
        rng = np.random.default_rng(99)
        m, v = c['m0'], c['v0']
        for s in range(int(c['sweeps'])):
            stats = {}
            m, v = step(m, v, kappa, W, rng, I, stats=stats)
            assert (m == c['m'][s]).all() and (v == c['v'][s]).all()
            assert stats['accepted'] == int(c['accepted'][s])
            assert stats['acceptance'] / int(c['N']) ** 2 == pytest.approx(float(c['acceptance'][s]), rel=1e-13)


def test_dense_checkerboard_restatement_reproduces_reference(golden_worldline_checkerboard):
    """The per-plaquette restatement that the GPU kernel implements == the reference chains."""
    for c in golden_worldline_checkerboard:
        kind, kappa, W = str(c['kind']), float(c['kappa']), int(c['W'])
        m, v = c['m0'], c['v0']
        for s in range(int(c['sweeps'])):
            draws = {'u': c['u'][s], 'a': c['a'][s], 'b': np.zeros_like(c['a'][s])}
            stats = {}
            m, v = WL.checkerboard_step_dense(m, v, kappa, W, draws, kind, stats=stats)
            assert (m == c['m'][s]).all() and (v == c['v'][s]).all(), (kind, int(c['N']), W, s)
            assert stats['accepted'] == int(c['accepted'][s])


def test_draw_replay_matches_golden(golden_worldline_checkerboard):
    for c in golden_worldline_checkerboard:
        rng = np.random.default_rng(99)
        for s in range(int(c['sweeps'])):
            d = WL.draw_checkerboard(rng, int(c['N']), str(c['kind']), int(c['interval']))
            assert (d['u'] == c['u'][s]).all() and (d['a'] == c['a'][s]).all()


def test_sequential_plaquette_step_reproduces_reference(golden_worldline_plaquette):
    for c in golden_worldline_plaquette:
        rng = np.random.default_rng(99)
        np.random.seed(7)                      # test/test_plaquette_update.py:31
        m, v = c['m0'], c['v0']
        for s in range(int(c['sweeps'])):
            stats = {}
            m, v = WL.plaquette_step(m, v, float(c['kappa']), int(c['W']), rng, stats=stats)
            assert (m == c['m'][s]).all() and (v == c['v'][s]).all()
            assert stats['accepted'] == int(c['accepted'][s])
            assert stats['acceptance'] == pytest.approx(float(c['acceptance'][s]), rel=1e-13)
            assert WL.valid(m)
            assert WL.action(m, v, float(c['kappa']), int(c['W'])) == pytest.approx(float(c['action'][s]), rel=1e-13)


def test_observables_match_reference(golden_worldline_observables):
    for c in golden_worldline_observables:
        m, v, kappa, W = c['m'], c['v'], float(c['kappa']), int(c['W'])
        assert (WL.links(m, v, W) == c['links']).all()
        assert WL.action(m, v, kappa, W) == pytest.approx(float(c['action']), rel=1e-14)
        assert WL.action_density(m, v, kappa, W) == pytest.approx(float(c['ActionDensity']), rel=1e-13)
        assert WL.internal_energy_density(m, v, kappa, W) == pytest.approx(float(c['InternalEnergyDensity']), rel=1e-13)
        assert WL.internal_energy_density_squared(m, v, kappa, W) == pytest.approx(float(c['InternalEnergyDensitySquared']), rel=1e-12)
        assert WL.winding_squared(m, v, kappa, W) == pytest.approx(float(c['WindingSquared']), rel=1e-12)
        assert (WL.torus_wrapping(m) == c['TorusWrapping']).all()
        np.testing.assert_allclose(WL.vortex_vortex(v, W), c['Vortex_Vortex'], rtol=0, atol=1e-13)


def test_action_raises_on_constraint_violation():
    m, v = WL.hot_start(np.random.default_rng(0), 4)
    m = m.copy(); m[0, 0, 0] += 1
    with pytest.raises(ValueError):
        WL.action(m, v, 0.5, 1)


@pytest.mark.parametrize('mode', ['joint', 'vortex', 'coexact'])
def test_delta_S_equals_action_difference_and_constraint_kept(mode):
    """Mirror of test/test_delta_s.py:150-247 for the three checkerboard moves, and
    test/test_validity.py: delta m = 0 is preserved."""
    N, kappa, W = 4, 0.7, 2
    m, v = WL.hot_start(np.random.default_rng(3), N)
    draws = WL.draw_checkerboard(np.random.default_rng(99), N, mode)
    dS = np.zeros((N, N))
    WL.checkerboard_step_dense(m, v, kappa, W, draws | {'u': np.ones((N, N))}, mode, dS_out=dS)
    S0 = WL.action(m, v, kappa, W)
    for x0 in range(N):
        for x1 in range(N):
            one = {'u': np.ones((N, N)), 'a': draws['a'], 'b': draws['b']}
            one['u'] = one['u'].copy(); one['u'][x0, x1] = 0.0     # force-accept exactly this plaquette...
            m2, v2 = m.copy(), v.copy()
            a, b = int(draws['a'][x0, x1]), int(draws['b'][x0, x1])
            p0, p1 = (x0 + 1) % N, (x1 + 1) % N
            if mode in ('joint', 'coexact'):
                m2[0, x0, x1] += a; m2[1, p0, x1] += a; m2[0, x0, p1] -= a; m2[1, x0, x1] -= a
            if mode == 'joint':
                v2[0, x0, x1] += b
            if mode == 'vortex':
                v2[0, x0, x1] += a
            assert WL.valid(m2)
            assert abs(dS[x0, x1] - (WL.action(m2, v2, kappa, W) - S0)) < 1e-10
    # a full accepted sweep keeps the constraint
    m3, v3 = WL.checkerboard_step_dense(m, v, kappa, W, draws | {'u': np.zeros((N, N))}, mode)
    assert WL.valid(m3)


def test_wrapping_oracle_reproduces_reference(golden_worldline_wrapping):
    for c in golden_worldline_wrapping:
        N, W, kappa, I = int(c['N']), int(c['W']), float(c['kappa']), int(c['interval'])
        rng = np.random.default_rng(99)
        m, v = c['m0'], c['v0']
        md = c['m0']
        for s in range(int(c['sweeps'])):
            st = {}
            m, v = WL.wrapping_step(m, v, kappa, W, rng, I, stats=st)
            md, _ = WL.wrapping_step_dense(md, c['v0'], kappa, W, {'u': c['u'][s], 'cm': c['cm'][s]})
            assert (m == c['m'][s]).all() and (md == c['m'][s]).all() and WL.valid(m)
            assert st['accepted'] == int(c['accepted'][s])
