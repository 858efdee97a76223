"""Pin the C oracle (oracle/svb_oracle.c) against the reference's golden vectors and the numpy oracle.  CPU only."""
import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import lattice_np as lat
from oracle import philox_np as P
from oracle import villain_np as V
from oracle import worldline_np as WL


def test_philox_and_colours():
    assert [int(x) for x in C.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    for N in (3, 4, 5, 8, 9):
        assert (C.colour_map(N) == lat.colour_map(N)).all()


def test_villain_dense_reproduces_reference_chain(golden_villain_neighborhood):
    for c in golden_villain_neighborhood:
        phi, n, kappa = c['phi0'], c['n0'], float(c['kappa'])
        for s in range(int(c['sweeps'])):
            draws = {k: c[k][s] for k in ('u', 'dphi', 'dn_fwd', 'dn_bwd')}
            phi, n, acc, accp = C.villain_sweep_dense(phi, n, kappa, draws)
            assert (n == c['n'][s]).all() and (phi == c['phi'][s]).all()
            assert acc == int(c['accepted'][s])
            assert accp / int(c['N']) ** 2 == pytest.approx(float(c['acceptance'][s]), rel=1e-12)
            assert C.villain_action(phi, n, kappa) == pytest.approx(float(c['action'][s]), rel=1e-13)


def test_villain_philox_driver_equals_numpy_oracle():
    N, chains, kappa, W = 8, 3, 0.4, 2
    phi0, n0 = V.hot_start(np.random.default_rng(0), N, chains)
    phi, n, acc, accp = C.villain_sweep_philox(phi0, n0, kappa, W=W, n_sweeps=3, seed=77, sweep0=2, chain0=5)
    for c in range(chains):
        p, q = phi0[c], n0[c]
        for s in range(3):
            p, q = V.neighborhood_step_dense(p, q, kappa, P.villain_draws(77, 5 + c, 2 + s, N, W=W))
        assert (phi[c] == p).all() and (n[c] == q).all()


def test_worldline_dense_reproduces_reference_chains(golden_worldline_checkerboard):
    for c in golden_worldline_checkerboard:
        kind, kappa, W = str(c['kind']), float(c['kappa']), int(c['W'])
        m, v = c['m0'], c['v0']
        for s in range(int(c['sweeps'])):
            draws = {'u': c['u'][s], 'a': c['a'][s], 'b': np.zeros_like(c['a'][s])}
            m, v, acc, accp = C.worldline_sweep_dense(m, v, kappa, W, draws, kind)
            assert (m == c['m'][s]).all() and (v == c['v'][s]).all()
            assert acc == int(c['accepted'][s])


@pytest.mark.parametrize('mode', ['joint', 'vortex', 'coexact'])
def test_worldline_philox_driver_equals_numpy_oracle(mode):
    N, chains, kappa, W = 6, 2, 0.5, 3
    m0, v0 = WL.hot_start(np.random.default_rng(1), N, chains)
    m, v, acc, accp = C.worldline_sweep_philox(m0, v0, kappa, W=W, mode=mode, interval=2, n_sweeps=3, seed=9, chain0=1)
    for c in range(chains):
        a, b = m0[c], v0[c]
        for s in range(3):
            a, b = WL.checkerboard_step_dense(a, b, kappa, W, P.worldline_draws(9, 1 + c, s, N, mode, 2), mode)
        assert (m[c] == a).all() and (v[c] == b).all()
