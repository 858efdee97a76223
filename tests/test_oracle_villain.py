"""The numpy oracle (oracle/villain_np.py, oracle/lattice_np.py) against golden vectors produced
by the reference itself (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import lattice_np as lat
from oracle import villain_np as V


def test_colour_maps_match_reference(golden_lattice_forms):
    _, extras = golden_lattice_forms
    for N in (3, 4, 5, 6, 7, 8, 9, 32):
        assert (lat.colour_map(N) == extras[f'colour_N{N}']).all()
        order = np.concatenate([np.stack(c, 0) for c in lat.colour_sites(N)], axis=1)
        assert (order == extras[f'colour_order_N{N}']).all()


def test_colour_map_N5_ground_truth():
    # SURVEY.md App. A.2 table
    expect = np.array([[0, 1, 0, 2, 3], [1, 0, 1, 3, 2], [0, 1, 0, 2, 3], [2, 3, 2, 0, 1], [3, 2, 3, 1, 0]])
    assert (lat.colour_map(5) == expect).all()


def test_form_operators_bitexact(golden_lattice_forms):
    cases, _ = golden_lattice_forms
    checked = 0
    for c in cases:
        p = int(c['p'])
        for op in ('d', 'delta', 'face_sum', 'coface_sum'):
            if op in c:
                got = lat.form_op(op, p, c['in'])
                assert got.dtype == c[op].dtype          # dtype preserving, test/test_field_dtypes.py
                assert (got == c[op]).all(), (op, p, int(c['N']))
                checked += 1
    assert checked == 4 * 2 * 8   # 4 sizes x {float,int} x (2 ops on 0-forms + 4 on 1-forms + 2 on 2-forms)


def test_vectorised_step_reproduces_reference_chain(golden_villain_neighborhood):
    for c in golden_villain_neighborhood:
        rng = np.random.default_rng(99)
        phi, n = c['phi0'], c['n0']
        for s in range(int(c['sweeps'])):
            stats = {}
            phi, n = V.neighborhood_step(phi, n, float(c['kappa']), int(c['W']), rng, stats=stats)
            assert (n == c['n'][s]).all()
            assert (phi == c['phi'][s]).all()            # bitwise
            assert stats['accepted'] == int(c['accepted'][s])
            N = int(c['N'])
            assert stats['acceptance'] / N**2 == pytest.approx(float(c['acceptance'][s]), rel=1e-13)


def test_draw_replay_matches_golden_draws(golden_villain_neighborhood):
    for c in golden_villain_neighborhood:
        rng = np.random.default_rng(99)
        for s in range(int(c['sweeps'])):
            d = V.draw_neighborhood(rng, int(c['N']), W=int(c['W']))
            assert (d['u'] == c['u'][s]).all() and (d['dphi'] == c['dphi'][s]).all()
            assert (d['dn_fwd'] == c['dn_fwd'][s]).all() and (d['dn_bwd'] == c['dn_bwd'][s]).all()


def test_dense_scalar_step_reproduces_reference_chain(golden_villain_neighborhood):
    for c in golden_villain_neighborhood:
        if int(c['N']) > 16:
            continue
        phi, n = c['phi0'], c['n0']
        for s in range(int(c['sweeps'])):
            draws = {k: c[k][s] for k in ('u', 'dphi', 'dn_fwd', 'dn_bwd')}
            stats = {}
            phi, n = V.neighborhood_step_dense(phi, n, float(c['kappa']), draws, stats=stats)
            assert (n == c['n'][s]).all()
            assert (phi == c['phi'][s]).all()
            assert stats['accepted'] == int(c['accepted'][s])
            assert V.action(phi, n, float(c['kappa'])) == pytest.approx(float(c['action'][s]), rel=1e-13)


def test_observables_match_reference(golden_villain_observables):
    for c in golden_villain_observables:
        phi, n, kappa = c['phi'], c['n'], float(c['kappa'])
        assert (V.links(phi, n) == c['links']).all()
        assert V.action(phi, n, kappa) == pytest.approx(float(c['action']), rel=1e-14)
        assert V.action_density(phi, n, kappa) == pytest.approx(float(c['ActionDensity']), rel=1e-14)
        assert V.internal_energy_density(phi, n, kappa) == pytest.approx(float(c['InternalEnergyDensity']), rel=1e-14)
        assert V.internal_energy_density(phi, n, kappa) ** 2 == pytest.approx(float(c['InternalEnergyDensitySquared']), rel=1e-13)
        assert V.winding_squared(n) == pytest.approx(float(c['WindingSquared']), rel=1e-14)
        assert (V.torus_wrapping(n) == c['TorusWrapping']).all()
        assert V.wrapping_squared(n) == float(c['WrappingSquared'])
        assert (lat.d1(n) == c['dn']).all()
        np.testing.assert_allclose(V.spin_spin(phi), c['Spin_Spin'], rtol=0, atol=1e-13)
        np.testing.assert_allclose(V.winding_winding(n), c['Winding_Winding'], rtol=0, atol=1e-12)


def test_delta_S_formula_equals_action_difference():
    """Mirror of the reference's test/test_delta_s.py:115-144 on the oracle: the per-site fast
    dS equals S(new) - S(old) to 1e-10."""
    N, kappa = 4, 0.7
    phi, n = V.hot_start(np.random.default_rng(0), N)
    rng = np.random.default_rng(99)
    draws = V.draw_neighborhood(rng, N)
    dS = np.zeros((N, N))
    V.neighborhood_step_dense(phi, n, kappa, draws | {'u': np.ones((N, N))}, dS_out=dS)  # u=1: nothing accepted
    for x0 in range(N):
        for x1 in range(N):
            p2, n2 = phi.copy(), n.copy()
            p2[0, x0, x1] += draws['dphi'][x0, x1]
            n2[0, x0, x1] += draws['dn_fwd'][0, x0, x1]; n2[0, (x0 - 1) % N, x1] += draws['dn_bwd'][0, x0, x1]
            n2[1, x0, x1] += draws['dn_fwd'][1, x0, x1]; n2[1, x0, (x1 - 1) % N] += draws['dn_bwd'][1, x0, x1]
            assert abs(dS[x0, x1] - (V.action(p2, n2, kappa) - V.action(phi, n, kappa))) < 1e-10


def test_decoupled_updates_reproduce_reference_chains(golden_villain_decoupled):
    """SiteUpdate / LinkUpdate / ExactUpdate restatements (oracle/villain_np.py) against chains produced by the
    UNMODIFIED reference with rng = default_rng(99): vectorised form with the replayed numpy stream, and the dense
    per-site / per-link form (the arithmetic the kernels perform) with the stored draws -- fields bit for bit."""
    for c in golden_villain_decoupled:
        kind = ['site', 'link', 'exact'][int(c['kind'])]
        N, kappa, W, sweeps = int(c['N']), float(c['kappa']), int(c['W']), int(c['sweeps'])
        interval = float(c['interval']) if kind == 'site' else int(c['interval'])
        rng = np.random.default_rng(99)
        phi, n = c['phi0'].copy(), c['n0'].copy()
        pd, nd = c['phi0'].copy(), c['n0'].copy()
        for s in range(sweeps):
            st, sd = {}, {}
            if kind == 'site':
                phi, n = V.site_step(phi, n, kappa, rng, interval_phi=interval, stats=st)
                draws = {'u': c['u'][s], 'dphi': c['a'][s], 'dn_fwd': np.zeros((2, N, N), dtype=np.int64),
                         'dn_bwd': np.zeros((2, N, N), dtype=np.int64)}
                pd, nd = V.neighborhood_step_dense(pd, nd, kappa, draws, stats=sd)
            elif kind == 'link':
                phi, n = V.link_step(phi, n, kappa, W, rng, interval_n=interval, stats=st)
                pd, nd = V.link_step_dense(pd, nd, kappa, {'u': c['u'][s], 'a': c['a'][s]}, stats=sd)
            else:
                phi, n = V.exact_step(phi, n, kappa, rng, interval_z=interval, stats=st)
                pd, nd = V.exact_step_dense(pd, nd, kappa, {'u': c['u'][s], 'a': c['a'][s]}, stats=sd)
            for (p_, n_, stats) in ((phi, n, st), (pd, nd, sd)):
                assert (n_ == c['n'][s]).all(), (kind, N, s)
                assert (p_ == c['phi'][s]).all(), (kind, N, s)
                assert stats['accepted'] == int(c['accepted'][s])
                norm = 2 * N * N if kind == 'link' else N * N
                assert stats['acceptance'] / norm == pytest.approx(float(c['acceptance'][s]), rel=1e-12)


def test_cohomology_update_reproduces_reference_chains(golden_villain_cohomology):
    """CohomologyUpdate restatement against chains of the UNMODIFIED reference (rng = default_rng(99)), with the replayed
    numpy stream and with the stored draws."""
    for c in golden_villain_cohomology:
        N, kappa, interval = int(c['N']), float(c['kappa']), int(c['interval'])
        rng = np.random.default_rng(99)
        phi = c['phi0']
        n, nd = c['n0'].copy(), c['n0'].copy()
        for s in range(int(c['sweeps'])):
            st, sd = {}, {}
            _, n = V.cohomology_step(phi, n, kappa, rng, interval_h=interval, stats=st)
            _, nd = V.cohomology_step(phi, nd, kappa, None, interval_h=interval, stats=sd, draws=(c['u'][s], c['h'][s]))
            for n_, stats in ((n, st), (nd, sd)):
                assert (n_ == c['n'][s]).all(), (N, s)
                assert stats['accepted'] == int(c['accepted'][s])
                assert stats['acceptance'] / 2 == pytest.approx(float(c['acceptance'][s]), rel=1e-12)


def test_autocorrelation_restatement_matches_reference(golden_autocorrelation):
    from oracle import lattice_np as lat
    for c in golden_autocorrelation:
        C, tau = lat.autocorrelation(c['data'])
        assert (C == c['C']).all() and tau == int(c['tau'])
        C, tau = lat.autocorrelation(c['data'], mean=3.0)
        assert (C == c['C_mean3']).all() and tau == int(c['tau_mean3'])
    with pytest.raises(ValueError):
        lat.autocorrelation(np.ones(16))


def test_blocking_and_bootstrap_restatements_match_reference(golden_resampling):
    """oracle.lattice_np.block_mean / bootstrap_mean against Blocking._block and Bootstrap._resample of the unmodified
    reference (bit for bit: the same numpy expressions), with unit and non-unit weights."""
    for c in golden_resampling:
        w = c['weight']
        assert c['drop'] == len(c['data']) % int(c['width'])
        assert (lat.block_mean(c['data'], int(c['width']), w) == c['blocked']).all()
        assert (lat.bootstrap_mean(c['data'], c['indices'], w) == c['resampled']).all()
        batch = np.stack([c['data'], 2 * c['data'] + 1])
        np.testing.assert_allclose(lat.block_mean(batch, int(c['width']), w)[1], 2 * c['blocked'] + w[int(c['drop']):].reshape(-1, int(c['width'])).mean(axis=1), rtol=1e-13)
        np.testing.assert_allclose(lat.bootstrap_mean(batch, c['indices'], w)[1], 2 * c['resampled'] + 1, rtol=1e-13)


def test_taxicab_observables_restatement_matches_reference(golden_taxicab):
    """oracle.lattice_np.spin_spin_worldline / vortex_vortex_villain (explicit taxicab paths) against Spin_Spin.Worldline
    (observable/spin.py:50-224) and Vortex_Vortex.Villain (observable/vortex.py:63-189) of the unmodified reference on
    random link fields, even and odd N: 1e-12 relative (the values span 100 orders of magnitude)."""
    for c in golden_taxicab:
        kappa = float(c['kappa'])
        np.testing.assert_allclose(lat.spin_spin_worldline(c['links_w'], kappa), c['spin_spin'], rtol=1e-12)
        np.testing.assert_allclose(lat.vortex_vortex_villain(c['links_v'], kappa), c['vortex_vortex'], rtol=1e-12)
        assert c['spin_spin'][0, 0] == 1 and c['vortex_vortex'][0, 0] == 1
