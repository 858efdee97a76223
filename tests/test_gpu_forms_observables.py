"""Form operators, action and observables on the GPU against the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import lattice_np as lat
from oracle import villain_np as V

pytestmark = pytest.mark.gpu

import supervillain_b200 as svb                      # noqa: E402
from supervillain_b200 import ops                    # noqa: E402
from supervillain_b200.generator.villain import villain_inline_values   # noqa: E402


def test_form_operators_bitexact_vs_reference(golden_lattice_forms):
    """d, delta, face_sum, coface_sum == the reference's results with `==` (test/test_lattice_kernels.py:17-34),
    dtype preserved (test/test_field_dtypes.py:11-31), scalar 0 at the ends of the complex (:37-42)."""
    cases, _ = golden_lattice_forms
    for c in cases:
        N, p = int(c['N']), int(c['p'])
        L = svb.Lattice2D(N)
        F = svb.Form(c['in'], degree=p, lattice=L)
        for name, fn in (('d', svb.d), ('delta', svb.delta), ('face_sum', lambda f: f.face_sum()),
                         ('coface_sum', lambda f: f.coface_sum())):
            got = fn(F)
            if name in c:
                assert isinstance(got, svb.Form) and got.dtype == c[name].dtype
                assert (np.asarray(got) == c[name]).all(), (name, p, N)
            else:
                assert isinstance(got, int) and got == 0


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32, torch.int32, torch.int64])
def test_form_operators_batched_all_dtypes(dtype):
    N, chains = 12, 5
    rng = np.random.default_rng(0)
    for p, C in ((0, 1), (1, 2), (2, 1)):
        if dtype.is_floating_point:
            a = rng.uniform(-3, 3, (chains, C, N, N))
        else:
            a = rng.integers(-9, 10, (chains, C, N, N))
        t = torch.from_numpy(a).to(dtype).cuda()
        for op in ('d', 'delta', 'face_sum', 'coface_sum'):
            out = ops.form_op(op, p, t)
            if (op, p) in lat._OPS:
                ref = lat.form_op(op, p, t.cpu().numpy())
                assert out.dtype == dtype and (out.cpu().numpy() == ref).all(), (op, p, dtype)
            else:
                assert out is None


def test_exterior_calculus_identities():
    """d^2 = 0, delta^2 = 0, <d a, b> = <a, delta b> (test/test_lattice.py) on integer forms, exactly."""
    N, chains = 16, 8
    rng = np.random.default_rng(1)
    a = torch.from_numpy(rng.integers(-5, 6, (chains, 1, N, N))).cuda()
    b = torch.from_numpy(rng.integers(-5, 6, (chains, 2, N, N))).cuda()
    v = torch.from_numpy(rng.integers(-5, 6, (chains, 1, N, N))).cuda()
    assert (ops.form_op('d', 1, ops.form_op('d', 0, a)) == 0).all()
    assert (ops.form_op('delta', 1, ops.form_op('delta', 2, v)) == 0).all()
    assert ((ops.form_op('d', 0, a) * b).sum() == (a * ops.form_op('delta', 1, b)).sum()).item()
    assert ((ops.form_op('d', 1, b) * v).sum() == (b * ops.form_op('delta', 2, v)).sum()).item()


def test_villain_action_and_observables_match_reference(golden_villain_observables):
    for c in golden_villain_observables:
        N, kappa = int(c['N']), float(c['kappa'])
        S = svb.Villain(svb.Lattice2D(N), kappa)
        assert S(c['phi'], c['n']) == pytest.approx(float(c['action']), rel=1e-12)
        assert (np.asarray(S.links(svb.Form(c['phi'], degree=0, lattice=S.Lattice), c['n'])) == c['links']).all()
        vals = villain_inline_values(S.observables(c['phi'], c['n']).cpu().numpy()[0], N, kappa)
        for name in ('ActionDensity', 'InternalEnergyDensity', 'InternalEnergyDensitySquared', 'WindingSquared',
                     'WrappingSquared'):
            assert vals[name] == pytest.approx(float(c[name]), rel=1e-12), name
        assert (vals['TorusWrapping'] == c['TorusWrapping']).all()
        phi = torch.from_numpy(c['phi'][None]).cuda()
        C = ops.villain_spin_spin(phi).cpu().numpy()[0]
        np.testing.assert_allclose(C, c['Spin_Spin'], rtol=0, atol=1e-12)
        dn = ops.form_op('d', 1, torch.from_numpy(c['n'][None]).cuda()).cpu().numpy()[0]
        assert (dn == c['dn']).all()


def test_gauge_invariance_of_observables():
    """phi -> phi + 2 pi k, n -> n + dk leaves every observable unchanged to 1e-12 (test/test_gauge-invariance.py)."""
    N, kappa, chains = 16, 0.6, 4
    rng = np.random.default_rng(3)
    phi, n = V.hot_start(rng, N, chains)
    k = rng.integers(-3, 4, (chains, 1, N, N))
    phi2, n2 = phi + 2 * np.pi * k, n + lat.d0(k)
    S = svb.Villain(svb.Lattice2D(N), kappa)
    r1, r2 = S.observables(phi, n).cpu().numpy(), S.observables(phi2, n2).cpu().numpy()
    np.testing.assert_allclose(r1[:, :2], r2[:, :2], rtol=1e-11)
    assert (r1[:, 2:4] == r2[:, 2:4]).all()              # sum of dk vanishes on the torus
    C1 = ops.villain_spin_spin(torch.from_numpy(phi).cuda()).cpu().numpy()
    C2 = ops.villain_spin_spin(torch.from_numpy(phi2).cuda()).cpu().numpy()
    np.testing.assert_allclose(C1, C2, atol=1e-12)


def test_ensemble_reference_protocol_config1():
    """BASELINE config 1: L=5 (odd: four colours), kappa=0.5, one chain, cold start, the call of test/end-to-end.py:53."""
    S = svb.Villain(svb.Lattice2D(5), 0.5, W=1)
    G = svb.generator.villain.NeighborhoodUpdate(S, seed=5, inline=('ActionDensity', 'WindingSquared', 'TorusWrapping'))
    E = svb.Ensemble(S).generate(200, G, 'cold')
    assert np.asarray(E.phi).shape == (200, 1, 5, 5) and np.asarray(E.n).dtype == np.int64
    assert len(E) == 200 and E.configuration.ActionDensity.shape == (200,)
    k = 137
    assert E.configuration.ActionDensity[k] == pytest.approx(float(V.action_density(np.asarray(E.phi[k]), np.asarray(E.n[k]), 0.5)), rel=1e-12)
    assert E.configuration.WindingSquared[k] == pytest.approx(float(V.winding_squared(np.asarray(E.n[k]))), rel=1e-12)
    assert (E.configuration.TorusWrapping[k] == V.torus_wrapping(np.asarray(E.n[k]))).all()
    assert 'neighborhood proposals accepted of 5000 proposed updates' in G.report()
    cut = E.cut(50).every(10)
    assert len(cut) == 15 and cut.index_stride == 10
    more = svb.Ensemble.continue_from(cut, 5)
    assert len(more) == 5 and more.index[0] == cut.index[-1] + 10


def test_winding_and_vortex_correlators_match_reference(golden_villain_observables, golden_worldline_observables):
    """Winding_Winding.Villain (winding.py:77-86) and Vortex_Vortex.Worldline (vortex.py:22-37) to 1e-12, and the
    reference's identity Winding_Winding[origin] == WindingSquared (test/test_winding.py:21-43)."""
    for c in golden_villain_observables:
        n = torch.from_numpy(c['n'][None]).to(torch.int32).cuda()
        C = ops.correlation('winding', n).cpu().numpy()[0]
        np.testing.assert_allclose(C, c['Winding_Winding'], rtol=0, atol=1e-12)
        assert C[0, 0].real == pytest.approx(float(c['WindingSquared']), rel=1e-12)
    for c in golden_worldline_observables:
        v = torch.from_numpy(c['v'][None]).to(torch.int32).cuda()
        C = ops.correlation('vortex', v, W=int(c['W'])).cpu().numpy()[0]
        np.testing.assert_allclose(C, c['Vortex_Vortex'], rtol=0, atol=1e-12)


def test_autocorrelation_matches_reference(golden_autocorrelation):
    """svb_autocorrelation (direct circular sums) against supervillain.analysis.autocorrelation (FFT): C within 1e-11 of
    C(0) = 1, tau identical; one series per chain at once; the reference's ValueError for a flat series."""
    import supervillain_b200 as svb
    from oracle import lattice_np as lat
    for c in golden_autocorrelation:
        C, tau = svb.analysis.autocorrelation(c['data'])
        np.testing.assert_allclose(C, c['C'], rtol=0, atol=1e-11)
        assert tau == int(c['tau'])
        C, tau = svb.analysis.autocorrelation(c['data'], mean=3.0)
        np.testing.assert_allclose(C, c['C_mean3'], rtol=0, atol=1e-11)
        assert tau == int(c['tau_mean3'])
        assert svb.analysis.autocorrelation_time(c['data']) == int(c['tau'])
    # a batch: the ActionDensity columns of many chains
    S = svb.Villain(svb.Lattice2D(8), 0.4)
    G = svb.generator.villain.NeighborhoodUpdate(S, seed=3)
    E = svb.BatchedEnsemble(S, 48).generate(300, G, 'hot', start_seed=1)
    C, tau = svb.analysis.autocorrelation(E.ActionDensity)
    for k in range(48):
        Cr, tr = lat.autocorrelation(E.ActionDensity[k])
        np.testing.assert_allclose(C[k], Cr, rtol=0, atol=1e-11)
        assert tau[k] == tr
    with pytest.raises(ValueError):
        svb.analysis.autocorrelation(np.ones(32))


def test_batched_ensemble_autocorrelation_time_matches_per_chain_reference_definition():
    from oracle import lattice_np as lat
    S = svb.Villain(svb.Lattice2D(8), 0.4)
    G = svb.generator.villain.NeighborhoodUpdate(S, seed=8)
    E = svb.BatchedEnsemble(S, 24).generate(400, G, 'hot', start_seed=2)
    every = E.autocorrelation_time(every=True)
    expect = 0
    for name in ('ActionDensity', 'WindingSquared', 'WrappingSquared'):
        for k in range(24):
            try:
                tau = lat.autocorrelation(getattr(E, name)[k])[1]
            except ValueError:
                tau = -1
            assert every[name][k] == tau, (name, k)
            expect = max(expect, tau)
    assert 'TorusWrapping' not in every                      # not a scalar column
    total = E.autocorrelation_time()
    assert total >= expect and total == max(int(v.max()) for v in every.values())


@pytest.mark.parametrize('N', [16, 32, 64, 128, 256, 1024])
def test_fft_correlators_match_the_reference_formula(N):
    """The shared-memory FFT route (N = 16, 32, 64) and the three-launch route (power-of-two N from 128 to 4096: configs 4
    and 5) for all three correlators against the restated Lattice.correlation (compact.py:465-536), 1e-12, on a batch."""
    rng = np.random.default_rng(N)
    chains = 5 if N <= 256 else 2
    phi = rng.uniform(-7, 7, (chains, 1, N, N))
    n = rng.integers(-3, 4, (chains, 2, N, N))
    v = rng.integers(-5, 6, (chains, 1, N, N))
    Cs = ops.villain_spin_spin(torch.from_numpy(phi).cuda()).cpu().numpy()
    Cw = ops.correlation('winding', torch.from_numpy(n).to(torch.int32).cuda()).cpu().numpy()
    Cv = ops.correlation('vortex', torch.from_numpy(v).to(torch.int32).cuda(), W=3).cpu().numpy()
    for c in range(chains):
        s = np.exp(1j * phi[c, 0])
        np.testing.assert_allclose(Cs[c], lat.correlation(s, s), rtol=0, atol=1e-12)
        dn = lat.d1(n[c])[0].astype(np.float64)
        np.testing.assert_allclose(Cw[c], lat.correlation(dn, dn), rtol=0, atol=1e-12 * max(1.0, np.abs(dn).max() ** 2))
        e = np.exp(2j * np.pi * v[c, 0] / 3)
        np.testing.assert_allclose(Cv[c], lat.correlation(e, e), rtol=0, atol=1e-12)


@pytest.mark.parametrize('N', [128, 256, 512, 1024, 2048])
def test_split_column_transforms_match_the_reference_formula(N, monkeypatch):
    """The transforms of the biggest lattices (N >= 1024) run in two steps, r = 64 r1 + r2 (correlation_split_*_kernel).
    With the threshold lowered the same kernels serve sizes that can be compared element by element with the restated
    Lattice.correlation (compact.py:465-536): n1 = N / 64 = 2, 4, 8 rows per outer transform (radix-2 stages) and 16, 32 (two
    radix passes, as the 64 of L = 4096), 1e-12."""
    monkeypatch.setenv('SVB_CORR_SPLIT_MIN_N', '128')
    rng = np.random.default_rng(N + 1)
    chains = 3 if N <= 256 else 1
    phi = rng.uniform(-7, 7, (chains, 1, N, N))
    n = rng.integers(-3, 4, (chains, 2, N, N))
    Cs = ops.villain_spin_spin(torch.from_numpy(phi).cuda()).cpu().numpy()
    Cw = ops.correlation('winding', torch.from_numpy(n).to(torch.int32).cuda()).cpu().numpy()
    monkeypatch.setenv('SVB_CORR_R16', '0')               # the radix-8 row kernels (the default rows are radix-16)
    Cs8 = ops.villain_spin_spin(torch.from_numpy(phi).cuda()).cpu().numpy()
    np.testing.assert_allclose(Cs8, Cs, rtol=0, atol=1e-13)
    monkeypatch.delenv('SVB_CORR_R16')
    monkeypatch.setenv('SVB_CORR_SPLIT_MIN_N', '1000000')
    whole = ops.villain_spin_spin(torch.from_numpy(phi).cuda()).cpu().numpy()
    for c in range(chains):
        s = np.exp(1j * phi[c, 0])
        np.testing.assert_allclose(Cs[c], lat.correlation(s, s), rtol=0, atol=1e-12)
        dn = lat.d1(n[c])[0].astype(np.float64)
        np.testing.assert_allclose(Cw[c], lat.correlation(dn, dn), rtol=0, atol=1e-12 * max(1.0, np.abs(dn).max() ** 2))
    np.testing.assert_allclose(Cs, whole, rtol=0, atol=1e-13)


@pytest.mark.parametrize('N,chains', [(128, 5), (256, 3), (512, 2), (1024, 2), (2048, 1)])
def test_mid_lattice_correlators_match_the_reference_formula(N, chains, monkeypatch):
    """128 <= N <= 512 (config 4's lattices): the row kernels of the split with 4096 / N rows per item and the fused
    column kernel (correlation_columns_fused_kernel), all three kinds, against the restated Lattice.correlation
    (compact.py:465-536) to 1e-12 and against the radix-2 kernels they replace (SVB_CORR_ROUTE=legacy)."""
    monkeypatch.delenv('SVB_CORR_SPLIT_MIN_N', raising=False)
    monkeypatch.delenv('SVB_CORR_ROUTE', raising=False)
    rng = np.random.default_rng(N + 7)
    phi = rng.uniform(-7, 7, (chains, 1, N, N))
    n = rng.integers(-3, 4, (chains, 2, N, N))
    v = rng.integers(-4, 5, (chains, 1, N, N))
    tphi, tn, tv = torch.from_numpy(phi).cuda(), torch.from_numpy(n).to(torch.int32).cuda(), torch.from_numpy(v).to(torch.int32).cuda()
    Cs = ops.villain_spin_spin(tphi).cpu().numpy()
    Cs32 = ops.villain_spin_spin(tphi.to(torch.float32)).cpu().numpy()
    Cw = ops.correlation('winding', tn).cpu().numpy()
    Cv = ops.correlation('vortex', tv, W=3).cpu().numpy()
    monkeypatch.setenv('SVB_CORR_ROUTE', 'legacy')
    legacy = ops.villain_spin_spin(tphi).cpu().numpy()
    np.testing.assert_allclose(Cs, legacy, rtol=0, atol=1e-13)
    monkeypatch.delenv('SVB_CORR_ROUTE')
    # the radix-8 kernels throughout (0) and the radix-16 column kernel at every size (2): the same result
    for r16 in ('0', '2'):
        monkeypatch.setenv('SVB_CORR_R16', r16)
        np.testing.assert_allclose(ops.villain_spin_spin(tphi).cpu().numpy(), Cs, rtol=0, atol=1e-13)
        np.testing.assert_allclose(ops.correlation('winding', tn).cpu().numpy(), Cw, rtol=0, atol=1e-12)
    monkeypatch.delenv('SVB_CORR_R16')
    for c in range(chains):
        s = np.exp(1j * phi[c, 0])
        np.testing.assert_allclose(Cs[c], lat.correlation(s, s), rtol=0, atol=1e-12)
        s32 = np.exp(1j * phi[c, 0].astype(np.float32).astype(np.float64))
        np.testing.assert_allclose(Cs32[c], lat.correlation(s32, s32), rtol=0, atol=1e-12)
        dn = lat.d1(n[c])[0].astype(np.float64)
        np.testing.assert_allclose(Cw[c], lat.correlation(dn, dn), rtol=0, atol=1e-12 * max(1.0, np.abs(dn).max() ** 2))
        e = np.exp(2j * np.pi * v[c, 0] / 3)
        np.testing.assert_allclose(Cv[c], lat.correlation(e, e), rtol=0, atol=1e-12)


@pytest.mark.parametrize('N,chains', [(128, 333), (256, 90), (512, 21)])
def test_mid_lattice_correlators_many_items_per_cta(N, chains, monkeypatch):
    """More work items than resident CTAs (2 per SM x 148): every CTA of the TMA-fed kernels goes round its loop several
    times -- the next tile requested when the store of this one has read shared memory, the staged field of the next item,
    the mbarrier parities -- and an odd chain count leaves some CTAs one item short.  All three kinds against the
    radix-2 kernels (SVB_CORR_ROUTE=legacy), which share none of that machinery, and a few chains against numpy."""
    monkeypatch.delenv('SVB_CORR_SPLIT_MIN_N', raising=False)
    rng = np.random.default_rng(N)
    tphi = torch.from_numpy(rng.uniform(-7, 7, (chains, 1, N, N))).cuda()
    tn = torch.from_numpy(rng.integers(-3, 4, (chains, 2, N, N))).to(torch.int32).cuda()
    tv = torch.from_numpy(rng.integers(-4, 5, (chains, 1, N, N))).to(torch.int32).cuda()
    got = [ops.villain_spin_spin(tphi), ops.correlation('winding', tn), ops.correlation('vortex', tv, W=5)]
    monkeypatch.setenv('SVB_CORR_ROUTE', 'legacy')
    ref = [ops.villain_spin_spin(tphi), ops.correlation('winding', tn), ops.correlation('vortex', tv, W=5)]
    for g, r, tol in zip(got, ref, (1e-13, 1e-11, 1e-13)):
        assert float((g - r).abs().max()) < tol
    for c in (0, chains // 2, chains - 1):
        s = np.exp(1j * tphi[c, 0].cpu().numpy())
        np.testing.assert_allclose(got[0][c].cpu().numpy(), lat.correlation(s, s), rtol=0, atol=1e-12)


def test_warp_per_chain_l32_correlator_equals_the_cta_kernel(monkeypatch):
    """L = 32: correlation_fft32_warp_kernel (a warp per chain, 32-point transforms in registers; opt-in with
    SVB_CORR_FFT32_WARP=1, it measured slower) against the default correlation_fft_kernel<32> for all kinds and against
    numpy, with more chains than resident warps and an odd count: two independent implementations of the same transform."""
    N, chains = 32, 5003
    rng = np.random.default_rng(32)
    tphi = torch.from_numpy(rng.uniform(-7, 7, (chains, 1, N, N))).cuda()
    tn = torch.from_numpy(rng.integers(-3, 4, (chains, 2, N, N))).to(torch.int32).cuda()
    tv = torch.from_numpy(rng.integers(-4, 5, (chains, 1, N, N))).to(torch.int32).cuda()
    monkeypatch.setenv('SVB_CORR_FFT32_WARP', '1')
    got = [ops.villain_spin_spin(tphi), ops.villain_spin_spin(tphi.to(torch.float32)), ops.correlation('winding', tn), ops.correlation('vortex', tv, W=5)]
    monkeypatch.setenv('SVB_CORR_FFT32_WARP', '0')
    ref = [ops.villain_spin_spin(tphi), ops.villain_spin_spin(tphi.to(torch.float32)), ops.correlation('winding', tn), ops.correlation('vortex', tv, W=5)]
    for g, r, tol in zip(got, ref, (1e-13, 1e-13, 1e-11, 1e-13)):
        assert float((g - r).abs().max()) < tol
    for c in (0, 2500, chains - 1):
        s = np.exp(1j * tphi[c, 0].cpu().numpy())
        np.testing.assert_allclose(got[0][c].cpu().numpy(), lat.correlation(s, s), rtol=0, atol=1e-12)
        dn = lat.d1(tn[c].cpu().numpy().astype(np.int64))[0].astype(np.float64)
        np.testing.assert_allclose(got[2][c].cpu().numpy(), lat.correlation(dn, dn), rtol=0, atol=1e-11)


@pytest.mark.parametrize('N', [32, 128, 256])
def test_spin_correlator_for_large_angles(N):
    """phi is never wrapped by the reference's updates, so a long run can reach angles of any magnitude: the spin field
    exp(i phi) of the FFT kernels (the small-lattice kernel, radix-16 lines of 128 and of 256) against numpy's exponential for
    |phi| up to 1e12, 1e-12 on the correlator.  (A table-driven sincos, 30 instructions for the library's 85, passed this
    test too and bought 1 % at L = 4096 and nothing on the small lattices: not kept.)"""
    rng = np.random.default_rng(N)
    scales = (1.0, 1.0e2, 9.9e4, 3.0e6, 1.0e12)
    phi = np.stack([rng.uniform(-sc, sc, (1, N, N)) for sc in scales])
    C = ops.villain_spin_spin(torch.from_numpy(phi).cuda()).cpu().numpy()
    for c in range(len(scales)):
        s = np.exp(1j * phi[c, 0])
        np.testing.assert_allclose(C[c], lat.correlation(s, s), rtol=0, atol=1e-12, err_msg=f'scale {scales[c]}')


def test_fft_correlator_of_a_config5_lattice_properties(monkeypatch):
    """L = 4096 (config 5), too large to compare element by element in a test: size-independent properties of
    Lattice.correlation instead -- C[0] = mean |s|^2 = 1 for a spin field, C[-r] = conj(C[r]), sum_r C[r] = N^2 |mean s|^2 ...
    and a plane wave s = exp(i k.x), whose correlator is exp(-i k.r) exactly.  The radix-8 row kernels (SVB_CORR_R16=0)
    and the default radix-16 ones agree to 1e-13."""
    N = 4096
    rng = np.random.default_rng(5)
    phi = rng.uniform(-np.pi, np.pi, (1, 1, N, N))
    monkeypatch.setenv('SVB_CORR_R16', '0')
    C8 = ops.villain_spin_spin(torch.from_numpy(phi).cuda())[0]
    monkeypatch.delenv('SVB_CORR_R16')
    C = ops.villain_spin_spin(torch.from_numpy(phi).cuda())[0]
    assert float((C - C8).abs().max()) < 1e-13
    del C8
    s = torch.from_numpy(np.exp(1j * phi[0, 0])).cuda()
    assert abs(complex(C[0, 0]) - 1.0) < 1e-12
    flipped = torch.roll(torch.flip(C, (0, 1)), (1, 1), (0, 1))
    assert float((flipped - C.conj()).abs().max()) < 1e-12
    assert abs(complex(C.sum()) - N * N * abs(complex(s.mean())) ** 2) < 1e-8
    # one displacement checked against its definition  C[r] = mean_x conj(s[x]) s[x - r]
    r0, r1 = 1234, 77
    direct = (s.conj() * torch.roll(s, (r0, r1), (0, 1))).mean()
    assert abs(complex(C[r0, r1]) - complex(direct)) < 1e-12
    k0, k1 = 5, 1000
    x = np.arange(N)
    wave = 2 * np.pi * (k0 * x[:, None] + k1 * x[None, :]) / N
    Cw = ops.villain_spin_spin(torch.from_numpy(wave[None, None]).cuda())[0].cpu().numpy()
    np.testing.assert_allclose(Cw, np.exp(-1j * wave), rtol=0, atol=1e-9)


def test_blocking_and_bootstrap_match_reference(golden_resampling):
    """svb_block_mean / svb_bootstrap_mean against Blocking._block and Bootstrap._resample of the unmodified reference on
    its own resampling indices (1e-13: the bootstrap sums in numpy's order, block means differ from numpy's pairwise sums
    in the last bits), with and without weights; then the batched classes on an ensemble of chains against the oracle,
    including Bootstrap(Blocking(E))."""
    for c in golden_resampling:
        data = torch.from_numpy(c['data'][None]).cuda()
        unit = bool((c['weight'] == 1).all())
        w = None if unit else torch.from_numpy(c['weight']).cuda()
        blocked = ops.block_mean(data, int(c['width']), int(c['drop']), weight=w)[0].cpu().numpy()
        np.testing.assert_allclose(blocked, c['blocked'], rtol=1e-13, atol=1e-15)
        res = ops.bootstrap_mean(data, torch.from_numpy(c['indices']).cuda(), weight=w)[0].cpu().numpy()
        np.testing.assert_allclose(res, c['resampled'], rtol=1e-13, atol=1e-15)
    S = svb.Villain(svb.Lattice2D(8), 0.4)
    G = svb.generator.villain.NeighborhoodUpdate(S, seed=3)
    E = svb.BatchedEnsemble(S, 40).generate(203, G, 'hot', start_seed=1)
    assert len(E) == 203
    B = svb.analysis.Blocking(E, width=10)
    assert (B.drop, B.blocks, len(B)) == (3, 20, 20)
    np.testing.assert_allclose(B.ActionDensity, lat.block_mean(E.ActionDensity, 10), rtol=1e-13)
    np.testing.assert_allclose(B.index, np.asarray(E.index)[3:].reshape(-1, 10).mean(axis=1))
    np.random.seed(4)
    R = svb.analysis.Bootstrap(E, draws=25)
    np.random.seed(4)
    expect_idx = np.random.randint(0, 203, (203, 25))
    assert (R.indices == expect_idx).all()                    # drawn exactly as the reference draws them
    np.testing.assert_allclose(R.WindingSquared, lat.bootstrap_mean(E.WindingSquared, expect_idx), rtol=1e-13)
    mean, err = R.estimate('ActionDensity')
    assert mean.shape == err.shape == (40,) and (err > 0).all()
    assert np.abs(mean - E.ActionDensity.mean(axis=1)).max() < 5 * err.max()
    RB = svb.analysis.Bootstrap(B, draws=12, indices=np.random.randint(0, 20, (20, 12)))
    np.testing.assert_allclose(RB.ActionDensity, lat.bootstrap_mean(lat.block_mean(E.ActionDensity, 10), RB.indices), rtol=1e-13)
    with pytest.raises(AttributeError):
        R.NoSuchObservable
    auto = svb.analysis.Blocking(E)                          # width='auto': the ensemble's autocorrelation time
    assert auto.width == E.autocorrelation_time()


def test_taxicab_observables_match_reference(golden_taxicab):
    """Spin_Spin.Worldline and Vortex_Vortex.Villain (the taxicab reweighting observables) from the fields, against the
    unmodified reference on its own configurations (even and odd N; 1e-11 relative over 100 orders of magnitude), then
    against the oracle on a batch at a production size with per-chain kappa."""
    for c in golden_taxicab:
        kappa, W = float(c['kappa']), int(c['W'])
        m = torch.from_numpy(c['m'][None]).to(torch.int32).cuda()
        v = torch.from_numpy(c['v'][None]).to(torch.int32).cuda()
        spin = ops.worldline_spin_spin(m, v, kappa, W=W)[0].cpu().numpy()
        np.testing.assert_allclose(spin, c['spin_spin'], rtol=1e-11)
        phi = torch.from_numpy(c['phi'][None]).cuda()
        n = torch.from_numpy(c['n'][None]).to(torch.int32).cuda()
        vortex = ops.villain_vortex_vortex(phi, n, kappa)[0].cpu().numpy()
        np.testing.assert_allclose(vortex, c['vortex_vortex'], rtol=1e-11)
        assert spin[0, 0] == 1 and vortex[0, 0] == 1
    rng = np.random.default_rng(2)
    chains, N = 6, 16
    links = rng.normal(size=(chains, 2, N, N)) * 0.3
    kc = np.linspace(0.4, 1.5, chains)
    for kind, fn in (('spin', lat.spin_spin_worldline), ('vortex', lat.vortex_vortex_villain)):
        out = ops.taxicab_correlator(kind, torch.from_numpy(links).cuda(), 1.0, kappa_chain=torch.from_numpy(kc).cuda()).cpu().numpy()
        for k in range(chains):
            np.testing.assert_allclose(out[k], fn(links[k], kc[k]), rtol=1e-11)


def test_ensemble_measures_two_point_observables_of_kept_configurations():
    """BatchedEnsemble.measure: the reference's two-point observables by name on the kept draws of every chain, against
    the oracle's restatements applied to the same configurations."""
    from oracle import villain_np as V
    S = svb.Villain(svb.Lattice2D(8), 0.6)
    G = svb.generator.villain.NeighborhoodUpdate(S, seed=2)
    E = svb.BatchedEnsemble(S, 5).generate(6, G, 'hot', start_seed=4, keep_every=2)
    vv, ss, ww = E.measure('Vortex_Vortex'), E.measure('Spin_Spin'), E.measure('Winding_Winding')
    assert vv.shape == ss.shape == ww.shape == (5, 3, 8, 8)
    for c in range(5):
        for t in range(3):
            phi, n = E.configuration['phi'][c, t], E.configuration['n'][c, t]
            np.testing.assert_allclose(vv[c, t], lat.vortex_vortex_villain(V.links(phi, n), 0.6), rtol=1e-11)
            s = np.exp(1j * phi[0])
            np.testing.assert_allclose(ss[c, t], lat.correlation(s, s), atol=1e-12)
    with pytest.raises(NotImplementedError):
        E.measure('NoSuchCorrelator')
    Sw = svb.Worldline(svb.Lattice2D(8), 0.6)
    Ew = svb.BatchedEnsemble(Sw, 4).generate(4, svb.generator.worldline.PlaquetteUpdate(Sw, seed=1), 'cold', keep_every=2)
    sp = Ew.measure('Spin_Spin')
    for c in range(4):
        m, v = Ew.configuration['m'][c, 1], Ew.configuration['v'][c, 1]
        links = m - lat.delta2(v)
        np.testing.assert_allclose(sp[c, 1], lat.spin_spin_worldline(links, 0.6), rtol=1e-11)
