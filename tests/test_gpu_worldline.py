"""Parity of the CUDA worldline path with the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import lattice_np as lat
from oracle import philox_np as P
from oracle import worldline_np as WL

pytestmark = pytest.mark.gpu

import supervillain_b200 as svb                      # noqa: E402
from supervillain_b200 import ops                    # noqa: E402
from supervillain_b200._lib import (WOBS_ACCEPTANCE, WOBS_ACCEPTED, WOBS_COUNT, WOBS_DELTA_M_ABS, WOBS_SUM_DF2,  # noqa: E402
                                    WOBS_SUM_F2, WOBS_WRAP0, WOBS_WRAP1)
from supervillain_b200.generator.worldline import CoexactUpdate, PlaquetteUpdate, VortexUpdate   # noqa: E402


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize('path', ['smem', 'global'])
def test_injected_rng_reproduces_reference_checkerboard_chains(golden_worldline_checkerboard, path):
    """VortexUpdate / CoexactUpdate with rng=default_rng(99): identical m, v and accept counts to the
    reference's step_reference (the pattern of test/test_vortex_sparse.py, test/test_coexact_sparse.py)."""
    for c in golden_worldline_checkerboard:
        kind, N, W, kappa, I = str(c['kind']), int(c['N']), int(c['W']), float(c['kappa']), int(c['interval'])
        S = svb.Worldline(svb.Lattice2D(N), kappa, W=W)
        G = (VortexUpdate if kind == 'vortex' else CoexactUpdate)(S, I, path=path)
        G.rng = np.random.default_rng(99)
        cfg = {'m': c['m0'], 'v': c['v0']}
        for s in range(int(c['sweeps'])):
            before = (G.accepted, G.acceptance)
            cfg = cfg | G.step(cfg)
            assert (np.asarray(cfg['m']) == c['m'][s]).all(), (kind, N, W, s)
            assert (np.asarray(cfg['v']) == c['v'][s]).all(), (kind, N, W, s)
            assert G.accepted - before[0] == int(c['accepted'][s])
            assert G.acceptance - before[1] == pytest.approx(float(c['acceptance'][s]), rel=1e-12)


@pytest.mark.parametrize('mode', ['joint', 'vortex', 'coexact'])
@pytest.mark.parametrize('path', ['smem', 'global'])
@pytest.mark.parametrize('N,W,kappa', [(4, 1, 0.5), (5, 2, 0.5), (8, 3, 0.3), (7, 1, 0.7), (16, 1, 0.5), (64, 1, 0.5)])
def test_philox_mode_matches_oracle_replay(mode, path, N, W, kappa):
    chains, sweeps, seed, chain0, sweep0 = 2, 3, 42, 9, 4
    if N == 64:
        chains, sweeps = 1, 1
    m0, v0 = WL.hot_start(np.random.default_rng(N + W), N, chains)
    m, v = dev(m0, torch.int32), dev(v0, torch.int32)
    obs = torch.zeros((chains, WOBS_COUNT), dtype=torch.float64, device='cuda')
    dS = torch.zeros((chains, N, N), dtype=torch.float64, device='cuda')
    mask = torch.zeros((chains, N, N), dtype=torch.uint8, device='cuda')
    ops.worldline_sweep(m, v, kappa, W=W, mode=mode, interval=1, n_sweeps=sweeps, seed=seed, sweep0=sweep0,
                        chain0=chain0, path=path, obs=obs, dS_out=dS, accept_mask=mask)
    rec = obs.cpu().numpy()
    for c in range(chains):
        mr, vr = m0[c], v0[c]
        acc = 0; accp = 0.0
        for s in range(sweeps):
            draws = P.worldline_draws(seed, chain0 + c, sweep0 + s, N, mode, 1)
            st = {}; dS_ref = np.zeros((N, N)); mask_ref = np.zeros((N, N), dtype=bool)
            mr, vr = WL.checkerboard_step_dense(mr, vr, kappa, W, draws, mode, stats=st, dS_out=dS_ref, accept_mask=mask_ref)
            acc += st['accepted']; accp += st['acceptance']
        assert (m[c].cpu().numpy() == mr).all() and (v[c].cpu().numpy() == vr).all()
        assert (dS[c].cpu().numpy() == dS_ref).all()               # strict arithmetic: bitwise
        assert (mask[c].cpu().numpy().astype(bool) == mask_ref).all()
        assert rec[c, WOBS_ACCEPTED] == acc
        assert rec[c, WOBS_ACCEPTANCE] == pytest.approx(accp, rel=1e-5)      # fast W=1 kernel: fp32 statistic
        assert rec[c, WOBS_DELTA_M_ABS] in (0, -1) and WL.valid(mr)
        assert ops.worldline_observables(m, v, W=W)[c, WOBS_DELTA_M_ABS].item() == 0
        f = WL.links(mr, vr, W)
        assert rec[c, WOBS_SUM_F2] == pytest.approx(float((f ** 2).sum()), rel=1e-12)
        assert rec[c, WOBS_SUM_DF2] == pytest.approx(float((lat.d1(f) ** 2).sum()), rel=1e-12, abs=1e-12)
        assert rec[c, WOBS_WRAP0] == mr[0].sum() and rec[c, WOBS_WRAP1] == mr[1].sum()


def test_observables_and_action_match_reference(golden_worldline_observables):
    for c in golden_worldline_observables:
        N, kappa, W = int(c['N']), float(c['kappa']), int(c['W'])
        S = svb.Worldline(svb.Lattice2D(N), kappa, W=W)
        assert S(c['m'], c['v']) == pytest.approx(float(c['action']), rel=1e-12)
        assert S.valid({'m': c['m']})
        from supervillain_b200.generator.worldline import worldline_inline_values
        vals = worldline_inline_values(S.observables(c['m'], c['v']).cpu().numpy()[0], N, kappa)
        for name in ('ActionDensity', 'InternalEnergyDensity', 'InternalEnergyDensitySquared', 'WindingSquared'):
            assert vals[name] == pytest.approx(float(c[name]), rel=1e-11, abs=1e-12), name
        assert (vals['TorusWrapping'] == c['TorusWrapping']).all()


def test_action_raises_when_constraint_is_violated():
    S = svb.Worldline(svb.Lattice2D(6), 0.5)
    m, v = WL.hot_start(np.random.default_rng(0), 6)
    bad = m.copy(); bad[0, 1, 1] += 1
    with pytest.raises(ValueError):
        S(bad, v)
    assert not S.valid({'m': bad})


def test_full_size_config3_shard_properties():
    """BASELINE config 3 per-GPU shard (L=64, 1024 chains): smem path == global path bit for bit,
    delta m = 0 on every chain after every sweep (test/test_validity.py), wrapping sector conserved."""
    N, chains, kappa = 64, 1024, 0.5
    S = svb.Worldline(svb.Lattice2D(N), kappa)
    E = svb.BatchedEnsemble(S, chains)
    m0, v0 = E._start('hot', 7)
    a_m, a_v, b_m, b_v = m0.clone(), v0.clone(), m0.clone(), v0.clone()
    oa = torch.zeros((chains, WOBS_COUNT), dtype=torch.float64, device='cuda'); ob = torch.zeros_like(oa)
    ops.worldline_sweep(a_m, a_v, kappa, n_sweeps=3, seed=1, path='smem', obs=oa)
    ops.worldline_sweep(b_m, b_v, kappa, n_sweeps=3, seed=1, path='global', obs=ob)
    assert torch.equal(a_m, b_m) and torch.equal(a_v, b_v)
    assert torch.equal(oa[:, WOBS_ACCEPTED], ob[:, WOBS_ACCEPTED])
    assert (ops.worldline_observables(a_m, a_v)[:, WOBS_DELTA_M_ABS] == 0).all()
    w0 = ops.worldline_observables(m0, v0)
    assert torch.equal(w0[:, WOBS_WRAP0], oa[:, WOBS_WRAP0]) and torch.equal(w0[:, WOBS_WRAP1], oa[:, WOBS_WRAP1])
    rate = oa[:, WOBS_ACCEPTED].sum().item() / (3 * chains * N * N)
    assert 0.05 < rate < 0.9


def test_plaquette_update_generator_protocol():
    S = svb.Worldline(svb.Lattice2D(8), 0.5)
    G = PlaquetteUpdate(S, seed=3)
    E = svb.Ensemble(S).generate(30, G, 'cold')
    assert np.asarray(E.m).shape == (30, 2, 8, 8) and np.asarray(E.v).shape == (30, 1, 8, 8)
    assert all(S.valid(E.configuration[k]) for k in range(30))
    touched = np.zeros((1, 8, 8), dtype=bool)
    for k in range(1, 30):
        touched |= (E.v[k] != E.v[k - 1])
    assert touched.all()                                     # test/test_plaquette_update.py:33-41
    assert 'single-plaquette proposals accepted' in G.report()
    with pytest.raises(ValueError):
        PlaquetteUpdate(svb.Villain(svb.Lattice2D(8), 0.5))


@pytest.mark.parametrize('mode', ['joint', 'vortex', 'coexact'])
def test_full_size_config3_shard_bit_exact_against_c_oracle(mode):
    """BASELINE config 3, one GPU's shard at full size (L=64, 1024 chains, 2 sweeps): m and v identical to the C oracle."""
    from oracle import c_oracle as C
    N, chains, kappa, W = 64, 1024, 0.5, 1
    m0, v0 = WL.hot_start(np.random.default_rng(64), N, chains)
    m, v = dev(m0, torch.int32), dev(v0, torch.int32)
    obs = torch.zeros((chains, WOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.worldline_sweep(m, v, kappa, W=W, mode=mode, n_sweeps=2, seed=8, chain0=3072, obs=obs)
    m_ref, v_ref, acc, accp = C.worldline_sweep_philox(m0, v0, kappa, W=W, mode=mode, n_sweeps=2, seed=8, chain0=3072)
    assert (m.cpu().numpy() == m_ref).all() and (v.cpu().numpy() == v_ref).all()
    rec = obs.cpu().numpy()
    assert (rec[:, WOBS_ACCEPTED] == acc).all()
    np.testing.assert_allclose(rec[:, WOBS_ACCEPTANCE], accp, rtol=1e-5)
    assert (ops.worldline_observables(m, v, W=W)[:, WOBS_DELTA_M_ABS] == 0).all().item()


@pytest.mark.parametrize('mode', ['joint', 'vortex', 'coexact'])
def test_L128_table_kernel_bit_exact_against_c_oracle(mode):
    """L = 128: the chain (192 KiB of m and v) still fits one SM, so the table kernel serves it, one CTA per SM -- m and v
    identical to the C oracle, also through the overlapped-launch entry point."""
    from oracle import c_oracle as C
    N, chains, kappa = 128, 170, 0.6
    m0, v0 = WL.hot_start(np.random.default_rng(128), N, chains)
    m, v = dev(m0, torch.int32), dev(v0, torch.int32)
    obs = torch.zeros((chains, WOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.worldline_sweep(m, v, kappa, mode=mode, n_sweeps=2, seed=8, chain0=11, obs=obs)
    m_ref, v_ref, acc, accp = C.worldline_sweep_philox(m0, v0, kappa, W=1, mode=mode, n_sweeps=2, seed=8, chain0=11)
    assert (m.cpu().numpy() == m_ref).all() and (v.cpu().numpy() == v_ref).all()
    rec = obs.cpu().numpy()
    assert (rec[:, WOBS_ACCEPTED] == acc).all()
    np.testing.assert_allclose(rec[:, WOBS_ACCEPTANCE], accp, rtol=1e-5)
    m2, v2 = dev(m0, torch.int32), dev(v0, torch.int32)
    ov = ops.WorldlineOverlappedSweeps(m2, v2, kappa, mode=mode, seed=8, chain0=11)
    obs2 = torch.zeros_like(obs)
    ov.step(0, 1, obs2)
    ov.step(1, 1, obs2)
    torch.cuda.synchronize()
    assert torch.equal(m2, m) and torch.equal(v2, v)


def test_wrapping_update_reproduces_reference_chain(golden_worldline_wrapping):
    """WrappingUpdate with rng=default_rng(99): identical m and accept counts to the reference (wrapping.py:43-90)."""
    from supervillain_b200.generator.worldline import WrappingUpdate
    for c in golden_worldline_wrapping:
        N, W, kappa, I = int(c['N']), int(c['W']), float(c['kappa']), int(c['interval'])
        S = svb.Worldline(svb.Lattice2D(N), kappa, W=W)
        G = WrappingUpdate(S, I)
        G.rng = np.random.default_rng(99)
        cfg = {'m': c['m0'], 'v': c['v0']}
        for s in range(int(c['sweeps'])):
            before = G.accepted
            cfg = G.step(cfg)
            assert (np.asarray(cfg['m']) == c['m'][s]).all(), (N, W, s)
            assert G.accepted - before == int(c['accepted'][s])
            assert S.valid(cfg)
    assert 'single-wrapping proposals accepted' in G.report()


def test_wrapping_philox_matches_dense_oracle_and_changes_the_sector():
    N, chains, kappa = 8, 64, 0.3
    m0, v0 = WL.hot_start(np.random.default_rng(5), N, chains)
    m, v = dev(m0, torch.int32), dev(v0, torch.int32)
    dS = torch.zeros((chains, 2, N), dtype=torch.float64, device='cuda')
    cnt = torch.zeros((chains, 2), dtype=torch.float64, device='cuda')
    ops.worldline_wrapping(m, v, kappa, seed=3, sweep=1, chain0=10, counters=cnt, dS_out=dS)
    got = m.cpu().numpy()
    changed = 0
    for c in range(chains):
        # rebuild the kernel's draws: stream 3, counter word 0 = mu * N + k
        x, y, z, w = P.philox_site(3, 10 + c, 1, np.arange(2 * N, dtype=np.uint64), 3)
        ku = (x << np.uint64(12)) | (y >> np.uint64(20))
        u = ((ku.astype(np.float64) + 0.5) * 2.0 ** -44).reshape(2, N)
        idx = ((w * np.uint64(2)) >> np.uint64(32)).astype(np.int64)
        cm = np.where(idx < 1, idx - 1, idx).reshape(2, N)
        st = {}; dS_ref = np.zeros((2, N))
        m_ref, _ = WL.wrapping_step_dense(m0[c], v0[c], kappa, 1, {'u': u, 'cm': cm}, stats=st, dS_out=dS_ref)
        assert (got[c] == m_ref).all()
        np.testing.assert_allclose(dS[c].cpu().numpy(), dS_ref, rtol=1e-13, atol=1e-13)
        assert cnt[c, 0].item() == st['accepted']
        assert WL.valid(m_ref)
        changed += int((WL.torus_wrapping(m_ref) != WL.torus_wrapping(m0[c])).any())
    assert changed > 0


def test_sequentially_plaquette_and_wrapping_in_a_batched_ensemble():
    """The ergodic pairing of test/end-to-end.py:48-50, device resident, many chains."""
    from supervillain_b200.generator.combining import Sequentially
    from supervillain_b200.generator.worldline import WrappingUpdate
    S = svb.Worldline(svb.Lattice2D(8), 0.5)
    G = Sequentially((PlaquetteUpdate(S, seed=1), WrappingUpdate(S, seed=2)))
    E = svb.BatchedEnsemble(S, 128).generate(50, G, 'cold', sweeps_per_step=4)
    m, v = E.fields
    assert (ops.worldline_observables(m, v)[:, WOBS_DELTA_M_ABS] == 0).all().item()
    assert E.ActionDensity.shape == (128, 50) and np.isfinite(E.ActionDensity).all()
    assert (E.TorusWrapping != 0).any()                      # wrapping sectors are visited
    ref = WL.action_density(m.cpu().numpy().astype(np.int64), v.cpu().numpy().astype(np.int64), 0.5, 1)
    np.testing.assert_allclose(E.ActionDensity[:, -1], ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('mode', ['joint', 'vortex', 'coexact'])
@pytest.mark.parametrize('N,chains', [(64, 700), (32, 1500), (16, 2500)])
def test_overlapped_launches_equal_ordinary_launches(N, chains, mode):
    """svb_worldline_sweep_overlapped: K overlapped steps leave exactly the fields and records of K ordinary launches."""
    kappa, K = 0.5, 6
    S = svb.Worldline(svb.Lattice2D(N), kappa)
    m, v = svb.BatchedEnsemble(S, chains)._start('hot', 4)
    rm, rv = m.clone(), v.clone()
    rec_ref = torch.zeros((K, chains, WOBS_COUNT), dtype=torch.float64, device='cuda')
    for k in range(K):
        ops.worldline_sweep(rm, rv, kappa, mode=mode, interval=2, seed=13, sweep0=2 * k, n_sweeps=1 + (k % 2), obs=rec_ref[k])
    rec = torch.zeros_like(rec_ref)
    ov = ops.WorldlineOverlappedSweeps(m, v, kappa, mode=mode, interval=2, seed=13)
    for k in range(K):
        ov.step(2 * k, 1 + (k % 2), obs=rec[k])
    torch.cuda.synchronize()
    assert torch.equal(m, rm) and torch.equal(v, rv)
    assert torch.equal(rec, rec_ref)


@pytest.mark.parametrize('N', [16, 32, 64])
def test_staged_observable_kernel_equals_the_generic_one(N):
    """svb_worldline_observables: the TMA-staged integer kernel (W = 1, >= 64 chains) against the generic fp64 kernel (the
    same call on slices of < 64 chains), on random fields that also violate delta m = 0."""
    g = torch.Generator(device='cuda'); g.manual_seed(N)
    chains = 150
    m = torch.randint(-9, 10, (chains, 2, N, N), generator=g, device='cuda', dtype=torch.int32)
    v = torch.randint(-9, 10, (chains, 1, N, N), generator=g, device='cuda', dtype=torch.int32)
    fast = ops.worldline_observables(m, v)
    slow = torch.cat([ops.worldline_observables(m[lo:lo + 50].contiguous(), v[lo:lo + 50].contiguous()) for lo in range(0, chains, 50)])
    assert torch.equal(fast, slow)
    assert (fast[:, WOBS_DELTA_M_ABS] > 0).all()
