"""Parity of the CUDA Villain path (through the C ABI) with the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import lattice_np as lat
from oracle import philox_np as P
from oracle import villain_np as V

pytestmark = pytest.mark.gpu

import supervillain_b200 as svb                      # noqa: E402
from supervillain_b200 import ops                    # noqa: E402
from supervillain_b200._lib import (VOBS_ACCEPTANCE, VOBS_ACCEPTED, VOBS_ACTION, VOBS_COUNT, VOBS_SUM_DN2,  # noqa: E402
                                    VOBS_WRAP0, VOBS_WRAP1)
from supervillain_b200.generator.villain import NeighborhoodUpdate   # noqa: E402


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize('path', ['smem', 'global'])
def test_injected_rng_reproduces_reference_chain(golden_villain_neighborhood, path):
    """Level 1: the reference's own numpy draws injected -> n, phi and accept counts identical to the
    reference's NeighborhoodUpdate.step, sweep after sweep (golden generated with rng=default_rng(99))."""
    for c in golden_villain_neighborhood:
        N, kappa, W = int(c['N']), float(c['kappa']), int(c['W'])
        S = svb.Villain(svb.Lattice2D(N), kappa, W=W)
        G = NeighborhoodUpdate(S, path=path)
        G.rng = np.random.default_rng(99)
        cfg = {'phi': c['phi0'], 'n': c['n0']}
        for s in range(int(c['sweeps'])):
            before = (G.accepted, G.acceptance)
            cfg = G.step(cfg)
            assert np.asarray(cfg['n']).dtype == np.int64 and np.asarray(cfg['phi']).dtype == np.float64
            assert (np.asarray(cfg['n']) == c['n'][s]).all(), (N, kappa, W, s)
            assert (np.asarray(cfg['phi']) == c['phi'][s]).all(), (N, kappa, W, s)      # bitwise
            assert G.accepted - before[0] == int(c['accepted'][s])
            assert G.acceptance - before[1] == pytest.approx(float(c['acceptance'][s]), rel=1e-12)
            assert float(S(cfg['phi'], cfg['n'])) == pytest.approx(float(c['action'][s]), rel=1e-12)


def test_dS_and_accept_mask_match_oracle(golden_villain_neighborhood):
    """Level 2: dS of every proposal within 1e-12 (relative, fp64) of the oracle, same accept mask."""
    for c in golden_villain_neighborhood[:9]:
        N, kappa, W = int(c['N']), float(c['kappa']), int(c['W'])
        draws = {k: c[k][0] for k in ('u', 'dphi', 'dn_fwd', 'dn_bwd')}
        dS_ref = np.zeros((N, N)); mask_ref = np.zeros((N, N), dtype=bool)
        V.neighborhood_step_dense(c['phi0'], c['n0'], kappa, draws, accept_mask=mask_ref, dS_out=dS_ref)
        phi, n = dev(c['phi0'][None]), dev(c['n0'][None], torch.int32)
        inj = {'u': dev(draws['u'][None, None]), 'dphi': dev(draws['dphi'][None, None]),
               'dn_fwd': dev(draws['dn_fwd'][None, None], torch.int32), 'dn_bwd': dev(draws['dn_bwd'][None, None], torch.int32)}
        mask = torch.zeros((1, N, N), dtype=torch.uint8, device='cuda')
        dS = torch.zeros((1, N, N), dtype=torch.float64, device='cuda')
        ops.villain_sweep(phi, n, kappa, W=W, injected=inj, accept_mask=mask, dS_out=dS)
        assert (mask.cpu().numpy()[0].astype(bool) == mask_ref).all()
        # the kernel recomputes r from the current fields where the reference carries it incrementally
        # (neighborhood.py:129): identical up to a few ulp of r, i.e. ~1e-15 of the terms of dS
        np.testing.assert_allclose(dS.cpu().numpy()[0], dS_ref, rtol=1e-12, atol=1e-12)


def oracle_philox_chain(phi, n, kappa, W, seed, chain, sweep0, sweeps, interval_phi=np.pi, interval_n=1):
    stats_all = []
    for s in range(sweeps):
        draws = P.villain_draws(seed, chain, sweep0 + s, phi.shape[-1], W=W, interval_phi=interval_phi, interval_n=interval_n)
        st = {}
        phi, n = V.neighborhood_step_dense(phi, n, kappa, draws, stats=st)
        stats_all.append(st)
    return phi, n, stats_all


def test_philox_draw_mapping_matches_oracle():
    for (N, W, I, seed, sweep, chain0) in [(8, 1, 1, 1, 0, 0), (5, 2, 1, 2**40 + 17, 3, 5), (16, 1, 3, 99, 2**33 + 1, 2**32 + 3)]:
        u, dphi, dn = ops.villain_draws(3, N, W=W, interval_n=I, seed=seed, sweep=sweep, chain0=chain0)
        for c in range(3):
            ref = P.villain_draws(seed, chain0 + c, sweep, N, W=W, interval_n=I)
            assert (u[c].cpu().numpy() == ref['u']).all()
            assert (dphi[c].cpu().numpy() == ref['dphi']).all()
            got = dn[c].cpu().numpy()
            assert (got[0] == ref['dn_fwd'][0]).all() and (got[1] == ref['dn_bwd'][0]).all()
            assert (got[2] == ref['dn_fwd'][1]).all() and (got[3] == ref['dn_bwd'][1]).all()


@pytest.mark.parametrize('path', ['smem', 'global'])
@pytest.mark.parametrize('arith', ['strict', 'fast'])
@pytest.mark.parametrize('N,W,kappa', [(4, 1, 0.5), (5, 1, 0.3), (8, 2, 0.2), (7, 1, 0.1), (16, 1, 0.05), (32, 1, 0.5)])
def test_philox_mode_matches_oracle_replay(path, arith, N, W, kappa):
    """The production RNG path: regenerate the kernel's Philox draws with the oracle's independent
    Philox, run the restated reference algorithm on them, demand identical fields."""
    chains, sweeps, seed, chain0, sweep0 = 3, 3, 20260101, 11, 5
    phi0, n0 = V.hot_start(np.random.default_rng(N), N, chains)
    n0 = n0 * W
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(phi, n, kappa, W=W, n_sweeps=sweeps, seed=seed, sweep0=sweep0, chain0=chain0,
                      arithmetic=arith, path=path, obs=obs)
    rec = obs.cpu().numpy()
    for c in range(chains):
        p_ref, n_ref, st = oracle_philox_chain(phi0[c], n0[c], kappa, W, seed, chain0 + c, sweep0, sweeps)
        assert (n[c].cpu().numpy() == n_ref).all()
        assert (phi[c].cpu().numpy() == p_ref).all()
        assert rec[c, VOBS_ACCEPTED] == sum(s['accepted'] for s in st)
        assert rec[c, VOBS_ACCEPTANCE] == pytest.approx(sum(s['acceptance'] for s in st), rel=1e-12 if arith == 'strict' else 1e-5)
        assert rec[c, VOBS_ACTION] == pytest.approx(float(V.action(p_ref, n_ref, kappa)), rel=1e-12)
        assert rec[c, VOBS_SUM_DN2] == float((lat.d1(n_ref) ** 2).sum())
        assert (rec[c, [VOBS_WRAP0, VOBS_WRAP1]] == V.torus_wrapping(n_ref)).all()


def test_fused_sweeps_equal_repeated_single_sweeps():
    N, chains, kappa = 32, 64, 0.5
    phi0, n0 = V.hot_start(np.random.default_rng(1), N, chains)
    a_phi, a_n = dev(phi0), dev(n0, torch.int32)
    b_phi, b_n = dev(phi0), dev(n0, torch.int32)
    ops.villain_sweep(a_phi, a_n, kappa, n_sweeps=6, seed=7, sweep0=100)
    for s in range(6):
        ops.villain_sweep(b_phi, b_n, kappa, n_sweeps=1, seed=7, sweep0=100 + s, path='global' if s % 2 else 'smem')
    assert torch.equal(a_n, b_n) and torch.equal(a_phi, b_phi)


def test_chain_offset_makes_shards_independent_of_partition():
    """Chains [0,8) in one call == chains [0,4) and [4,8) in two calls with chain0 offsets (multi-GPU sharding)."""
    N, kappa = 16, 0.4
    phi0, n0 = V.hot_start(np.random.default_rng(2), N, 8)
    phi, n = dev(phi0), dev(n0, torch.int32)
    ops.villain_sweep(phi, n, kappa, n_sweeps=3, seed=5)
    for lo in (0, 4):
        p, q = dev(phi0[lo:lo + 4]), dev(n0[lo:lo + 4], torch.int32)
        ops.villain_sweep(p, q, kappa, n_sweeps=3, seed=5, chain0=lo)
        assert torch.equal(p, phi[lo:lo + 4]) and torch.equal(q, n[lo:lo + 4])


def test_full_size_config2_smem_equals_global_and_is_sane():
    """BASELINE config 2 (L=32, 4096 chains): the shared-memory path and the per-colour global path are
    independent code paths that must agree bit for bit; dn stays a multiple of W; acceptance ~0.5 %."""
    N, chains, kappa = 32, 4096, 0.5
    g = torch.Generator(device='cuda'); g.manual_seed(3)
    phi0 = (torch.rand((chains, 1, N, N), generator=g, device='cuda', dtype=torch.float64) * 2 - 1) * np.pi
    n0 = torch.randint(-2, 3, (chains, 2, N, N), generator=g, device='cuda', dtype=torch.int32)
    a_phi, a_n, b_phi, b_n = phi0.clone(), n0.clone(), phi0.clone(), n0.clone()
    oa = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda'); ob = torch.zeros_like(oa)
    ops.villain_sweep(a_phi, a_n, kappa, n_sweeps=4, seed=11, path='smem', obs=oa)
    ops.villain_sweep(b_phi, b_n, kappa, n_sweeps=4, seed=11, path='global', obs=ob)
    assert torch.equal(a_n, b_n) and torch.equal(a_phi, b_phi)
    assert torch.equal(oa[:, VOBS_ACCEPTED], ob[:, VOBS_ACCEPTED])
    torch.testing.assert_close(oa, ob, rtol=1e-5, atol=1e-9)
    changed = (a_n != n0).any().item() and (a_phi != phi0).any().item()
    assert changed
    # every changed phi comes from an accepted proposal (a site may be accepted more than once in 4 sweeps)
    assert int((a_phi != phi0).sum().item()) <= int(oa[:, VOBS_ACCEPTED].sum().item())


def test_cold_start_acceptance_and_observables_are_consistent():
    N, chains, kappa = 32, 256, 0.5
    S = svb.Villain(svb.Lattice2D(N), kappa)
    G = NeighborhoodUpdate(S, seed=1)
    E = svb.BatchedEnsemble(S, chains).generate(20, G, 'cold', sweeps_per_step=5)
    assert E.ActionDensity.shape == (chains, 20)
    rate = G.accepted / G.proposed
    assert 0.002 < rate < 0.02                     # ~0.4-0.5 % at kappa = 0.5
    phi, n = E.fields
    rec = ops.villain_observables(phi, n, kappa).cpu().numpy()
    ref = V.action(phi.cpu().numpy(), n.cpu().numpy().astype(np.int64), kappa)
    np.testing.assert_allclose(rec[:, VOBS_ACTION], ref, rtol=1e-12)
    np.testing.assert_allclose(E.ActionDensity[:, -1], ref / N**2, rtol=1e-12)


def test_fp32_mode_tracks_fp64_within_stated_tolerance():
    """fp32 phi: same Philox proposals; dS differs at float precision (stated tolerance: 2e-5 relative to
    the scale of the terms), so nearly all decisions agree after one sweep."""
    N, chains, kappa = 32, 32, 0.5
    phi0, n0 = V.hot_start(np.random.default_rng(5), N, chains)
    p64, n64 = dev(phi0), dev(n0, torch.int32)
    p32, n32 = dev(phi0).to(torch.float32), dev(n0, torch.int32)
    d64 = torch.zeros((chains, N, N), dtype=torch.float64, device='cuda'); d32 = torch.zeros_like(d64)
    ops.villain_sweep(p64, n64, kappa, seed=3, dS_out=d64)
    ops.villain_sweep(p32, n32, kappa, seed=3, dS_out=d32)
    scale = 200.0     # |terms| of dS are O(kappa/2 * (3 pi)^2 * 4) ~ 10^2
    assert (d64 - d32).abs().max().item() < 2e-5 * scale
    assert (n64 != n32).float().mean().item() < 1e-3


def test_argument_validation_maps_to_python_exceptions():
    phi = torch.zeros((2, 1, 8, 8), dtype=torch.float64, device='cuda')
    n = torch.zeros((2, 2, 8, 8), dtype=torch.int32, device='cuda')
    with pytest.raises(ValueError):
        ops.villain_sweep(phi, n, -1.0)
    with pytest.raises(ValueError):
        ops.villain_sweep(phi, n, 0.5, W=float('inf'))
    with pytest.raises(TypeError):
        ops.villain_sweep(phi, n.to(torch.int64), 0.5)
    with pytest.raises(ValueError):
        ops.villain_sweep(phi.cpu(), n, 0.5)
    with pytest.raises(ValueError):
        NeighborhoodUpdate(svb.Worldline(svb.Lattice2D(8), 0.5))


def test_full_size_config2_bit_exact_against_c_oracle():
    """BASELINE config 2 at full size (L=32, 4096 chains, 3 sweeps): every phi and n of every chain identical
    to the C oracle replaying the same Philox draws through the reference algorithm; counters too."""
    from oracle import c_oracle as C
    N, chains, kappa, sweeps, seed = 32, 4096, 0.5, 3, 20260101
    phi0, n0 = V.hot_start(np.random.default_rng(42), N, chains)
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(phi, n, kappa, n_sweeps=sweeps, seed=seed, sweep0=7, chain0=100, obs=obs)
    p_ref, n_ref, acc, accp = C.villain_sweep_philox(phi0, n0, kappa, n_sweeps=sweeps, seed=seed, sweep0=7, chain0=100)
    assert (n.cpu().numpy() == n_ref).all()
    assert (phi.cpu().numpy() == p_ref).all()
    rec = obs.cpu().numpy()
    assert (rec[:, VOBS_ACCEPTED] == acc).all()
    np.testing.assert_allclose(rec[:, VOBS_ACCEPTANCE], accp, rtol=1e-5)      # FAST arithmetic: fp32 statistic
    np.testing.assert_allclose(rec[:, VOBS_ACTION], V.action(p_ref, n_ref, kappa), rtol=1e-12)
    assert (rec[:, VOBS_SUM_DN2] == (lat.d1(n_ref) ** 2).sum(axis=(-3, -2, -1))).all()
    assert (rec[:, [VOBS_WRAP0, VOBS_WRAP1]] == V.torus_wrapping(n_ref)).all()


@pytest.mark.parametrize('arith', ['fast', 'strict'])
@pytest.mark.parametrize('sweeps', [1, 2])
@pytest.mark.parametrize('N,chains', [(16, 512), (32, 256), (64, 128), (128, 8), (48, 16)])
def test_other_shapes_bit_exact_against_c_oracle(N, chains, sweeps, arith):
    """Every kernel behind SVB_PATH_AUTO: the pipelined recompute kernel (1 sweep, fast) and the resident-residual kernel
    (fused sweeps, or strict) at the compile-time geometries 16 / 32 / 64, the global path (128), a generic even size (48)."""
    from oracle import c_oracle as C
    kappa, seed = 0.7, 5
    phi0, n0 = V.hot_start(np.random.default_rng(N), N, chains)
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(phi, n, kappa, n_sweeps=sweeps, seed=seed, obs=obs, arithmetic=arith)
    p_ref, n_ref, acc, accp = C.villain_sweep_philox(phi0, n0, kappa, n_sweeps=sweeps, seed=seed)
    assert (n.cpu().numpy() == n_ref).all() and (phi.cpu().numpy() == p_ref).all()
    rec = obs.cpu().numpy()
    assert (rec[:, VOBS_ACCEPTED] == acc).all()
    np.testing.assert_allclose(rec[:, VOBS_ACTION], V.action(p_ref, n_ref, kappa), rtol=1e-12)
    assert (rec[:, VOBS_SUM_DN2] == (lat.d1(n_ref) ** 2).sum(axis=(-3, -2, -1))).all()


def test_kappa_scan_per_chain_couplings():
    """BASELINE config 4 in miniature: per-chain kappa (a scan across the BKT region) in one launch."""
    from oracle import c_oracle as C
    N, per, kappas = 16, 4, np.linspace(0.3, 1.2, 8)
    chains = per * len(kappas)
    kc = np.repeat(kappas, per)
    phi0, n0 = V.hot_start(np.random.default_rng(9), N, chains)
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(phi, n, 1.0, n_sweeps=2, seed=3, kappa_chain=dev(kc), obs=obs)
    for c in range(chains):
        p_ref, n_ref, _, _ = C.villain_sweep_philox(phi0[c:c + 1], n0[c:c + 1], kc[c], n_sweeps=2, seed=3, chain0=c)
        assert (n[c].cpu().numpy() == n_ref[0]).all() and (phi[c].cpu().numpy() == p_ref[0]).all()
        assert obs[c, VOBS_ACTION].item() == pytest.approx(float(V.action(p_ref[0], n_ref[0], kc[c])), rel=1e-12)


def test_exp_clipped_agrees_with_libm_through_acceptance():
    """The kernel's own min(1, e^-dS) (degree-11 polynomial) against libm: sum of acceptance probabilities of a
    sweep with all proposals rejected (u = 1) equals the oracle's to 1e-13 relative on a wide spread of dS."""
    N, chains = 32, 8
    phi0, n0 = V.hot_start(np.random.default_rng(11), N, chains)
    for kappa in (0.01, 0.1, 1.0, 5.0):
        u = np.ones((1, chains, N, N))
        rng = np.random.default_rng(3)
        dphi = rng.uniform(-np.pi, np.pi, (1, chains, N, N))
        dnf = rng.integers(-1, 2, (1, chains, 2, N, N)); dnb = rng.integers(-1, 2, (1, chains, 2, N, N))
        phi, n = dev(phi0), dev(n0, torch.int32)
        obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
        ops.villain_sweep(phi, n, kappa, injected={'u': dev(u), 'dphi': dev(dphi), 'dn_fwd': dev(dnf, torch.int32),
                                                    'dn_bwd': dev(dnb, torch.int32)}, obs=obs)
        for c in range(chains):
            st = {}
            V.neighborhood_step_dense(phi0[c], n0[c], kappa, {'u': u[0, c], 'dphi': dphi[0, c], 'dn_fwd': dnf[0, c], 'dn_bwd': dnb[0, c]}, stats=st)
            assert obs[c, VOBS_ACCEPTANCE].item() == pytest.approx(st['acceptance'], rel=1e-13)
            assert obs[c, VOBS_ACCEPTED].item() == 0


def test_host_stepper_equals_resident_path():
    """The host-buffer API (svb_villain_sweep_host: chunked H2D -> sweep -> D2H over several streams) gives exactly the
    fields and observables of the device-resident call, for any chunking."""
    from supervillain_b200.hostpath import HostStepper
    N, chains, kappa = 32, 96, 0.5
    S = svb.Villain(svb.Lattice2D(N), kappa)
    phi0, n0 = V.hot_start(np.random.default_rng(8), N, chains)
    ref_phi, ref_n = dev(phi0), dev(n0, torch.int32)
    ref_obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(ref_phi, ref_n, kappa, n_sweeps=2, seed=77, sweep0=0, chain0=5, obs=ref_obs)
    ops.villain_sweep(ref_phi, ref_n, kappa, n_sweeps=2, seed=77, sweep0=2, chain0=5, obs=ref_obs)
    for chunks, streams in ((1, 1), (5, 2), (16, 4)):
        G = NeighborhoodUpdate(S, seed=77)
        st = HostStepper(G, chains, chain0=5, chunks=chunks, streams=streams)
        a, b = st.pinned_fields()
        a.copy_(torch.from_numpy(phi0)); b.copy_(torch.from_numpy(n0).to(torch.int32))
        st.step(a, b, n_sweeps=2)
        rec = st.step(a, b, n_sweeps=2)
        assert torch.equal(a, ref_phi.cpu()) and torch.equal(b, ref_n.cpu())
        assert torch.equal(rec, ref_obs.cpu())
        assert G.counter == 4


@pytest.mark.parametrize('arith', ['fast', 'strict'])
@pytest.mark.parametrize('N,chains,sweeps', [(32, 8, 1), (64, 6, 2), (128, 4, 3), (256, 2, 2)])
def test_tiled_path_equals_oracle_and_global_path(N, chains, sweeps, arith):
    """The single-pass tiled kernel (ghost zones + ping-pong, configs 4 and 5): identical to the C oracle and to the
    per-colour global path, independent of the tiling (N = 32 is a single tile whose ghost zone wraps onto itself)."""
    from oracle import c_oracle as C
    kappa, seed = 0.6, 31
    phi0, n0 = V.hot_start(np.random.default_rng(N + 1), N, chains)
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    mask = torch.zeros((chains, N, N), dtype=torch.uint8, device='cuda')
    ops.villain_sweep(phi, n, kappa, n_sweeps=sweeps, seed=seed, sweep0=3, chain0=7, obs=obs, arithmetic=arith, path='tiled',
                      accept_mask=mask)
    p_ref, n_ref, acc, accp = C.villain_sweep_philox(phi0, n0, kappa, n_sweeps=sweeps, seed=seed, sweep0=3, chain0=7)
    assert (n.cpu().numpy() == n_ref).all() and (phi.cpu().numpy() == p_ref).all()
    rec = obs.cpu().numpy()
    assert (rec[:, VOBS_ACCEPTED] == acc).all()
    np.testing.assert_allclose(rec[:, VOBS_ACCEPTANCE], accp, rtol=1e-12 if arith == 'strict' else 1e-5)
    np.testing.assert_allclose(rec[:, VOBS_ACTION], V.action(p_ref, n_ref, kappa), rtol=1e-12)
    assert (rec[:, VOBS_SUM_DN2] == (lat.d1(n_ref) ** 2).sum(axis=(-3, -2, -1))).all()
    g_phi, g_n = dev(phi0), dev(n0, torch.int32)
    g_mask = torch.zeros_like(mask)
    ops.villain_sweep(g_phi, g_n, kappa, n_sweeps=sweeps, seed=seed, sweep0=3, chain0=7, arithmetic=arith, path='global',
                      accept_mask=g_mask)
    assert torch.equal(g_phi, phi) and torch.equal(g_n, n) and torch.equal(g_mask, mask)


@pytest.mark.parametrize('chains,sweeps,interval_n,W', [(3, 1, 1, 1), (5, 3, 1, 1), (700, 2, 1, 1), (4, 2, 3, 2)])
def test_cluster_kernel_equals_oracle(chains, sweeps, interval_n, W):
    """L = 128 (config 4) through svb_villain_sweep: one chain per cluster of four CTAs, a 32-row strip each, strip
    boundaries through distributed shared memory (svb_villain_cluster.cuh).  Identical to the C oracle -- fields,
    accepted counts, records -- for fewer chains than clusters, more chains than clusters (every cluster loops), fused
    sweeps and general proposal widths; also with the records of the arriving state (obs_in)."""
    from oracle import c_oracle as C
    N, kappa, seed = 128, 0.55, 77
    phi0, n0 = V.hot_start(np.random.default_rng(chains), N, chains)
    kc = np.linspace(0.3, 1.1, chains)
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(phi, n, kappa, n_sweeps=sweeps, seed=seed, sweep0=5, chain0=2, obs=obs, kappa_chain=dev(kc), W=W,
                      interval_n=interval_n, interval_phi=2.5)
    per_chain = [C.villain_sweep_philox(phi0[i:i + 1], n0[i:i + 1], kc[i], n_sweeps=sweeps, seed=seed, sweep0=5, chain0=2 + i, W=W,
                                        interval_n=interval_n, interval_phi=2.5) for i in range(chains)]
    p_ref, n_ref, acc, accp = (np.concatenate([r[j] for r in per_chain]) for j in range(4))
    assert (n.cpu().numpy() == n_ref).all() and (phi.cpu().numpy() == p_ref).all()
    rec = obs.cpu().numpy()
    assert (rec[:, VOBS_ACCEPTED] == acc).all()
    np.testing.assert_allclose(rec[:, VOBS_ACCEPTANCE], accp, rtol=1e-5)
    np.testing.assert_allclose(rec[:, VOBS_ACTION], V.action(p_ref, n_ref, 1.0) * kc, rtol=1e-12)
    assert (rec[:, VOBS_SUM_DN2] == (lat.d1(n_ref) ** 2).sum(axis=(-3, -2, -1))).all()
    assert (rec[:, VOBS_WRAP0] == n_ref[:, 0].sum(axis=(-2, -1))).all() and (rec[:, VOBS_WRAP1] == n_ref[:, 1].sum(axis=(-2, -1))).all()
    # identical to the tiled path
    t_phi, t_n = dev(phi0), dev(n0, torch.int32)
    ops.villain_sweep(t_phi, t_n, kappa, n_sweeps=sweeps, seed=seed, sweep0=5, chain0=2, kappa_chain=dev(kc), W=W,
                      interval_n=interval_n, interval_phi=2.5, path='tiled')
    assert torch.equal(t_phi, phi) and torch.equal(t_n, n)


@pytest.mark.parametrize('N,chains', [(32, 2000), (16, 3000), (64, 300), (128, 150)])
def test_overlapped_launches_equal_ordinary_launches(N, chains):
    """svb_villain_sweep_overlapped: K steps whose launches overlap (per-chain epochs order the data) leave exactly the
    fields and per-step records of K ordinary launches; two chain sets interleaved on one stream stay independent."""
    kappa, K = 0.5, 7
    S = svb.Villain(svb.Lattice2D(N), kappa)
    sets = [svb.BatchedEnsemble(S, chains)._start('hot', 5 + r) for r in range(2)]
    ref = [(phi.clone(), n.clone()) for phi, n in sets]
    rec_ref = torch.zeros((2, K, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    for r, (phi, n) in enumerate(ref):
        for k in range(K):
            ops.villain_sweep(phi, n, kappa, seed=11 + r, sweep0=3 * k, n_sweeps=1 + (k % 2), obs=rec_ref[r, k])
    rec = torch.zeros_like(rec_ref)
    steppers = [ops.VillainOverlappedSweeps(phi, n, kappa, seed=11 + r) for r, (phi, n) in enumerate(sets)]
    for k in range(K):
        for r in range(2):
            steppers[r].step(3 * k, 1 + (k % 2), obs=rec[r, k])
    torch.cuda.synchronize()
    for r in range(2):
        assert torch.equal(sets[r][0], ref[r][0]) and torch.equal(sets[r][1], ref[r][1])
    assert torch.equal(rec, rec_ref)
    # a foreign write between steps needs fence(): the next launch then waits for the whole stream
    phi, n = sets[0]
    phi.add_(0.125); ref[0][0].add_(0.125)
    steppers[0].fence()
    steppers[0].step(100, 1)
    steppers[0].step(101, 1)
    ops.villain_sweep(ref[0][0], ref[0][1], kappa, seed=11, sweep0=100, n_sweeps=2)
    torch.cuda.synchronize()
    assert torch.equal(phi, ref[0][0]) and torch.equal(n, ref[0][1])
    assert int(steppers[0].epochs.min()) == steppers[0].epoch == K + 2


@pytest.mark.parametrize('N,chains', [(16, 1500), (32, 2500), (64, 400), (128, 90)])
def test_sparse_launches_equal_dense_launches(N, chains):
    """A single-sweep launch that owes no record of the state it leaves applies its accepted proposals to global memory
    as reductions and stores nothing back (SPARSE); a launch that must record the state it leaves, and a
    launch of several fused sweeps, store the chain densely.  Same fields, same acceptance counters, sweep after sweep."""
    kappa, K = 0.5, 5
    S = svb.Villain(svb.Lattice2D(N), kappa)
    phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 21)
    ophi, on = phi.clone(), n.clone()
    dphi, dn = phi.clone(), n.clone()
    fphi, fn = phi.clone(), n.clone()
    rec_in = torch.zeros((K + 1, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    rec = torch.zeros((K, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    stepper = ops.VillainOverlappedSweeps(ophi, on, kappa, seed=4)
    for k in range(K):
        ops.villain_sweep(phi, n, kappa, seed=4, sweep0=k, n_sweeps=1)                       # sparse: no record at all
        stepper.step(k, 1, obs=rec_in[k + 1], obs_in=rec_in[k])                              # sparse: record of the arriving state
        ops.villain_sweep(dphi, dn, kappa, seed=4, sweep0=k, n_sweeps=1, obs=rec[k])         # dense: record of the state it leaves
        assert torch.equal(phi, dphi) and torch.equal(n, dn)
        assert torch.equal(ophi, dphi) and torch.equal(on, dn)
    ops.villain_sweep(fphi, fn, kappa, seed=4, sweep0=0, n_sweeps=K)                         # dense: fused sweeps
    torch.cuda.synchronize()
    assert torch.equal(fphi, dphi) and torch.equal(fn, dn)
    assert torch.equal(rec_in[1:, :, 4], rec[:, :, 4])                                        # SVB_VOBS_ACCEPTED
    assert torch.allclose(rec_in[1:, :, 5], rec[:, :, 5], rtol=1e-6, atol=0)                 # SVB_VOBS_ACCEPTANCE: an fp32 monitor
    assert torch.equal(rec_in[1:K, :, 1:4], rec[:K - 1, :, 1:4])                              # integer state columns arrive one launch late
    assert torch.allclose(rec_in[1:K, :, 0], rec[:K - 1, :, 0], rtol=1e-13, atol=0)         # the action (summation order)


@pytest.mark.parametrize('chains,W,interval_n', [(150, 1, 1), (37, 2, 1), (301, 1, 0)])
def test_strips_kernel_equals_cluster_kernel(chains, W, interval_n, monkeypatch):
    """L = 128: single sparse sweeps run with one chain per CTA, phi and n streamed through a ring of strips
    (villain_strips_kernel); SVB_VILLAIN_KERNEL128=cluster sends the same launches to the cluster kernel.  Same fields, same
    records, launch after launch -- ordinary and overlapped launches, kappa per chain, more chains than CTAs."""
    N, kappa, K = 128, 0.6, 4
    S = svb.Villain(svb.Lattice2D(N), kappa, W=W)
    kc = torch.linspace(0.3, 1.2, chains, dtype=torch.float64, device='cuda')
    phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 9)
    n *= W
    results = {}
    for kernel in ('strips', 'cluster'):
        monkeypatch.setenv('SVB_VILLAIN_KERNEL128', kernel)
        a_phi, a_n, b_phi, b_n = phi.clone(), n.clone(), phi.clone(), n.clone()
        rec = torch.zeros((K + 1, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
        plain = torch.zeros((K, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
        st = ops.VillainOverlappedSweeps(a_phi, a_n, kappa, W=W, interval_n=interval_n, seed=6, chain0=3, kappa_chain=kc)
        for k in range(K):
            st.step(k, 1, obs=rec[k + 1], obs_in=rec[k])
            ops.villain_sweep(b_phi, b_n, kappa, W=W, interval_n=interval_n, seed=6, sweep0=k, chain0=3, kappa_chain=kc)
        torch.cuda.synchronize()
        assert torch.equal(a_phi, b_phi) and torch.equal(a_n, b_n), kernel
        results[kernel] = (a_phi, a_n, rec)
    s_phi, s_n, s_rec = results['strips']
    c_phi, c_n, c_rec = results['cluster']
    assert torch.equal(s_phi, c_phi) and torch.equal(s_n, c_n)
    assert not torch.equal(s_phi, phi)
    assert torch.equal(s_rec[:, :, 1:5], c_rec[:, :, 1:5])                              # sum dn^2, wrapping, accepted: integers
    assert torch.allclose(s_rec[:, :, 0], c_rec[:, :, 0], rtol=1e-13, atol=0)           # the action (summation order)
    assert torch.allclose(s_rec[:, :, 5], c_rec[:, :, 5], rtol=1e-6, atol=0)            # the fp32 acceptance monitor


def test_overlapped_launches_reject_what_they_do_not_serve():
    S = svb.Villain(svb.Lattice2D(48), 0.5)
    phi, n = svb.BatchedEnsemble(S, 4)._start('cold', 0)
    with pytest.raises(NotImplementedError):
        ops.VillainOverlappedSweeps(phi, n, 0.5)


def test_observables_of_the_arriving_state_complete_the_previous_record():
    """obs_in: the state columns computed while the residuals are built equal, bit for bit, what the separate pass of
    the previous launch would have written; BatchedEnsemble.generate uses it and reproduces the ordinary records."""
    N, chains, K, kappa = 32, 1500, 6, 0.5
    S = svb.Villain(svb.Lattice2D(N), kappa)
    phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 3)
    rphi, rn = phi.clone(), n.clone()
    rec_ref = torch.zeros((K, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    for k in range(K):
        ops.villain_sweep(rphi, rn, kappa, seed=21, sweep0=k, obs=rec_ref[k])
    rec = torch.full((K, chains, VOBS_COUNT), -7.0, dtype=torch.float64, device='cuda')
    scratch = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ov = ops.VillainOverlappedSweeps(phi, n, kappa, seed=21)
    for k in range(K):
        ov.step(k, 1, obs=rec[k], obs_in=rec[k - 1] if k else scratch)
    rec[K - 1, :, :4] = ops.villain_observables(phi, n, kappa)[:, :4]
    torch.cuda.synchronize()
    assert torch.equal(phi, rphi) and torch.equal(n, rn)
    # the integer columns and the counters bit for bit; the action is the same sum of squares in another order of
    # summation where the two records come from different passes (1e-13: far inside north_star's 1e-12)
    assert torch.equal(rec[:K - 1, :, 1:], rec_ref[:K - 1, :, 1:])
    torch.testing.assert_close(rec[:K - 1, :, 0], rec_ref[:K - 1, :, 0], rtol=1e-13, atol=0)
    assert torch.equal(rec[K - 1, :, 1:], rec_ref[K - 1, :, 1:])              # integers and this launch's counters
    # the last state's action comes from svb_villain_observables (another summation order): 1e-13, not bitwise
    torch.testing.assert_close(rec[K - 1, :, 0], rec_ref[K - 1, :, 0], rtol=1e-13, atol=0)
    # the ensemble driver: same records as ordinary launches record by record
    G1 = NeighborhoodUpdate(S, seed=5)
    E1 = svb.BatchedEnsemble(S, 64).generate(5, G1, start='hot', start_seed=9, sweeps_per_step=2)
    G2 = NeighborhoodUpdate(S, seed=5, path='smem')        # path != 'auto': ordinary launches
    E2 = svb.BatchedEnsemble(S, 64).generate(5, G2, start='hot', start_seed=9, sweeps_per_step=2)
    assert np.array_equal(E1.record[:, :-1], E2.record[:, :-1]) and np.array_equal(E1.record[:, -1, 1:], E2.record[:, -1, 1:])
    np.testing.assert_allclose(E1.record[:, -1, 0], E2.record[:, -1, 0], rtol=1e-13)
    assert G1.accepted == G2.accepted and G1.acceptance == G2.acceptance


def test_cluster_kernel_records_of_the_arriving_state():
    """obs_in at L = 128 (the cluster kernel): each strip's share of the sums rides along with the residual build and
    rank 0 gathers them -- bit for bit the records of ordinary launches."""
    N, chains, K, kappa = 128, 160, 5, 0.7
    S = svb.Villain(svb.Lattice2D(N), kappa)
    phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 3)
    rphi, rn = phi.clone(), n.clone()
    rec_ref = torch.zeros((K, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    for k in range(K):
        ops.villain_sweep(rphi, rn, kappa, seed=21, sweep0=2 * k, n_sweeps=2, obs=rec_ref[k])
    rec = torch.full((K, chains, VOBS_COUNT), -7.0, dtype=torch.float64, device='cuda')
    scratch = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ov = ops.VillainOverlappedSweeps(phi, n, kappa, seed=21)
    for k in range(K):
        ov.step(2 * k, 2, obs=rec[k], obs_in=rec[k - 1] if k else scratch)
    torch.cuda.synchronize()
    assert torch.equal(phi, rphi) and torch.equal(n, rn)
    assert torch.equal(rec[:K - 1, :, 1:], rec_ref[:K - 1, :, 1:])
    torch.testing.assert_close(rec[:K - 1, :, 0], rec_ref[:K - 1, :, 0], rtol=1e-13, atol=0)
    assert torch.equal(rec[K - 1, :, 4:], rec_ref[K - 1, :, 4:])
    assert int(ov.epochs.min()) == ov.epoch == K


def test_host_stepper_two_steps_in_flight():
    """step_async: two steps in flight on alternating host buffer pairs give the records and fields of waited steps."""
    from supervillain_b200.hostpath import HostStepper
    N, chains, kappa = 32, 200, 0.5
    S = svb.Villain(svb.Lattice2D(N), kappa)
    starts = [V.hot_start(np.random.default_rng(20 + r), N, chains) for r in range(2)]
    G1, G2 = NeighborhoodUpdate(S, seed=3), NeighborhoodUpdate(S, seed=3)
    st1, st2 = HostStepper(G1, chains, chunks=7, streams=3), HostStepper(G2, chains, chunks=7, streams=3)
    sets1 = [st1.pinned_fields() for _ in range(2)]
    sets2 = [st2.pinned_fields() for _ in range(2)]
    for r in range(2):
        for sets in (sets1, sets2):
            sets[r][0].copy_(torch.from_numpy(starts[r][0])); sets[r][1].copy_(torch.from_numpy(starts[r][1]).to(torch.int32))
    waited = [st1.step(*sets1[k % 2]).clone() for k in range(6)]
    records, pending = [], None
    for k in range(6):
        issued = st2.step_async(*sets2[k % 2])
        if pending is not None:
            records.append(pending.wait().clone())
        pending = issued
    records.append(pending.wait().clone())
    for r in range(2):
        assert torch.equal(sets1[r][0], sets2[r][0]) and torch.equal(sets1[r][1], sets2[r][1])
    for a, b in zip(waited, records):
        assert torch.equal(a, b)


@pytest.mark.parametrize('N,chains,W,interval_n,interval_phi,kappa', [(32, 300, 2, 1, np.pi, 0.3), (32, 300, 1, 3, 1.0, 0.1), (16, 600, 3, 2, 2.0, 0.05),
                                                                       (64, 80, 2, 2, np.pi, 0.2), (128, 4, 2, 2, 1.5, 0.1),
                                                                       (16, 600, 3, 1, 2.0, 0.05), (64, 80, 2, 1, 1.0, 0.2), (128, 4, 2, 1, 1.5, 0.1),
                                                                       (32, 100, 1, 0, 2.0, 0.4)])
def test_general_proposals_bit_exact_against_c_oracle(N, chains, W, interval_n, interval_phi, kappa):
    """W > 1, wider dn intervals and narrower dphi intervals, smem and tiled: fields identical to the C oracle.  With
    interval_n <= 1 these are the fp32-filtered kernels in the run-time form of the constants config 2 has at compile time,
    also through the overlapped-launch entry point; interval_n >= 2 ("wide": the uniform's leading bits come from the
    refinement block, svb_villain.cu) is served by the generic kernels and refused by the overlapped entry point."""
    from oracle import c_oracle as C
    seed, sweeps = 17, 2
    phi0, n0 = V.hot_start(np.random.default_rng(N + W), N, chains)
    n0 = n0 * W
    kw = dict(W=W, interval_phi=interval_phi, interval_n=interval_n)
    p_ref, n_ref, acc, accp = C.villain_sweep_philox(phi0, n0, kappa, n_sweeps=sweeps, seed=seed, sweep0=4, chain0=9, **kw)
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(phi, n, kappa, n_sweeps=sweeps, seed=seed, sweep0=4, chain0=9, obs=obs, **kw)
    assert (n.cpu().numpy() == n_ref).all() and (phi.cpu().numpy() == p_ref).all()
    rec = obs.cpu().numpy()
    assert (rec[:, VOBS_ACCEPTED] == acc).all()
    np.testing.assert_allclose(rec[:, VOBS_ACCEPTANCE], accp, rtol=1e-5)
    np.testing.assert_allclose(rec[:, VOBS_ACTION], V.action(p_ref, n_ref, kappa), rtol=1e-12)
    assert acc.sum() > 0
    if N <= 64:
        phi, n = dev(phi0), dev(n0, torch.int32)
        if interval_n > 1:
            with pytest.raises(NotImplementedError):
                ops.VillainOverlappedSweeps(phi, n, kappa, seed=seed, chain0=9, **kw)
            return
        ov = ops.VillainOverlappedSweeps(phi, n, kappa, seed=seed, chain0=9, **kw)
        ov.step(4, 1)
        ov.step(5, 1)
        torch.cuda.synchronize()
        assert (n.cpu().numpy() == n_ref).all() and (phi.cpu().numpy() == p_ref).all()


@pytest.mark.parametrize('phi_scale,n_scale,kappa', [(50.0, 1, 0.5), (1.0, 12, 0.02), (300.0, 40, 0.001), (1.0, 1, 8.0)])
def test_filtered_decisions_stay_exact_far_from_equilibrium(phi_scale, n_scale, kappa):
    """The fp32 filter must never decide differently from fp64: large |phi| (a field that has wandered), large residuals
    (|r| ~ 250), tiny and large couplings.  The error band scales with max |r|; whatever falls inside it takes the exact
    path.  Fields identical to the C oracle."""
    from oracle import c_oracle as C
    N, chains, seed = 32, 400, 23
    phi0, n0 = V.hot_start(np.random.default_rng(int(phi_scale) + n_scale), N, chains)
    phi0, n0 = phi0 * phi_scale, n0 * n_scale
    p_ref, n_ref, acc, _ = C.villain_sweep_philox(phi0, n0, kappa, n_sweeps=3, seed=seed)
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ops.villain_sweep(phi, n, kappa, n_sweeps=3, seed=seed, obs=obs)
    assert (n.cpu().numpy() == n_ref).all() and (phi.cpu().numpy() == p_ref).all()
    assert (obs.cpu().numpy()[:, VOBS_ACCEPTED] == acc).all()


@pytest.mark.parametrize('N,path,sweeps_per_step', [(256, 'auto', 1), (256, 'auto', 2), (128, 'tiled', 3), (64, 'tiled', 1)])
def test_swapping_tiled_sweeps_equal_the_in_place_ones(N, path, sweeps_per_step):
    """svb_villain_sweep_tiled_swap (the state alternates between two buffer pairs, no copy back) through
    BatchedEnsemble.generate against svb_villain_sweep_tiled in place: fields and records bit for bit, for odd and even
    sweep counts, and a continued run picks the state up from whichever pair holds it."""
    import supervillain_b200 as svb
    from supervillain_b200 import ops
    from supervillain_b200._lib import VOBS_COUNT
    from supervillain_b200.generator.villain import NeighborhoodUpdate
    chains, steps, kappa = 3, 5, 0.6
    S = svb.Villain(svb.Lattice2D(N), kappa)
    G = NeighborhoodUpdate(S, seed=21)
    G.path = path
    assert callable(G.swapping_device(*svb.BatchedEnsemble(S, chains)._start('hot', 2)))
    E = svb.BatchedEnsemble(S, chains, chain0=7).generate(steps, G, 'hot', start_seed=2, sweeps_per_step=sweeps_per_step)
    phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 2)
    rec = torch.empty((steps, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    for k in range(steps):
        ops.villain_sweep(phi, n, kappa, n_sweeps=sweeps_per_step, seed=21, sweep0=k * sweeps_per_step, chain0=7, path='tiled',
                          obs=rec[k])
    assert torch.equal(E.fields[0], phi) and torch.equal(E.fields[1], n)
    want = rec.cpu().numpy().transpose(1, 0, 2)
    exact = [VOBS_SUM_DN2, VOBS_WRAP0, VOBS_WRAP1, VOBS_ACCEPTED]            # integer-valued columns
    np.testing.assert_array_equal(E.record[..., exact], want[..., exact])
    # the tiles of a chain add their fp64 partial sums atomically, in an order that varies from launch to launch
    np.testing.assert_allclose(E.record, want, rtol=1e-13)
    assert G.counter == steps * sweeps_per_step
    E.generate(2, G, 'continue', sweeps_per_step=sweeps_per_step)
    for k in range(steps, steps + 2):
        ops.villain_sweep(phi, n, kappa, n_sweeps=sweeps_per_step, seed=21, sweep0=k * sweeps_per_step, chain0=7, path='tiled')
    assert torch.equal(E.fields[0], phi) and torch.equal(E.fields[1], n)


@pytest.mark.parametrize('kind', ['villain', 'villain128', 'worldline'])
def test_overlapped_launch_protocol_soak(kind):
    """compute-sanitizer cannot run on this pool, so the epoch protocol of the overlapped launches is soaked instead: 2000
    back-to-back steps on each of TWO interleaved chain sets whose chain count (5000, 3000) is far above the grid (1184
    CTAs: every CTA loops over several chains, remainder chains included), with obs_in records in flight.  A lost or early
    epoch would load a chain before its previous store landed; the final fields and the accepted totals must equal those
    of 2000 ordinary (fully serialised) launches bit for bit."""
    K, kappa, seed = 2000, 0.5, 31
    if kind.startswith('villain'):
        # N = 128 is the strips kernel (one chain per CTA of a 148-CTA grid, the next chain's first strips requested while
        # the last passes of this one still run): 400 steps of 700 and 300 chains
        N = 128 if kind == 'villain128' else 32
        if N == 128:
            K = 400
        S = svb.Villain(svb.Lattice2D(N), kappa)
        make = lambda a, b, c0: ops.VillainOverlappedSweeps(a, b, kappa, seed=seed, chain0=c0)
        plain = lambda a, b, c0, k, obs: ops.villain_sweep(a, b, kappa, seed=seed, sweep0=k, chain0=c0, obs=obs)
        nobs = VOBS_COUNT
    else:
        N = 64
        S = svb.Worldline(svb.Lattice2D(N), kappa)
        make = lambda a, b, c0: ops.WorldlineOverlappedSweeps(a, b, kappa, seed=seed, chain0=c0)
        plain = lambda a, b, c0, k, obs: ops.worldline_sweep(a, b, kappa, seed=seed, sweep0=k, chain0=c0, obs=obs)
        nobs = 7
    counts = {'villain': (5000, 3000), 'villain128': (700, 300), 'worldline': (1500, 700)}[kind]
    sets = [svb.BatchedEnsemble(S, c)._start('hot', 50 + i) for i, c in enumerate(counts)]
    refs = [(a.clone(), b.clone()) for a, b in sets]
    steppers = [make(a, b, 1000 * i) for i, (a, b) in enumerate(sets)]
    # one record per step and set (no kernel of any other kind is enqueued between the overlapped launches: an ordinary
    # kernel in between would serialise them and the soak would test nothing)
    recs = [torch.zeros((K, c, nobs), dtype=torch.float64, device='cuda') for c in counts]
    scratch = [torch.zeros((c, nobs), dtype=torch.float64, device='cuda') for c in counts]
    for k in range(K):
        for i, st in enumerate(steppers):
            if kind.startswith('villain'):
                st.step(k, 1, obs=recs[i][k], obs_in=recs[i][k - 1] if k else scratch[i])
            else:
                st.step(k, 1, obs=recs[i][k])
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(refs):
        obs = torch.zeros((counts[i], nobs), dtype=torch.float64, device='cuda')
        total = torch.zeros((counts[i],), dtype=torch.float64, device='cuda')
        for k in range(K):
            plain(a, b, 1000 * i, k, obs)
            total += obs[:, 4]
        assert torch.equal(a, sets[i][0]) and torch.equal(b, sets[i][1]), (kind, i)
        assert torch.equal(total, recs[i][:, :, 4].sum(dim=0)) and float(total.sum()) > 0
        assert int(steppers[i].epochs.min()) == steppers[i].epoch == K


@pytest.mark.parametrize('N,chains,W,kappa,launches', [(256, 3, 1, 0.5, 'passes'), (128, 5, 2, 0.3, 'passes'), (144, 2, 1, 0.4, 'passes'),
                                                        (48, 7, 1, 0.6, 'passes'), (384, 1, 1, 0.8, 'passes'),
                                                        (256, 3, 1, 0.5, 'wavefront'), (128, 5, 2, 0.3, 'wavefront'),
                                                        (384, 1, 1, 0.8, 'wavefront'), (128, 37, 1, 0.7, 'wavefront')])
def test_inplace_colour_passes_equal_the_c_oracle(N, chains, W, kappa, launches):
    """svb_villain_sweep_inplace (svb_villain_stream.cuh): one launch per colour pass, in place, only accepted proposals written;
    svb_villain_sweep_wavefront (launches = 'wavefront'): the same phases in ONE launch, following one another through L2.
    N a multiple of 128 runs on TMA tensor-staged tiles (tiles on the edge of the lattice patch the periodic wrap), any other
    multiple of 16 straight from global memory.  Fields identical to the C oracle; the records of the ARRIVING state (obs_in:
    action and wrapping from the first colour pass, sum (dn)^2 from a pass over n) complete the previous step's record, the
    record of the state AFTER the sweeps (no obs_in) comes from one more read -- both equal to the restated reference."""
    from oracle import c_oracle as C
    seed, K = 29, 3
    phi0, n0 = V.hot_start(np.random.default_rng(N + chains), N, chains)
    n0 = n0 * W
    kc = np.linspace(kappa, kappa + 0.2, chains)
    refs = []
    p, q = phi0, n0
    for k in range(K):
        out = [C.villain_sweep_philox(p[i:i + 1], q[i:i + 1], kc[i], W=W, n_sweeps=1, seed=seed, sweep0=k, chain0=4 + i) for i in range(chains)]
        p, q = np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])
        refs.append((p, q, np.concatenate([o[2] for o in out]), np.concatenate([o[3] for o in out])))
    phi, n = dev(phi0), dev(n0, torch.int32)
    st = ops.VillainInplaceSweeps(phi, n, kappa, W=W, seed=seed, chain0=4, kappa_chain=dev(kc), launches=launches)
    assert st.launches == launches
    rec = torch.full((K, chains, VOBS_COUNT), -7.0, dtype=torch.float64, device='cuda')
    scratch = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    for k in range(K):
        st.step(k, 1, obs=rec[k], obs_in=rec[k - 1] if k else scratch)
    torch.cuda.synchronize()
    p_ref, n_ref = refs[-1][0], refs[-1][1]
    assert (n.cpu().numpy() == n_ref).all() and (phi.cpu().numpy() == p_ref).all()
    got = rec.cpu().numpy()
    for k in range(K):
        assert (got[k, :, VOBS_ACCEPTED] == refs[k][2]).all()
        np.testing.assert_allclose(got[k, :, VOBS_ACCEPTANCE], refs[k][3], rtol=1e-5)
        if k < K - 1:                                   # completed by the launch after
            pk, qk = refs[k][0], refs[k][1]
            np.testing.assert_allclose(got[k, :, VOBS_ACTION], V.action(pk, qk, 1.0) * kc, rtol=1e-12)
            assert (got[k, :, VOBS_SUM_DN2] == (lat.d1(qk) ** 2).sum(axis=(-3, -2, -1))).all()
            assert (got[k, :, VOBS_WRAP0] == qk[:, 0].sum(axis=(-2, -1))).all() and (got[k, :, VOBS_WRAP1] == qk[:, 1].sum(axis=(-2, -1))).all()
    # the full record of the state after the sweeps, two sweeps fused into one call
    phi2, n2 = dev(phi0), dev(n0, torch.int32)
    full = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    st2 = ops.VillainInplaceSweeps(phi2, n2, kappa, W=W, seed=seed, chain0=4, kappa_chain=dev(kc), launches=launches)
    st2.step(0, 2, obs=full)
    if launches == 'wavefront':
        assert int(st.workspace.abs().sum()) == 0 and int(st2.workspace.abs().sum()) == 0      # every call leaves its counters zero
    p2, q2 = refs[1][0], refs[1][1]
    assert (n2.cpu().numpy() == q2).all() and (phi2.cpu().numpy() == p2).all()
    f = full.cpu().numpy()
    np.testing.assert_allclose(f[:, VOBS_ACTION], V.action(p2, q2, 1.0) * kc, rtol=1e-12)
    assert (f[:, VOBS_SUM_DN2] == (lat.d1(q2) ** 2).sum(axis=(-3, -2, -1))).all()
    assert (f[:, VOBS_ACCEPTED] == refs[0][2] + refs[1][2]).all()
    # the ensemble driver steps big lattices this way
    if N == 256 and launches == 'passes':
        S = svb.Villain(svb.Lattice2D(N), kappa)
        E = svb.BatchedEnsemble(S, chains, chain0=4).generate(K, NeighborhoodUpdate(S, seed=seed), start={'phi': phi0, 'n': n0}, kappa_chain=kc)
        assert (E.fields[1].cpu().numpy() == n_ref).all() and (E.fields[0].cpu().numpy() == p_ref).all()
        np.testing.assert_allclose(E.record[:, :K - 1, VOBS_ACTION].T, got[:K - 1, :, VOBS_ACTION], rtol=1e-13)
        np.testing.assert_allclose(E.record[:, K - 1, VOBS_ACTION], V.action(p_ref, n_ref, 1.0) * kc, rtol=1e-12)


@pytest.mark.parametrize('N,chains,fused', [(128, 300, 1), (128, 64, 3), (512, 2, 2), (1024, 1, 1), (256, 9, 4)])
def test_wavefront_launches_equal_colour_pass_launches(N, chains, fused):
    """svb_villain_sweep_wavefront against svb_villain_sweep_inplace over a run of steps: fields identical after every step,
    integer records identical, the action to summation order; `fused` sweeps of a multi-sweep step share one launch (its
    colour phases chase one another through L2 as the two colours of one sweep do); hot and cold starts; kappa per chain."""
    kappa, K = 0.6, 4
    S = svb.Villain(svb.Lattice2D(N), kappa)
    kc = torch.linspace(0.3, 1.1, chains, dtype=torch.float64, device='cuda')
    for start in ('hot', 'cold'):
        phi, n = svb.BatchedEnsemble(S, chains)._start(start, 5)
        wphi, wn = phi.clone(), n.clone()
        a = ops.VillainInplaceSweeps(phi, n, kappa, seed=8, chain0=2, kappa_chain=kc, launches='passes')
        b = ops.VillainInplaceSweeps(wphi, wn, kappa, seed=8, chain0=2, kappa_chain=kc, launches='wavefront', fused_sweeps=fused)
        ra = torch.zeros((K + 1, chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
        rb = torch.zeros_like(ra)
        for k in range(K):
            a.step(fused * k, fused, obs=ra[k + 1], obs_in=ra[k])
            b.step(fused * k, fused, obs=rb[k + 1], obs_in=rb[k])
            assert torch.equal(phi, wphi) and torch.equal(n, wn), (start, k)
        assert torch.equal(ra[:, :, 1:5], rb[:, :, 1:5])                                  # sum dn^2, wrapping, accepted
        assert torch.allclose(ra[:, :, 0], rb[:, :, 0], rtol=1e-12, atol=0)
        assert torch.allclose(ra[:, :, 5], rb[:, :, 5], rtol=1e-5, atol=0)
        assert int(b.workspace.abs().sum()) == 0
        assert float(ra[1:, :, 4].sum()) > 0


def test_config5_full_size_bit_exact_against_c_oracle():
    """Config 5 at its named size: ONE L = 4096 lattice (16.7 M sites: 8192 tiles of 16 x 128 per colour pass, 32-bit offsets up
    to 2^24, every kind of edge tile), two sweeps through both entry points that serve it -- svb_villain_sweep_inplace and
    svb_villain_sweep_tiled_swap -- against the scalar C restatement of neighborhood.py:93-129 with the same Philox draws:
    phi and n identical, accepted counts identical, action to 1e-12."""
    from oracle import c_oracle as C
    N, kappa, seed = 4096, 0.5, 77
    rng = np.random.default_rng(5)
    phi0 = rng.uniform(-np.pi, np.pi, (1, 1, N, N))
    n0 = rng.integers(-2, 3, (1, 2, N, N))
    p_ref, n_ref, acc, accp = C.villain_sweep_philox(phi0, n0, kappa, n_sweeps=2, seed=seed, sweep0=3, chain0=1)
    assert acc[0] > 100000
    for launches in ('wavefront', 'passes'):
        phi, n = dev(phi0), dev(n0, torch.int32)
        obs = torch.zeros((1, VOBS_COUNT), dtype=torch.float64, device='cuda')
        st = ops.VillainInplaceSweeps(phi, n, kappa, seed=seed, chain0=1, launches=launches)
        assert st.launches == launches
        st.step(3, 2, obs=obs)
        assert torch.equal(n.cpu(), torch.from_numpy(n_ref).to(torch.int32)) and torch.equal(phi.cpu(), torch.from_numpy(p_ref)), launches
        rec = obs.cpu().numpy()
        assert rec[0, VOBS_ACCEPTED] == acc[0]
        np.testing.assert_allclose(rec[0, VOBS_ACCEPTANCE], accp[0], rtol=1e-5)
        np.testing.assert_allclose(rec[0, VOBS_ACTION], float(V.action(p_ref, n_ref, kappa)[0]), rtol=1e-12)
        assert rec[0, VOBS_SUM_DN2] == (lat.d1(n_ref) ** 2).sum()
        assert rec[0, VOBS_WRAP0] == n_ref[0, 0].sum() and rec[0, VOBS_WRAP1] == n_ref[0, 1].sum()
    # the record protocol of a running ensemble at this size: the arriving state's columns from the wavefront's own phases
    phi_w, n_w = dev(phi0), dev(n0, torch.int32)
    stw = ops.VillainInplaceSweeps(phi_w, n_w, kappa, seed=seed, chain0=1, launches='wavefront')
    recs = torch.zeros((3, 1, VOBS_COUNT), dtype=torch.float64, device='cuda')
    stw.step(3, 1, obs=recs[0], obs_in=recs[2])
    stw.step(4, 1, obs=recs[1], obs_in=recs[0])
    assert torch.equal(n_w, n) and torch.equal(phi_w, phi)
    r = recs.cpu().numpy()
    assert r[0, 0, VOBS_ACCEPTED] + r[1, 0, VOBS_ACCEPTED] == acc[0]
    assert r[2, 0, VOBS_SUM_DN2] == (lat.d1(n0) ** 2).sum() and r[2, 0, VOBS_WRAP0] == n0[0, 0].sum()
    np.testing.assert_allclose(r[2, 0, VOBS_ACTION], float(V.action(phi0, n0, kappa)[0]), rtol=1e-12)
    # the swapping entry point (one sweep per call, the state may change buffer pairs)
    sw = ops.VillainSwappingSweeps(dev(phi0), dev(n0, torch.int32), kappa, seed=seed, chain0=1)
    sw.step(3, 1)
    f_phi, f_n = sw.step(4, 1)
    assert torch.equal(f_n, n) and torch.equal(f_phi, phi)
