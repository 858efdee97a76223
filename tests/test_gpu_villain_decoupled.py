"""SiteUpdate / LinkUpdate / ExactUpdate on the GPU (svb_villain_decoupled) against the reference's golden chains and
the oracle."""
import numpy as np
import pytest
import torch

from oracle import philox_np as P
from oracle import villain_np as V

pytestmark = pytest.mark.gpu

import supervillain_b200 as svb                      # noqa: E402
from supervillain_b200 import ops                    # noqa: E402
from supervillain_b200._lib import VOBS_ACCEPTANCE, VOBS_ACCEPTED, VOBS_ACTION, VOBS_COUNT   # noqa: E402
from supervillain_b200.generator.villain import ExactUpdate, LinkUpdate, SiteUpdate          # noqa: E402

KINDS = ['site', 'link', 'exact']


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize('path', ['smem', 'global'])
def test_injected_rng_reproduces_reference_chains(golden_villain_decoupled, path):
    """Level 1: the reference's own numpy draws replayed and injected -> phi and n identical, bit for bit, to the
    reference's SiteUpdate / LinkUpdate / ExactUpdate chains, sweep after sweep (golden made with rng=default_rng(99))."""
    for c in golden_villain_decoupled:
        kind = KINDS[int(c['kind'])]
        N, kappa, W = int(c['N']), float(c['kappa']), int(c['W'])
        S = svb.Villain(svb.Lattice2D(N), kappa, W=W)
        G = {'site': lambda: SiteUpdate(S, interval_phi=float(c['interval']), path=path),
             'link': lambda: LinkUpdate(S, interval_n=int(c['interval']), path=path),
             'exact': lambda: ExactUpdate(S, interval_z=int(c['interval']), path=path)}[kind]()
        G.rng = np.random.default_rng(99)
        cfg = {'phi': c['phi0'], 'n': c['n0']}
        for s in range(int(c['sweeps'])):
            before = (G.accepted, G.acceptance)
            cfg = G.step(cfg)
            assert (np.asarray(cfg['n']) == c['n'][s]).all(), (kind, N, s)
            assert (np.asarray(cfg['phi']) == c['phi'][s]).all(), (kind, N, s)          # bitwise
            assert G.accepted - before[0] == int(c['accepted'][s])
            assert G.acceptance - before[1] == pytest.approx(float(c['acceptance'][s]), rel=1e-12)
        assert 'proposals accepted' in G.report()


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('N,W,kappa,interval', [(4, 1, 0.5, 1), (5, 2, 0.3, 2), (8, 1, 0.1, 1), (16, 3, 0.2, 1), (32, 1, 0.5, 2), (64, 2, 0.4, 3), (128, 1, 0.6, 1)])
def test_philox_mode_matches_oracle_replay(kind, N, W, kappa, interval):
    """Production RNG: the oracle regenerates the kernels' Philox draws and runs the restated reference algorithm."""
    chains, sweeps, seed = 3, 3, 77
    rng = np.random.default_rng(N)
    phi0 = np.stack([V.hot_start(rng, N)[0] for _ in range(chains)])
    n0 = np.stack([V.hot_start(rng, N)[1] * W for _ in range(chains)])
    phi, n = dev(phi0), dev(n0, torch.int32)
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    ipi = 1.25
    ops.villain_decoupled(kind, phi, n, kappa, W=W, interval_phi=ipi, interval=interval, n_sweeps=sweeps, seed=seed, sweep0=5,
                          chain0=2, obs=obs)
    for c in range(chains):
        p, q = phi0[c].copy(), n0[c].copy()
        acc, accp = 0, 0.0
        for s in range(sweeps):
            st = {}
            if kind == 'site':
                d = P.villain_draws(seed, 2 + c, 5 + s, N, W=W, interval_phi=ipi, interval_n=0, kind='site')
                p, q = V.neighborhood_step_dense(p, q, kappa, d, stats=st)
            elif kind == 'link':
                p, q = V.link_step_dense(p, q, kappa, P.villain_link_draws(seed, 2 + c, 5 + s, N, W=W, interval_n=interval), stats=st)
            else:
                p, q = V.exact_step_dense(p, q, kappa, P.villain_exact_draws(seed, 2 + c, 5 + s, N, interval_z=interval), stats=st)
            acc += st['accepted']; accp += st['acceptance']
        assert (n[c].cpu().numpy() == q).all(), (kind, N, c)
        assert (phi[c].cpu().numpy() == p).all(), (kind, N, c)
        assert int(obs[c, VOBS_ACCEPTED]) == acc
        # the filtered kernels sum the acceptance probabilities in fp32 (a diagnostic, as in the production sweep)
        assert float(obs[c, VOBS_ACCEPTANCE]) == pytest.approx(accp, rel=1e-5)
        assert float(obs[c, VOBS_ACTION]) == pytest.approx(float(V.action(p, q, kappa)), rel=1e-12)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('N,chains,kappa,W,interval', [(32, 1500, 0.5, 1, 1), (16, 2000, 2.5, 2, 3), (64, 200, 0.05, 1, 2), (128, 90, 0.4, 2, 2)])
def test_filtered_kernels_decide_like_the_strict_kernels(kind, N, chains, kappa, W, interval):
    """SiteUpdate, ExactUpdate and LinkUpdate with Philox draws run on fp32-filtered shared-memory kernels (L = 128: Site
    and Exact on the cluster kernel) whose cold path decides in STRICT arithmetic; a debug output (accept_mask / dS_out) sends the same call to the STRICT fp64 kernels.
    Millions of proposals, hot and cold regimes: identical fields and accepted counts, i.e. the filter never answers
    differently."""
    S = svb.Villain(svb.Lattice2D(N), kappa, W=W)
    phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 11)
    n *= W
    rphi, rn = phi.clone(), n.clone()
    obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
    robs = torch.zeros_like(obs)
    mask = torch.zeros((chains, N, N), dtype=torch.uint8, device='cuda')
    debug = dict(dS_out=torch.zeros((chains, 2, N, N), dtype=torch.float64, device='cuda')) if kind == 'link' else dict(accept_mask=mask)
    for s in range(4):
        ops.villain_decoupled(kind, phi, n, kappa, W=W, interval_phi=2.0, interval=interval, seed=5, sweep0=s, obs=obs)
        ops.villain_decoupled(kind, rphi, rn, kappa, W=W, interval_phi=2.0, interval=interval, seed=5, sweep0=s, obs=robs, **debug)
        assert torch.equal(phi, rphi) and torch.equal(n, rn)
        assert torch.equal(obs[:, VOBS_ACCEPTED], robs[:, VOBS_ACCEPTED])
        if kind != 'link':
            assert float(obs[:, VOBS_ACCEPTED].sum()) == float(mask.sum())
        torch.testing.assert_close(obs[:, VOBS_ACCEPTANCE], robs[:, VOBS_ACCEPTANCE], rtol=1e-5, atol=0)
        torch.testing.assert_close(obs[:, :4], robs[:, :4], rtol=1e-12, atol=0)
    assert float(obs[:, VOBS_ACCEPTED].sum()) > 0


def test_wrong_action_raises_like_the_reference():
    S = svb.Worldline(svb.Lattice2D(4), 0.5)
    for cls in (SiteUpdate, LinkUpdate, ExactUpdate):
        with pytest.raises(ValueError):
            cls(S)


def test_cohomology_update_reproduces_reference_chains(golden_villain_cohomology):
    """CohomologyUpdate on the GPU with the reference's draws injected: n identical to the reference chain."""
    from supervillain_b200.generator.villain import CohomologyUpdate
    for c in golden_villain_cohomology:
        N, kappa, W = int(c['N']), float(c['kappa']), int(c['W'])
        S = svb.Villain(svb.Lattice2D(N), kappa, W=W)
        G = CohomologyUpdate(S, interval_h=int(c['interval']))
        G.rng = np.random.default_rng(99)
        cfg = {'phi': c['phi0'], 'n': c['n0']}
        for s in range(int(c['sweeps'])):
            before = (G.accepted, G.acceptance)
            cfg = G.step(cfg)
            assert (np.asarray(cfg['n']) == c['n'][s]).all(), (N, s)
            assert G.accepted - before[0] == int(c['accepted'][s])
            assert G.acceptance - before[1] == pytest.approx(float(c['acceptance'][s]), rel=1e-12)
    assert sum(int(c['accepted'].sum()) for c in golden_villain_cohomology) > 10


def test_cohomology_philox_matches_oracle_and_changes_the_sector():
    N, kappa, chains, interval = 16, 0.01, 64, 2
    rng = np.random.default_rng(3)
    phi0 = np.stack([V.hot_start(rng, N)[0] for _ in range(chains)])
    n0 = np.stack([V.hot_start(rng, N)[1] for _ in range(chains)])
    phi, n = dev(phi0), dev(n0, torch.int32)
    counters = torch.zeros((chains, 2), dtype=torch.float64, device='cuda')
    dS = torch.zeros((chains, 2), dtype=torch.float64, device='cuda')
    for s in range(3):
        ops.villain_cohomology(phi, n, kappa, interval=interval, seed=9, sweep=s, chain0=4, counters=counters, dS_out=dS)
    changed = 0
    for c in range(chains):
        q, acc, dref = n0[c].copy(), 0, np.zeros(2)
        for s in range(3):
            st = {}
            _, q = V.cohomology_step(phi0[c], q, kappa, None, interval_h=interval, stats=st, dS_out=dref,
                                     draws=P.villain_cohomology_draws(9, 4 + c, s, interval))
            acc += st['accepted']
        assert (n[c].cpu().numpy() == q).all()
        assert int(counters[c, 0]) == acc
        np.testing.assert_allclose(dS[c].cpu().numpy(), dref, rtol=1e-12, atol=1e-12)
        changed += int((q.sum(axis=(1, 2)) != n0[c].sum(axis=(1, 2))).any())
    assert changed > 5          # the winding sector does move
