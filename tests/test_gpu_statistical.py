"""Level-3 correctness: ensemble observables of the GPU generators agree with the UNMODIFIED reference generators.

Anchors: tests/golden/statistical_anchors.json (means and binned errors produced by the reference's own generators,
tests/golden/make_statistical_anchors.py).  GPU errors come from the scatter of independent chains, so they are honest
about autocorrelation.  Tolerance: 5 combined standard errors (stated below)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import supervillain_b200 as svb                                                     # noqa: E402
from supervillain_b200.generator.combining import Sequentially                      # noqa: E402
from supervillain_b200.generator.villain import NeighborhoodUpdate                  # noqa: E402
from supervillain_b200.generator.worldline import PlaquetteUpdate, WrappingUpdate   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ANCHORS = json.load(open(os.path.join(HERE, 'golden', 'statistical_anchors.json')))
NSIGMA = 5.0


def anchor(action, generator, N, kappa):
    for a in ANCHORS:
        if (a['action'], a['generator'], a['N'], a['kappa']) == (action, generator, N, kappa):
            return a
    raise KeyError((action, generator, N, kappa))


def chain_statistics(E, name, discard):
    per_chain = getattr(E, name)[:, discard:].mean(axis=1)              # one number per independent chain
    return per_chain.mean(), per_chain.std(ddof=1) / np.sqrt(len(per_chain))


def check(E, ref, discard):
    for name in ('ActionDensity', 'WindingSquared'):
        mean, err = chain_statistics(E, name, discard)
        ref_mean, ref_err = ref[name]
        sigma = np.hypot(err, ref_err)
        assert abs(mean - ref_mean) < NSIGMA * sigma, (name, mean, err, ref_mean, ref_err)
        assert err < 0.01 * abs(ref_mean) + 1e-3                           # the GPU estimate is itself precise


@pytest.mark.parametrize('N,kappa', [(8, 0.3)])
def test_villain_neighborhood_matches_reference_hammer(N, kappa):
    """NeighborhoodUpdate (ergodic on its own for W=1) vs the reference's villain Hammer and, by duality
    (BASELINE.md section 5), the far more precise worldline Hammer."""
    S = svb.Villain(svb.Lattice2D(N), kappa)
    G = NeighborhoodUpdate(S, seed=424242)
    E = svb.BatchedEnsemble(S, 2048).generate(120, G, 'cold', sweeps_per_step=500)
    check(E, anchor('Villain', 'Hammer', N, kappa), discard=40)
    check(E, anchor('Worldline', 'Hammer', N, kappa), discard=40)
    assert 0.005 < G.accepted / G.proposed < 0.2


@pytest.mark.parametrize('N,kappa', [(8, 0.3), (8, 0.5), (5, 0.5)])
def test_worldline_plaquette_plus_wrapping_matches_reference_hammer(N, kappa):
    """Sequentially((PlaquetteUpdate, WrappingUpdate)) -- the ergodic pairing of test/end-to-end.py:48-50 -- vs the
    reference's worldline Hammer.  N=5 exercises the four-colour sweep."""
    S = svb.Worldline(svb.Lattice2D(N), kappa)
    G = Sequentially((PlaquetteUpdate(S, seed=11), WrappingUpdate(S, seed=12)))
    E = svb.BatchedEnsemble(S, 1024).generate(150, G, 'cold', sweeps_per_step=20)
    check(E, anchor('Worldline', 'Hammer', N, kappa), discard=50)


@pytest.mark.parametrize('kappa', [0.3, 0.5])
def test_worldline_plaquette_only_matches_reference_in_the_trivial_sector(kappa):
    """PlaquetteUpdate alone never leaves the wrapping sector of its start (plaquette.py:16-19).  From a cold start
    that is the sector the reference's Vortex+Coexact checkerboard pair samples too."""
    S = svb.Worldline(svb.Lattice2D(8), kappa)
    G = PlaquetteUpdate(S, seed=5)
    E = svb.BatchedEnsemble(S, 2048).generate(100, G, 'cold', sweeps_per_step=50)
    assert (E.TorusWrapping == 0).all()
    check(E, anchor('Worldline', 'Vortex+Coexact', 8, kappa), discard=30)
    if kappa == 0.5:
        mean, err = chain_statistics(E, 'ActionDensity', 30)
        ref = anchor('Worldline', 'PlaquetteUpdate', 8, 0.5)                # the reference's sequential PlaquetteUpdate
        assert abs(mean - ref['ActionDensity'][0]) < NSIGMA * np.hypot(err, ref['ActionDensity'][1])


def test_villain_decoupled_updates_match_reference_hammer():
    """Sequentially((SiteUpdate, LinkUpdate, ExactUpdate, CohomologyUpdate)) -- the reference's villain Hammer without the
    worm -- samples the same ensemble as NeighborhoodUpdate and as the reference's Hammers."""
    from supervillain_b200.generator.villain import CohomologyUpdate, ExactUpdate, LinkUpdate, SiteUpdate
    N, kappa = 8, 0.3
    S = svb.Villain(svb.Lattice2D(N), kappa)
    G = Sequentially((SiteUpdate(S, seed=1), LinkUpdate(S, seed=2), ExactUpdate(S, seed=3), CohomologyUpdate(S, seed=4)))
    E = svb.BatchedEnsemble(S, 1024).generate(120, G, 'cold', sweeps_per_step=40)
    check(E, anchor('Villain', 'Hammer', N, kappa), discard=40)
    check(E, anchor('Worldline', 'Hammer', N, kappa), discard=40)


def test_villain_neighborhood_with_wide_dn_proposals_matches_reference_hammer():
    """interval_n = 2 (K = 5, K^4 = 625 > 256: the "wide" draw mapping, where the uniform's leading bits come from the
    refinement block instead of the digits' remainder, svb_villain.cu) samples the same ensemble: the target distribution
    does not depend on the proposal width.  Acceptance is ~5 x lower than at interval_n = 1, hence the longer run."""
    N, kappa = 8, 0.3
    S = svb.Villain(svb.Lattice2D(N), kappa)
    G = NeighborhoodUpdate(S, interval_n=2, seed=777)
    E = svb.BatchedEnsemble(S, 2048).generate(100, G, 'cold', sweeps_per_step=2000)
    check(E, anchor('Villain', 'Hammer', N, kappa), discard=40)
    check(E, anchor('Worldline', 'Hammer', N, kappa), discard=40)


def test_generators_sharing_one_seed_still_sample_the_target():
    """Every generator kind draws from its own Philox stream pair (svb_common.cuh), so handing ONE seed to all members of a
    Sequentially -- whose sweep counters then advance in lockstep -- is safe: the composite chain samples the same ensemble
    as with hand-picked distinct seeds."""
    from supervillain_b200.generator.villain import CohomologyUpdate, ExactUpdate, LinkUpdate, SiteUpdate
    from supervillain_b200.generator.worldline import CoexactUpdate, VortexUpdate
    N, kappa, seed = 8, 0.3, 2026
    S = svb.Villain(svb.Lattice2D(N), kappa)
    G = Sequentially((NeighborhoodUpdate(S, seed=seed), SiteUpdate(S, seed=seed), LinkUpdate(S, seed=seed), ExactUpdate(S, seed=seed),
                      CohomologyUpdate(S, seed=seed)))
    E = svb.BatchedEnsemble(S, 1024).generate(120, G, 'cold', sweeps_per_step=40)
    check(E, anchor('Villain', 'Hammer', N, kappa), discard=40)
    check(E, anchor('Worldline', 'Hammer', N, kappa), discard=40)
    Wl = svb.Worldline(svb.Lattice2D(N), kappa)
    H = Sequentially((PlaquetteUpdate(Wl, seed=seed), VortexUpdate(Wl, seed=seed), CoexactUpdate(Wl, seed=seed), WrappingUpdate(Wl, seed=seed)))
    F = svb.BatchedEnsemble(Wl, 1024).generate(150, H, 'cold', sweeps_per_step=10)
    check(F, anchor('Worldline', 'Hammer', N, kappa), discard=50)
