"""Continuation and checkpoint/resume of batched ensembles: the generator state is the Philox (seed, sweep counter), so an
interrupted run continues as the run that was never interrupted (the batched form of Ensemble.continue_from,
supervillain/ensemble.py:103-142)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb                                                                   # noqa: E402
from supervillain_b200._lib import VOBS_ACTION                                                    # noqa: E402

pytestmark = pytest.mark.gpu


def _make(kind, N):
    if kind == 'villain':
        S = svb.Villain(svb.Lattice2D(N), 0.5)
        return S, svb.generator.villain.NeighborhoodUpdate(S, seed=13)
    S = svb.Worldline(svb.Lattice2D(N), 0.5)
    return S, svb.generator.worldline.PlaquetteUpdate(S, seed=13)


@pytest.mark.parametrize('kind,N', [('villain', 32), ('villain', 12), ('worldline', 16)])
def test_save_load_continue_equals_the_uninterrupted_run(kind, N, tmp_path):
    chains = 37
    S, G = _make(kind, N)
    whole = svb.BatchedEnsemble(S, chains, chain0=5).generate(10, G, 'hot', start_seed=3, sweeps_per_step=2)
    S2, G2 = _make(kind, N)
    first = svb.BatchedEnsemble(S2, chains, chain0=5).generate(6, G2, 'hot', start_seed=3, sweeps_per_step=2)
    path = tmp_path / 'checkpoint.npz'
    first.save(path)
    del first, G2
    resumed = svb.BatchedEnsemble.continue_from(path, 4)
    assert torch.equal(resumed.fields[0], whole.fields[0]) and torch.equal(resumed.fields[1], whole.fields[1])
    assert (resumed.index == whole.index[6:]).all() and resumed.generator.counter == G.counter == 20
    # records: bit for bit, except that the state columns of a run's LAST sample come from a separate kernel (another
    # summation order of the action) where an uninterrupted run fuses them into the next launch
    np.testing.assert_array_equal(resumed.record, whole.record[:, 6:])
    loaded = svb.BatchedEnsemble.load(path)
    np.testing.assert_array_equal(loaded.record[:, :5], whole.record[:, :5])
    np.testing.assert_allclose(loaded.record[:, 5], whole.record[:, 5], rtol=1e-13)
    rest = np.delete(np.arange(whole.record.shape[-1]), VOBS_ACTION)
    np.testing.assert_array_equal(loaded.record[:, 5][:, rest], whole.record[:, 5][:, rest])
    np.testing.assert_allclose(loaded.ActionDensity, whole.ActionDensity[:, :6], rtol=1e-13)
    # continuing in memory is the same thing
    S3, G3 = _make(kind, N)
    a = svb.BatchedEnsemble(S3, chains, chain0=5).generate(6, G3, 'hot', start_seed=3, sweeps_per_step=2)
    b = svb.BatchedEnsemble.continue_from(a, 4)
    assert torch.equal(b.fields[0], whole.fields[0]) and torch.equal(b.fields[1], whole.fields[1])
    np.testing.assert_array_equal(b.record, resumed.record)
    with pytest.raises(ValueError):
        svb.BatchedEnsemble.continue_from(svb.BatchedEnsemble(S3, chains), 2)


def test_kept_configurations_line_up_with_their_inline_columns(tmp_path):
    """What `BatchedEnsemble.to_reference` relies on: draw t of the kept configurations is the state after step
    keep_every * (t + 1), and row keep_every * (t + 1) - 1 of every inline column describes exactly that state;
    `keep_every` survives a checkpoint."""
    N, chains, keep = 16, 9, 3
    S = svb.Villain(svb.Lattice2D(N), 0.6)
    G = svb.generator.villain.NeighborhoodUpdate(S, seed=4)
    E = svb.BatchedEnsemble(S, chains).generate(12, G, 'hot', start_seed=1, sweeps_per_step=2, keep_every=keep)
    assert E.keep_every == keep and E.configuration['phi'].shape == (chains, 4, 1, N, N)
    assert E.configuration['n'].dtype == np.int64
    kept = keep * (1 + np.arange(4)) - 1
    for t, row in enumerate(kept):
        phi = torch.from_numpy(E.configuration['phi'][:, t]).cuda()
        n = torch.from_numpy(E.configuration['n'][:, t]).to(device='cuda', dtype=torch.int32)
        obs = svb.ops.villain_observables(phi, n, 0.6).cpu().numpy()
        np.testing.assert_allclose(E.ActionDensity[:, row], obs[:, VOBS_ACTION] / (N * N), rtol=1e-12)
        np.testing.assert_array_equal(E.record[:, row, 1:4], obs[:, 1:4])
    assert torch.equal(torch.from_numpy(E.configuration['phi'][:, -1]).cuda(), E.fields[0])
    path = tmp_path / 'kept.npz'
    E.save(path)
    L = svb.BatchedEnsemble.load(path)
    assert L.keep_every == keep
    np.testing.assert_array_equal(L.configuration['n'], E.configuration['n'])
