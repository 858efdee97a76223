"""The UNMODIFIED reference package driving the GPU generators, on hardware.

`supervillain` is imported from oracle/_ref/ -- the copy of the reference's package tree that
`oracle/stage_reference.py` stages at build time and that ships to the GPU box with the snapshot
(never /root/reference, which does not exist there).  Nothing here restates the reference: its own
`Ensemble.generate` loop (supervillain/ensemble.py:74-98), its `Sequentially`, its `Batch` lossless-cast
guard (batch.py:206-227), its observables and its action's validity check do the driving and the checking.
"""
import numpy as np
import pytest

from oracle import refimport

pytestmark = pytest.mark.gpu

if not refimport.available():          # the staged copy is a build product: without it these tests cannot run at all
    pytest.skip('oracle/_ref is not staged (run __graft_entry__.build() where /root/reference is mounted)', allow_module_level=True)

sv = refimport.import_reference()

import supervillain_b200 as svb                                                    # noqa: E402
from supervillain_b200.generator.villain import NeighborhoodUpdate                # noqa: E402
from supervillain_b200.generator.worldline import PlaquetteUpdate, WrappingUpdate  # noqa: E402


def arr(batch):
    return np.asarray(sv.batch.Batch.as_array(batch))


def test_reference_ensemble_generates_with_the_gpu_neighborhood_update():
    """supervillain.Ensemble(S).generate(200, <GPU NeighborhoodUpdate>, 'cold') -- config 1 (L=5, kappa=0.5, one chain,
    test/end-to-end.py) -- stores float64 phi / int64 n Form columns, and the reference's own measurements of the stored
    configurations equal the inline columns the device reduced."""
    N, kappa, steps = 5, 0.5, 200
    S = sv.action.Villain(sv.lattice.Lattice2D(N), kappa, 1)
    G = NeighborhoodUpdate(S, seed=314159, inline=('ActionDensity', 'WindingSquared', 'TorusWrapping'))
    E = sv.Ensemble(S).generate(steps, G, start='cold')
    assert type(E).__module__.startswith('supervillain') and len(E) == steps
    phi, n = arr(E.configuration.fields['phi']), arr(E.configuration.fields['n'])
    assert phi.shape == (steps, 1, N, N) and phi.dtype == np.float64
    assert n.shape == (steps, 2, N, N) and n.dtype == np.int64
    assert type(E.configuration[3]['phi']).__name__ == 'Form' and E.configuration[3]['n'].degree == 1
    assert G.sweeps == steps and G.proposed == steps * N * N and 0 < G.accepted < G.proposed
    assert 'neighborhood proposals accepted' in G.report()
    # the chain moved, and it is the chain the batched path produces for the same (seed, chain 0)
    assert np.abs(phi[-1]).max() > 0
    B = svb.BatchedEnsemble(svb.Villain(svb.Lattice2D(N), kappa), 1).generate(
        steps, NeighborhoodUpdate(svb.Villain(svb.Lattice2D(N), kappa), seed=314159), start='cold', keep_every=1)
    assert (B.configuration['phi'][0] == phi).all() and (B.configuration['n'][0] == n).all()
    # inline columns short-circuit the reference's measurement (observable/observable.py:49-54); measured afresh from the
    # stored fields by the reference they agree to 1e-12 (north_star level 2)
    bare = sv.Ensemble(S).from_configurations(sv.configurations.Configurations(
        {k: E.configuration.fields[k] for k in ('phi', 'n')}))
    for name in ('ActionDensity', 'WindingSquared', 'TorusWrapping'):
        inline, measured = arr(getattr(E, name)), arr(getattr(bare, name))
        assert inline.shape == measured.shape, name
        assert np.allclose(inline, measured, rtol=1e-12, atol=1e-12), name


@pytest.mark.parametrize('N,W,kappa', [(5, 1, 0.5), (8, 2, 0.3), (6, 1, 0.7)])
def test_reference_ensemble_replays_its_own_chain_through_the_gpu(N, W, kappa):
    """Level 1 end to end, nothing restated: the reference's Ensemble loop run twice with rng = default_rng(99) -- once with
    the reference's NeighborhoodUpdate, once with the GPU one -- stores bit-identical phi and n and the same counters."""
    steps = 40
    S = sv.action.Villain(sv.lattice.Lattice2D(N), kappa, W)
    R = sv.generator.villain.NeighborhoodUpdate(S)
    R.rng = np.random.default_rng(99)
    G = NeighborhoodUpdate(S)
    G.rng = np.random.default_rng(99)
    rng = np.random.default_rng(N)
    hot = {'phi': S.Lattice.form(0), 'n': S.Lattice.form(1, dtype=int)}
    hot['phi'][...] = rng.uniform(-np.pi, np.pi, hot['phi'].shape)
    hot['n'][...] = rng.integers(-2, 3, hot['n'].shape)
    Er = sv.Ensemble(S).generate(steps, R, start={k: v.copy() for k, v in hot.items()})
    Eg = sv.Ensemble(S).generate(steps, G, start={k: v.copy() for k, v in hot.items()})
    for k in ('phi', 'n'):
        assert (arr(Er.configuration.fields[k]) == arr(Eg.configuration.fields[k])).all(), k
    assert int(R.accepted) == int(G.accepted) and int(R.proposed) == int(G.proposed)
    assert float(G.acceptance) == pytest.approx(float(R.acceptance), rel=1e-12)


def test_reference_sequentially_of_gpu_plaquette_and_wrapping_updates():
    """Sequentially((PlaquetteUpdate, WrappingUpdate)) -- the reference's combinator and Ensemble (combining.py:9-52,
    test/end-to-end.py:48-50) around the two GPU worldline generators: int64 m, v columns, delta m = 0 for every stored
    configuration by the reference's own check (worldline.py:54-70), its action evaluates on them, and the reference's
    observables on the stored fields equal the GPU generator's inline columns."""
    N, kappa, steps = 8, 0.5, 120
    S = sv.action.Worldline(sv.lattice.Lattice2D(N), kappa, 1)
    P = PlaquetteUpdate(S, seed=7, inline=('ActionDensity', 'WindingSquared'))
    H = WrappingUpdate(S, seed=8)
    G = sv.generator.combining.Sequentially((P, H))
    E = sv.Ensemble(S).generate(steps, G, start='cold')
    m, v = arr(E.configuration.fields['m']), arr(E.configuration.fields['v'])
    assert m.shape == (steps, 2, N, N) and m.dtype == np.int64 and v.shape == (steps, 1, N, N) and v.dtype == np.int64
    assert np.abs(m).max() > 0 and np.abs(v).max() > 0
    for t in (0, steps // 2, steps - 1):
        cfg = E.configuration[t]
        assert S.valid(cfg)
        assert np.isfinite(S(cfg['m'], cfg['v']))
    assert P.sweeps == steps and H.sweeps == steps
    assert 'PlaquetteUpdate' in str(G) and 'WrappingUpdate' in str(G)
    # the wrapping update runs AFTER the plaquette update within a step, so the stored fields are not the ones the inline
    # columns describe unless the wrapping proposal was rejected: compare on those draws (at kappa=0.5, L=8 most of them)
    bare = sv.Ensemble(S).from_configurations(sv.configurations.Configurations(
        {k: E.configuration.fields[k] for k in ('m', 'v')}))
    alone = sv.Ensemble(S).generate(60, PlaquetteUpdate(S, seed=9, inline=('ActionDensity', 'WindingSquared')), start='cold')
    fresh = sv.Ensemble(S).from_configurations(sv.configurations.Configurations(
        {k: alone.configuration.fields[k] for k in ('m', 'v')}))
    for name in ('ActionDensity', 'WindingSquared'):
        assert np.allclose(arr(getattr(alone, name)), arr(getattr(fresh, name)), rtol=1e-12, atol=1e-12), name
        assert arr(getattr(bare, name)).shape == (steps,)


def test_to_reference_round_trips_through_hdf5_in_the_reference_layout(tmp_path):
    """BatchedEnsemble.to_reference(chain).to_h5 -> supervillain.Ensemble.from_h5 (h5/strategy/batch.py:20-37,
    h5/extendable.py:34-42).  Runs wherever h5py is installed; this image has none, so here it skips."""
    h5py = pytest.importorskip('h5py')
    if getattr(h5py, '__svb_stub__', False):
        pytest.skip('only the h5py stub (oracle/stubs) is present: no HDF5 in this image')
    N, kappa, chains = 8, 0.5, 4
    S = svb.Villain(svb.Lattice2D(N), kappa)
    B = svb.BatchedEnsemble(S, chains).generate(20, NeighborhoodUpdate(S, seed=5), start='cold', sweeps_per_step=2, keep_every=5)
    R = B.to_reference(2, supervillain=sv)
    path = tmp_path / 'chain2.h5'
    with h5py.File(path, 'w') as f:
        R.to_h5(f.create_group('ensemble'))
    with h5py.File(path, 'r') as f:
        back = sv.Ensemble.from_h5(f['ensemble'])
    for k in ('phi', 'n'):
        assert (arr(back.configuration.fields[k]) == arr(R.configuration.fields[k])).all()
    assert (arr(back.index) == arr(R.index)).all() and back.index_stride == R.index_stride
    assert np.allclose(arr(back.ActionDensity), arr(R.ActionDensity), rtol=0, atol=0)


def test_inline_two_point_observables_short_circuit_the_reference_measurement():
    """inline=('Spin_Spin', 'Winding_Winding') / ('Vortex_Vortex',): the generators return the correlators of the configuration
    they have just produced, computed on the device (FFT kernels), under the reference's observable names -- so the reference
    reads the stored column instead of measuring (observable/observable.py:49-54).  Measured afresh by the reference from the
    stored fields (observable/spin.py:28-42, winding.py:77-86, vortex.py:22-37 via Lattice.correlation) they agree to 1e-12."""
    N, kappa, steps = 16, 0.5, 12
    S = sv.action.Villain(sv.lattice.Lattice2D(N), kappa, 1)
    G = NeighborhoodUpdate(S, seed=11, inline=('Spin_Spin', 'Winding_Winding', 'ActionDensity'))
    rng = np.random.default_rng(3)
    hot = {'phi': S.Lattice.form(0), 'n': S.Lattice.form(1, dtype=int)}
    hot['phi'][...] = rng.uniform(-np.pi, np.pi, hot['phi'].shape)
    hot['n'][...] = rng.integers(-2, 3, hot['n'].shape)
    E = sv.Ensemble(S).generate(steps, G, start=hot)
    bare = sv.Ensemble(S).from_configurations(sv.configurations.Configurations({k: E.configuration.fields[k] for k in ('phi', 'n')}))
    for name in ('Spin_Spin', 'Winding_Winding'):
        inline, measured = arr(getattr(E, name)), arr(getattr(bare, name))
        assert inline.shape == measured.shape == (steps, N, N) and np.iscomplexobj(inline), name
        assert np.abs(inline - measured).max() <= 1e-12 * max(1.0, np.abs(measured).max()), name
    Wl = sv.action.Worldline(sv.lattice.Lattice2D(N), kappa, 1)
    P = PlaquetteUpdate(Wl, seed=12, inline=('Vortex_Vortex',))
    F = sv.Ensemble(Wl).generate(steps, P, start='cold')
    bare = sv.Ensemble(Wl).from_configurations(sv.configurations.Configurations({k: F.configuration.fields[k] for k in ('m', 'v')}))
    inline, measured = arr(F.Vortex_Vortex), arr(bare.Vortex_Vortex)
    assert inline.shape == measured.shape == (steps, N, N)
    assert np.abs(inline - measured).max() <= 1e-12 * max(1.0, np.abs(measured).max())


def test_batched_ensemble_measures_two_point_observables_on_the_resident_fields():
    """BatchedEnsemble.generate(..., correlate_every=k): the FFT kernels run on the resident fields every k-th step into device
    columns (no configuration crosses to the host); the columns equal the correlators of the kept configurations."""
    import torch
    from supervillain_b200 import ops
    N, kappa, chains, steps, k = 32, 0.5, 6, 9, 3
    S = svb.Villain(svb.Lattice2D(N), kappa)
    E = svb.BatchedEnsemble(S, chains).generate(steps, NeighborhoodUpdate(S, seed=21), start='hot', start_seed=4, keep_every=k,
                                                correlate_every=k)
    assert set(E.two_point) == {'Spin_Spin', 'Winding_Winding'} and E.two_point['Spin_Spin'].is_cuda
    assert E.Spin_Spin.shape == (chains, steps // k, N, N) and np.iscomplexobj(E.Spin_Spin)
    for t in range(steps // k):
        phi = torch.from_numpy(E.configuration['phi'][:, t]).cuda()
        n = torch.from_numpy(E.configuration['n'][:, t]).to(device='cuda', dtype=torch.int32)
        assert np.array_equal(E.Spin_Spin[:, t], ops.villain_spin_spin(phi).cpu().numpy())
        assert np.array_equal(E.Winding_Winding[:, t], ops.correlation('winding', n).cpu().numpy())
    assert abs(E.Spin_Spin[:, :, 0, 0] - 1.0).max() < 1e-12                       # C[0] = 1
    # one big lattice (the in-place colour passes), one correlator
    S2 = svb.Villain(svb.Lattice2D(256), kappa)
    E2 = svb.BatchedEnsemble(S2, 1).generate(4, NeighborhoodUpdate(S2, seed=22), start='hot', start_seed=5, correlate_every=2,
                                             correlators=('Spin_Spin',))
    assert E2.Spin_Spin.shape == (1, 2, 256, 256)
    assert np.array_equal(E2.Spin_Spin[:, -1], ops.villain_spin_spin(E2.fields[0]).cpu().numpy())
    Wl = svb.Worldline(svb.Lattice2D(N), kappa)
    E3 = svb.BatchedEnsemble(Wl, chains).generate(4, PlaquetteUpdate(Wl, seed=23), start='hot', start_seed=6, correlate_every=4)
    assert np.array_equal(E3.Vortex_Vortex[:, 0], ops.correlation('vortex', E3.fields[1], W=1).cpu().numpy())
    with pytest.raises(NotImplementedError):
        svb.BatchedEnsemble(Wl, chains).generate(2, PlaquetteUpdate(Wl, seed=23), start='cold', correlate_every=1, correlators=('Spin_Spin',))
