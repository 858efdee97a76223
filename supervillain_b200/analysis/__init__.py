"""The step after the path: analysis of the scalar columns a batch of chains produced, on the device.

`autocorrelation` / `autocorrelation_time` mirror supervillain.analysis.autocorrelation
(supervillain/analysis/autocorrelation.py:7-66) -- same definition (circular autocorrelation normalised to C(0) = 1, tau_int
= ceil of its sum up to the first zero), same ValueError for a series without fluctuations -- for one series or for one
series per chain at once.

`Blocking` and `Bootstrap` mirror supervillain.analysis.Blocking / Bootstrap (analysis/blocking.py, analysis/bootstrap.py)
for a `BatchedEnsemble`: every scalar observable column of every chain is blocked / resampled in one launch
(svb_block_mean, svb_bootstrap_mean).  Results keep the batched convention, chains first: (chains, blocks) and
(chains, draws) where the reference returns (blocks,) and (draws,) for its single chain."""
import numpy as np
import torch

from .. import ops


def autocorrelation(data, mean=None, _cutoff=1e-16):
    """`data`: a time series (T,) -> (C (T,) numpy, tau int), like the reference; or (series, T) -> (C (series, T), tau
    (series,)).  numpy arrays, Batches and torch tensors are accepted; the sums run on the GPU."""
    if _cutoff != 1e-16:
        raise NotImplementedError('the device kernel uses the reference default cutoff 1e-16')
    t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64)))
    single = t.dim() == 1
    t = t.reshape(1, -1) if single else t
    t = t.to(device='cuda', dtype=torch.float64).contiguous()
    m = None
    if mean is not None:
        m = torch.as_tensor(np.broadcast_to(np.asarray(mean, dtype=np.float64), (t.shape[0],)).copy()).cuda()
    C, tau = ops.autocorrelation(t, m)
    C, tau = C.cpu().numpy(), tau.cpu().numpy()
    if (tau < 0).any():
        raise ValueError('The fluctuations are too small to reliably determine an autocorrelation.')
    return (C[0], int(tau[0])) if single else (C, tau.astype(np.int64))


def autocorrelation_time(data, mean=None):
    """Just like `autocorrelation` but only returns tau_int (autocorrelation.py:62-66)."""
    return autocorrelation(data, mean)[1]


class _ColumnAnalysis:
    """Shared plumbing: scalar columns (chains, T) of the underlying object, analysed lazily and cached by name."""

    def _column(self, name):
        col = getattr(self.Ensemble, name)
        col = np.asarray(col, dtype=np.float64)
        if col.ndim != 2:
            raise NotImplementedError(f'{name}: only scalar observables (one number per chain and sample) are analysed on the device')
        return torch.from_numpy(np.ascontiguousarray(col)).cuda()

    def __getattr__(self, name):
        if name.startswith('_') or name in ('Ensemble',):
            raise AttributeError(name)
        try:
            col = self._column(name)
        except AttributeError as e:
            raise AttributeError(f"... and so '{type(self).__name__}' object has no attribute '{name}'") from e
        value = self._analyse(col).cpu().numpy()
        self.__dict__[name] = value
        return value


class Blocking(_ColumnAnalysis):
    """supervillain.analysis.Blocking (analysis/blocking.py:12-111) for a batch of chains: consecutive samples averaged in
    blocks of `width` ('auto': the ensemble's autocorrelation time), the first `len % width` samples dropped."""

    def __init__(self, ensemble, width='auto'):
        self.Ensemble = ensemble
        self.Action = ensemble.Action
        self.chains = ensemble.chains
        self.width = int(ensemble.autocorrelation_time() if width == 'auto' else width)
        cfgs = len(ensemble)
        self.drop = cfgs % self.width
        self.blocks = (cfgs - self.drop) // self.width
        self.weight = np.ones(self.blocks)
        index = np.asarray(getattr(ensemble, 'index', np.arange(cfgs)))
        self.index = index[self.drop:].reshape(-1, self.width).mean(axis=1)
        self.steps = self.blocks

    def __len__(self):
        return self.blocks

    def _analyse(self, col):
        return ops.block_mean(col, self.width, self.drop)

    def autocorrelation_time(self, *args, **kwargs):
        return self.Ensemble.autocorrelation_time(*args, **kwargs)


class Bootstrap(_ColumnAnalysis):
    """supervillain.analysis.Bootstrap (analysis/bootstrap.py:12-67) for a batch of chains (or a `Blocking` of one): `draws`
    resamplings, the same for every chain and every observable -- `indices` (samples, draws) are drawn exactly as the
    reference draws them, np.random.randint(0, samples, (samples, draws)), unless given."""

    def __init__(self, ensemble, draws=100, indices=None):
        self.Ensemble = ensemble
        self.Action = ensemble.Action
        self.chains = ensemble.chains
        self.draws = int(draws)
        cfgs = len(ensemble)
        self.indices = np.random.randint(0, cfgs, (cfgs, self.draws)) if indices is None else np.asarray(indices, dtype=np.int64)
        if self.indices.shape != (cfgs, self.draws):
            raise ValueError(f'indices must have shape ({cfgs}, {self.draws}); got {self.indices.shape}')
        self._idx = torch.from_numpy(np.ascontiguousarray(self.indices, dtype=np.int64)).cuda()

    def __len__(self):
        return self.draws

    def _analyse(self, col):
        return ops.bootstrap_mean(col, self._idx)

    def estimate(self, name):
        """(mean, std) over the draws for every chain: the central value and its uncertainty (bootstrap.py:36-38)."""
        v = getattr(self, name)
        return v.mean(axis=1), v.std(axis=1)
