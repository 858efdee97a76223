"""The step after the path: analysis of the scalar columns a batch of chains produced, on the device.

`autocorrelation` / `autocorrelation_time` mirror supervillain.analysis.autocorrelation
(supervillain/analysis/autocorrelation.py:7-66) -- same definition (circular autocorrelation normalised to C(0) = 1, tau_int
= ceil of its sum up to the first zero), same ValueError for a series without fluctuations -- for one series or for one
series per chain at once."""
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
