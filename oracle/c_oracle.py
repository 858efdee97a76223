"""ctypes wrapper of oracle/svb_oracle.c (TEST INFRASTRUCTURE: the fast checker for full-size parity)."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        vp, i, i64, u64, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
        L.svo_villain_sweep_dense.restype = i
        L.svo_villain_sweep_dense.argtypes = [vp, vp, i, d, vp, vp, vp, vp, vp, vp, vp, vp]
        L.svo_villain_sweep_philox.restype = i
        L.svo_villain_sweep_philox.argtypes = [vp, vp, i64, i, d, i, d, i, i, u64, u64, u64, vp, vp]
        L.svo_villain_action.restype = d
        L.svo_villain_action.argtypes = [vp, vp, i, d]
        L.svo_worldline_sweep_dense.restype = i
        L.svo_worldline_sweep_dense.argtypes = [vp, vp, i, d, i, i, vp, vp, vp, vp, vp, vp]
        L.svo_worldline_sweep_philox.restype = i
        L.svo_worldline_sweep_philox.argtypes = [vp, vp, i64, i, d, i, i, i, i, u64, u64, u64, vp, vp]
        L.svo_philox4x32_10.restype = None
        L.svo_philox4x32_10.argtypes = [vp, vp, vp]
        L.svo_colour.restype = i
        L.svo_colour.argtypes = [i, i, i]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32); k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().svo_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def colour_map(N):
    return np.array([[lib().svo_colour(a, b, N) for b in range(N)] for a in range(N)])


def villain_sweep_dense(phi, n, kappa, draws, want_debug=False):
    """One sweep of one chain; phi (1,N,N) f64, n (2,N,N) int64.  Returns (phi, n, accepted, acceptance[, mask, dS])."""
    N = phi.shape[-1]
    p = np.ascontiguousarray(phi, dtype=np.float64).copy()
    q = np.ascontiguousarray(n, dtype=np.int64).copy()
    u = np.ascontiguousarray(draws['u'], dtype=np.float64); dphi = np.ascontiguousarray(draws['dphi'], dtype=np.float64)
    dnf = np.ascontiguousarray(draws['dn_fwd'], dtype=np.int64); dnb = np.ascontiguousarray(draws['dn_bwd'], dtype=np.int64)
    scratch = np.empty(2 * N * N)
    mask = np.zeros((N, N), dtype=np.uint8) if want_debug else None
    dS = np.zeros((N, N)) if want_debug else None
    accp = ctypes.c_double(0.0)
    acc = lib().svo_villain_sweep_dense(_p(p), _p(q), N, float(kappa), _p(u), _p(dphi), _p(dnf), _p(dnb), _p(scratch),
                                        _p(mask), _p(dS), ctypes.addressof(accp))
    if want_debug:
        return p, q, acc, accp.value, mask.astype(bool), dS
    return p, q, acc, accp.value


def villain_sweep_philox(phi, n, kappa, *, W=1, interval_phi=np.pi, interval_n=1, n_sweeps=1, seed=0, sweep0=0, chain0=0):
    """Batched: phi (chains,1,N,N), n (chains,2,N,N).  Returns (phi, n, accepted[chains], acceptance[chains])."""
    chains, N = phi.shape[0], phi.shape[-1]
    p = np.ascontiguousarray(phi, dtype=np.float64).copy()
    q = np.ascontiguousarray(n, dtype=np.int64).copy()
    acc = np.zeros(chains, dtype=np.int64); accp = np.zeros(chains)
    rc = lib().svo_villain_sweep_philox(_p(p), _p(q), chains, N, float(kappa), int(W), float(interval_phi), int(interval_n),
                                        int(n_sweeps), int(seed) & (2**64 - 1), int(sweep0), int(chain0), _p(acc), _p(accp))
    assert rc == 0
    return p, q, acc, accp


def villain_action(phi, n, kappa):
    N = phi.shape[-1]
    return lib().svo_villain_action(_p(np.ascontiguousarray(phi, dtype=np.float64)), _p(np.ascontiguousarray(n, dtype=np.int64)),
                                    N, float(kappa))


_MODES = {'joint': 0, 'vortex': 1, 'coexact': 2}


def worldline_sweep_dense(m, v, kappa, W, draws, mode, want_debug=False):
    N = m.shape[-1]
    a = np.ascontiguousarray(m, dtype=np.int64).copy(); b = np.ascontiguousarray(v, dtype=np.int64).copy()
    u = np.ascontiguousarray(draws['u'], dtype=np.float64)
    da = np.ascontiguousarray(draws['a'], dtype=np.int64); db = np.ascontiguousarray(draws['b'], dtype=np.int64)
    mask = np.zeros((N, N), dtype=np.uint8) if want_debug else None
    dS = np.zeros((N, N)) if want_debug else None
    accp = ctypes.c_double(0.0)
    acc = lib().svo_worldline_sweep_dense(_p(a), _p(b), N, float(kappa), int(W), _MODES[mode], _p(u), _p(da), _p(db), _p(mask),
                                          _p(dS), ctypes.addressof(accp))
    if want_debug:
        return a, b, acc, accp.value, mask.astype(bool), dS
    return a, b, acc, accp.value


def worldline_sweep_philox(m, v, kappa, *, W=1, mode='joint', interval=1, n_sweeps=1, seed=0, sweep0=0, chain0=0):
    chains, N = m.shape[0], m.shape[-1]
    a = np.ascontiguousarray(m, dtype=np.int64).copy(); b = np.ascontiguousarray(v, dtype=np.int64).copy()
    acc = np.zeros(chains, dtype=np.int64); accp = np.zeros(chains)
    rc = lib().svo_worldline_sweep_philox(_p(a), _p(b), chains, N, float(kappa), int(W), _MODES[mode], int(interval), int(n_sweeps),
                                          int(seed) & (2**64 - 1), int(sweep0), int(chain0), _p(acc), _p(accp))
    assert rc == 0
    return a, b, acc, accp
