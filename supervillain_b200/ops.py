"""Thin, validated Python wrappers over the C ABI of libsvb200.so.

Arguments are `torch` CUDA tensors used purely as device storage (`data_ptr()` + the current
stream); every function enqueues work on `torch.cuda.current_stream()` and returns immediately.
Nothing here computes on the CPU -- a missing library or device raises.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import (ARITH_FAST, ARITH_STRICT, PATH_AUTO, PATH_GLOBAL, PATH_SMEM, RNG_INJECTED, RNG_PHILOX,  # noqa: F401
                   VOBS_COUNT, WOBS_COUNT, WL_COEXACT, WL_JOINT, WL_VORTEX)

_PATHS = {'auto': PATH_AUTO, 'smem': PATH_SMEM, 'global': PATH_GLOBAL}
_TILE = 32
_SMEM_MAX_N = 96            # largest fp64 lattice whose chain fits one SM's shared memory
_CLUSTER_N = (128,)         # lattices svb_villain_sweep spreads over a thread-block cluster (svb_villain_cluster.cuh)
_workspaces = {}            # (device, stream) -> (chains, N, phi_ws, n_ws): the tiled path's ping-pong workspace, ONE PER STREAM
_ARITH = {'strict': ARITH_STRICT, 'fast': ARITH_FAST}
_WL_MODES = {'joint': WL_JOINT, 'vortex': WL_VORTEX, 'coexact': WL_COEXACT}
_OPS = {'d': _lib.OP_D, 'delta': _lib.OP_DELTA, 'face_sum': _lib.OP_FACE_SUM, 'coface_sum': _lib.OP_COFACE_SUM}
_DTYPES = {torch.float64: _lib.F64, torch.float32: _lib.F32, torch.int32: _lib.I32, torch.int64: _lib.I64}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev(t, name, dtypes, shape=None):
    """Validate a device tensor argument and return its pointer."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name} must be a torch.Tensor on a CUDA device, got {type(t).__name__}')
    if not t.is_cuda:
        raise ValueError(f'{name} must live on a CUDA device (supervillain_b200 has no CPU path)')
    if t.dtype not in dtypes:
        raise TypeError(f'{name} has dtype {t.dtype}; expected one of {dtypes}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be C-contiguous')
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f'{name} has shape {tuple(t.shape)}; expected {tuple(shape)}')
    return t.data_ptr()


def _opt(t, name, dtypes, shape):
    return None if t is None else _dev(t, name, dtypes, shape)


def _fields_shape(t, name, comps):
    if t.dim() != 4 or t.shape[1] != comps or t.shape[2] != t.shape[3]:
        raise ValueError(f'{name} must have shape (chains, {comps}, N, N); got {tuple(t.shape)}')
    return int(t.shape[0]), int(t.shape[2])


def villain_sweep(phi, n, kappa, *, W=1, interval_phi=math.pi, interval_n=1, n_sweeps=1, seed=0, sweep0=0,
                  chain0=0, injected=None, arithmetic='fast', path='auto', kappa_chain=None, obs=None,
                  accept_mask=None, dS_out=None):
    """`n_sweeps` NeighborhoodUpdate sweeps on every chain, in place (svb_villain_sweep).

    phi (chains,1,N,N) float64|float32, n (chains,2,N,N) int32.  `injected` is None (in-kernel
    Philox) or a dict of device tensors u, dphi (n_sweeps,chains,N,N) f64 and dn_fwd, dn_bwd
    (n_sweeps,chains,2,N,N) int32 -- the reference's own draws, see generator.villain.
    """
    lib = _lib.load()
    chains, N = _fields_shape(phi, 'phi', 1)
    p_phi = _dev(phi, 'phi', (torch.float64, torch.float32))
    p_n = _dev(n, 'n', (torch.int32,), (chains, 2, N, N))
    if W != W or W == float('inf') or int(W) != W:
        raise ValueError('the Villain NeighborhoodUpdate needs a finite integer W (the reference yields nan for W=inf, '
                         'neighborhood.py:105)')
    tiled_ok = injected is None and phi.dtype == torch.float64 and N % _TILE == 0
    if path == 'tiled' and not tiled_ok:
        raise NotImplementedError('the tiled path needs fp64 phi, Philox draws and N a multiple of 32')
    cluster_ok = (tiled_ok and N in _CLUSTER_N and arithmetic == 'fast' and accept_mask is None and dS_out is None)
    if path == 'tiled' or (path == 'auto' and tiled_ok and N > _SMEM_MAX_N and not cluster_ok):
        # The workspace belongs to the stream the sweep is enqueued on: two streams sweeping at once (HostStepper's chunks)
        # never share one, and a workspace is only ever replaced by the stream that used it -- it was allocated with that
        # stream current, so the caching allocator hands its memory on in that stream's order.
        key = (phi.device, _stream())
        held = _workspaces.get(key)
        if held is None or held[0] != chains or held[1] != N:
            held = _workspaces[key] = (chains, N, torch.empty_like(phi), torch.empty_like(n))
        ws = held[2:]
        _lib.check(lib.svb_villain_sweep_tiled(
            p_phi, p_n, ws[0].data_ptr(), ws[1].data_ptr(), chains, N, float(kappa),
            _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)), int(W), float(interval_phi), int(interval_n),
            int(n_sweeps), int(seed) & (2**64 - 1), int(sweep0), int(chain0), _ARITH[arithmetic],
            _opt(obs, 'obs', (torch.float64,), (chains, VOBS_COUNT)),
            _opt(accept_mask, 'accept_mask', (torch.uint8,), (chains, N, N)),
            _opt(dS_out, 'dS_out', (torch.float64,), (chains, N, N)), _stream()))
        return
    if injected is None:
        rng_mode, pu, pd, pf, pb = RNG_PHILOX, None, None, None, None
    else:
        rng_mode = RNG_INJECTED
        arithmetic = 'strict'
        pu = _dev(injected['u'], 'injected[u]', (torch.float64,), (n_sweeps, chains, N, N))
        pd = _dev(injected['dphi'], 'injected[dphi]', (torch.float64,), (n_sweeps, chains, N, N))
        pf = _dev(injected['dn_fwd'], 'injected[dn_fwd]', (torch.int32,), (n_sweeps, chains, 2, N, N))
        pb = _dev(injected['dn_bwd'], 'injected[dn_bwd]', (torch.int32,), (n_sweeps, chains, 2, N, N))
    code = lib.svb_villain_sweep(
        p_phi, _DTYPES[phi.dtype], p_n, chains, N, float(kappa),
        _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)), int(W),
        float(interval_phi), int(interval_n), int(n_sweeps), int(seed) & (2**64 - 1), int(sweep0), int(chain0),
        rng_mode, _ARITH[arithmetic], _PATHS[path], pu, pd, pf, pb,
        _opt(obs, 'obs', (torch.float64,), (chains, VOBS_COUNT)),
        _opt(accept_mask, 'accept_mask', (torch.uint8,), (chains, N, N)),
        _opt(dS_out, 'dS_out', (torch.float64,), (chains, N, N)),
        _stream())
    _lib.check(code)


class VillainSwappingSweeps:
    """Tiled sweeps (lattices beyond one SM's shared memory and beyond the cluster kernel) for a caller that lets the
    state move between two buffer pairs (svb_villain_sweep_tiled_swap): an odd number of sweeps ends in the other pair
    and nothing is copied back.  `fields` is the pair that holds the state now; the pair passed in is used as is
    (its tensors stay valid, but after `step` the state may live in the other pair)."""

    def __init__(self, phi, n, kappa, *, W=1, interval_phi=math.pi, interval_n=1, seed=0, chain0=0, arithmetic='fast',
                 kappa_chain=None, force_tiled=False):
        self.lib = _lib.load()
        self.chains, self.N = _fields_shape(phi, 'phi', 1)
        if phi.dtype != torch.float64:
            raise NotImplementedError('swapping sweeps need fp64 phi')
        _dev(phi, 'phi', (torch.float64,))
        _dev(n, 'n', (torch.int32,), (self.chains, 2, self.N, self.N))
        if self.N % _TILE or (not force_tiled and (self.N <= _SMEM_MAX_N or self.N in _CLUSTER_N)):
            raise NotImplementedError('swapping sweeps serve the tiled path: N a multiple of 32 above 96 that the cluster kernel does not take')
        if W != W or W == float('inf') or int(W) != W:
            raise ValueError('the Villain NeighborhoodUpdate needs a finite integer W')
        self.fields = (phi, n)
        self.spare = (torch.empty_like(phi), torch.empty_like(n))
        self.kappa_chain = kappa_chain
        self.p_kappa_chain = _opt(kappa_chain, 'kappa_chain', (torch.float64,), (self.chains,))
        self.args = (float(kappa), int(W), float(interval_phi), int(interval_n))
        self.tail = (int(seed) & (2**64 - 1), int(chain0), _ARITH[arithmetic])
        self.flag = ctypes.c_int(0)

    def step(self, sweep0, n_sweeps=1, obs=None):
        kappa, W, interval_phi, interval_n = self.args
        seed, chain0, arith = self.tail
        (phi, n), (wphi, wn) = self.fields, self.spare
        _lib.check(self.lib.svb_villain_sweep_tiled_swap(
            phi.data_ptr(), n.data_ptr(), wphi.data_ptr(), wn.data_ptr(), self.chains, self.N, kappa, self.p_kappa_chain, W,
            interval_phi, interval_n, int(n_sweeps), seed, int(sweep0), chain0, arith,
            _opt(obs, 'obs', (torch.float64,), (self.chains, VOBS_COUNT)), ctypes.byref(self.flag), _stream()))
        if self.flag.value:
            self.fields, self.spare = self.spare, self.fields
        return self.fields


def villain_sweep_plan(phi, n, kappa, *, W=1, interval_phi=math.pi, interval_n=1, seed=0, chain0=0, arithmetic='fast',
                       path='auto', kappa_chain=None, obs=None):
    """Validate the arguments of a Philox-mode `villain_sweep` ONCE and return `run(sweep0, n_sweeps=1)`.

    Launch-bound loops (one 50 us kernel per step) should not pay Python-side validation per launch; the returned
    closure does nothing but the C call.  The tensors must stay alive and in place for the life of the plan.
    """
    lib = _lib.load()
    chains, N = _fields_shape(phi, 'phi', 1)
    p_phi = _dev(phi, 'phi', (torch.float64, torch.float32))
    p_n = _dev(n, 'n', (torch.int32,), (chains, 2, N, N))
    if W != W or W == float('inf') or int(W) != W:
        raise ValueError('the Villain NeighborhoodUpdate needs a finite integer W')
    p_kc = _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,))
    p_obs = _opt(obs, 'obs', (torch.float64,), (chains, VOBS_COUNT))
    fn = lib.svb_villain_sweep
    dt, kappa, W, interval_phi, interval_n = _DTYPES[phi.dtype], float(kappa), int(W), float(interval_phi), int(interval_n)
    seed, chain0, arith, pth = int(seed) & (2**64 - 1), int(chain0), _ARITH[arithmetic], _PATHS[path]
    keep = (phi, n, kappa_chain, obs)
    current_stream = torch.cuda.current_stream

    def run(sweep0, n_sweeps=1, _keep=keep):
        code = fn(p_phi, dt, p_n, chains, N, kappa, p_kc, W, interval_phi, interval_n, n_sweeps, seed, sweep0, chain0,
                  RNG_PHILOX, arith, pth, None, None, None, None, p_obs, None, None, current_stream().cuda_stream)
        if code:
            _lib.check(code)
    return run


OVERLAP_SIZES = (16, 32, 64, 128)        # worldline table kernels (128: one chain per SM)
VILLAIN_OVERLAP_SIZES = (16, 32, 64, 128)  # filtered kernels; 128 is the cluster kernel


class VillainOverlappedSweeps:
    """Back-to-back Philox sweeps of ONE chain set as overlapped launches (svb_villain_sweep_overlapped).

    Each `step` is one kernel launch that may begin while the previous step's launch is still running; chains are
    ordered individually through a per-chain epoch word, so results are identical to `villain_sweep` call by call.
    The FIRST step after construction or `fence()` waits for all earlier work in the stream; call `fence()` whenever
    something other than these steps has written phi, n or kappa_chain (anything that only READS them -- a normal
    kernel or copy enqueued afterwards -- waits for the overlapped launches by ordinary stream order and needs nothing).
    Steps of several instances (different chain sets) may be interleaved on one stream.
    """

    def __init__(self, phi, n, kappa, *, W=1, interval_phi=math.pi, interval_n=1, seed=0, chain0=0, kappa_chain=None):
        self.lib = _lib.load()
        self.chains, self.N = _fields_shape(phi, 'phi', 1)
        if self.N not in VILLAIN_OVERLAP_SIZES or phi.dtype != torch.float64 or int(interval_n) > 1:
            raise NotImplementedError('overlapped sweeps need fp64 phi, N in (16, 32, 64, 128) and interval_n <= 1')
        self.p_phi = _dev(phi, 'phi', (torch.float64,))
        self.p_n = _dev(n, 'n', (torch.int32,), (self.chains, 2, self.N, self.N))
        if W != W or W == float('inf') or int(W) != W:
            raise ValueError('the Villain NeighborhoodUpdate needs a finite integer W')
        self.p_kc = _opt(kappa_chain, 'kappa_chain', (torch.float64,), (self.chains,))
        self.args = (float(kappa), self.p_kc, int(W), float(interval_phi), int(interval_n))
        self.seed, self.chain0 = int(seed) & (2**64 - 1), int(chain0)
        self.epochs = torch.zeros((self.chains,), dtype=torch.int32, device=phi.device)
        self.p_epochs = self.epochs.data_ptr()
        self.epoch = 0
        self.fenced = True
        self._keep = (phi, n, kappa_chain)
        self._fn = self.lib.svb_villain_sweep_overlapped
        self._stream = torch.cuda.current_stream
        self._records, self._record_keep = {}, {}

    def fence(self):
        """The next step waits for everything enqueued before it (use after foreign writes to the fields)."""
        self.fenced = True

    def _record(self, t, name):
        if t is None:
            return None
        key = id(t)
        ptr = self._records.get(key)
        if ptr is None:                            # validated once: tight loops pay nothing per launch
            ptr = _dev(t, name, (torch.float64,), (self.chains, VOBS_COUNT))
            if len(self._records) > 64:
                self._records.clear()
                self._record_keep.clear()
            self._records[key] = ptr
            self._record_keep[key] = t
        return ptr

    def step(self, sweep0, n_sweeps=1, obs=None, obs_in=None):
        """One launch of `n_sweeps` sweeps.  obs: this step's record.  obs_in: if given, the state columns of the chains
        AS THEY ARRIVE go there (pass the previous step's record) and `obs` receives only this launch's counters."""
        p_obs, p_obs_in = self._record(obs, 'obs'), self._record(obs_in, 'obs_in')
        e = self.epoch
        code = self._fn(self.p_phi, self.p_n, self.chains, self.N, *self.args, int(n_sweeps), self.seed, int(sweep0), self.chain0,
                        p_obs, p_obs_in, self.p_epochs, e & 0xFFFFFFFF, (e + 1) & 0xFFFFFFFF, 0 if self.fenced else _lib.OVERLAP_PREDECESSOR,
                        self._stream().cuda_stream)
        if code:
            _lib.check(code)
        self.epoch = e + 1
        self.fenced = False


class VillainInplaceSweeps:
    """Back-to-back Philox sweeps of lattices beyond a CTA (config 5: L = 4096; any N that is a multiple of 16), IN PLACE:
    no second copy of the state, only accepted proposals are written.  The `step` / `fence` interface of
    `VillainOverlappedSweeps`, including the records of the arriving state (`obs_in`).

    launches='passes' (the default): one launch per colour pass and one over n for sum (dn)^2 (svb_villain_sweep_inplace).
    launches='wavefront' (N a multiple of 128): ONE launch per step, the phases of the step following one another down the
    lattice through L2 (svb_villain_sweep_wavefront) -- one DRAM read of the state per step instead of 2.5, but no faster on
    a B200, where the colour passes are bound by instruction issue, not by DRAM; SVB_VILLAIN_STEP=wavefront
    makes it the default for an A/B.  Identical fields either way.  `fused_sweeps`: how many sweeps of a multi-sweep step the
    wavefront workspace has room to put in one launch."""

    def __init__(self, phi, n, kappa, *, W=1, interval_phi=math.pi, interval_n=1, seed=0, chain0=0, kappa_chain=None,
                 launches='auto', fused_sweeps=2):
        self.lib = _lib.load()
        self.chains, self.N = _fields_shape(phi, 'phi', 1)
        if self.N % 16 or phi.dtype != torch.float64 or int(interval_n) > 1:
            raise NotImplementedError('in-place sweeps need fp64 phi, N a multiple of 16 and interval_n <= 1')
        self.p_phi = _dev(phi, 'phi', (torch.float64,))
        self.p_n = _dev(n, 'n', (torch.int32,), (self.chains, 2, self.N, self.N))
        if W != W or W == float('inf') or int(W) != W:
            raise ValueError('the Villain NeighborhoodUpdate needs a finite integer W')
        self.p_kc = _opt(kappa_chain, 'kappa_chain', (torch.float64,), (self.chains,))
        self.args = (float(kappa), self.p_kc, int(W), float(interval_phi), int(interval_n))
        self.seed, self.chain0 = int(seed) & (2**64 - 1), int(chain0)
        self._keep = (phi, n, kappa_chain)
        self._stream = torch.cuda.current_stream
        if launches not in ('auto', 'wavefront', 'passes'):
            raise ValueError("launches must be 'auto', 'wavefront' or 'passes'")
        serves = self.N % 128 == 0 and 2 * self.chains < 2**31 - 1
        if launches == 'wavefront' and not serves:
            raise NotImplementedError('wavefront launches need N a multiple of 128')
        if launches == 'auto':
            launches = 'wavefront' if serves and os.environ.get('SVB_VILLAIN_STEP', 'passes') == 'wavefront' else 'passes'
        self.launches = launches
        if launches == 'wavefront':
            ints = int(self.lib.svb_villain_wavefront_workspace(self.chains, self.N, max(1, int(fused_sweeps)), 1))
            self.workspace = torch.zeros((ints,), dtype=torch.int32, device=phi.device)      # zero before first use; calls leave it zero
            self._tail = (self.workspace.data_ptr(), ints)
            self._fn = self.lib.svb_villain_sweep_wavefront
        else:
            self._tail = ()
            self._fn = self.lib.svb_villain_sweep_inplace

    def fence(self):
        pass

    def step(self, sweep0, n_sweeps=1, obs=None, obs_in=None):
        p_obs = None if obs is None else _dev(obs, 'obs', (torch.float64,), (self.chains, VOBS_COUNT))
        p_obs_in = None if obs_in is None else _dev(obs_in, 'obs_in', (torch.float64,), (self.chains, VOBS_COUNT))
        _lib.check(self._fn(self.p_phi, self.p_n, self.chains, self.N, *self.args, int(n_sweeps), self.seed, int(sweep0), self.chain0,
                            p_obs, p_obs_in, *self._tail, self._stream().cuda_stream))


class WorldlineOverlappedSweeps:
    """Back-to-back Philox worldline sweeps (W = 1; mode 'joint', or 'vortex' / 'coexact' with interval <= 2) of ONE chain
    set as overlapped launches (svb_worldline_sweep_overlapped); the protocol and the `fence()` rule are those of
    `VillainOverlappedSweeps`."""

    def __init__(self, m, v, kappa, *, mode='joint', interval=1, seed=0, chain0=0, kappa_chain=None):
        self.lib = _lib.load()
        self.chains, self.N = _fields_shape(m, 'm', 2)
        if self.N not in OVERLAP_SIZES or (mode != 'joint' and interval > 2):
            raise NotImplementedError('overlapped sweeps need N in (16, 32, 64, 128) and interval <= 2')
        self.mode, self.interval = _WL_MODES[mode], int(interval)
        self.p_m = _dev(m, 'm', (torch.int32,))
        self.p_v = _dev(v, 'v', (torch.int32,), (self.chains, 1, self.N, self.N))
        self.p_kc = _opt(kappa_chain, 'kappa_chain', (torch.float64,), (self.chains,))
        self.kappa = float(kappa)
        self.seed, self.chain0 = int(seed) & (2**64 - 1), int(chain0)
        self.epochs = torch.zeros((self.chains,), dtype=torch.int32, device=m.device)
        self.p_epochs = self.epochs.data_ptr()
        self.epoch = 0
        self.fenced = True
        self._keep = (m, v, kappa_chain)
        self._fn = self.lib.svb_worldline_sweep_overlapped
        self._stream = torch.cuda.current_stream

    def fence(self):
        self.fenced = True

    def step(self, sweep0, n_sweeps=1, obs=None):
        p_obs = None if obs is None else _dev(obs, 'obs', (torch.float64,), (self.chains, WOBS_COUNT))
        e = self.epoch
        code = self._fn(self.p_m, self.p_v, self.chains, self.N, self.kappa, self.p_kc, self.mode, self.interval, int(n_sweeps),
                        self.seed, int(sweep0),
                        self.chain0, p_obs, self.p_epochs, e & 0xFFFFFFFF, (e + 1) & 0xFFFFFFFF,
                        0 if self.fenced else _lib.OVERLAP_PREDECESSOR, self._stream().cuda_stream)
        if code:
            _lib.check(code)
        self.epoch = e + 1
        self.fenced = False


_VU_KINDS = {'site': _lib.VU_SITE, 'link': _lib.VU_LINK, 'exact': _lib.VU_EXACT}


def villain_decoupled(kind, phi, n, kappa, *, W=1, interval_phi=math.pi, interval=1, n_sweeps=1, seed=0, sweep0=0, chain0=0,
                      injected=None, path='auto', kappa_chain=None, obs=None, accept_mask=None, dS_out=None):
    """`n_sweeps` sweeps of SiteUpdate ('site'), LinkUpdate ('link') or ExactUpdate ('exact') on every chain, in place
    (svb_villain_decoupled).  phi (chains,1,N,N) float64, n (chains,2,N,N) int32.  `injected`: dict with u and, for
    'site', dphi (n_sweeps,chains,N,N); for 'link', u and a shaped (n_sweeps,chains,2,N,N); for 'exact', a = z."""
    lib = _lib.load()
    chains, N = _fields_shape(phi, 'phi', 1)
    p_phi = _dev(phi, 'phi', (torch.float64,))
    p_n = _dev(n, 'n', (torch.int32,), (chains, 2, N, N))
    if W != W or W == float('inf') or int(W) != W:
        raise ValueError('the Villain updates need a finite integer W')
    per = (chains, 2, N, N) if kind == 'link' else (chains, N, N)
    pu = pd = pa = None
    rng_mode = RNG_PHILOX
    if injected is not None:
        rng_mode = RNG_INJECTED
        pu = _dev(injected['u'], 'injected[u]', (torch.float64,), (n_sweeps,) + per)
        if kind == 'site':
            pd = _dev(injected['dphi'], 'injected[dphi]', (torch.float64,), (n_sweeps,) + per)
        else:
            pa = _dev(injected['a'], 'injected[a]', (torch.int32,), (n_sweeps,) + per)
    code = lib.svb_villain_decoupled(
        _VU_KINDS[kind], p_phi, p_n, chains, N, float(kappa), _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)),
        int(W), float(interval_phi), int(interval), int(n_sweeps), int(seed) & (2**64 - 1), int(sweep0), int(chain0),
        rng_mode, _PATHS[path], pu, pd, pa, _opt(obs, 'obs', (torch.float64,), (chains, VOBS_COUNT)),
        _opt(accept_mask, 'accept_mask', (torch.uint8,), per), _opt(dS_out, 'dS_out', (torch.float64,), per), _stream())
    _lib.check(code)


def villain_cohomology(phi, n, kappa, *, interval=1, seed=0, sweep=0, chain0=0, injected=None, kappa_chain=None,
                       counters=None, dS_out=None):
    """One CohomologyUpdate step (one slice proposal per direction and chain), in place on n (svb_villain_cohomology).
    `injected`: dict of u (chains,2) f64 and h (chains,2) int32.  counters (chains,2), accumulated: accepted, acceptance."""
    lib = _lib.load()
    chains, N = _fields_shape(phi, 'phi', 1)
    p_phi = _dev(phi, 'phi', (torch.float64,))
    p_n = _dev(n, 'n', (torch.int32,), (chains, 2, N, N))
    if injected is None:
        rng_mode, pu, ph = RNG_PHILOX, None, None
    else:
        rng_mode = RNG_INJECTED
        pu = _dev(injected['u'], 'injected[u]', (torch.float64,), (chains, 2))
        ph = _dev(injected['h'], 'injected[h]', (torch.int32,), (chains, 2))
    _lib.check(lib.svb_villain_cohomology(
        p_phi, p_n, chains, N, float(kappa), _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)), int(interval),
        int(seed) & (2**64 - 1), int(sweep), int(chain0), rng_mode, pu, ph,
        _opt(counters, 'counters', (torch.float64,), (chains, 2)), _opt(dS_out, 'dS_out', (torch.float64,), (chains, 2)),
        _stream()))


def villain_observables(phi, n, kappa, *, kappa_chain=None, obs=None):
    """Per-chain action / sum dn^2 / wrapping sums of the current state -> (chains, VOBS_COUNT) f64."""
    lib = _lib.load()
    chains, N = _fields_shape(phi, 'phi', 1)
    if obs is None:
        obs = torch.empty((chains, VOBS_COUNT), dtype=torch.float64, device=phi.device)
    code = lib.svb_villain_observables(
        _dev(phi, 'phi', (torch.float64, torch.float32)), _DTYPES[phi.dtype],
        _dev(n, 'n', (torch.int32,), (chains, 2, N, N)), chains, N, float(kappa),
        _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)),
        _dev(obs, 'obs', (torch.float64,), (chains, VOBS_COUNT)), _stream())
    _lib.check(code)
    return obs


def worldline_sweep(m, v, kappa, *, W=1, mode='joint', interval=1, n_sweeps=1, seed=0, sweep0=0, chain0=0,
                    injected=None, path='auto', kappa_chain=None, obs=None, accept_mask=None, dS_out=None):
    """`n_sweeps` checkerboard plaquette sweeps on every chain, in place (svb_worldline_sweep).

    m (chains,2,N,N) int32, v (chains,1,N,N) int32.  mode: 'joint' (PlaquetteUpdate's move),
    'vortex' (v only) or 'coexact' (m only).  `injected`: dict of u (n_sweeps,chains,N,N) f64,
    a (same shape) int32 and, for 'joint', b.
    """
    lib = _lib.load()
    chains, N = _fields_shape(m, 'm', 2)
    p_m = _dev(m, 'm', (torch.int32,))
    p_v = _dev(v, 'v', (torch.int32,), (chains, 1, N, N))
    if W != W or W == float('inf') or int(W) != W:
        raise ValueError('the GPU worldline sweep needs a finite integer W (integer-valued v)')
    if injected is None:
        rng_mode, pu, pa, pb = RNG_PHILOX, None, None, None
    else:
        rng_mode = RNG_INJECTED
        pu = _dev(injected['u'], 'injected[u]', (torch.float64,), (n_sweeps, chains, N, N))
        pa = _dev(injected['a'], 'injected[a]', (torch.int32,), (n_sweeps, chains, N, N))
        pb = _opt(injected.get('b'), 'injected[b]', (torch.int32,), (n_sweeps, chains, N, N))
    code = lib.svb_worldline_sweep(
        p_m, p_v, chains, N, float(kappa), _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)), int(W),
        _WL_MODES[mode], int(interval), int(n_sweeps), int(seed) & (2**64 - 1), int(sweep0), int(chain0),
        rng_mode, _PATHS[path], pu, pa, pb,
        _opt(obs, 'obs', (torch.float64,), (chains, WOBS_COUNT)),
        _opt(accept_mask, 'accept_mask', (torch.uint8,), (chains, N, N)),
        _opt(dS_out, 'dS_out', (torch.float64,), (chains, N, N)),
        _stream())
    _lib.check(code)


def worldline_wrapping(m, v, kappa, *, W=1, interval=1, seed=0, sweep=0, chain0=0, injected=None, kappa_chain=None,
                       counters=None, dS_out=None):
    """One WrappingUpdate step (2N torus-cycle proposals per chain), in place on m (svb_worldline_wrapping).

    `injected`: dict of u (chains,2,N) f64 and c (chains,2,N) int32.  counters (chains,2): accepted, sum of acceptance.
    """
    lib = _lib.load()
    chains, N = _fields_shape(m, 'm', 2)
    p_m = _dev(m, 'm', (torch.int32,))
    p_v = _dev(v, 'v', (torch.int32,), (chains, 1, N, N))
    if injected is None:
        rng_mode, pu, pc = RNG_PHILOX, None, None
    else:
        rng_mode = RNG_INJECTED
        pu = _dev(injected['u'], 'injected[u]', (torch.float64,), (chains, 2, N))
        pc = _dev(injected['c'], 'injected[c]', (torch.int32,), (chains, 2, N))
    _lib.check(lib.svb_worldline_wrapping(
        p_m, p_v, chains, N, float(kappa), _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)), int(W), int(interval),
        int(seed) & (2**64 - 1), int(sweep), int(chain0), rng_mode, pu, pc,
        _opt(counters, 'counters', (torch.float64,), (chains, 2)), _opt(dS_out, 'dS_out', (torch.float64,), (chains, 2, N)),
        _stream()))


def worldline_observables(m, v, *, W=1, obs=None):
    lib = _lib.load()
    chains, N = _fields_shape(m, 'm', 2)
    if obs is None:
        obs = torch.empty((chains, WOBS_COUNT), dtype=torch.float64, device=m.device)
    code = lib.svb_worldline_observables(
        _dev(m, 'm', (torch.int32,)), _dev(v, 'v', (torch.int32,), (chains, 1, N, N)), chains, N, int(W),
        _dev(obs, 'obs', (torch.float64,), (chains, WOBS_COUNT)), _stream())
    _lib.check(code)
    return obs


_OUT_COMPS = {('d', 0): 2, ('d', 1): 1, ('coface_sum', 0): 2, ('coface_sum', 1): 1,
              ('delta', 1): 1, ('delta', 2): 2, ('face_sum', 1): 1, ('face_sum', 2): 2}
_IN_COMPS = {0: 1, 1: 2, 2: 1}


def form_op(op, degree, f, out=None):
    """d / delta / face_sum / coface_sum of a batch of forms f (chains, C, N, N), dtype-preserving.

    Returns None at the ends of the complex where the reference returns the scalar 0
    (compact.py:999-1000, 1035-1036, 865-866, 888-889).
    """
    lib = _lib.load()
    if op not in _OPS:
        raise ValueError(f'unknown form operator {op!r}')
    if (op, degree) not in _OUT_COMPS:
        return None
    chains, N = _fields_shape(f, 'f', _IN_COMPS[degree])
    p_in = _dev(f, 'f', tuple(_DTYPES))
    if out is None:
        out = torch.empty((chains, _OUT_COMPS[(op, degree)], N, N), dtype=f.dtype, device=f.device)
    p_out = _dev(out, 'out', (f.dtype,), (chains, _OUT_COMPS[(op, degree)], N, N))
    _lib.check(lib.svb_form_op(_OPS[op], int(degree), _DTYPES[f.dtype], p_in, p_out, chains, N, _stream()))
    return out


def villain_spin_spin(phi, out=None):
    """Spin_Spin.Villain for every chain -> complex128 (chains, N, N)."""
    lib = _lib.load()
    chains, N = _fields_shape(phi, 'phi', 1)
    if out is None:
        out = torch.empty((chains, N, N, 2), dtype=torch.float64, device=phi.device)
    _lib.check(lib.svb_villain_spin_spin(_dev(phi, 'phi', (torch.float64, torch.float32)), _DTYPES[phi.dtype],
                                         chains, N, _dev(out, 'out', (torch.float64,), (chains, N, N, 2)), _stream()))
    return torch.view_as_complex(out)


def correlation(kind, field, W=1, out=None):
    """Translation-averaged correlator of every chain -> complex128 (chains, N, N).

    kind 'spin' (field = phi), 'winding' (field = n, int32) or 'vortex' (field = v, int32; needs W).
    """
    lib = _lib.load()
    kinds = {'spin': (_lib.CORR_SPIN, 1), 'winding': (_lib.CORR_WINDING, 2), 'vortex': (_lib.CORR_VORTEX, 1)}
    if kind not in kinds:
        raise ValueError(f'unknown correlator {kind!r}')
    code, comps = kinds[kind]
    chains, N = _fields_shape(field, 'field', comps)
    dtypes = (torch.float64, torch.float32) if kind == 'spin' else (torch.int32,)
    if out is None:
        out = torch.empty((chains, N, N, 2), dtype=torch.float64, device=field.device)
    _lib.check(lib.svb_correlation(code, _dev(field, 'field', dtypes), _DTYPES[field.dtype], chains, N, int(W),
                                   _dev(out, 'out', (torch.float64,), (chains, N, N, 2)), _stream()))
    return torch.view_as_complex(out)


def villain_draws(chains, N, *, W=1, interval_phi=math.pi, interval_n=1, seed=0, sweep=0, chain0=0, device='cuda'):
    """The Philox draw mapping evaluated on the device (for tests): u, dphi (chains,N,N), dn (chains,4,N,N)."""
    lib = _lib.load()
    u = torch.empty((chains, N, N), dtype=torch.float64, device=device)
    dphi = torch.empty_like(u)
    dn = torch.empty((chains, 4, N, N), dtype=torch.int32, device=device)
    _lib.check(lib.svb_villain_draws(chains, N, int(W), float(interval_phi), int(interval_n), int(seed) & (2**64 - 1),
                                     int(sweep), int(chain0), u.data_ptr(), dphi.data_ptr(), dn.data_ptr(), _stream()))
    return u, dphi, dn


def philox4x32_10(ctr, key):
    """One Philox4x32-10 block on the host (known-answer tests)."""
    lib = _lib.load()
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib.svb_philox4x32_10_host(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def autocorrelation(data, mean=None, *, want_C=True):
    """supervillain.analysis.autocorrelation for every row of `data` (series, T) float64 on the device
    (svb_autocorrelation) -> (C (series, T) or None, tau (series,) int32; -1 where the series does not fluctuate)."""
    lib = _lib.load()
    if data.dim() != 2:
        raise ValueError(f'data must have shape (series, T); got {tuple(data.shape)}')
    series, T = int(data.shape[0]), int(data.shape[1])
    p_data = _dev(data, 'data', (torch.float64,))
    C = torch.empty((series, T), dtype=torch.float64, device=data.device) if want_C else None
    tau = torch.empty((series,), dtype=torch.int32, device=data.device)
    _lib.check(lib.svb_autocorrelation(p_data, series, T, _opt(mean, 'mean', (torch.float64,), (series,)),
                                       None if C is None else C.data_ptr(), tau.data_ptr(), _stream()))
    return C, tau


def block_mean(data, width, drop=0, weight=None):
    """Blocking._block (analysis/blocking.py:54-66) for every row of `data` (series, T) float64 on the device
    (svb_block_mean) -> (series, (T - drop) // width)."""
    lib = _lib.load()
    if data.dim() != 2:
        raise ValueError(f'data must have shape (series, T); got {tuple(data.shape)}')
    series, T = int(data.shape[0]), int(data.shape[1])
    width, drop = int(width), int(drop)
    if width < 1 or drop < 0 or drop > T or (T - drop) % width:
        raise ValueError(f'(T - drop) must be a multiple of width; got T={T} drop={drop} width={width}')
    out = torch.empty((series, (T - drop) // width), dtype=torch.float64, device=data.device)
    _lib.check(lib.svb_block_mean(_dev(data, 'data', (torch.float64,)), _opt(weight, 'weight', (torch.float64,), (T,)), series, T,
                                  width, drop, out.data_ptr(), _stream()))
    return out


def bootstrap_mean(data, indices, weight=None):
    """Bootstrap._resample (analysis/bootstrap.py:57-67) for every row of `data` (series, T) float64 on the device
    (svb_bootstrap_mean); `indices` (T, draws) int64 as numpy drew them -> (series, draws)."""
    lib = _lib.load()
    if data.dim() != 2:
        raise ValueError(f'data must have shape (series, T); got {tuple(data.shape)}')
    series, T = int(data.shape[0]), int(data.shape[1])
    if indices.dim() != 2 or int(indices.shape[0]) != T:
        raise ValueError(f'indices must have shape (T, draws) = ({T}, draws); got {tuple(indices.shape)}')
    draws = int(indices.shape[1])
    out = torch.empty((series, draws), dtype=torch.float64, device=data.device)
    _lib.check(lib.svb_bootstrap_mean(_dev(data, 'data', (torch.float64,)), _opt(weight, 'weight', (torch.float64,), (T,)), series, T,
                                      _dev(indices, 'indices', (torch.int64,)), draws, out.data_ptr(), _stream()))
    return out


TAXI_SPIN, TAXI_VORTEX = 0, 1


def taxicab_correlator(kind, links, kappa, kappa_chain=None):
    """The taxicab reweighting observables from the gauge-invariant links (chains, 2, N, N) float64 -> (chains, N, N)
    (svb_taxicab_correlator): kind 'spin' = Spin_Spin.Worldline (links = m - delta(v)/W, observable/spin.py:50-224),
    kind 'vortex' = Vortex_Vortex.Villain (links = d(phi) - 2 pi n, observable/vortex.py:63-189)."""
    lib = _lib.load()
    chains, N = _fields_shape(links, 'links', 2)
    out = torch.empty((chains, N, N), dtype=torch.float64, device=links.device)
    _lib.check(lib.svb_taxicab_correlator({'spin': TAXI_SPIN, 'vortex': TAXI_VORTEX}[kind], _dev(links, 'links', (torch.float64,)),
                                          chains, N, float(kappa), _opt(kappa_chain, 'kappa_chain', (torch.float64,), (chains,)),
                                          out.data_ptr(), _stream()))
    return out


def worldline_spin_spin(m, v, kappa, W=1, kappa_chain=None):
    """Spin_Spin.Worldline for every chain: Links.Worldline = m - delta(v)/W (observable/links.py:35-45) by the form kernel,
    then the taxicab reweighting."""
    links = m.to(torch.float64)
    if W != float('inf'):
        links = links - form_op('delta', 2, v).to(torch.float64) / float(W)
    return taxicab_correlator('spin', links.contiguous(), kappa, kappa_chain)


def villain_vortex_vortex(phi, n, kappa, kappa_chain=None):
    """Vortex_Vortex.Villain for every chain: Links.Villain = d(phi) - 2 pi n (observable/links.py:18-32), then the taxicab
    reweighting."""
    links = form_op('d', 0, phi.to(torch.float64)) - (2 * math.pi) * n.to(torch.float64)
    return taxicab_correlator('vortex', links.contiguous(), kappa, kappa_chain)
