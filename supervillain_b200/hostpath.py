"""Host-buffer stepping: the `generator.step(cfg)` contract at batch scale.

The reference's `step` takes host arrays and returns host arrays (neighborhood.py:59-137).  For a
batch of thousands of chains that means, per step, one host->device copy of the fields, the sweep,
and one device->host copy of the fields plus the per-chain observable record.  `HostStepper` does
exactly that on pinned host buffers, split into chunks of chains pipelined over CUDA streams so the
H2D copy of chunk i+1, the sweep of chunk i and the D2H copy of chunk i-1 overlap (PCIe is
full-duplex).  The kernels are the same ones the resident path uses.
"""
import ctypes

import torch

from . import _lib
from ._lib import VOBS_COUNT, WOBS_COUNT


class HostStepper:
    def __init__(self, generator, chains, *, chain0=0, chunks=16, streams=8, device=None):
        self.generator = generator
        self.chains = int(chains)
        self.chain0 = int(chain0)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.kind = type(generator.Action).__name__
        N = generator.Lattice.N
        self.N = N
        chunks = max(1, min(int(chunks), self.chains))
        bounds = [round(i * self.chains / chunks) for i in range(chunks + 1)]
        self.bounds = [(lo, hi) for lo, hi in zip(bounds[:-1], bounds[1:]) if hi > lo]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, int(streams)))]
        if self.kind == 'Villain':
            self.comps, self.dtypes, self.nobs = (1, 2), (getattr(generator, 'dtype', torch.float64), torch.int32), VOBS_COUNT
        else:
            self.comps, self.dtypes, self.nobs = (2, 1), (torch.int32, torch.int32), WOBS_COUNT
        self.dev_a = torch.empty((self.chains, self.comps[0], N, N), dtype=self.dtypes[0], device=self.device)
        self.dev_b = torch.empty((self.chains, self.comps[1], N, N), dtype=self.dtypes[1], device=self.device)
        self.dev_obs = torch.empty((self.chains, self.nobs), dtype=torch.float64, device=self.device)
        self._host_obs = [torch.empty((self.chains, self.nobs), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        self._flip = 0
        self.host_obs = self._host_obs[0]
        self.h2d_bytes = self.dev_a.numel() * self.dev_a.element_size() + self.dev_b.numel() * self.dev_b.element_size()
        self.d2h_bytes = self.h2d_bytes + self.host_obs.numel() * 8

    def pinned_fields(self, from_device=None):
        """Allocate pinned host buffers for the two fields, optionally filled from device tensors."""
        a = torch.empty(tuple(self.dev_a.shape), dtype=self.dtypes[0], pin_memory=True)
        b = torch.empty(tuple(self.dev_b.shape), dtype=self.dtypes[1], pin_memory=True)
        if from_device is not None:
            a.copy_(from_device[0])
            b.copy_(from_device[1])
        return a, b

    class Pending:
        """A step in flight (`step_async`): `wait()` returns the pinned observable record once the host fields are final."""

        def __init__(self, events, record):
            self.events, self.record = events, record

        def wait(self):
            for e in self.events:
                e.synchronize()
            return self.record

    def step(self, host_a, host_b, n_sweeps=1):
        """One step on host fields, IN PLACE in the pinned host buffers; returns the pinned observable record.

        host_a, host_b: (phi, n) for a Villain generator, (m, v) for a worldline generator, pinned,
        shaped (chains, C, N, N) with the device dtypes (float64/float32 phi, int32 integer fields).
        """
        return self.step_async(host_a, host_b, n_sweeps).wait()

    def step_async(self, host_a, host_b, n_sweeps=1):
        """`step` without the final wait: returns a `Pending`.  Issue the next step on ANOTHER pair of host buffers before
        waiting for this one and its copies overlap this one's (chunk i of consecutive steps shares a stream, so the
        device staging buffers are reused in order); at most two steps should be in flight (two observable records)."""
        for t, d in ((host_a, self.dev_a), (host_b, self.dev_b)):
            if tuple(t.shape) != tuple(d.shape) or t.dtype != d.dtype or t.is_cuda:
                raise ValueError(f'host field must be a CPU tensor of shape {tuple(d.shape)} and dtype {d.dtype}')
            if not t.is_pinned():
                raise ValueError('host fields must be pinned (use HostStepper.pinned_fields)')
        G = self.generator
        sweep0 = G.counter
        current = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(current)
        if self.kind == 'Villain' and getattr(G, 'rng', None) is None and getattr(G, 'path', 'auto') == 'auto':
            # the whole chunked H2D -> sweep -> D2H pipeline is issued by one C call (svb_villain_sweep_host)
            lib = _lib.load()
            handles = (ctypes.c_void_p * len(self.streams))(*[s.cuda_stream for s in self.streams])
            _lib.check(lib.svb_villain_sweep_host(
                host_a.data_ptr(), _lib.F64 if self.dtypes[0] == torch.float64 else _lib.F32, host_b.data_ptr(),
                self.host_obs.data_ptr(), self.dev_a.data_ptr(), self.dev_b.data_ptr(), self.dev_obs.data_ptr(),
                self.chains, self.N, float(G.kappa), int(G.Action.W), float(G.interval_phi), int(G.interval_n),
                int(n_sweeps), int(G.seed) & (2**64 - 1), int(sweep0), self.chain0,
                _lib.ARITH_FAST if G.arithmetic == 'fast' else _lib.ARITH_STRICT,
                len(self.bounds), handles, len(self.streams)))
            G.counter = sweep0 + n_sweeps
            return self._pending()
        for i, (lo, hi) in enumerate(self.bounds):
            st = self.streams[i % len(self.streams)]
            with torch.cuda.stream(st):
                self.dev_a[lo:hi].copy_(host_a[lo:hi], non_blocking=True)
                self.dev_b[lo:hi].copy_(host_b[lo:hi], non_blocking=True)
                G.counter = sweep0
                G.sweep_device(self.dev_a[lo:hi], self.dev_b[lo:hi], n_sweeps, obs=self.dev_obs[lo:hi],
                               chain0=self.chain0 + lo)
                host_a[lo:hi].copy_(self.dev_a[lo:hi], non_blocking=True)
                host_b[lo:hi].copy_(self.dev_b[lo:hi], non_blocking=True)
                self.host_obs[lo:hi].copy_(self.dev_obs[lo:hi], non_blocking=True)
        G.counter = sweep0 + n_sweeps
        return self._pending()

    def _pending(self):
        events = []
        for s in self.streams:
            e = torch.cuda.Event()
            e.record(s)
            events.append(e)
        record = self.host_obs
        self._flip ^= 1
        self.host_obs = self._host_obs[self._flip]
        return HostStepper.Pending(events, record)
