"""Worldline generators on the GPU: `PlaquetteUpdate`, `VortexUpdate`, `CoexactUpdate`.

Drop-ins for supervillain.generator.worldline.{PlaquetteUpdate, VortexUpdate, CoexactUpdate}
(worldline/plaquette.py:9-113, worldline/vortex.py:13-209, worldline/coexact.py:13-197).

`PlaquetteUpdate` makes the reference's move -- (dm on the four boundary links, dv) per plaquette,
dm in {-1,+1}, dv in {-1,0,+1}, dS = df/kappa (f1+f2-f3-f4+2 df) -- but visits the plaquettes in
red/black checkerboard order instead of the reference's random sequential order
(plaquette.py:63): an equally valid Markov chain with the same stationary distribution, NOT the
same chain draw for draw.  `VortexUpdate` and `CoexactUpdate` are already checkerboarded in the
reference, and with an injected numpy `rng` they reproduce the reference chains bit for bit.
"""
import numpy as np
import torch

from .. import ops
from .._lib import (WOBS_ACCEPTANCE, WOBS_ACCEPTED, WOBS_COUNT, WOBS_SUM_DF2, WOBS_SUM_F2, WOBS_WRAP0, WOBS_WRAP1)
from ..action import to_device
from ..batch import Batch
from ..lattice import Form
from . import _replay
from .generator import Generator, fresh_seed

INLINE_NAMES = ('ActionDensity', 'InternalEnergyDensity', 'InternalEnergyDensitySquared', 'WindingSquared',
                'TorusWrapping', 'WrappingSquared', 'Vortex_Vortex')
# the two-point observable among them: (N, N) complex per draw, evaluated on the device from the v the sweep has just left there
# (observable/vortex.py:22-37 via Lattice.correlation, compact.py:465-536)
TWO_POINT = ('Vortex_Vortex',)


def _is_worldline(action):
    return type(action).__name__ == 'Worldline' and all(hasattr(action, a) for a in ('Lattice', 'kappa', 'W'))


def worldline_inline_values(rec, N, kappa):
    """Reference worldline observables (observable/action.py:35-47, energy.py:68-102, winding.py:40-52,
    wrapping.py:28-39,58-59) from a per-chain device record."""
    sites, nlinks = N * N, 2 * N * N
    f2 = rec[..., WOBS_SUM_F2]
    p1 = (nlinks / 2 - 0.5 / kappa * f2) / kappa
    p2 = (f2 / kappa - nlinks / 2) / kappa ** 2
    wrap = np.stack([rec[..., WOBS_WRAP0], rec[..., WOBS_WRAP1]], axis=-1) / N
    return {
        'ActionDensity': (nlinks / 2 - 0.5 / kappa * f2) / sites,
        'InternalEnergyDensity': (nlinks / 2 - 0.5 / kappa * f2) / (sites * kappa),
        'InternalEnergyDensitySquared': (p1 ** 2 - p2) / sites ** 2,
        'WindingSquared': 1 / (np.pi ** 2 * kappa) - (rec[..., WOBS_SUM_DF2] / sites) / (2 * np.pi * kappa) ** 2,
        'TorusWrapping': wrap,
        'WrappingSquared': (wrap ** 2).sum(axis=-1),
    }


class _CheckerboardWorldline(Generator):
    mode = None
    noun = None

    def __init__(self, action, interval=1, *, seed=None, inline=(), path='auto'):
        if not _is_worldline(action):
            raise ValueError(self._wrong_action)
        if not (action.W < float('inf')):
            raise NotImplementedError('the GPU worldline generators implement finite W (integer v) only')
        self.Action = action
        self.Lattice = action.Lattice
        self.kappa = action.kappa
        self.interval = interval
        self.rng = None
        self.seed = fresh_seed() if seed is None else int(seed)
        self.counter = 0
        self.path = path
        unknown = set(inline) - set(INLINE_NAMES)
        if unknown:
            raise ValueError(f'unknown inline observables {sorted(unknown)}')
        self.inline = tuple(inline)
        self.accepted = 0
        self.proposed = 0
        self.acceptance = 0.
        self.sweeps = 0

    def sweep_device(self, m, v, n_sweeps=1, *, obs=None, chain0=0, kappa_chain=None, injected=None,
                     accept_mask=None, dS_out=None):
        """`n_sweeps` sweeps in place on device tensors m (chains,2,N,N), v (chains,1,N,N), both int32."""
        ops.worldline_sweep(m, v, self.kappa, W=self.Action.W, mode=self.mode, interval=self.interval,
                            n_sweeps=n_sweeps, seed=self.seed, sweep0=self.counter, chain0=chain0, injected=injected,
                            path=self.path, kappa_chain=kappa_chain, obs=obs, accept_mask=accept_mask, dS_out=dS_out)
        if injected is None:
            self.counter += n_sweeps

    def overlapped_device(self, m, v, *, chain0=0, kappa_chain=None):
        """Overlapped-launch stepping of one resident chain set (ops.WorldlineOverlappedSweeps): returns
        `step(n_sweeps=1, obs=None)` advancing the Philox counter, with `step.fence()` for foreign writes to the fields.
        Serves W = 1 and Philox draws (interval <= 2 for vortex / coexact); raises NotImplementedError otherwise."""
        if self.rng is not None or self.Action.W != 1 or self.path != 'auto':
            raise NotImplementedError('overlapped launches serve W = 1, Philox draws, path="auto"')
        ov = ops.WorldlineOverlappedSweeps(m, v, self.kappa, mode=self.mode, interval=self.interval, seed=self.seed, chain0=chain0,
                                           kappa_chain=kappa_chain)

        def step(n_sweeps=1, obs=None, obs_in=None):
            if obs_in is not None:
                raise NotImplementedError('the worldline record is complete at the end of its own launch')
            ov.step(self.counter, n_sweeps, obs)
            self.counter += n_sweeps
        step.fence = ov.fence
        step.complete_records = True          # no obs_in pipelining
        return step

    def _injected_draws(self, chains, n_sweeps):
        N = self.Lattice.N
        u = np.empty((n_sweeps, chains, N, N))
        a = np.empty((n_sweeps, chains, N, N), dtype=np.int32)
        b = np.empty_like(a)
        for c in range(chains):
            for s in range(n_sweeps):
                u[s, c], a[s, c], b[s, c] = _replay.worldline_checkerboard(self.rng, self.Lattice, self.mode, self.interval)
        return {k: torch.from_numpy(x).cuda() for k, x in dict(u=u, a=a, b=b).items()}

    def _count(self, rec, chains, n_sweeps):
        plaquettes = self.Lattice.cells_of_degree[2]
        self.sweeps += n_sweeps * chains
        self.proposed += plaquettes * n_sweeps * chains
        self.accepted += int(round(float(rec[:, WOBS_ACCEPTED].sum())))
        self._accumulate_acceptance(float(rec[:, WOBS_ACCEPTANCE].sum()), plaquettes)

    def _accumulate_acceptance(self, total, plaquettes):
        self.acceptance += total / plaquettes          # vortex.py:132, coexact.py:123

    def step(self, cfg, n_sweeps=1):
        N = self.Lattice.N
        m, single = to_device(cfg['m'], torch.int32, 2, N)
        v, _ = to_device(cfg['v'], torch.int32, 1, N)
        if isinstance(cfg['m'], torch.Tensor):
            m = m.clone()
        if isinstance(cfg['v'], torch.Tensor):
            v = v.clone()
        chains = m.shape[0]
        injected = self._injected_draws(chains, n_sweeps) if self.rng is not None else None
        obs = torch.empty((chains, WOBS_COUNT), dtype=torch.float64, device=m.device)
        self.sweep_device(m, v, n_sweeps, obs=obs, injected=injected)
        rec = obs.cpu().numpy()
        self._count(rec, chains, n_sweeps)
        L = self.Lattice
        result = {}
        if self.mode != 'vortex':
            out_m = m.cpu().numpy().astype(np.int64)
            result['m'] = Form(out_m[0], degree=1, lattice=L) if single else out_m
        if self.mode != 'coexact':
            out_v = v.cpu().numpy().astype(np.int64)
            result['v'] = Form(out_v[0], degree=2, lattice=L) if single else out_v
        if self.inline:
            vals = worldline_inline_values(rec, N, self.kappa)
            for name in self.inline:
                if name in TWO_POINT:
                    val = ops.correlation('vortex', v, W=self.Action.W).cpu().numpy()
                else:
                    val = vals[name]
                result[name] = val[0] if single else val
        return cfg | result

    def inline_observables(self, steps):
        N = self.Lattice.N
        shapes = {'TorusWrapping': (2,), 'Vortex_Vortex': (N, N)}
        dtypes = {'Vortex_Vortex': complex}
        return {name: Batch(steps, shape=shapes.get(name, ()), dtype=dtypes.get(name, float)) for name in self.inline}

    def report(self):
        return (
            f'There were {self.accepted} {self.noun} proposals accepted of {self.proposed} proposed updates.'
            + '\n' +
            f'    {self.accepted/self.proposed:.6f} acceptance rate'
            + '\n' +
            f'    {self._mean_acceptance():.6f} average Metropolis acceptance probability.'
        )

    def _mean_acceptance(self):
        return self.acceptance / self.sweeps


class PlaquetteUpdate(_CheckerboardWorldline):
    """(dm, dv) per plaquette in checkerboard order; constructor `PlaquetteUpdate(action)` as in the reference."""
    mode = 'joint'
    noun = 'single-plaquette'
    _wrong_action = 'The PlaquetteUpdate requires the Worldline action.'

    def __init__(self, action, *, seed=None, inline=(), path='auto'):
        super().__init__(action, 1, seed=seed, inline=inline, path=path)

    def __str__(self):
        return 'PlaquetteUpdate'

    def _accumulate_acceptance(self, total, plaquettes):
        self.acceptance += total                          # plaquette.py:88: summed per proposal

    def _mean_acceptance(self):
        return self.acceptance / self.proposed            # plaquette.py:112


class VortexUpdate(_CheckerboardWorldline):
    """v-only updates, dv in [-interval_v, interval_v] minus {0}; `VortexUpdate(action, interval_v=1)`."""
    mode = 'vortex'
    noun = 'vortex'
    _wrong_action = 'Need a Worldline action'

    def __init__(self, action, interval_v=1, *, seed=None, inline=(), path='auto'):
        super().__init__(action, interval_v, seed=seed, inline=inline, path=path)
        self.interval_v = interval_v
        self.vs = _replay._nonzero_choices(interval_v)

    def __str__(self):
        return 'VortexUpdate'


class CoexactUpdate(_CheckerboardWorldline):
    """m += delta t with t in [-interval_t, interval_t] minus {0}; `CoexactUpdate(action, interval_t=1)`."""
    mode = 'coexact'
    noun = 'coexact'
    _wrong_action = 'Need a Worldline action'

    def __init__(self, action, interval_t=1, *, seed=None, inline=(), path='auto'):
        super().__init__(action, interval_t, seed=seed, inline=inline, path=path)
        self.interval_t = interval_t
        self.ts = _replay._nonzero_choices(interval_t)

    def __str__(self):
        return 'CoexactUpdate'


class WrappingUpdate(Generator):
    """Torus-cycle updates of m (supervillain/generator/worldline/wrapping.py:9-99); `WrappingUpdate(action, interval_w=1)`.

    Needed next to `PlaquetteUpdate` for ergodicity: plaquette moves cannot change the wrapping sector
    (plaquette.py:16-19; test/end-to-end.py:48-50 always pairs them).
    """

    def __init__(self, action, interval_w=1, *, seed=None):
        if not _is_worldline(action):
            raise ValueError('The WrappingUpdate requires the Worldline action.')
        if not (action.W < float('inf')):
            raise NotImplementedError('the GPU worldline generators implement finite W (integer v) only')
        self.Action = action
        self.Lattice = action.Lattice
        self.kappa = action.kappa
        self.interval_w = interval_w
        self.w = _replay._nonzero_choices(interval_w)
        self.rng = None
        self.seed = fresh_seed() if seed is None else int(seed)
        self.counter = 0
        self.accepted = 0
        self.proposed = 0
        self.sweeps = 0
        self.acceptance = 0.

    def __str__(self):
        return 'WrappingUpdate'

    def sweep_device(self, m, v, n_sweeps=1, *, obs=None, chain0=0, kappa_chain=None, injected=None, counters=None,
                     dS_out=None):
        """`n_sweeps` wrapping steps in place on m (chains,2,N,N) int32; `obs`, if given, is refreshed from the final
        state with svb_worldline_observables (a wrapping move changes every link observable)."""
        for s in range(n_sweeps):
            ops.worldline_wrapping(m, v, self.kappa, W=self.Action.W, interval=self.interval_w, seed=self.seed,
                                   sweep=self.counter + s, chain0=chain0, injected=injected, kappa_chain=kappa_chain,
                                   counters=counters, dS_out=dS_out)
        if injected is None:
            self.counter += n_sweeps
        if obs is not None:
            ops.worldline_observables(m, v, W=self.Action.W, obs=obs)

    def step(self, cfg):
        N = self.Lattice.N
        m, single = to_device(cfg['m'], torch.int32, 2, N)
        v, _ = to_device(cfg['v'], torch.int32, 1, N)
        if isinstance(cfg['m'], torch.Tensor):
            m = m.clone()
        chains = m.shape[0]
        injected = None
        if self.rng is not None:
            u = np.empty((chains, 2, N)); c = np.empty((chains, 2, N), dtype=np.int32)
            for k in range(chains):
                u[k], c[k] = _replay.worldline_wrapping(self.rng, self.Lattice, self.interval_w)
            injected = {'u': torch.from_numpy(u).cuda(), 'c': torch.from_numpy(c).cuda()}
        counters = torch.zeros((chains, 2), dtype=torch.float64, device=m.device)
        self.sweep_device(m, v, 1, injected=injected, counters=counters)
        rec = counters.cpu().numpy()
        n_cycles = 2 * N
        self.proposed += n_cycles * chains
        self.accepted += int(round(float(rec[:, 0].sum())))
        self.acceptance += float(rec[:, 1].sum()) / n_cycles
        self.sweeps += chains
        out_m = m.cpu().numpy().astype(np.int64)
        return cfg | {'m': Form(out_m[0], degree=1, lattice=self.Lattice) if single else out_m}

    def report(self):
        return (
            f'There were {self.accepted} single-wrapping proposals accepted of {self.proposed} proposed updates.'
            + '\n' +
            f'    {self.accepted / self.proposed:.6f} acceptance rate'
            + '\n' +
            f'    {self.acceptance / self.sweeps:.6f} average Metropolis acceptance probability.'
        )
