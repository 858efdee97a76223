"""`Ensemble` (the reference's per-chain driver) and `BatchedEnsemble` (thousands of resident chains).

`Ensemble` mirrors supervillain/ensemble.py:17-336 for the calls on the hot path: `generate`,
`continue_from`, `cut`, `every`, field access by attribute.  It advances ONE chain through the
generator protocol, so any generator of this package -- or of the reference -- can drive it.

`BatchedEnsemble` is the fast path the reference does not have: a chain
axis.  All chains live in HBM for the whole run; one kernel launch advances every chain by
`sweeps_per_step` sweeps and reduces the observables; only the per-chain observable record (a few
doubles per chain) crosses PCIe per step, and configurations only at the stride asked for.
"""
import json
import os

import numpy as np
import torch

from . import ops
from ._lib import VOBS_COUNT, WOBS_COUNT
from .batch import Batch, Configurations
from .generator.combining import KeepEvery
from .generator.villain import villain_inline_values
from .generator.worldline import worldline_inline_values
from .lattice import Form


def _no_op(x, **kwargs):
    return x


class Ensemble:
    """An ensemble of configurations of one Markov chain, importance-sampled according to `action`."""

    def __init__(self, action):
        self.Action = action

    def from_configurations(self, configurations):
        self.configuration = configurations
        return self

    def generate(self, steps, generator, start='cold', progress=_no_op, starting_index=0, index_stride=1):
        self.configuration = self.Action.configurations(steps)
        self.configuration |= generator.inline_observables(steps)
        self.index_stride = index_stride
        self.index = Batch(starting_index + index_stride * np.arange(steps))
        self.weight = Batch(np.ones(steps))
        if isinstance(start, str) and start == 'cold':
            seed = self.Action.configurations(1)[0]
        elif type(start) is dict:
            seed = start
        else:
            raise ValueError(f'Not sure how to transform a {type(start)} into a starting configuration.')
        self.configuration[0] = generator.step(seed)
        for k in progress(range(1, steps), desc='Generation'):
            self.configuration[k] = generator.step(self.configuration[k - 1])
        self.start = start
        self.generator = generator
        return self

    @classmethod
    def continue_from(cls, ensemble, steps, progress=_no_op):
        if not isinstance(ensemble, Ensemble):
            raise ValueError('ensemble should be a supervillain_b200.Ensemble.')
        try:
            generator, action = ensemble.generator, ensemble.Action
            last = ensemble.configuration[-1]
            index = ensemble.index[-1] + ensemble.index_stride
        except Exception:
            raise ValueError('The ensemble must provide a generator, an Action, and at least one configuration.')
        return Ensemble(action).generate(steps, generator, last, progress=progress, starting_index=index,
                                         index_stride=ensemble.index_stride)

    def __len__(self):
        return len(self.configuration)

    def _sub(self, sl, stride=1):
        e = Ensemble(self.Action).from_configurations(self.configuration[sl])
        e.index = self.index[sl]
        e.index_stride = self.index_stride * stride
        e.weight = self.weight[sl]
        return e

    def cut(self, start):
        e = self._sub(slice(start, None))
        e.generator = self.generator
        return e

    def every(self, stride):
        e = self._sub(slice(None, None, stride), stride)
        e.generator = KeepEvery(stride, self.generator, blocked_inline=False)
        return e

    def __getattr__(self, name):
        if name in ('configuration', 'Action'):
            raise AttributeError(name)
        try:
            return getattr(self.configuration, name)
        except AttributeError:
            raise AttributeError(name) from None


class BatchedEnsemble:
    """`chains` independent Markov chains of one action, resident on one GPU.

    >>> S = Villain(Lattice2D(32), kappa=0.5)
    >>> E = BatchedEnsemble(S, chains=4096).generate(1000, NeighborhoodUpdate(S), start='cold', sweeps_per_step=10)
    >>> E.ActionDensity.shape      # (chains, steps)
    """

    def __init__(self, action, chains, *, device=None, dtype=torch.float64, chain0=0):
        self.Action = action
        self.chains = int(chains)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        self.chain0 = int(chain0)      # global id of the first chain: shards of one job draw disjoint Philox streams
        self.kind = type(action).__name__
        if self.kind not in ('Villain', 'Worldline'):
            raise ValueError('BatchedEnsemble needs a Villain or Worldline action')
        self.fields = None

    # -- starting configurations ---------------------------------------------------------------
    def _start(self, start, seed):
        N, C = self.Action.Lattice.N, self.chains
        dev = self.device
        if isinstance(start, dict):
            keys = ('phi', 'n') if self.kind == 'Villain' else ('m', 'v')
            comps = (1, 2) if self.kind == 'Villain' else (2, 1)
            out = []
            for key, comp in zip(keys, comps):
                dt = self.dtype if key == 'phi' else torch.int32
                t = start[key] if isinstance(start[key], torch.Tensor) else torch.from_numpy(np.ascontiguousarray(start[key]))
                if t.dim() == 3:
                    t = t[None].expand(C, comp, N, N)
                if tuple(t.shape) != (C, comp, N, N):
                    raise ValueError(f'start[{key!r}] must have shape ({C}, {comp}, {N}, {N}) or ({comp}, {N}, {N})')
                out.append(t.to(device=dev, dtype=dt).contiguous().clone())
            return tuple(out)
        if start == 'cold':
            if self.kind == 'Villain':
                return (torch.zeros((C, 1, N, N), dtype=self.dtype, device=dev),
                        torch.zeros((C, 2, N, N), dtype=torch.int32, device=dev))
            return (torch.zeros((C, 2, N, N), dtype=torch.int32, device=dev),
                    torch.zeros((C, 1, N, N), dtype=torch.int32, device=dev))
        if start == 'hot':
            # synthetic hot start of SURVEY.md 8(d): phi ~ U(-pi, pi), n ~ integers(-2, 3);
            # worldline: m = delta t with t ~ integers(-2, 3) (so delta m = 0), v ~ integers(-2, 3)
            g = torch.Generator(device=dev)
            g.manual_seed(int(seed))
            if self.kind == 'Villain':
                phi = (torch.rand((C, 1, N, N), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
                n = torch.randint(-2, 3, (C, 2, N, N), generator=g, device=dev, dtype=torch.int32)
                return phi.to(self.dtype), n
            t = torch.randint(-2, 3, (C, 1, N, N), generator=g, device=dev, dtype=torch.int32)
            m = ops.form_op('delta', 2, t)
            v = torch.randint(-2, 3, (C, 1, N, N), generator=g, device=dev, dtype=torch.int32)
            return m, v
        raise ValueError(f'Not sure how to transform {start!r} into a starting configuration.')

    # -- generation ----------------------------------------------------------------------------
    def generate(self, steps, generator, start='cold', *, sweeps_per_step=1, keep_every=0, start_seed=0,
                 kappa_chain=None, progress=_no_op, correlate_every=0, correlators=None):
        """Advance every chain `steps` times by `sweeps_per_step` sweeps.

        keep_every: 0 keeps no configurations (observables only); k > 0 copies the fields to the host
        every k-th step into reference-layout columns `(chain, draw, C, N, N)`.
        kappa_chain: optional per-chain couplings (kappa scans), device or host array of length `chains`.
        correlate_every: k > 0 measures the two-point observables `correlators` (default: every one the formulation has on
        the FFT path -- Villain: Spin_Spin, Winding_Winding; Worldline: Vortex_Vortex) INLINE on every k-th step: the kernels
        run on the resident fields and write into device columns `self.two_point[name]` of shape (draws, chains, N, N)
        complex128; nothing crosses to the host until the column is asked for (`E.Spin_Spin` -> (chains, draws, N, N) numpy).
        """
        if self.fields is None or start != 'continue':
            self.fields = self._start(start, start_seed)
        if kappa_chain is not None and not isinstance(kappa_chain, torch.Tensor):
            kappa_chain = torch.as_tensor(np.asarray(kappa_chain, dtype=np.float64))
        if kappa_chain is not None:
            kappa_chain = kappa_chain.to(device=self.device, dtype=torch.float64).contiguous()
        self.kappa_chain = kappa_chain
        nobs = VOBS_COUNT if self.kind == 'Villain' else WOBS_COUNT
        record = torch.empty((steps, self.chains, nobs), dtype=torch.float64, device=self.device)
        kept = []
        a, b = self.fields
        two_point_fns = {}
        if correlate_every:
            N_ = self.Action.Lattice.N
            if self.kind == 'Villain':
                available = {'Spin_Spin': lambda phi, n, out: ops.villain_spin_spin(phi, out=out),
                             'Winding_Winding': lambda phi, n, out: ops.correlation('winding', n, out=out)}
            else:
                W_ = self.Action.W
                available = {'Vortex_Vortex': lambda m, v, out: ops.correlation('vortex', v, W=W_, out=out)}
            for name in (tuple(available) if correlators is None else tuple(correlators)):
                if name not in available:
                    raise NotImplementedError(f'{name} is not an inline two-point observable of the {self.kind} formulation')
                two_point_fns[name] = available[name]
            draws_2pt = steps // int(correlate_every)
            self.two_point = {name: torch.empty((draws_2pt, self.chains, N_, N_, 2), dtype=torch.float64, device=self.device)
                              for name in two_point_fns}
            self.correlate_every = int(correlate_every)
        overlapped = None
        if hasattr(generator, 'overlapped_device'):
            try:          # launches that overlap their predecessor; each step writes its own row of the record
                overlapped = generator.overlapped_device(a, b, chain0=self.chain0, kappa_chain=kappa_chain)
            except NotImplementedError:
                overlapped = None
        scratch = torch.empty((self.chains, nobs), dtype=torch.float64, device=self.device) if overlapped is not None else None
        swapping = None
        if overlapped is None and hasattr(generator, 'swapping_device'):
            try:          # tiled path: the state alternates between two buffer pairs instead of being copied back
                swapping = generator.swapping_device(a, b, chain0=self.chain0, kappa_chain=kappa_chain)
            except NotImplementedError:
                swapping = None
        for k in progress(range(steps), desc='Generation'):
            if overlapped is not None and getattr(overlapped, 'complete_records', False):
                overlapped(sweeps_per_step, obs=record[k])
            elif overlapped is not None:
                # the state columns of draw k - 1 ride along with launch k (they describe the chains as they arrive);
                # launch k's own counters go to row k
                overlapped(sweeps_per_step, obs=record[k], obs_in=record[k - 1] if k else scratch)
            elif swapping is not None:
                a, b = swapping(sweeps_per_step, obs=record[k])
            else:
                generator.sweep_device(a, b, sweeps_per_step, obs=record[k], chain0=self.chain0, kappa_chain=kappa_chain)
            if keep_every and (k + 1) % keep_every == 0:
                kept.append((a.cpu().numpy(), b.cpu().numpy()))          # reads only: stream order suffices
            if two_point_fns and (k + 1) % correlate_every == 0:
                # ordinary launches that only READ the fields: they wait for the sweep by stream order, and the next sweep launch
                # (also an overlapped one: its predecessor in the stream is now this kernel) waits for them
                for name, fn in two_point_fns.items():
                    fn(a, b, self.two_point[name][(k + 1) // correlate_every - 1])
        if overlapped is not None and steps and not getattr(overlapped, 'complete_records', False):
            last = ops.villain_observables(a, b, self.Action.kappa, kappa_chain=kappa_chain)     # the final state's columns
            record[steps - 1, :, :4] = last[:, :4]
        self.fields = (a, b)
        self.record = record.cpu().numpy().transpose(1, 0, 2)          # (chains, steps, nobs): ONE D2H
        self.steps = steps
        self.sweeps_per_step = sweeps_per_step
        self.generator = generator
        self.keep_every = int(keep_every)
        self.index = sweeps_per_step * (1 + np.arange(steps))
        if kept:
            names = ('phi', 'n') if self.kind == 'Villain' else ('m', 'v')
            self.configuration = {
                names[0]: np.stack([x[0] for x in kept], axis=1),
                names[1]: np.stack([x[1] for x in kept], axis=1).astype(np.int64),
            }
        N = self.Action.Lattice.N
        kappa = self.Action.kappa if kappa_chain is None else kappa_chain.cpu().numpy()[:, None]
        values = villain_inline_values if self.kind == 'Villain' else worldline_inline_values
        self.observables = values(self.record, N, kappa)
        account = getattr(generator, '_count', None)       # single GPU generators keep the reference's counters
        if account is not None:
            flat = self.record.reshape(-1, self.record.shape[-1])
            account(flat, self.chains, steps * sweeps_per_step)
        return self

    # -- continuation and checkpoints ------------------------------------------------------------
    @classmethod
    def continue_from(cls, ensemble, steps, *, keep_every=0, progress=_no_op):
        """The batched form of `Ensemble.continue_from` (supervillain/ensemble.py:103-142): `steps` more samples of every
        chain from the last state of `ensemble` (a BatchedEnsemble or the path of a checkpoint written by `save`) with its
        generator.  The generator's Philox state is (seed, sweep counter), so the continuation is the run that was never
        interrupted: fields and records are bit for bit those of one longer `generate`."""
        e = cls.load(ensemble) if isinstance(ensemble, (str, os.PathLike)) else ensemble
        if not isinstance(e, BatchedEnsemble) or e.fields is None or 'generator' not in e.__dict__:
            raise ValueError('The ensemble must be a BatchedEnsemble (or a checkpoint of one) that has generated at least one sample.')
        new = cls(e.Action, e.chains, device=e.device, dtype=e.dtype, chain0=e.chain0)
        new.fields = tuple(f.clone() for f in e.fields)
        new.generate(steps, e.generator, start='continue', sweeps_per_step=e.sweeps_per_step, keep_every=keep_every,
                     kappa_chain=e.__dict__.get('kappa_chain'), progress=progress)
        new.index = int(e.index[-1]) + new.index
        return new

    def save(self, path):
        """Checkpoint: the resident fields, the records, the sample index and the (action, generator) pair -- whose state is
        the Philox (seed, counter) -- in one .npz.  Everything is plain data: the action and the generator are stored as JSON
        (class names from a whitelist, constructor arguments, counters; generator/_state.py), so reading a checkpoint never
        executes code.  (The reference stores ensembles as HDF5, supervillain/h5/; h5py is not part of this environment.)"""
        if self.fields is None:
            raise ValueError('nothing to save: generate first')
        from .generator import _state
        state = {'format': 1, 'action': _state.describe_action(self.Action),
                 'generator': _state.describe_generator(self.__dict__.get('generator'))}
        blob = np.frombuffer(json.dumps(state).encode('utf-8'), dtype=np.uint8)
        arrays = dict(field0=self.fields[0].cpu().numpy(), field1=self.fields[1].cpu().numpy(), record=self.record, index=self.index,
                      meta=np.array([self.chains, self.chain0, self.sweeps_per_step, self.steps, self.__dict__.get('keep_every', 0)],
                                    dtype=np.int64), state_json=blob)
        if self.__dict__.get('kappa_chain') is not None:
            arrays['kappa_chain'] = self.kappa_chain.cpu().numpy()
        for k, v in self.__dict__.get('configuration', {}).items():
            arrays['configuration_' + k] = v
        with open(path, 'wb') as f:
            np.savez(f, **arrays)

    @classmethod
    def load(cls, path, device=None):
        """A BatchedEnsemble as `save` left it (fields back on the device), ready for `continue_from`.  The action and the
        generator are rebuilt from plain JSON through a whitelist of classes; a checkpoint of an earlier build (which held
        a pickle) is refused rather than unpickled."""
        from .generator import _state
        with np.load(path, allow_pickle=False) as z:
            if 'state_json' not in z.files:
                raise ValueError('this checkpoint holds a pickled (action, generator) pair from an earlier build; it is not read '
                                 '(unpickling executes code) -- regenerate it with this version')
            state = json.loads(z['state_json'].tobytes().decode('utf-8'))
            action = _state.rebuild_action(state['action'])
            generator = _state.rebuild_generator(state['generator'], action)
            chains, chain0, sweeps_per_step, steps = (int(x) for x in z['meta'][:4])
            keep_every = int(z['meta'][4]) if len(z['meta']) > 4 else 0       # checkpoints of earlier builds have four entries
            e = cls(action, chains, device=device, chain0=chain0)
            e.fields = tuple(torch.from_numpy(z[k]).to(e.device) for k in ('field0', 'field1'))
            e.dtype = e.fields[0].dtype if e.kind == 'Villain' else e.dtype
            e.record, e.index = z['record'], z['index']
            e.steps, e.sweeps_per_step, e.generator = steps, sweeps_per_step, generator
            e.keep_every = keep_every
            e.kappa_chain = torch.from_numpy(z['kappa_chain']).to(e.device) if 'kappa_chain' in z.files else None
            cfg = {k[len('configuration_'):]: z[k] for k in z.files if k.startswith('configuration_')}
            if cfg:
                e.configuration = cfg
        N = action.Lattice.N
        kappa = action.kappa if e.kappa_chain is None else e.kappa_chain.cpu().numpy()[:, None]
        e.observables = (villain_inline_values if e.kind == 'Villain' else worldline_inline_values)(e.record, N, kappa)
        return e

    def to_reference(self, chain, supervillain=None):
        """One chain as an ensemble OF THE REFERENCE PACKAGE, so that everything downstream of generation -- `to_h5` in the
        reference's HDF5 layout (supervillain/h5/), `Ensemble.from_h5`, the observables and `analysis/` --
        consumes it unchanged.  Needs kept configurations (`generate(..., keep_every=k)`) and the reference importable
        (pass the module as `supervillain` or have it on sys.path); nothing of the reference is used on the sampling path.

        The result is what `supervillain.Ensemble(S).generate(draws, KeepEvery(k * sweeps_per_step, G))` stores
        (ensemble.py:74-78, 94-95): `configuration.fields` = the field columns (phi float64 / integer fields int64, as Forms)
        plus the inline scalar observables of the kept draws under the reference's observable names, so the reference reads
        them instead of measuring again (observable/observable.py:49-54); `index` counts sweeps, `index_stride` is the
        sweeps between kept draws, `weight` is one."""
        if 'configuration' not in self.__dict__:
            raise ValueError('no configurations were kept; call generate(..., keep_every=k)')
        chain = int(chain)
        if not 0 <= chain < self.chains:
            raise IndexError(f'chain {chain} is not one of {self.chains}')
        if supervillain is None:
            import importlib
            supervillain = importlib.import_module('supervillain')
        sv = supervillain
        A = self.Action
        L = sv.lattice.Lattice2D(A.Lattice.N)
        kappa = float(A.kappa if self.__dict__.get('kappa_chain') is None else self.kappa_chain[chain].item())
        S = getattr(sv.action, self.kind)(L, kappa, int(A.W))
        names = ('phi', 'n') if self.kind == 'Villain' else ('m', 'v')
        draws = self.configuration[names[0]].shape[1]
        k = self.__dict__.get('keep_every', 0) or self.steps // max(draws, 1)
        kept = k * (1 + np.arange(draws)) - 1                     # the steps whose configurations were kept
        cfgs = S.configurations(draws)
        for name in names:
            column = cfgs.fields[name]
            for t in range(draws):
                column[t] = self.configuration[name][chain, t]     # Batch.__setitem__ checks the cast is lossless (batch.py:206-227)
        inline = {}
        for name, values in self.observables.items():
            v = np.asarray(values)[chain]
            if v.shape[0] != self.steps:
                continue
            inline[name] = sv.batch.Batch(np.ascontiguousarray(v[kept]))
        cfgs |= inline
        E = sv.Ensemble(S).from_configurations(cfgs)
        stride = k * self.sweeps_per_step
        E.index_stride = stride
        E.index = sv.batch.Batch(np.asarray(self.index)[kept].astype(np.int64))
        E.weight = sv.batch.Batch(np.ones(draws))
        E.start = 'cold'
        return E

    def __len__(self):
        """The number of recorded samples per chain (what `len(Ensemble)` is for the reference's single chain)."""
        return int(self.__dict__.get('steps', 0))

    def __getattr__(self, name):
        obs = self.__dict__.get('observables')
        if obs is not None and name in obs:
            return obs[name]
        two = self.__dict__.get('two_point')
        if two is not None and name in two:             # (draws, chains, N, N, 2) on the device -> (chains, draws, N, N) complex
            return torch.view_as_complex(two[name]).cpu().numpy().transpose(1, 0, 2, 3)
        raise AttributeError(name)

    def measure(self, name):
        """A two-point observable of every kept configuration (needs keep_every > 0) -> (chains, draws, N, N), computed on
        the device, named as the reference names them:
          Villain    Spin_Spin (spin.py:28-42, complex), Winding_Winding (winding.py:77-86, complex),
                     Vortex_Vortex (vortex.py:63-189, the taxicab reweighting)
          Worldline  Vortex_Vortex (vortex.py:22-37, complex), Spin_Spin (spin.py:50-224, the taxicab reweighting)"""
        from . import ops
        if 'configuration' not in self.__dict__:
            raise ValueError('no configurations were kept; call generate(..., keep_every=k)')
        kappa, W = self.Action.kappa, self.Action.W
        cfg = self.configuration
        dev = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x)).to(device=self.device, dtype=dt)
        if self.kind == 'Villain':
            fn = {'Spin_Spin': lambda phi, n: ops.villain_spin_spin(phi),
                  'Winding_Winding': lambda phi, n: ops.correlation('winding', n),
                  'Vortex_Vortex': lambda phi, n: ops.villain_vortex_vortex(phi, n, kappa)}.get(name)
            fields = ('phi', torch.float64), ('n', torch.int32)
        else:
            fn = {'Vortex_Vortex': lambda m, v: ops.correlation('vortex', v, W=W),
                  'Spin_Spin': lambda m, v: ops.worldline_spin_spin(m, v, kappa, W=W)}.get(name)
            fields = ('m', torch.int32), ('v', torch.int32)
        if fn is None:
            raise NotImplementedError(f'{name} is not a two-point observable of the {self.kind} formulation on this path')
        draws = cfg[fields[0][0]].shape[1]
        out = [fn(*(dev(cfg[k][:, t], dt) for k, dt in fields)).cpu().numpy() for t in range(draws)]
        return np.stack(out, axis=1)

    def autocorrelation_time(self, observables=None, every=False):
        """The batched form of `Ensemble.autocorrelation_time` (supervillain/ensemble.py:184-239): the integrated
        autocorrelation time of every scalar observable column of every chain, computed on the device
        (`svb_autocorrelation`), maximised over chains.  Columns that do not fluctuate enough are left out, as the
        reference does; with nothing left, half the length.  every=True: a dict name -> (chains,) array of times
        (-1 where a chain's column does not fluctuate)."""
        from . import ops
        names = [n for n, v in self.observables.items() if v.ndim == 2] if observables is None else list(observables)
        auto = {}
        for name in names:
            col = np.ascontiguousarray(getattr(self, name), dtype=np.float64)
            if col.ndim != 2:
                continue
            _, tau = ops.autocorrelation(torch.from_numpy(col).to(self.device), want_C=False)
            auto[name] = tau.cpu().numpy().astype(np.int64)
        if every:
            return auto
        times = [int(t[t >= 0].max()) for t in auto.values() if (t >= 0).any()]
        return max(times) if times else int(np.ceil(self.steps / 2))

    def chain(self, c):
        """Chain `c` as a reference-layout `Ensemble` (needs keep_every > 0): fields `(draw, C, N, N)` Batches
        plus the inline observable columns, exactly what `Ensemble.generate` would have stored."""
        if 'configuration' not in self.__dict__:
            raise ValueError('no configurations were kept; call generate(..., keep_every=k)')
        L = self.Action.Lattice
        degree = {'phi': 0, 'n': 1, 'm': 1, 'v': 2}
        cols = {k: Batch(v[c], cls=Form, degree=degree[k], lattice=L) for k, v in self.configuration.items()}
        draws = len(next(iter(cols.values())))
        stride = self.__dict__.get('keep_every', 0) or self.steps // max(draws, 1)
        kept = stride * (1 + np.arange(draws)) - 1                  # the steps whose configurations were kept
        for name, v in self.observables.items():
            cols[name] = Batch(np.ascontiguousarray(v[c][kept]))
        e = Ensemble(self.Action).from_configurations(Configurations(cols))
        e.index = Batch(np.asarray(self.index)[kept])
        e.index_stride = stride * self.sweeps_per_step
        e.weight = Batch(np.ones(len(e.index)))
        e.generator = self.generator
        return e
