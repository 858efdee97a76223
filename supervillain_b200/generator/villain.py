"""`NeighborhoodUpdate` on the GPU, drop-in for supervillain.generator.villain.NeighborhoodUpdate
(supervillain/generator/villain/neighborhood.py:12-150)."""
import numpy as np
import torch

from .. import ops
from .._lib import VOBS_ACCEPTANCE, VOBS_ACCEPTED, VOBS_ACTION, VOBS_COUNT, VOBS_SUM_DN2, VOBS_WRAP0, VOBS_WRAP1
from ..action import to_device
from ..batch import Batch
from ..lattice import Form
from . import _replay
from .generator import Generator, fresh_seed

INLINE_NAMES = ('ActionDensity', 'InternalEnergyDensity', 'InternalEnergyDensitySquared', 'WindingSquared',
                'TorusWrapping', 'WrappingSquared', 'Spin_Spin', 'Winding_Winding')
# the two-point observables among them: (N, N) complex per draw, evaluated on the device from the fields the sweep has just left
# there (observable/spin.py:28-42, observable/winding.py:77-86 via Lattice.correlation, compact.py:465-536)
TWO_POINT = {'Spin_Spin': lambda phi, n: ops.villain_spin_spin(phi),
             'Winding_Winding': lambda phi, n: ops.correlation('winding', n)}


def _is_villain(action):
    return type(action).__name__ == 'Villain' and all(hasattr(action, a) for a in ('Lattice', 'kappa', 'W'))


def villain_inline_values(rec, N, kappa):
    """Reference observables (observable/{action,energy,winding,wrapping}.py) from a per-chain device record."""
    sites = N * N
    S = rec[..., VOBS_ACTION]
    wrap = np.stack([rec[..., VOBS_WRAP0], rec[..., VOBS_WRAP1]], axis=-1)
    return {
        'ActionDensity': S / sites,
        'InternalEnergyDensity': S / (sites * kappa),
        'InternalEnergyDensitySquared': (S / (sites * kappa)) ** 2,
        'WindingSquared': rec[..., VOBS_SUM_DN2] / sites,
        'TorusWrapping': np.rint(wrap).astype(np.int64),
        'WrappingSquared': (wrap ** 2).sum(axis=-1),
    }


class NeighborhoodUpdate(Generator):
    r"""Checkerboard Metropolis on (phi_x, the four n touching x), every chain of a batch at once.

    Same constructor, attributes and `step`/`report` contract as the reference class.  Extras:

    rng
        `None` (default): proposals come from the in-kernel Philox4x32-10 stream keyed by `seed`.
        Assign a `numpy.random.Generator` -- as the reference's tests do -- and the numpy draws are
        replayed in the reference's order and injected, reproducing the reference chain bit for bit.
    inline
        names from `INLINE_NAMES` to return as inline observables (observable/observable.py:49-54).
    """

    def __init__(self, action, interval_phi=np.pi, interval_n=1, *, seed=None, inline=(), arithmetic='fast',
                 path='auto', dtype=torch.float64):
        if not _is_villain(action):
            raise ValueError('The Neighborhood Metropolis update requires the Villain action.')
        self.Action = action
        self.Lattice = action.Lattice
        self.kappa = action.kappa
        self.interval_phi = interval_phi
        self.interval_n = interval_n
        self.rng = None
        self.seed = fresh_seed() if seed is None else int(seed)
        self.counter = 0               # Philox sweep counter: resume = (seed, counter)
        self.arithmetic = arithmetic
        self.path = path
        self.dtype = dtype
        unknown = set(inline) - set(INLINE_NAMES)
        if unknown:
            raise ValueError(f'unknown inline observables {sorted(unknown)}')
        self.inline = tuple(inline)
        self.n_changes = np.arange(-interval_n, 1 + interval_n)
        self.accepted = 0
        self.proposed = 0
        self.acceptance = 0.
        self.sweeps = 0

    def __str__(self):
        return 'NeighborhoodUpdate'

    # -- device API -------------------------------------------------------------------------
    def sweep_device(self, phi, n, n_sweeps=1, *, obs=None, chain0=0, kappa_chain=None, injected=None,
                     accept_mask=None, dS_out=None):
        """`n_sweeps` sweeps in place on device tensors phi (chains,1,N,N), n (chains,2,N,N) int32."""
        ops.villain_sweep(phi, n, self.kappa, W=self.Action.W, interval_phi=self.interval_phi,
                          interval_n=self.interval_n, n_sweeps=n_sweeps, seed=self.seed, sweep0=self.counter,
                          chain0=chain0, injected=injected, arithmetic=self.arithmetic, path=self.path,
                          kappa_chain=kappa_chain, obs=obs, accept_mask=accept_mask, dS_out=dS_out)
        if injected is None:
            self.counter += n_sweeps

    def plan_device(self, phi, n, *, obs=None, chain0=0, kappa_chain=None):
        """Pre-validated form of `sweep_device` for tight loops: returns `run(n_sweeps=1)` advancing the Philox counter."""
        run = ops.villain_sweep_plan(phi, n, self.kappa, W=self.Action.W, interval_phi=self.interval_phi,
                                     interval_n=self.interval_n, seed=self.seed, chain0=chain0, arithmetic=self.arithmetic,
                                     path=self.path, kappa_chain=kappa_chain, obs=obs)

        def step(n_sweeps=1):
            run(self.counter, n_sweeps)
            self.counter += n_sweeps
        return step

    def swapping_device(self, phi, n, *, chain0=0, kappa_chain=None):
        """Stepping of one resident chain set on the tiled path without the copy back (ops.VillainSwappingSweeps): returns
        `step(n_sweeps=1, obs=None) -> (phi, n)`, the tensors that hold the state after the step (they alternate between the
        pair passed in and a second pair the stepper owns).  Raises NotImplementedError where the tiled path does not apply."""
        if self.rng is not None or self.path not in ('auto', 'tiled'):
            raise NotImplementedError('swapping sweeps serve Philox draws on path "auto" or "tiled"')
        sw = ops.VillainSwappingSweeps(phi, n, self.kappa, W=self.Action.W, interval_phi=self.interval_phi,
                                       interval_n=self.interval_n, seed=self.seed, chain0=chain0, arithmetic=self.arithmetic,
                                       kappa_chain=kappa_chain, force_tiled=(self.path == 'tiled'))

        def step(n_sweeps=1, obs=None):
            fields = sw.step(self.counter, n_sweeps, obs)
            self.counter += n_sweeps
            return fields
        return step

    def overlapped_device(self, phi, n, *, chain0=0, kappa_chain=None):
        """Overlapped-launch stepping of one resident chain set (ops.VillainOverlappedSweeps): returns `step(n_sweeps=1,
        obs=None, obs_in=None)` advancing the Philox counter, with `step.fence()` for foreign writes to the fields.
        obs_in: the state columns of the chains as they ARRIVE are written there (the previous step's record) and `obs`
        gets this launch's counters only.  Raises NotImplementedError where the overlapped kernel does not apply."""
        if self.rng is not None or self.arithmetic != 'fast' or self.path != 'auto':
            raise NotImplementedError('overlapped launches serve Philox draws, FAST arithmetic, path="auto"')
        kw = dict(W=self.Action.W, interval_phi=self.interval_phi, interval_n=self.interval_n, seed=self.seed, chain0=chain0,
                  kappa_chain=kappa_chain)
        N = int(phi.shape[-1])
        if N not in ops.VILLAIN_OVERLAP_SIZES and N > 128 and N % 16 == 0:
            # lattices beyond a CTA (config 5): in-place colour passes with the same step / obs_in protocol
            ov = ops.VillainInplaceSweeps(phi, n, self.kappa, **kw)
        else:
            ov = ops.VillainOverlappedSweeps(phi, n, self.kappa, **kw)

        def step(n_sweeps=1, obs=None, obs_in=None):
            ov.step(self.counter, n_sweeps, obs, obs_in)
            self.counter += n_sweeps
        step.fence = ov.fence
        return step

    def _injected_draws(self, chains, n_sweeps):
        L, W = self.Lattice, self.Action.W
        N = L.N
        u = np.empty((n_sweeps, chains, N, N)); dphi = np.empty_like(u)
        dn_fwd = np.empty((n_sweeps, chains, 2, N, N), dtype=np.int32); dn_bwd = np.empty_like(dn_fwd)
        for c in range(chains):              # chain-major: chain c consumes its n_sweeps sweeps consecutively
            for s in range(n_sweeps):
                u[s, c], dphi[s, c], dn_fwd[s, c], dn_bwd[s, c] = _replay.villain_neighborhood(
                    self.rng, L, W, self.interval_phi, self.interval_n)
        return {k: torch.from_numpy(v).cuda() for k, v in
                dict(u=u, dphi=dphi, dn_fwd=dn_fwd, dn_bwd=dn_bwd).items()}

    def _count(self, rec, chains, n_sweeps):
        sites = self.Lattice.sites
        self.sweeps += n_sweeps * chains
        self.proposed += sites * n_sweeps * chains
        self.accepted += int(round(float(rec[:, VOBS_ACCEPTED].sum())))
        self.acceptance += float(rec[:, VOBS_ACCEPTANCE].sum()) / sites

    # -- reference protocol -----------------------------------------------------------------
    def step(self, cfg, n_sweeps=1):
        r"""A volume's worth of single-site updates (`n_sweeps` of them) on cfg['phi'], cfg['n'].

        cfg holds one configuration -- phi (1,N,N), n (2,N,N) -- or a batch with a leading chain axis.
        Returns `cfg | {'phi': ..., 'n': ...}` as float64 / int64 `Form`s like the reference (:137).
        """
        N = self.Lattice.N
        phi, single = to_device(cfg['phi'], self.dtype, 1, N)
        n, _ = to_device(cfg['n'], torch.int32, 2, N)
        if not isinstance(cfg['phi'], torch.Tensor):
            pass
        else:
            phi, n = phi.clone(), n.clone()     # reference copy semantics (:84-85)
        chains = phi.shape[0]
        injected = self._injected_draws(chains, n_sweeps) if self.rng is not None else None
        obs = torch.empty((chains, VOBS_COUNT), dtype=torch.float64, device=phi.device)
        self.sweep_device(phi, n, n_sweeps, obs=obs, injected=injected)
        rec = obs.cpu().numpy()
        self._count(rec, chains, n_sweeps)
        out_phi = phi.to(torch.float64).cpu().numpy()
        out_n = n.cpu().numpy().astype(np.int64)
        L = self.Lattice
        if single:
            result = {'phi': Form(out_phi[0], degree=0, lattice=L), 'n': Form(out_n[0], degree=1, lattice=L)}
        else:
            result = {'phi': out_phi, 'n': out_n}
        if self.inline:
            vals = villain_inline_values(rec, N, self.kappa)
            for name in self.inline:
                if name in TWO_POINT:
                    # measured where the configuration is: only the (N, N) correlator crosses to the host
                    v = TWO_POINT[name](phi if phi.dtype == torch.float64 else phi.to(torch.float64), n).cpu().numpy()
                else:
                    v = vals[name]
                result[name] = v[0] if single else v
        return cfg | result

    def inline_observables(self, steps):
        N = self.Lattice.N
        shapes = {'TorusWrapping': (2,), 'Spin_Spin': (N, N), 'Winding_Winding': (N, N)}
        dtypes = {'TorusWrapping': int, 'Spin_Spin': complex, 'Winding_Winding': complex}
        return {name: Batch(steps, shape=shapes.get(name, ()), dtype=dtypes.get(name, float)) for name in self.inline}

    def report(self):
        return (
            f'There were {self.accepted} neighborhood proposals accepted of {self.proposed} proposed updates.'
            + '\n' +
            f'    {self.accepted/self.proposed:.6f} acceptance rate'
            + '\n' +
            f'    {self.acceptance / self.sweeps:.6f} average Metropolis acceptance probability.'
        )


class _DecoupledVillain(Generator):
    """Shared machinery of SiteUpdate / LinkUpdate / ExactUpdate on the GPU (svb_villain_decoupled)."""
    kind = None
    noun = None

    def _init(self, action, seed, path):
        if not _is_villain(action):
            raise ValueError(self._wrong_action)
        self.Action = action
        self.Lattice = action.Lattice
        self.kappa = action.kappa
        self.rng = None
        self.seed = fresh_seed() if seed is None else int(seed)
        self.counter = 0
        self.path = path
        self.accepted = 0
        self.proposed = 0
        self.acceptance = 0.
        self.sweeps = 0

    def _per_sweep(self):
        return self.Lattice.sites * (2 if self.kind == 'link' else 1)

    def _interval(self):
        return 1

    def sweep_device(self, phi, n, n_sweeps=1, *, obs=None, chain0=0, kappa_chain=None, injected=None, accept_mask=None,
                     dS_out=None):
        ops.villain_decoupled(self.kind, phi, n, self.kappa, W=self.Action.W, interval_phi=getattr(self, 'interval_phi', np.pi),
                              interval=self._interval(), n_sweeps=n_sweeps, seed=self.seed, sweep0=self.counter, chain0=chain0,
                              injected=injected, path=self.path, kappa_chain=kappa_chain, obs=obs, accept_mask=accept_mask,
                              dS_out=dS_out)
        if injected is None:
            self.counter += n_sweeps

    def _count(self, rec, chains, n_sweeps):
        self.sweeps += n_sweeps * chains
        self.proposed += self._per_sweep() * n_sweeps * chains
        self.accepted += int(round(float(rec[:, VOBS_ACCEPTED].sum())))
        self.acceptance += float(rec[:, VOBS_ACCEPTANCE].sum()) / self._per_sweep()

    def step(self, cfg, n_sweeps=1):
        N = self.Lattice.N
        phi, single = to_device(cfg['phi'], torch.float64, 1, N)
        n, _ = to_device(cfg['n'], torch.int32, 2, N)
        if isinstance(cfg['phi'], torch.Tensor):
            phi, n = phi.clone(), n.clone()
        chains = phi.shape[0]
        injected = self._injected_draws(chains, n_sweeps) if self.rng is not None else None
        obs = torch.empty((chains, VOBS_COUNT), dtype=torch.float64, device=phi.device)
        self.sweep_device(phi, n, n_sweeps, obs=obs, injected=injected)
        self._count(obs.cpu().numpy(), chains, n_sweeps)
        L = self.Lattice
        result = {}
        if self.kind == 'site':
            out = phi.cpu().numpy()
            result['phi'] = Form(out[0], degree=0, lattice=L) if single else out
        else:
            out = n.cpu().numpy().astype(np.int64)
            result['n'] = Form(out[0], degree=1, lattice=L) if single else out
        return cfg | result

    def inline_observables(self, steps):
        return {}

    def report(self):
        return (
            f'There were {self.accepted} {self.noun} proposals accepted of {self.proposed} proposed updates.'
            + '\n' +
            f'    {self.accepted/self.proposed:.6f} acceptance rate'
            + '\n' +
            f'    {self.acceptance / self.sweeps:.6f} average Metropolis acceptance probability.'
        )


class SiteUpdate(_DecoupledVillain):
    """Checkerboard Metropolis on phi alone; drop-in for supervillain.generator.villain.SiteUpdate (site.py:12-132)."""
    kind, noun, _wrong_action = 'site', 'single-phi', 'Need a Villain action'

    def __init__(self, action, interval_phi=np.pi, *, seed=None, path='auto'):
        self._init(action, seed, path)
        self.interval_phi = interval_phi

    def __str__(self):
        return 'SiteUpdate'

    def _injected_draws(self, chains, n_sweeps):
        N = self.Lattice.N
        u = np.empty((n_sweeps, chains, N, N)); dphi = np.empty_like(u)
        for c in range(chains):
            for s in range(n_sweeps):
                u[s, c], dphi[s, c] = _replay.villain_site(self.rng, self.Lattice, self.interval_phi)
        return {'u': torch.from_numpy(u).cuda(), 'dphi': torch.from_numpy(dphi).cuda()}


class LinkUpdate(_DecoupledVillain):
    """Every n independently against the frozen phi; drop-in for supervillain.generator.villain.LinkUpdate (link.py:12-114)."""
    kind, noun, _wrong_action = 'link', 'single-link', 'The LinkUpdate requires the Villain action.'

    def __init__(self, action, interval_n=1, *, seed=None, path='auto'):
        self._init(action, seed, path)
        self.interval_n = interval_n
        self.n_changes = tuple(n for n in range(-interval_n, 0)) + tuple(n for n in range(1, interval_n + 1))

    def __str__(self):
        return 'LinkUpdate'

    def _interval(self):
        return self.interval_n

    def _injected_draws(self, chains, n_sweeps):
        N = self.Lattice.N
        u = np.empty((n_sweeps, chains, 2, N, N)); a = np.empty((n_sweeps, chains, 2, N, N), dtype=np.int32)
        for c in range(chains):
            for s in range(n_sweeps):
                u[s, c], a[s, c] = _replay.villain_link(self.rng, self.Lattice, self.Action.W, self.interval_n)
        return {'u': torch.from_numpy(u).cuda(), 'a': torch.from_numpy(a).cuda()}


class ExactUpdate(_DecoupledVillain):
    """n += d z, z on one colour at a time; drop-in for supervillain.generator.villain.ExactUpdate (exact.py:12-141)."""
    kind, noun, _wrong_action = 'exact', 'exact', 'Need a Villain action'

    def __init__(self, action, interval_z=1, *, seed=None, path='auto'):
        self._init(action, seed, path)
        self.interval_z = interval_z
        self.zs = tuple(z for z in range(-interval_z, 0)) + tuple(z for z in range(1, interval_z + 1))

    def __str__(self):
        return 'ExactUpdate'

    def _interval(self):
        return self.interval_z

    def _injected_draws(self, chains, n_sweeps):
        N = self.Lattice.N
        u = np.empty((n_sweeps, chains, N, N)); a = np.empty((n_sweeps, chains, N, N), dtype=np.int32)
        for c in range(chains):
            for s in range(n_sweeps):
                u[s, c], a[s, c] = _replay.villain_exact(self.rng, self.Lattice, self.interval_z)
        return {'u': torch.from_numpy(u).cuda(), 'a': torch.from_numpy(a).cuda()}


class CohomologyUpdate(Generator):
    """Winding-sector moves of n (supervillain/generator/villain/cohomology.py:12-130): `CohomologyUpdate(action,
    interval_h=1)`; one slice proposal per direction and chain."""

    def __init__(self, action, interval_h=1, *, seed=None):
        if not _is_villain(action):
            raise ValueError('Need a Villain action')
        self.Action = action
        self.Lattice = action.Lattice
        self.kappa = action.kappa
        self.interval_h = interval_h
        self.h = tuple(h for h in range(-interval_h, 0)) + tuple(h for h in range(1, interval_h + 1))
        self.rng = None
        self.seed = fresh_seed() if seed is None else int(seed)
        self.counter = 0
        self.accepted = 0
        self.proposed = 0
        self.acceptance = 0.
        self.sweeps = 0

    def __str__(self):
        return 'CohomologyUpdate'

    def sweep_device(self, phi, n, n_sweeps=1, *, obs=None, chain0=0, kappa_chain=None, injected=None, counters=None,
                     dS_out=None):
        """`n_sweeps` cohomology steps in place on n (chains,2,N,N) int32; `obs`, if given, is refreshed from the final
        state with svb_villain_observables (its two counter columns are zeroed: use `counters`)."""
        for s in range(n_sweeps):
            ops.villain_cohomology(phi, n, self.kappa, interval=self.interval_h, seed=self.seed, sweep=self.counter + s,
                                   chain0=chain0, injected=injected, kappa_chain=kappa_chain, counters=counters, dS_out=dS_out)
        if injected is None:
            self.counter += n_sweeps
        if obs is not None:
            ops.villain_observables(phi, n, self.kappa, kappa_chain=kappa_chain, obs=obs)

    def step(self, cfg):
        N = self.Lattice.N
        phi, single = to_device(cfg['phi'], torch.float64, 1, N)
        n, _ = to_device(cfg['n'], torch.int32, 2, N)
        if isinstance(cfg['n'], torch.Tensor):
            n = n.clone()
        chains = phi.shape[0]
        injected = None
        if self.rng is not None:
            u = np.empty((chains, 2)); h = np.empty((chains, 2), dtype=np.int32)
            for k in range(chains):
                u[k], h[k] = _replay.villain_cohomology(self.rng, self.interval_h)
            injected = {'u': torch.from_numpy(u).cuda(), 'h': torch.from_numpy(h).cuda()}
        counters = torch.zeros((chains, 2), dtype=torch.float64, device=phi.device)
        self.sweep_device(phi, n, 1, injected=injected, counters=counters)
        rec = counters.cpu().numpy()
        self.proposed += 2 * chains
        self.accepted += int(round(float(rec[:, 0].sum())))
        self.acceptance += float(rec[:, 1].sum()) / 2
        self.sweeps += chains
        out_n = n.cpu().numpy().astype(np.int64)
        return cfg | {'n': Form(out_n[0], degree=1, lattice=self.Lattice) if single else out_n}

    def inline_observables(self, steps):
        return {}

    def report(self):
        return (
            f'There were {self.accepted} cohomology proposals accepted of {self.proposed} proposed updates.'
            + '\n' +
            f'    {self.accepted/self.proposed:.6f} acceptance rate'
            + '\n' +
            f'    {self.acceptance / self.sweeps:.6f} average Metropolis acceptance probability.'
        )
