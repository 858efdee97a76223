"""`Sequentially` and `KeepEvery` (supervillain/generator/combining.py:9-116)."""
import inspect

from .generator import Generator


class Sequentially(Generator):
    """Apply each generator's `step` one after the next (combining.py:9-52)."""

    def __init__(self, generators):
        self.generators = generators

    def __str__(self):
        return 'Sequentially((' + ', '.join(str(g) for g in self.generators) + '))'

    def step(self, cfg):
        result = cfg
        for g in self.generators:
            result = g.step(result)
        return result

    def inline_observables(self, steps):
        combined = dict()
        for g in self.generators:
            combined |= g.inline_observables(steps)
        return combined

    def sweep_device(self, a, b, n_sweeps=1, *, obs=None, chain0=0, kappa_chain=None):
        """Device-resident counterpart of `step` for GPU generators (used by `BatchedEnsemble`): `n_sweeps` rounds of
        every generator in order, in place; the observable record is written by the last generator of the last round."""
        gens = list(self.generators)
        for s in range(n_sweeps):
            for k, g in enumerate(gens):
                final = (s == n_sweeps - 1) and (k == len(gens) - 1)
                g.sweep_device(a, b, 1, obs=obs if final else None, chain0=chain0, kappa_chain=kappa_chain)

    def report(self):
        return '\n\n'.join(g.report() for g in self.generators)


class KeepEvery(Generator):
    """Apply `generator` n times per emitted configuration (combining.py:57-116).

    When the wrapped generator is one of the GPU generators and no blocked inline averaging is
    requested, the n sweeps are fused into ONE kernel launch with the lattice resident in shared
    memory (`step(cfg, n_sweeps=n)`), instead of n host round trips.
    """

    def __init__(self, n, generator, blocked_inline=True):
        self.stride = n
        self.generator = generator
        self.blocked_inline = blocked_inline

    def __str__(self):
        return f'KeepEvery({self.stride}, {str(self.generator)})'

    def _fusable(self):
        try:
            return 'n_sweeps' in inspect.signature(self.generator.step).parameters
        except (TypeError, ValueError):
            return False

    def step(self, cfg):
        blocked = self.inline_observables(1) if self.blocked_inline else dict()
        if not blocked and self._fusable():
            return self.generator.step(cfg, n_sweeps=self.stride)
        for o in blocked:
            blocked[o] = blocked[o][0]
        result = cfg
        for _ in range(self.stride):
            result = self.generator.step(result)
            for o in blocked:
                blocked[o] = blocked[o] + result[o] / self.stride
        return result | blocked

    def inline_observables(self, steps):
        return self.generator.inline_observables(steps)

    def report(self):
        return self.generator.report()
