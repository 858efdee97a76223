"""The duck-typed generator protocol `Ensemble.generate` drives (supervillain/generator/generator.py:3-33;
ensemble.py:75,89-92,97): `step(cfg) -> cfg`, `inline_observables(steps) -> {name: Batch}`, `report() -> str`."""
import os


class Generator:
    def step(self, configuration):
        raise NotImplementedError

    def inline_observables(self, steps):
        return dict()

    def report(self):
        return ''


def fresh_seed():
    """64 random bits for the Philox key when the caller does not fix a seed."""
    return int.from_bytes(os.urandom(8), 'little')
