"""Plain-data (JSON) state of the actions and generators of this package, for checkpoints.

A checkpoint must not execute code when it is read (no pickle), and must not depend on Python class paths.  The state of
every generator here is a handful of scalars -- constructor arguments, the Philox (seed, sweep counter) and the report
counters -- so `describe` turns an (action, generator) pair into nested dicts of numbers and strings and `rebuild`
reconstructs it through the constructors of a fixed whitelist of classes.  (The reference stores the generator's
`__dict__` in HDF5 with the rng pickled, h5/readwriteable.py:37-45; the plain fields are the same.)
"""
import numpy as np
import torch

from ..action import Villain, Worldline
from ..lattice import Lattice2D
from . import combining, villain, worldline

_ACTIONS = {'Villain': Villain, 'Worldline': Worldline}

# class -> (constructor keyword -> attribute that holds it)
_LEAVES = {
    'villain.NeighborhoodUpdate': (villain.NeighborhoodUpdate, dict(interval_phi='interval_phi', interval_n='interval_n', inline='inline',
                                                                    arithmetic='arithmetic', path='path', dtype='dtype')),
    'villain.SiteUpdate': (villain.SiteUpdate, dict(interval_phi='interval_phi', path='path')),
    'villain.LinkUpdate': (villain.LinkUpdate, dict(interval_n='interval_n', path='path')),
    'villain.ExactUpdate': (villain.ExactUpdate, dict(interval_z='interval_z', path='path')),
    'villain.CohomologyUpdate': (villain.CohomologyUpdate, dict(interval_h='interval_h')),
    'worldline.PlaquetteUpdate': (worldline.PlaquetteUpdate, dict(inline='inline', path='path')),
    'worldline.VortexUpdate': (worldline.VortexUpdate, dict(interval_v='interval_v', inline='inline', path='path')),
    'worldline.CoexactUpdate': (worldline.CoexactUpdate, dict(interval_t='interval_t', inline='inline', path='path')),
    'worldline.WrappingUpdate': (worldline.WrappingUpdate, dict(interval_w='interval_w')),
}
_COUNTERS = ('counter', 'accepted', 'proposed', 'acceptance', 'sweeps')
_DTYPES = {'float64': torch.float64, 'float32': torch.float32}


def _plain(x):
    if isinstance(x, torch.dtype):
        return str(x).replace('torch.', '')
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (tuple, list)):
        return [_plain(v) for v in x]
    if isinstance(x, (bool, int, float, str)) or x is None:
        return x
    raise TypeError(f'{type(x).__name__} is not plain checkpoint data')


def describe_action(action):
    name = type(action).__name__
    if name not in _ACTIONS:
        raise TypeError(f'cannot checkpoint an action of type {name}')
    return {'class': name, 'N': int(action.Lattice.N), 'kappa': float(action.kappa), 'W': _plain(action.W)}


def rebuild_action(d):
    return _ACTIONS[d['class']](Lattice2D(int(d['N'])), float(d['kappa']), W=d['W'])


def describe_generator(g):
    """Nested plain data for a generator of this package (or None)."""
    if g is None:
        return None
    if isinstance(g, combining.Sequentially):
        return {'class': 'combining.Sequentially', 'generators': [describe_generator(x) for x in g.generators]}
    if isinstance(g, combining.KeepEvery):
        return {'class': 'combining.KeepEvery', 'stride': int(g.stride), 'blocked_inline': bool(g.blocked_inline),
                'generator': describe_generator(g.generator)}
    for name, (cls, kwargs) in _LEAVES.items():
        if type(g) is cls:
            d = {'class': name, 'seed': int(g.seed), 'kwargs': {k: _plain(getattr(g, attr)) for k, attr in kwargs.items()}}
            d.update({k: _plain(getattr(g, k)) for k in _COUNTERS})
            rng = getattr(g, 'rng', None)
            if rng is not None:                     # a numpy Generator's state is plain data too (a dict of ints and strings)
                d['rng'] = rng.bit_generator.state
            return d
    raise TypeError(f'cannot checkpoint a generator of type {type(g).__name__}: not one of this package\'s generators')


def rebuild_generator(d, action):
    if d is None:
        return None
    name = d['class']
    if name == 'combining.Sequentially':
        return combining.Sequentially(tuple(rebuild_generator(x, action) for x in d['generators']))
    if name == 'combining.KeepEvery':
        return combining.KeepEvery(int(d['stride']), rebuild_generator(d['generator'], action), blocked_inline=bool(d['blocked_inline']))
    if name not in _LEAVES:
        raise ValueError(f'checkpoint names an unknown generator class {name!r}')
    cls, _ = _LEAVES[name]
    kwargs = dict(d['kwargs'])
    if 'inline' in kwargs:
        kwargs['inline'] = tuple(kwargs['inline'])
    if 'dtype' in kwargs:
        kwargs['dtype'] = _DTYPES[kwargs['dtype']]
    g = cls(action, seed=int(d['seed']), **kwargs)
    for k in _COUNTERS:
        setattr(g, k, d[k])
    if d.get('rng') is not None:
        bit = getattr(np.random, d['rng']['bit_generator'])()
        bit.state = d['rng']
        g.rng = np.random.Generator(bit)
    return g
