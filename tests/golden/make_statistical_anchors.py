#!/usr/bin/env python
"""Expectation values from the UNMODIFIED reference generators, for the statistical-equivalence tests (level 3).

Run in the build container:  python tests/golden/make_statistical_anchors.py
Writes tests/golden/statistical_anchors.json: for each case the mean and a binned standard error (50 bins after
discarding the thermalisation) of ActionDensity and WindingSquared.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refimport  # noqa: E402

sv = refimport.import_reference()


def binned(x, bins=50):
    x = np.asarray(x, dtype=float)
    n = len(x) // bins * bins
    b = x[:n].reshape(bins, -1).mean(axis=1)
    return float(b.mean()), float(b.std(ddof=1) / np.sqrt(bins))


def run(action_name, gen_name, N, kappa, W, sweeps, therm, seed):
    np.random.seed(seed)
    L = sv.lattice.Lattice2D(N)
    if action_name == 'Villain':
        S = sv.action.Villain(L, kappa, W=W)
        G = {'Hammer': lambda: sv.generator.villain.Hammer(S),
             'NeighborhoodUpdate': lambda: sv.generator.villain.NeighborhoodUpdate(S)}[gen_name]()
    else:
        S = sv.action.Worldline(L, kappa, W=W)
        G = {'Hammer': lambda: sv.generator.worldline.Hammer(S),
             'PlaquetteUpdate': lambda: sv.generator.worldline.PlaquetteUpdate(S),
             'Vortex+Coexact': lambda: sv.generator.combining.Sequentially(
                 (sv.generator.worldline.VortexUpdate(S), sv.generator.worldline.CoexactUpdate(S)))}[gen_name]()
    for k, g in enumerate(getattr(G, 'generators', [G])):
        if hasattr(g, 'rng'):
            g.rng = np.random.default_rng([seed, k])         # distinct, reproducible stream per sub-generator
    t0 = time.time()
    E = sv.Ensemble(S).generate(sweeps, G, start='cold')
    cut = E.cut(therm)
    out = {'action': action_name, 'generator': gen_name, 'N': N, 'kappa': kappa, 'W': W, 'sweeps': sweeps, 'therm': therm}
    for name in ('ActionDensity', 'WindingSquared'):
        m, e = binned(np.asarray(getattr(cut, name).array if hasattr(getattr(cut, name), 'array') else getattr(cut, name)))
        out[name] = [m, e]
    out['seconds'] = round(time.time() - t0, 1)
    print(out, flush=True)
    return out


if __name__ == '__main__':
    cases = [
        # the full ensembles (all wrapping sectors): what an ergodic sampler must reproduce
        ('Worldline', 'Hammer', 8, 0.3, 1, 40000, 2000, 1),
        ('Worldline', 'Hammer', 8, 0.5, 1, 40000, 2000, 2),
        ('Worldline', 'Hammer', 5, 0.5, 1, 40000, 2000, 3),
        ('Villain', 'Hammer', 8, 0.3, 1, 40000, 2000, 4),
        # the sector a plaquette-only chain from a cold start stays in (plaquette.py:16-19)
        ('Worldline', 'Vortex+Coexact', 8, 0.5, 1, 60000, 2000, 5),
        ('Worldline', 'Vortex+Coexact', 8, 0.3, 1, 60000, 2000, 6),
        ('Worldline', 'PlaquetteUpdate', 8, 0.5, 1, 12000, 1000, 7),
    ]
    results = [run(*c) for c in cases]
    with open(os.path.join(HERE, 'statistical_anchors.json'), 'w') as f:
        json.dump(results, f, indent=1)
