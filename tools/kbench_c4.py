#!/usr/bin/env python
"""Decompose the cost of a config-4 (L = 128) launch of the cluster kernel: records on/off, sweeps per call, chains."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200 import ops

def timeit(fn, n=10, reps=3):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best * 1e3

N = 128
for CH in (71, 142, 284, 1024):
    S = svb.Villain(svb.Lattice2D(N), 0.5)
    phi, n = svb.BatchedEnsemble(S, CH)._start('hot', 1)
    obs = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda')
    obs_b = torch.zeros_like(obs)
    ov = ops.VillainOverlappedSweeps(phi, n, 0.5, seed=1)
    for label, kw in (('no records', dict()), ('records', dict(obs=obs))):
        for sw in (1, 2, 5):
            k = [0]
            def f():
                k[0] += sw
                ops.villain_sweep(phi, n, 0.5, seed=1, sweep0=k[0], n_sweeps=sw, **kw)
            print(f'chains {CH:5d}  {label:12s} {sw} sweep(s): {timeit(f):8.1f} us', flush=True)
    k = [0]
    def g():
        k[0] += 1
        ov.step(k[0], 1, obs if k[0] & 1 else obs_b, obs_b if k[0] & 1 else obs)
    print(f'chains {CH:5d}  overlapped + obs_in, 1 sweep: {timeit(g):8.1f} us', flush=True)
