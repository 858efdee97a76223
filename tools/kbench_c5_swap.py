#!/usr/bin/env python
"""C5 (one L=4096 lattice): single-sweep steps in place (copy back after the odd sweep) against swapping buffer pairs."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200 import ops
from supervillain_b200._lib import VOBS_COUNT

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = svb.Villain(svb.Lattice2D(N), 0.5)
phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 1)
obs = torch.zeros((chains, VOBS_COUNT), dtype=torch.float64, device='cuda')
ops.villain_sweep(phi, n, 0.5, n_sweeps=int(os.environ.get('KB_THERM', 300)), seed=9, sweep0=10**6, path='tiled')   # thermalise


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best * 1e3


use_obs = os.environ.get('KB_OBS', '1') == '1'
if not use_obs: obs = None
k = [0]
def in_place():
    ops.villain_sweep(phi, n, 0.5, n_sweeps=1, seed=3, sweep0=k[0], path='tiled', obs=obs); k[0] += 1
sw = ops.VillainSwappingSweeps(phi, n, 0.5, seed=3, force_tiled=True)
def swapping():
    sw.step(k[0], 1, obs); k[0] += 1
for name, fn in (('in place (copy back)', in_place), ('swapping buffer pairs', swapping)):
    t = timeit(fn)
    upd = chains * N * N
    print(f'villain L={N} x {chains}, tiled, 1 sweep/step, {name:24s} {t:8.1f} us  {upd / t * 1e6:.3e} upd/s  {upd * 32 / t * 1e6 / 6538.6e9 * 100:5.1f}% of HBM roofline')
