#!/usr/bin/env python
"""Quick kernel timing of the config-2 Villain sweep for a library variant (SVB200_LIB env var)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200 import ops

L, CH, R = int(os.environ.get('KB_L', 32)), int(os.environ.get('KB_CHAINS', 4096)), 4
S = svb.Villain(svb.Lattice2D(L), 0.5)
sets = [svb.BatchedEnsemble(S, CH)._start('hot', 100 + r) for r in range(R)]
obs = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda')
use_obs = os.environ.get('KB_OBS', '1') == '1'
sweeps = int(os.environ.get('KB_SWEEPS', 1))
overlap = os.environ.get('KB_OVERLAP', '0') == '1'
obs_sets = [torch.zeros_like(obs) for _ in range(R)]
obs_prev = [torch.zeros_like(obs) for _ in range(R)]
obs_in = os.environ.get('KB_OBSIN', '0') == '1'
steppers = [ops.VillainOverlappedSweeps(phi, n, 0.5, seed=1) for phi, n in sets] if overlap else None
def step(k):
    phi, n = sets[k % R]
    if overlap:
        steppers[k % R].step(k * sweeps, sweeps, obs=obs_sets[k % R] if use_obs else None,
                             obs_in=obs_prev[k % R] if (use_obs and obs_in) else None)
    else:
        ops.villain_sweep(phi, n, 0.5, n_sweeps=sweeps, seed=1, sweep0=k * sweeps, obs=obs if use_obs else None)
therm = int(os.environ.get('KB_THERM', 200))     # untimed sweeps: the acceptance rate of a fresh hot start is far from equilibrium
for phi, n in sets:
    if therm: ops.villain_sweep(phi, n, 0.5, n_sweeps=therm, seed=7, sweep0=10**6)
for k in range(10): step(k)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(100): step(k)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 100)
upd = CH * L * L * sweeps / (best * 1e-3)
print(f"{os.environ.get('SVB200_LIB','default'):40s} L={L} obs={use_obs} sweeps={sweeps} overlap={overlap} obs_in={obs_in}: {best*1e3:8.2f} us/step  {upd:.3e} upd/s  {upd*32/6538.6e9*100:5.1f}% of HBM roofline")
