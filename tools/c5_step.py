#!/usr/bin/env python
"""One config-5 step after another (svb_villain_sweep_inplace with the obs_in record protocol), for launch lists and profiles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200 import ops
N = int(os.environ.get('KB_L', 4096)); chains = int(os.environ.get('KB_CHAINS', 1))
S = svb.Villain(svb.Lattice2D(N), 0.5)
phi, n = svb.BatchedEnsemble(S, chains)._start('hot', 1)
st = ops.VillainInplaceSweeps(phi, n, 0.5, seed=1)
a = torch.zeros((chains, 6), dtype=torch.float64, device='cuda'); b = torch.zeros_like(a)
st.step(0, int(os.environ.get('KB_THERM', 300)))
for k in range(int(os.environ.get('KB_STEPS', 12))):
    st.step(1000 + k, 1, obs=a, obs_in=b); a, b = b, a
torch.cuda.synchronize()
