#!/usr/bin/env python
"""One L = 4096 spin-spin correlator after another (svb_villain_spin_spin), for launch lists and profiles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supervillain_b200 import ops
N = int(os.environ.get('KB_L', 4096))
phi = torch.rand((1, 1, N, N), dtype=torch.float64, device='cuda')
out = torch.empty((1, N, N, 2), dtype=torch.float64, device='cuda')
for _ in range(int(os.environ.get('KB_STEPS', 3))):
    ops.villain_spin_spin(phi, out=out)
torch.cuda.synchronize()
