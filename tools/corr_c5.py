#!/usr/bin/env python
"""One L = 4096 spin-spin correlator after another (svb_villain_spin_spin), for launch lists and profiles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supervillain_b200 import ops
N = int(os.environ.get('KB_L', 4096))
chains = int(os.environ.get('KB_CHAINS', 1))
phi = torch.rand((chains, 1, N, N), dtype=torch.float64, device='cuda')
out = torch.empty((chains, N, N, 2), dtype=torch.float64, device='cuda')
steps = int(os.environ.get('KB_STEPS', 3))
ops.villain_spin_spin(phi, out=out)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0.record()
for _ in range(steps):
    ops.villain_spin_spin(phi, out=out)
t1.record()
torch.cuda.synchronize()
print(f'L={N} x {chains} spin_spin {1e3 * t0.elapsed_time(t1) / steps:.1f} us per call ({steps} calls, lib {os.environ.get("SVB200_LIB", "default")})')
