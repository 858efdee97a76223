#!/usr/bin/env python
"""Build container only: the oracle's numpy port of NeighborhoodUpdate.step (what bench.py times as `cpu_baseline`,
kind "port") beside the UNMODIFIED reference generator on the same core and shape (L=32, kappa=0.5, one chain).
The reference cannot travel to the GPU box, so this is where the two are compared."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refimport, villain_np as V                                                   # noqa: E402

L, KAPPA, SECONDS = 32, 0.5, float(sys.argv[1]) if len(sys.argv) > 1 else 5.0

sv = refimport.import_reference()
S = sv.action.Villain(sv.lattice.Lattice2D(L), KAPPA)
G = sv.generator.villain.NeighborhoodUpdate(S)
cfg = S.configurations(1)[0]
for _ in range(20):                       # numba JIT and caches
    cfg = G.step(cfg)
t0, sweeps = time.perf_counter(), 0
while time.perf_counter() - t0 < SECONDS:
    for _ in range(50):
        cfg = G.step(cfg)
    sweeps += 50
ref_rate = sweeps * L * L / (time.perf_counter() - t0)

rng = np.random.default_rng(1)
phi, n = np.zeros((1, L, L)), np.zeros((2, L, L), dtype=np.int64)
for _ in range(20):
    phi, n = V.neighborhood_step(phi, n, KAPPA, 1, rng)
t0, sweeps = time.perf_counter(), 0
while time.perf_counter() - t0 < SECONDS:
    for _ in range(50):
        phi, n = V.neighborhood_step(phi, n, KAPPA, 1, rng)
    sweeps += 50
port_rate = sweeps * L * L / (time.perf_counter() - t0)
print(f'L={L} kappa={KAPPA}, one core: unmodified reference {ref_rate:.3e} site-updates/s, oracle port {port_rate:.3e} site-updates/s, '
      f'port / reference = {port_rate / ref_rate:.2f}')
