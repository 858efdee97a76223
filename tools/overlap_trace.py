#!/usr/bin/env python
"""Evidence for the overlapped-launch timing (no nsys on this image): a build of the library with -DSVB_TRACE makes every CTA of
an overlapped config-2 launch record %globaltimer at its start and after it has published its chains.  This script runs the
bench's stepping (4 chain sets rotated, obs_in records) for 100 steady-state steps and prints, per launch: when its first CTA
started, when its last CTA ended, the start-to-start spacing to the next launch and by how much the next launch's first CTA
started BEFORE this launch's last CTA ended (the overlap).

    tools/build_variants.sh "trace:-DSVB_TRACE=1" && SVB200_LIB=$PWD/variants/libsvb200_trace.so python tools/overlap_trace.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb                                                   # noqa: E402
from supervillain_b200 import _lib, ops                                           # noqa: E402

L, CH, R, STEPS = 32, 4096, 4, 100
lib = _lib.load()
fn = getattr(ctypes.CDLL(_lib.LIB_PATH), 'svb_debug_trace_read')
S = svb.Villain(svb.Lattice2D(L), 0.5)
sets = [svb.BatchedEnsemble(S, CH)._start('hot', 100 + r) for r in range(R)]
for phi, n in sets:
    ops.villain_sweep(phi, n, 0.5, n_sweeps=200, seed=7, sweep0=10**6)
steppers = [ops.VillainOverlappedSweeps(phi, n, 0.5, seed=1) for phi, n in sets]
obs = [[torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda') for _ in range(2)] for _ in range(R)]
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(STEPS + 20):
    if k == 20:
        start.record()
    r = k % R
    steppers[r].step(k, 1, obs=obs[r][k // R % 2], obs_in=obs[r][(k // R + 1) % 2])
stop.record()
torch.cuda.synchronize()
ms = start.elapsed_time(stop) / STEPS
launches, ctas = ctypes.c_int(0), ctypes.c_int(0)
fn(None, ctypes.byref(launches), ctypes.byref(ctas))
buf = np.zeros((launches.value, ctas.value, 2), dtype=np.uint64)
fn(buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(launches), ctypes.byref(ctas))
# launch k of set r carries signal epoch k // R + 1; the ring index is the epoch -> launches of the 4 sets share a slot,
# so trace ONE set: its launches are every R-th step, and between two of them run the three other sets' launches.  To see
# consecutive launches, re-run with a single set below.
print(f'CUDA events: {ms * 1e3:.2f} us per step over {STEPS} steps ({R} chain sets rotated)')

# consecutive launches on ONE chain set (every launch waits for the chains its predecessor publishes: the hardest case)
phi, n = sets[0]
st = ops.VillainOverlappedSweeps(phi, n, 0.5, seed=3)
o = [torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda') for _ in range(2)]
torch.cuda.synchronize()
start.record()
for k in range(STEPS):
    st.step(k, 1, obs=o[k % 2], obs_in=o[(k + 1) % 2])
stop.record()
torch.cuda.synchronize()
ms1 = start.elapsed_time(stop) / STEPS
fn(buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(launches), ctypes.byref(ctas))
grid = int((buf[1, :, 0] > 0).sum())
t = buf[:, :grid, :].astype(np.int64)
first_start = t[:, :, 0].min(axis=1)
last_start = t[:, :, 0].max(axis=1)
last_end = t[:, :, 1].max(axis=1)
epochs = [e % launches.value for e in range(STEPS - 60 + 1, STEPS + 1)]          # the last 60 launches (epoch = step + 1)
print(f'one chain set, {STEPS} consecutive overlapped launches of {grid} CTAs: CUDA events {ms1 * 1e3:.2f} us per step')
print('launch  first CTA start  last CTA start  last CTA end   start-to-start   overlap with next   (us, relative to the first row)')
t0 = first_start[epochs[0]]
spacing, overlap = [], []
for a, b in zip(epochs[:-1], epochs[1:]):
    sp = (first_start[b] - first_start[a]) * 1e-3
    ov = (last_end[a] - first_start[b]) * 1e-3
    spacing.append(sp); overlap.append(ov)
    print(f'{a:5d}  {(first_start[a] - t0) * 1e-3:14.2f}  {(last_start[a] - t0) * 1e-3:14.2f}  {(last_end[a] - t0) * 1e-3:12.2f}  {sp:14.2f}  {ov:16.2f}')
print(f'mean start-to-start spacing {np.mean(spacing):.2f} us (CUDA events: {ms1 * 1e3:.2f} us per step); '
      f'mean overlap of consecutive launches {np.mean(overlap):.2f} us; every launch overlaps its successor: {bool(np.min(overlap) > 0)}')
