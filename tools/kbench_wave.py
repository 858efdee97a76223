#!/usr/bin/env python
"""Wavefront steps (svb_villain_sweep_wavefront) against colour-pass steps (svb_villain_sweep_inplace) and, at L = 128, the
overlapped launches of svb_villain_sweep_overlapped (kind 'cluster': the strips kernel, or the cluster kernel under
SVB_VILLAIN_KERNEL128=cluster): CUDA-event time per single-sweep step with the obs_in record protocol.
    KB_SHAPES='4096x1,128x8192'  KB_LAGS='0,8,16,32'  (0 = the launcher's choice)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200 import ops

PEAK = 6538.6


def timeit(fn, n, reps=3):
    for _ in range(5):
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


shapes = [tuple(int(v) for v in s.split('x')) for s in os.environ.get('KB_SHAPES', '4096x1,128x8192,128x1024,512x64').split(',')]
lags = [int(v) for v in os.environ.get('KB_LAGS', '0').split(',')]
kinds_only = os.environ.get('KB_KINDS', '')
for N, CH in shapes:
    S = svb.Villain(svb.Lattice2D(N), 0.5)
    phi, n = svb.BatchedEnsemble(S, CH)._start('hot', 1)
    a = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda'); b = torch.zeros_like(a)
    sites = CH * N * N
    reps = max(5, min(200, int(2e9 / sites)))
    therm = ops.VillainInplaceSweeps(phi, n, 0.5, seed=1, launches='passes')
    therm.step(0, 100)
    rows = []
    for kind in ['passes'] + [f'wavefront lag={l}' for l in lags] + (['cluster'] if N == 128 else []):
        if kinds_only and not kind.startswith(kinds_only):
            continue
        if kind.startswith('wavefront'):
            lag = int(kind.split('=')[1])
            if lag:
                os.environ['SVB_WAVE_LAG'] = str(lag)
            else:
                os.environ.pop('SVB_WAVE_LAG', None)
            st = ops.VillainInplaceSweeps(phi, n, 0.5, seed=1, launches='wavefront')
        elif kind == 'passes':
            st = ops.VillainInplaceSweeps(phi, n, 0.5, seed=1, launches='passes')
        else:
            st = ops.VillainOverlappedSweeps(phi, n, 0.5, seed=1)
        k = [1000]
        for with_rec in (True, False):
            def f():
                k[0] += 1
                if with_rec:
                    st.step(k[0], 1, obs=a if k[0] & 1 else b, obs_in=b if k[0] & 1 else a)
                else:
                    st.step(k[0], 1)
            us = timeit(f, reps)
            print(f'L={N:5d} chains={CH:5d} {kind:20s} {"records" if with_rec else "no records":10s}: {us:9.1f} us/step  '
                  f'{sites / us * 1e-3:7.1f} G site-updates/s  {32 * sites / us * 1e-3 / PEAK:5.3f} of the roofline', flush=True)
