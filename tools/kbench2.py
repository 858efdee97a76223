#!/usr/bin/env python
"""Timing of the other BASELINE shapes: worldline C3 shard, Villain C4 shard (L=128), C5 (L=4096), form ops."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200 import ops

PEAK = 6538.6e9


def timeit(fn, n=20, reps=3):
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
    return best * 1e-3


def report(name, updates, bytes_per, t):
    print(f'{name:58s} {t*1e6:9.1f} us  {updates/t:.3e} upd/s  {updates*bytes_per/t/PEAK*100:5.1f}% of HBM roofline ({bytes_per} B/upd)')


which = sys.argv[1:] or ['c3', 'c4', 'c5', 'forms']
if 'c3' in which:
    N, CH = 64, 1024
    S = svb.Worldline(svb.Lattice2D(N), 0.5)
    sets = [svb.BatchedEnsemble(S, CH)._start('hot', r) for r in range(6)]      # 6 x 48 MiB > L2
    obs = torch.zeros((CH, ops.WOBS_COUNT), dtype=torch.float64, device='cuda')
    for mode in ('joint', 'vortex', 'coexact'):
        k = [0]
        def f():
            m, v = sets[k[0] % 6]; k[0] += 1
            ops.worldline_sweep(m, v, 0.5, mode=mode, seed=1, sweep0=k[0], obs=obs)
        report(f'worldline {mode} L=64 x 1024 chains (C3 shard), smem path', CH * N * N, 24, timeit(f))
    obs6 = [torch.zeros_like(obs) for _ in range(6)]
    steppers = [ops.WorldlineOverlappedSweeps(m, v, 0.5, seed=1) for m, v in sets]
    k = [0]
    def g():
        i = k[0] % 6; k[0] += 1
        steppers[i].step(k[0], 1, obs6[i])
    report('worldline joint L=64 x 1024 chains (C3 shard), overlapped launches', CH * N * N, 24, timeit(g, n=60))
if 'c4' in which:
    N, CH = 128, 1024                                   # 1/8 of a GPU's C4 shard; 256 MiB of state
    S = svb.Villain(svb.Lattice2D(N), 0.5)
    phi, n = svb.BatchedEnsemble(S, CH)._start('hot', 1)
    obs = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda')
    kc = torch.linspace(0.3, 1.2, CH, dtype=torch.float64, device='cuda')
    for path, sw in (('auto', 1), ('auto', 2), ('auto', 10), ('global', 1), ('tiled', 1), ('tiled', 2), ('tiled', 10)):
        k = [0]
        def f():
            k[0] += sw
            ops.villain_sweep(phi, n, 0.5, seed=1, sweep0=k[0], kappa_chain=kc, obs=obs, path=path, n_sweeps=sw)
        report(f'villain L=128 x 1024 chains (C4), {path}, {sw} sweep(s)/call', CH * N * N * sw, 32, timeit(f, n=5))
    ov = ops.VillainOverlappedSweeps(phi, n, 0.5, seed=1, kappa_chain=kc)
    obs2 = [torch.zeros_like(obs) for _ in range(2)]
    k = [0]
    def g():
        k[0] += 1
        ov.step(k[0], 1, obs2[k[0] & 1], obs2[(k[0] & 1) ^ 1])
    report('villain L=128 x 1024 chains (C4), overlapped launches + obs_in, 1 sweep/call', CH * N * N, 32, timeit(g, n=20))
if 'c5' in which:
    N, CH = 4096, 1
    S = svb.Villain(svb.Lattice2D(N), 0.5)
    phi, n = svb.BatchedEnsemble(S, CH)._start('hot', 1)
    obs = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda')
    for path, sw in (('global', 1), ('tiled', 1), ('tiled', 2), ('tiled', 10)):
        k = [0]
        def f():
            k[0] += sw
            ops.villain_sweep(phi, n, 0.5, seed=1, sweep0=k[0], obs=obs, path=path, n_sweeps=sw)
        report(f'villain L=4096 x 1 chain (C5), {path}, {sw} sweep(s)/call', CH * N * N * sw, 32, timeit(f, n=5))
    swp = ops.VillainSwappingSweeps(phi, n, 0.5, seed=1)       # the state alternates between two buffer pairs: no copy back
    k = [0]
    def g():
        k[0] += 1
        swp.step(k[0], 1, obs)
    report('villain L=4096 x 1 chain (C5), tiled, swapping buffer pairs, 1 sweep/call', CH * N * N, 32, timeit(g, n=5))
if 'forms' in which:
    N, CH = 32, 65536                                    # 512 MiB f64 0-forms
    a = torch.randn((CH, 1, N, N), dtype=torch.float64, device='cuda')
    out = torch.empty((CH, 2, N, N), dtype=torch.float64, device='cuda')
    report('d(0-form f64) L=32 x 65536', CH * N * N, 24, timeit(lambda: ops.form_op('d', 0, a, out)))
    b = torch.randint(-3, 4, (CH, 2, N, N), dtype=torch.int32, device='cuda')
    out1 = torch.empty((CH, 1, N, N), dtype=torch.int32, device='cuda')
    report('d(1-form i32) L=32 x 65536', CH * N * N, 12, timeit(lambda: ops.form_op('d', 1, b, out1)))
    report('delta(1-form i32) L=32 x 65536', CH * N * N, 12, timeit(lambda: ops.form_op('delta', 1, b, out1)))
    phi = a
    n = b
    obs = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda')
    report('villain_observables L=32 x 65536', CH * N * N, 16, timeit(lambda: ops.villain_observables(phi, n, 0.5, obs=obs)))
if 'forms' in which:
    N, CH = 64, 16384
    m = torch.randint(-3, 4, (CH, 2, N, N), dtype=torch.int32, device='cuda')
    v = torch.randint(-3, 4, (CH, 1, N, N), dtype=torch.int32, device='cuda')
    wobs = torch.zeros((CH, ops.WOBS_COUNT), dtype=torch.float64, device='cuda')
    report('worldline_observables L=64 x 16384', CH * N * N, 12, timeit(lambda: ops.worldline_observables(m, v, obs=wobs)))
    N, CH = 32, 8192
    phi = torch.rand((CH, 1, N, N), dtype=torch.float64, device='cuda')
    report('spin_spin correlator L=32 x 8192 (shared-memory FFT)', CH * N * N, 24, timeit(lambda: ops.villain_spin_spin(phi), n=3, reps=2))
    for N, CH in ((128, 512), (4096, 1)):
        phi = torch.rand((CH, 1, N, N), dtype=torch.float64, device='cuda')
        out = torch.empty((CH, N, N, 2), dtype=torch.float64, device='cuda')
        # three launches: field in (8 B), rows out (16), columns in + out (32), rows in + out (32) per site
        report(f'spin_spin correlator L={N} x {CH} ({"three" if N <= 512 else "five"}-launch FFT)', CH * N * N, 88, timeit(lambda: ops.villain_spin_spin(phi, out=out), n=3, reps=2))
if 'dec' in which:
    N, CH = 32, 4096                                     # config-2 shape: the decoupled Villain updates and the Hammer sequence
    S = svb.Villain(svb.Lattice2D(N), 0.5)
    sets = [svb.BatchedEnsemble(S, CH)._start('hot', 1 + r) for r in range(4)]
    obs = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda')
    for kind, sites in (('site', N * N), ('link', 2 * N * N), ('exact', N * N)):
        k = [0]
        def f():
            phi, n = sets[k[0] % 4]; k[0] += 1
            ops.villain_decoupled(kind, phi, n, 0.5, seed=1, sweep0=k[0], obs=obs)
        report(f'villain {kind} update L=32 x 4096 chains', CH * sites, (24 if kind == 'link' else 32) * N * N / sites, timeit(f))
    cnt = torch.zeros((CH, 2), dtype=torch.float64, device='cuda')
    k = [0]
    def f():
        phi, n = sets[k[0] % 4]; k[0] += 1
        ops.villain_cohomology(phi, n, 0.5, seed=1, sweep=k[0], counters=cnt)
    report('villain cohomology update L=32 x 4096 chains (2 slice proposals per chain; reads 2 N links + sites, writes N links on accept)', CH * 2 * N, 32, timeit(f))
if 'wl128' in which:
    N, CH = 128, 512
    S = svb.Worldline(svb.Lattice2D(N), 0.5)
    m, v = svb.BatchedEnsemble(S, CH)._start('hot', 1)
    obs = torch.zeros((CH, ops.WOBS_COUNT), dtype=torch.float64, device='cuda')
    for mode in ('joint', 'vortex'):
        k = [0]
        def f():
            k[0] += 1
            ops.worldline_sweep(m, v, 0.5, mode=mode, seed=1, sweep0=k[0], obs=obs)
        report(f'worldline {mode} L=128 x 512 chains, table kernel (one chain per SM)', CH * N * N, 24, timeit(f))
    k = [0]
    def g():
        k[0] += 1
        ops.worldline_sweep(m, v, 0.5, mode='joint', seed=1, sweep0=k[0], obs=obs, path='global')
    report('worldline joint L=128 x 512 chains, global path', CH * N * N, 24, timeit(g, n=5))
if 'dec128' in which:
    N, CH = 128, 1024
    S = svb.Villain(svb.Lattice2D(N), 0.5)
    phi, n = svb.BatchedEnsemble(S, CH)._start('hot', 1)
    obs = torch.zeros((CH, ops.VOBS_COUNT), dtype=torch.float64, device='cuda')
    for kind, sites in (('site', N * N), ('link', 2 * N * N), ('exact', N * N)):
        k = [0]
        def f():
            k[0] += 1
            ops.villain_decoupled(kind, phi, n, 0.5, seed=1, sweep0=k[0], obs=obs)
        report(f'villain {kind} update L=128 x 1024 chains', CH * sites, (24 if kind == 'link' else 32) * N * N / sites, timeit(f, n=5))
if 'taxi' in which:
    N, CH = 32, 256
    links = torch.randn((CH, 2, N, N), dtype=torch.float64, device='cuda') * 0.3
    report('taxicab correlator (Vortex_Vortex.Villain) L=32 x 256 chains: N^4 exponentials per chain', CH * N ** 4, 0.016, timeit(lambda: ops.taxicab_correlator('vortex', links, 0.5), n=2, reps=2))
