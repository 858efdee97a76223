#!/usr/bin/env python
"""PCIe probe for the host-buffer path: raw H2D / D2H / duplex bandwidth and HostStepper chunking variants."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200.hostpath import HostStepper
from supervillain_b200.generator.villain import NeighborhoodUpdate

MB = 64
h = torch.empty(MB << 20, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(MB << 20, dtype=torch.uint8, pin_memory=True)
d = torch.empty(MB << 20, dtype=torch.uint8, device='cuda')
d2 = torch.empty(MB << 20, dtype=torch.uint8, device='cuda')
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
print(f'H2D {MB} MiB: {MB/1024/t(lambda: d.copy_(h, non_blocking=True)):.1f} GiB/s')
print(f'D2H {MB} MiB: {MB/1024/t(lambda: h.copy_(d, non_blocking=True)):.1f} GiB/s')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def duplex():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print(f'duplex {MB}+{MB} MiB: {2*MB/1024/t(duplex):.1f} GiB/s total')
S = svb.Villain(svb.Lattice2D(32), 0.5)
G = NeighborhoodUpdate(S, seed=1)
fields = svb.BatchedEnsemble(S, 4096)._start('hot', 1)
for chunks, streams in ((1, 1), (4, 2), (8, 3), (16, 4), (32, 4), (16, 8), (64, 8)):
    st = HostStepper(G, 4096, chunks=chunks, streams=streams)
    a, b = st.pinned_fields(from_device=fields)
    dt = t(lambda: st.step(a, b), n=10)
    print(f'HostStepper chunks={chunks:3d} streams={streams}: {dt*1e3:.3f} ms/step  {4096*1024/dt:.3e} upd/s  {(st.h2d_bytes+st.d2h_bytes)/dt/2**30:.1f} GiB/s both ways')
