import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb
from supervillain_b200.generator.villain import NeighborhoodUpdate
S = svb.Villain(svb.Lattice2D(32), 0.5); G = NeighborhoodUpdate(S, seed=1)
phi, n = svb.BatchedEnsemble(S, 2)._start('hot', 1)
obs = torch.zeros((2, 6), dtype=torch.float64, device='cuda')
ov = G.overlapped_device(phi, n)
for name, fn in (('sweep_device', lambda: G.sweep_device(phi, n, 1, obs=obs)), ('plan', G.plan_device(phi, n, obs=obs)),
                 ('overlapped', lambda: ov(1, obs))):
    for _ in range(100): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5000): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'{name:14s}: {1e6*(t1-t0)/5000:.1f} us per call issued (CPU), {1e6*(t2-t0)/5000:.1f} us incl. drain')
