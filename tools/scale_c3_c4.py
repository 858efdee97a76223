#!/usr/bin/env python
"""Configs 3 and 4 at their full per-GPU shard sizes on N GPUs (torch.distributed.run, one rank per GPU): chains are
sharded with sharding.shard_chains / kappa_scan, no collective inside the timed region, CUDA-event time per step, MAX
over ranks.  C3: worldline L=64, 8192 chains over 8 GPUs (1024 per GPU).  C4: L=128, 64 kappa x 1024 chains over 8 GPUs
(8 kappa x 1024 = 8192 chains, 2 GiB per GPU).  With fewer GPUs the per-GPU shard stays the same (weak scaling)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supervillain_b200 as svb                                                                  # noqa: E402
from supervillain_b200 import ops, sharding                                                      # noqa: E402

PEAK = 6538.6e9
world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', '1'), ('RANK', '0'), ('LOCAL_RANK', '0')))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)


def timed(step, steps=60, warmup=10):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def line(name, updates_per_rank, bytes_per, t, extra):
    if rank == 0:
        print(json.dumps({'config': name, 'n_gpus': world, 'us_per_step_max_over_ranks': t * 1e6,
                          'site_updates_per_s': world * updates_per_rank / t,
                          'roofline_frac_per_gpu': updates_per_rank * bytes_per / t / PEAK, **extra}), flush=True)


# ---- C3: worldline L=64, 1024 chains per GPU, overlapped launches, 6 chain sets rotated (6 x 48 MiB > L2) ----
N, per_gpu = 64, 1024
chain0, count = sharding.shard_chains(per_gpu * world, world, rank)
S = svb.Worldline(svb.Lattice2D(N), 0.5)
sets = [svb.BatchedEnsemble(S, count, chain0=chain0)._start('hot', 100 * rank + r) for r in range(6)]
obs = [torch.zeros((count, ops.WOBS_COUNT), dtype=torch.float64, device=dev) for _ in range(6)]
steppers = [ops.WorldlineOverlappedSweeps(m, v, 0.5, seed=1, chain0=chain0) for m, v in sets]
k = [0]
def c3():
    i = k[0] % 6; k[0] += 1
    steppers[i].step(k[0], 1, obs[i])
t = timed(c3)
line('C3 worldline L=64 PlaquetteUpdate (checkerboard), 1024 chains/GPU, 1 sweep + observables per step', count * N * N, 24, t,
     {'chains_total': per_gpu * world})
del sets, steppers, obs
torch.cuda.empty_cache()

# ---- C4: L=128, kappa scan, 8 kappa x 1024 chains per GPU (2 GiB of state: far beyond L2), overlapped launches ----
N, per_kappa, kappas_per_gpu = 128, 1024, 8
kappas = 0.3 + 0.9 * np.arange(kappas_per_gpu * world) / max(kappas_per_gpu * world - 1, 1)
chain0, count, kc = sharding.kappa_scan(kappas, per_kappa, world, rank)
S = svb.Villain(svb.Lattice2D(N), 0.5)
phi, n = svb.BatchedEnsemble(S, count, chain0=chain0)._start('hot', 7 + rank)
kc = torch.as_tensor(np.asarray(kc, dtype=np.float64)).to(dev)
ov = ops.VillainOverlappedSweeps(phi, n, 0.5, seed=1, chain0=chain0, kappa_chain=kc)
obs2 = [torch.zeros((count, ops.VOBS_COUNT), dtype=torch.float64, device=dev) for _ in range(2)]
k = [0]
def c4():
    k[0] += 1
    ov.step(k[0], 1, obs2[k[0] & 1], obs2[(k[0] & 1) ^ 1])
t = timed(c4, steps=20, warmup=5)
line('C4 Villain L=128 kappa scan, 8 kappa x 1024 chains/GPU, 1 sweep + observables per step', count * N * N, 32, t,
     {'chains_total': count * world, 'kappas_total': len(kappas)})
if world > 1:
    dist.destroy_process_group()
