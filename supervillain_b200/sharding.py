"""Multi-GPU plumbing: chains shard across ranks, nothing else does.

One process per GPU (`torch.distributed`, NCCL on the GPUs, gloo in the CPU tests).  Chains are
independent Markov chains, so the hot path has NO collective: every rank sweeps a contiguous
block of the global chain axis, and the Philox counter carries the GLOBAL chain id (`chain0`),
which makes the sampled configurations independent of the number of GPUs.  The only inter-GPU
traffic is the final gather of the per-chain observable columns.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_chains(total_chains, world_size, rank):
    """Contiguous block of the chain axis owned by `rank`: (chain0, count).  Remainders go to the low ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(int(total_chains), int(world_size))
    count = base + (1 if rank < extra else 0)
    chain0 = rank * base + min(rank, extra)
    return chain0, count


def kappa_scan(kappas, chains_per_kappa, world_size, rank):
    """BASELINE config 4 partitioning: the chains of a kappa scan, kappa-major, sharded contiguously.

    Returns (chain0, count, kappa_chain) where kappa_chain (count,) float64 is this rank's per-chain coupling,
    so that a whole block of chains shares a kappa and CTAs see uniform couplings.
    """
    kappas = np.asarray(kappas, dtype=np.float64)
    total = len(kappas) * int(chains_per_kappa)
    chain0, count = shard_chains(total, world_size, rank)
    ids = np.arange(chain0, chain0 + count)
    return chain0, count, kappas[ids // int(chains_per_kappa)]


def gather_columns(local, group=None):
    """Concatenate per-chain columns (chains_local, ...) from every rank along the chain axis, on every rank.

    Uneven shards are padded to the largest shard for the collective and trimmed afterwards.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    biggest = max(counts)
    padded = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
