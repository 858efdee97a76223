"""The N>1 path on CPU: world_size-2 gloo processes shard the chain axis, 'measure', and gather.  No GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from supervillain_b200 import sharding


def test_shard_chains_partitions_exactly():
    for total, world in ((4096, 1), (4096, 8), (8192, 8), (10, 4), (3, 8), (65536, 8)):
        blocks = [sharding.shard_chains(total, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
        for (a0, ac), (b0, _) in zip(blocks, blocks[1:]):
            assert a0 + ac == b0
    with pytest.raises(ValueError):
        sharding.shard_chains(8, 2, 2)


def test_kappa_scan_config4_layout():
    kappas = 0.3 + 0.9 * np.arange(64) / 63                      # C4
    seen = []
    for r in range(8):
        chain0, count, kc = sharding.kappa_scan(kappas, 1024, 8, r)
        assert count == 8192 and chain0 == r * 8192
        assert len(np.unique(kc)) == 8 and (kc.reshape(8, 1024) == kc.reshape(8, 1024)[:, :1]).all()
        seen.append(kc)
    assert (np.concatenate(seen) == np.repeat(kappas, 1024)).all()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    chain0, count = sharding.shard_chains(total, world, rank)
    ids = torch.arange(chain0, chain0 + count, dtype=torch.float64)
    local = torch.stack([ids, ids ** 2, -ids], dim=1)            # a stand-in observable record per GLOBAL chain id
    full = sharding.gather_columns(local)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                     # the bench's max-over-ranks timing reduction
    np.save(os.path.join(out_dir, f'rank{rank}.npy'), np.concatenate([full.numpy().ravel(), t.numpy()]))
    dist.destroy_process_group()


@pytest.mark.parametrize('total', [8, 7])
def test_two_rank_gather_reassembles_the_global_chain_axis(tmp_path, total):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    ids = np.arange(total, dtype=np.float64)
    expect = np.concatenate([np.stack([ids, ids ** 2, -ids], axis=1).ravel(), [2.0]])
    for r in range(world):
        assert (np.load(tmp_path / f'rank{r}.npy') == expect).all()
