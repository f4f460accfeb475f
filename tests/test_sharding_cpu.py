"""Multi-GPU sampling shards the list of volumes by rank with no data-path collective.  The
partition logic is host code: exercised here with a real 2-process gloo group on the CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fcwdm.pipeline import shard_indices


def test_shard_indices_partition_properties():
    for n in (0, 1, 7, 8, 13, 64):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))                                  # disjoint, ordered cover
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, n_items):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_items, rank, world)
    # the only communication of the sampling path: a barrier and a max-reduction of the per-rank time
    mark = torch.zeros(n_items)
    mark[mine] = 1
    dist.all_reduce(mark)                      # test-only: every item owned exactly once across ranks
    assert bool((mark == 1).all())
    elapsed = torch.tensor([float(10 + rank)])
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    assert float(elapsed) == 10 + world - 1
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 13), nprocs=2, join=True)
