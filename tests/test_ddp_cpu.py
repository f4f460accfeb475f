"""Data-parallel training exchanges gradients with ONE bucketed all-reduce (mean) per step.  The
bucketing / readiness logic of fcwdm.ddp.GradSync is host code over a flat tensor: exercised here with a real
2-process gloo group on the CPU (on the GPU box the same object runs over NCCL on a side stream)."""
import os
import socket

import pytest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fcwdm.ddp import GradSync


def _offsets(sizes):
    offs, off = [], 0
    for n in sizes:
        offs.append((off, off + n))
        off += (n + 3) // 4 * 4
    return offs, off


def test_bucket_partition_covers_every_parameter_once():
    sizes = [5, 1000, 3, 64, 4096, 7, 7, 20000, 1]
    offs, total = _offsets(sizes)
    gs = GradSync(torch.zeros(total), offs, bucket_bytes=4096 * 4)
    assert sum(b[2] for b in gs.buckets) == len(sizes)
    assert gs.buckets[0][0] == 0 and gs.buckets[-1][1] == offs[-1][1]
    for a, b in zip(gs.buckets, gs.buckets[1:]):
        assert a[1] <= b[0]
    for lo, hi in offs:                                    # every parameter lies inside exactly one bucket
        b = gs.buckets[gs._bucket_of(lo)]
        assert b[0] <= lo and hi <= b[1]


def _worker(rank, world, port, wire="fp32"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      FCWDM_DDP_GRAD_DTYPE=wire)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [5, 1000, 3, 64, 4096, 7, 7, 20000, 1]
    offs, total = _offsets(sizes)
    flat = torch.zeros(total)
    gs = GradSync(flat, offs, bucket_bytes=4096 * 4)
    assert gs.wire_dtype == (torch.bfloat16 if wire == "bf16" else torch.float32)
    for step in range(2):                                  # two steps: begin() re-arms the buckets
        gs.begin()
        for i, (lo, hi) in enumerate(offs):
            flat[lo:hi] = float((rank + 1) * (i + 1) * (step + 1))
        order = list(reversed(range(len(offs))))           # backward order: last parameter first
        order.remove(2)                                    # one parameter never reports (unused): finish() covers it
        for i in order:
            gs.ready(*offs[i])
        gs.finish()
        mean_rank = sum(r + 1 for r in range(world)) / world
        for i, (lo, hi) in enumerate(offs):
            want = mean_rank * (i + 1) * (step + 1)
            assert torch.allclose(flat[lo:hi], torch.full((hi - lo,), want)), (rank, i, step)
        assert gs.launched == len(gs.buckets)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("wire", ["fp32", "bf16"])
def test_two_rank_gloo_gradient_mean(wire):
    """wire = bf16: the bucket crosses the wire as bf16 and is widened back into the fp32 flat gradient (the test's values
    and their sums are small integers and halves, exact in bf16)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, wire), nprocs=2, join=True)
