"""world_size-2 gloo test (CPU) of the host side of the data-parallel step: bucketed SUM all-reduce of
slices of one flat gradient buffer (animal2vec_b200.trainer.BucketReducer) -- the N>1 path of bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from animal2vec_b200.trainer import BucketReducer

    n = 1000
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    red = BucketReducer(flat)
    assert red.enabled
    # buckets become final out of order (reverse block order in the real backward), then the tail ranges
    for lo, hi in [(600, 900), (300, 600), (0, 300), (900, 1000)]:
        red.reduce_range(lo, hi)
    red.finish()
    expect = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(flat, expect) and red.bytes_reduced == n * 4
    # packed statistics all-reduce (loss sum, sample size, column sums) in float64
    stats = torch.tensor([1.5 * (rank + 1), 100.0], dtype=torch.float64)
    dist.all_reduce(stats)
    ok = ok and stats.tolist() == [1.5 * 3, 200.0]
    out[rank] = ok
    dist.destroy_process_group()


def test_bucket_reducer_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_bucket_reducer_single_process_is_a_noop():
    from animal2vec_b200.trainer import BucketReducer

    flat = torch.ones(16)
    red = BucketReducer(flat)
    assert not red.enabled
    red.reduce_range(0, 16)
    red.finish()
    assert torch.equal(flat, torch.ones(16))
