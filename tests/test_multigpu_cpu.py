"""Host-side logic of the multi-GPU path, on CPU: sharding is a pure function, and the
one-process-per-GPU helpers are exercised with a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest

from pyjpegdecoder_b200.multigpu import reduce_max, shard_by_bytes, shard_range


@pytest.mark.parametrize("n,world", [(0, 1), (1, 8), (7, 2), (4096, 8), (4097, 8), (10, 3)])
def test_shard_range_partitions(n, world):
    seen = []
    sizes = []
    for r in range(world):
        lo, hi = shard_range(n, world, r)
        assert 0 <= lo <= hi <= n
        seen.extend(range(lo, hi))
        sizes.append(hi - lo)
    assert seen == list(range(n))
    assert max(sizes) - min(sizes) <= 1


def test_shard_by_bytes_balances():
    rng = np.random.default_rng(0)
    sizes = list(rng.integers(1_000, 20_000_000, 64))
    shards = shard_by_bytes(sizes, 8)
    assert sorted(i for s in shards for i in s) == list(range(64))
    loads = [sum(sizes[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(sizes)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(101, world, rank)
        m = reduce_max(float(rank + 1) * 10.0)
        # every rank reports its shard; rank 0 checks the partition
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi))
        dist.barrier()
        if rank == 0:
            out.put((m, gathered))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_group():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    m, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert m == 20.0
    assert gathered == [(0, 51), (51, 101)]
