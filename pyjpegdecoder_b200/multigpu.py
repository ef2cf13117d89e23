"""Batch sharding across the GPUs of one box.

Images are independent: image i of a batch goes to one GPU, every GPU runs the whole
device pipeline on its shard, and there is NO collective on the data path (no NCCL).  Two entry points:

  * decode_files_multi_gpu(): single process, one host thread + one CUDA stream per GPU;
  * shard_range() / reduce_max(): helpers for the one-process-per-GPU launch used by bench.py, where
    torch.distributed is only used to line the ranks up for timing (barrier, max over ranks).
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank`: sizes differ by at most one, earlier ranks get
    the larger shards, every item belongs to exactly one rank."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_by_bytes(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy balance by compressed size (largest first): returns the item indices of every rank.
    Used when images differ a lot in size, e.g. the mixed 8192x8192 batch of BASELINE.json configs[4]."""
    order = sorted(range(len(sizes)), key=lambda i: -sizes[i])
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: loads[k])
        out[r].append(i)
        loads[r] += sizes[i]
    for lst in out:
        lst.sort()
    return out


def reduce_max(value: float, device=None) -> float:
    """Max of a per-rank scalar over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def decode_files_multi_gpu(files: Sequence, devices: Optional[Sequence] = None, balance: str = "bytes"):
    """Decode a list of files on several GPUs of this box from ONE process: the list is sharded, each
    GPU gets a host thread that parses its shard and runs the device pipeline on its own stream.
    Returns JpegDecoder-like objects in the order of `files` (pixels stay on their GPU)."""
    import torch
    from .decoder import JpegDecoder, _read
    from .pipeline import decode_batch_on_device
    if devices is None:
        devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    if not devices:
        from .errors import NativeLibraryError
        raise NativeLibraryError("no CUDA device is visible: the B200 decode path has no CPU fallback")
    datas = [_read(f) for f in files]
    world = min(len(devices), max(1, len(datas)))
    if balance == "bytes":
        shards = shard_by_bytes([len(d) for d in datas], world)
    else:
        shards = [list(range(*shard_range(len(datas), world, r))) for r in range(world)]
    results: List = [None] * len(datas)
    errors: List = []

    def work(r: int):
        try:
            idx = shards[r]
            if not idx:
                return
            dev = torch.device(devices[r])
            stream = torch.cuda.Stream(dev)
            batch = decode_batch_on_device([datas[i] for i in idx], device=dev, stream=stream)
            stream.synchronize()
            for k, i in enumerate(idx):
                results[i] = JpegDecoder(files[i], _batch=batch, _index=k)
        except BaseException as e:  # re-raised in the caller
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results
