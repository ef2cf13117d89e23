"""Streaming front end for very large inputs (the step before the hot path).

`decode_stream(files, chunk=512)` yields the decoded images chunk by chunk.  File reading, the threaded gather
into pinned memory and the host-side planning of chunk k+1 run on a worker thread while the GPU decodes chunk k
(the C helpers and numpy release the GIL), so the host work disappears behind the device time for all but the
first chunk.  Two pinned staging buffers alternate; a chunk's buffer is reused only after that chunk has been
synchronised (its status words were read back).  No CPU decode path: the worker only prepares inputs."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Iterable, Iterator, List, Optional, Sequence, Union

import torch

from .decoder import JpegDecoder, _read
from .parser import parse_jpeg
from .pipeline import FAST_PLAN_MIN_FILES, BatchPlan, decode_batch_on_device, pack_files, raise_for_errors
from .stages import require_cuda

Source = Union[str, Path, bytes, bytearray, memoryview]


def _chunks(files: Iterable[Source], n: int) -> Iterator[List[Source]]:
    buf: List[Source] = []
    for f in files:
        buf.append(f)
        if len(buf) == n:
            yield buf
            buf = []
    if buf:
        yield buf


_N_SLOTS = 3
_WORKERS = 1


class _Uploader:
    """Pinned staging slots + a copy stream: the upload of chunk k+1 runs while chunk k is being decoded."""

    def __init__(self, device):
        self.dev = require_cuda(device)
        with torch.cuda.device(self.dev):
            self.stream = torch.cuda.Stream(self.dev)
        self.done: List[Optional[torch.cuda.Event]] = [None] * _N_SLOTS

    def prepare(self, files: Sequence[Source], k: int, read_threads: int = 8):
        """Host side of chunk k: read, gather into a pinned slot, start the H2D copy, plan."""
        if any(not isinstance(f, (bytes, bytearray, memoryview)) for f in files) and len(files) > 1:
            with ThreadPoolExecutor(min(read_threads, len(files))) as ex:
                datas = list(ex.map(_read, files))
        else:
            datas = [_read(f) for f in files]
        slot = k % _N_SLOTS
        if self.done[slot] is not None:
            self.done[slot].synchronize()          # the copy that last read this pinned slot has finished
        packed = pack_files(datas, pin=True, reuse_slot=("loader", id(self), slot), walk=len(datas) >= FAST_PLAN_MIN_FILES)
        raw_host, offsets = packed
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            raw_dev = torch.empty(raw_host.numel(), dtype=torch.uint8, device=self.dev)
            raw_dev.copy_(raw_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.done[slot] = ev
        if len(datas) >= FAST_PLAN_MIN_FILES:
            from .fastplan import plan_batch
            plan = plan_batch(raw_host, offsets, [len(d) for d in datas], walked=getattr(raw_host, "_bj_walk", None))
        else:
            plan = BatchPlan([parse_jpeg(d) for d in datas], offsets, raw_host.numel())
        return list(files), packed, plan, raw_dev, ev


    def close(self) -> None:
        from .pipeline import _PINNED_POOL
        for slot in range(_N_SLOTS):
            if self.done[slot] is not None:
                self.done[slot].synchronize()
            _PINNED_POOL.pop(("loader", id(self), slot), None)


def _check(batch) -> None:
    raise_for_errors(batch.stats["_pipe"].err.cpu().numpy())


def decode_stream(files: Iterable[Source], chunk: int = 512, device: Optional[Union[str, torch.device]] = None
                  ) -> Iterator[List[JpegDecoder]]:
    """Decode an arbitrarily long sequence of files `chunk` at a time; yields one list of JpegDecoder objects
    per chunk, in order.  Errors of a file (NotJpeg, CorruptedJpeg, ...) are raised when its chunk is reached.
    Three things overlap: the worker thread prepares and uploads chunk k+1, the GPU decodes chunk k (its kernels
    are enqueued before chunk k-1 is checked), and the caller consumes chunk k-1."""
    if chunk < 1:
        raise ValueError("chunk must be positive")
    up = _Uploader(device)
    it = _chunks(files, chunk)
    depth = _N_SLOTS - 1                     # chunks prepared ahead of the one being enqueued
    try:
        yield from _stream(up, it, depth, device)
    finally:
        up.close()


def _stream(up, it, depth, device):
    with ThreadPoolExecutor(_WORKERS) as worker:
        queue = []
        k = 0

        def refill():
            nonlocal k
            while len(queue) < depth:
                nxt = next(it, None)
                if nxt is None:
                    return
                queue.append(worker.submit(up.prepare, nxt, k))
                k += 1

        refill()
        prev = None
        while queue:
            chunk_files, packed, plan, raw_dev, ev = queue.pop(0).result()
            refill()
            raw_dev.record_stream(torch.cuda.current_stream(up.dev))
            batch = decode_batch_on_device(None, device=device, packed=packed, plan=plan, check=False,
                                           raw_dev=raw_dev, raw_ready=ev)
            if prev is not None:
                _check(prev[0])
                yield [JpegDecoder(f, _batch=prev[0], _index=i) for i, f in enumerate(prev[1])]
            prev = (batch, chunk_files)
        if prev is not None:
            _check(prev[0])
            yield [JpegDecoder(f, _batch=prev[0], _index=i) for i, f in enumerate(prev[1])]
