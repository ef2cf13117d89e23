#!/usr/bin/env python3
"""Throughput of the PUBLIC batch entry point, everything included: list of file bytes in, device tensors out
(pack into pinned memory, marker walk + plan on the host, H2D, all kernels, status read-back)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    files = bench.make_files(32)
    datas = [files[i % len(files)] for i in range(n)]
    import torch
    from pyjpegdecoder_b200.pipeline import decode_batch_on_device
    for _ in range(2):
        r = decode_batch_on_device(datas, device="cuda:0")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 3
    for _ in range(k):
        r = decode_batch_on_device(datas, device="cuda:0")
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / k
    print(f"public API: {n} files in {dt*1e3:.1f} ms = {n*1920*1080/1e6/dt:.0f} MP/s ({dt/n*1e6:.1f} us/file)")


if __name__ == "__main__":
    main()
