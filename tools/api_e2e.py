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


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "stream"):
    main()


def stream_main():
    """decode_stream over n files in chunks: host preparation of chunk k+1 overlaps the GPU work on chunk k."""
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    files = bench.make_files(32)
    datas = [files[i % len(files)] for i in range(n)]
    import torch
    from pyjpegdecoder_b200 import decode_stream
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cnt = 0
        for decs in decode_stream(datas, chunk=chunk, device="cuda:0"):
            cnt += len(decs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"decode_stream: {cnt} files, chunk {chunk}, {dt*1e3:.1f} ms = {cnt*1920*1080/1e6/dt:.0f} MP/s")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "stream":
    stream_main()
