#!/usr/bin/env python3
"""Wall-clock of the public batch call: python tools/api_e2e.py [files] [chunk]   (GPU box)"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else None
    import torch
    from pyjpegdecoder_b200 import decode_batch
    files = bench.make_files(64)
    datas = [files[i % len(files)] for i in range(n)]
    for rep in range(5):
        t0 = time.perf_counter()
        res = decode_batch(datas, device="cuda:0", chunk=chunk)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        del res
        t2 = time.perf_counter()
        print(f"rep {rep}: decode_batch {1e3 * (t1 - t0):7.1f} ms = {n * bench.W * bench.H / 1e6 / (t1 - t0):9.0f} MP/s   free {1e3 * (t2 - t1):.1f} ms")


if __name__ == "__main__":
    main()
