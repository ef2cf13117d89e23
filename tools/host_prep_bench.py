#!/usr/bin/env python3
"""Host side of the public batch call, stage by stage (GPU box): gather into pinned memory, marker walk + hash, plan.
usage: python tools/host_prep_bench.py [files per chunk]"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    import torch  # noqa: F401
    from pyjpegdecoder_b200.fastplan import plan_batch
    from pyjpegdecoder_b200.pipeline import pack_files
    files = bench.make_files(32)
    datas = [files[i % len(files)] for i in range(n)]
    mb = sum(map(len, datas)) / 1e6
    print(f"cpus {os.cpu_count()}  {n} files  {mb:.0f} MB")
    for rep in range(4):
        t0 = time.perf_counter()
        raw, offs = pack_files(datas, pin=True, reuse_slot="a", walk=False)
        t1 = time.perf_counter()
        raw2, offs2 = pack_files(datas, pin=True, reuse_slot="b", walk=True)
        t2 = time.perf_counter()
        plan_batch(raw2, offs2, [len(d) for d in datas], walked=getattr(raw2, "_bj_walk", None))
        t3 = time.perf_counter()
        print(f"gather {1e3 * (t1 - t0):6.2f} ms ({mb / (t1 - t0) / 1e3:5.1f} GB/s)   gather+walk+hash {1e3 * (t2 - t1):6.2f} ms   plan {1e3 * (t3 - t2):6.2f} ms")


if __name__ == "__main__":
    main()
