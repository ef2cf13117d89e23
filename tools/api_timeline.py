#!/usr/bin/env python3
"""Timeline of one public decode_batch() call (GPU box): when the worker thread prepares each sub-batch, when the main
thread enqueues and checks it.  usage: python tools/api_timeline.py [files] [chunk]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else None
    import torch
    from pyjpegdecoder_b200 import decode_batch, loader, pipeline
    files = bench.make_files(64)
    datas = [files[i % len(files)] for i in range(n)]
    log = []

    def wrap(obj, name, tag):
        fn = getattr(obj, name)

        def inner(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            log.append((tag, t0, time.perf_counter()))
            return r
        setattr(obj, name, inner)

    wrap(loader._Uploader, "pack_stage", "pack")
    wrap(loader._Uploader, "plan_stage", "plan")
    gpu = []
    real_dbod = loader.decode_batch_on_device

    def dbod(*a, **k):
        st = k["stream"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        r = real_dbod(*a, **k)
        e1.record(st)
        gpu.append((e0, e1, len(k["plan"].images) if hasattr(k["plan"], "images") else -1))
        return r
    loader.decode_batch_on_device = dbod
    wrap(loader, "decode_batch_on_device", "enqueue")
    wrap(loader, "_check", "check")
    runs = []
    for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
        log.clear()
        gpu.clear()
        t0 = time.perf_counter()
        res = decode_batch(datas, device="cuda:0", chunk=chunk)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        del res
        print(f"rep {rep}: {1e3 * (t1 - t0):.1f} ms")
        if rep:
            runs.append((t1 - t0, t0, list(log), [(a, b) for a, b, _ in gpu]))
    for title, (dt, t0, lg, gp) in (("slowest", max(runs, key=lambda r: r[0])), ("fastest", min(runs, key=lambda r: r[0]))):
        print(f"--- {title}: {1e3 * dt:.1f} ms")
        for tag, a, b in sorted(lg, key=lambda x: x[1]):
            print(f"{tag:10s} {1e3 * (a - t0):7.2f} -> {1e3 * (b - t0):7.2f}  ({1e3 * (b - a):5.2f} ms)")
        base = gp[0][0]
        for e0, e1 in gp:
            print(f"gpu sub-batch: start {base.elapsed_time(e0):7.2f}  end {base.elapsed_time(e1):7.2f}  ({e0.elapsed_time(e1):5.2f} ms)")


if __name__ == "__main__":
    main()
