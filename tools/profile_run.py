#!/usr/bin/env python3
"""Short, profiler-friendly run of the decode pipeline: N synthetic 1080p images, a few steps.
Used under ncu (see profiles/README.md); prints nothing that is a benchmark value."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    files = bench.make_files(a.distinct)
    import torch
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import BatchPlan, DevicePipeline, pack_files
    datas = [files[i % len(files)] for i in range(a.images)]
    pd = [parse_jpeg(d) for d in files]
    parsed = [pd[i % len(files)] for i in range(a.images)]
    raw, offs = pack_files(datas)
    plan = BatchPlan(parsed, offs, raw.numel())
    pipe = DevicePipeline(plan, "cuda:0")
    pipe.upload(raw)
    for _ in range(a.steps):
        pipe.launch()
    torch.cuda.synchronize()
    print("profile run done", int(pipe.err.abs().sum()))


if __name__ == "__main__":
    main()
