#!/usr/bin/env python3
"""Exact-path rates of the fused pixel kernel on the bench workload: share of blocks recomputed exactly (stats[0]) and of
pixels converted by the exact colour path (stats[1]).  usage (GPU box): python tools/pix_stats.py [images]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    import torch
    from pyjpegdecoder_b200 import _native
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import BatchPlan, DevicePipeline, pack_files
    from pyjpegdecoder_b200.stages import run_pixels
    files = bench.make_files(n)
    parsed = [parse_jpeg(d) for d in files]
    raw, offs = pack_files(files)
    plan = BatchPlan(parsed, offs, raw.numel())
    pipe = DevicePipeline(plan, "cuda:0")
    pipe.upload(raw)
    pipe.launch()
    stats = torch.zeros(4, dtype=torch.int32, device="cuda:0")
    run_pixels(pipe.dg, pipe.coef, _native.IN_COEF, _native.OUT_RGB, stats=stats)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    blocks = plan.geom.total_blocks
    pixels = sum(p.width * p.height for p in parsed)
    print(f"{n} images: {blocks} blocks, exact-recompute blocks {s[0]} ({100.0 * s[0] / blocks:.3f} %), "
          f"exact-colour pixels {s[1]} ({100.0 * s[1] / pixels:.4f} %)")


if __name__ == "__main__":
    main()
