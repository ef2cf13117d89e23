#!/usr/bin/env python3
"""Per-group device timing of progressive decodes: the reference's own
test image, a synthetic 4160x2340 4:2:0 progressive file, and batches of each."""
import io, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from pyjpegdecoder_b200.parser import parse_jpeg
from pyjpegdecoder_b200.pipeline import BatchPlan, DevicePipeline, pack_files


def run(name, datas, reps=3):
    packed, offs = pack_files(datas)
    plan = BatchPlan([parse_jpeg(d) for d in datas], offs, packed.numel())
    pipe = DevicePipeline(plan, "cuda:0")
    pipe.upload(packed)
    for _ in range(2):
        pipe.launch()
    torch.cuda.synchronize()
    ev = {}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        pipe.launch(events=ev)
    b.record()
    torch.cuda.synchronize()
    mp = sum(p.width * p.height for p in plan.parsed) / 1e6
    tot = a.elapsed_time(b) / reps
    print(f"{name}: {len(datas)} files, {mp:.1f} MP, {tot:.2f} ms/step = {mp / tot * 1e3:.0f} MP/s; groups={[(g.mode, g.count) for g in plan.groups]}")
    for k, v in ev.items():
        ts = [x.elapsed_time(y) for x, y in v]
        n = len(ts) // reps
        per = [np.mean(ts[i::n]) for i in range(n)]
        print("   ", k, " ".join(f"{t:.3f}" for t in per))


def main():
    from PIL import Image
    base = (ROOT / "tests/golden/base_image.jpg").read_bytes()
    run("base_image x1", [base])
    run("base_image x64", [base] * 64)
    if True:
        rng = np.random.default_rng(7)
        y, x = np.mgrid[0:2340, 0:4160]
        img = np.clip(np.stack([128 + 90 * np.sin(x / 37 + y / 53), 128 + 90 * np.cos(x / 29 - y / 41),
                                128 + 90 * np.sin((x + y) / 61)], -1) + rng.normal(0, 12, (2340, 4160, 3)), 0, 255).astype(np.uint8)
    # SURVEY.md section 8d generator for config C3: quality 75
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", quality=75, subsampling=2, progressive=True)
    run("C3 prog 4160x2340 q75 x1", [b.getvalue()])
    try:
        import oracle
        t0 = time.perf_counter()
        oracle.decode(b.getvalue(), want=("rgb",))
        print(f"    CPU oracle (1 core) on the same file: {1e3 * (time.perf_counter() - t0):.0f} ms")
    except Exception as e:  # noqa
        print("    oracle unavailable:", e)
    for kw, nm in ((dict(progressive=True), "C3 prog 4160x2340 q90 (dense worst case)"), (dict(progressive=True, restart_marker_rows=1), "C3 prog +DRI"),
                   (dict(), "C3-size baseline")):
        b = io.BytesIO()
        Image.fromarray(img).save(b, "JPEG", quality=90, subsampling=2, **kw)
        run(nm + " x1", [b.getvalue()])
        if "DRI" not in nm:
            run(nm + " x16", [b.getvalue()] * 16)


if __name__ == "__main__":
    main()
