#!/usr/bin/env python3
"""One large baseline image (4160x2340 4:2:0, optional restart interval): single-image latency profiling target."""
import io, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from pyjpegdecoder_b200.pipeline import decode_batch_on_device


def big_baseline(w=4160, h=2340, seed=7, **kw):
    from PIL import Image
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.clip(np.stack([128 + 90 * np.sin(x / 37 + y / 53), 128 + 90 * np.cos(x / 29 - y / 41),
                            128 + 90 * np.sin((x + y) / 61)], -1) + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", quality=90, subsampling=2, **kw)
    return b.getvalue()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    kw = {"restart_marker_rows": 1} if len(sys.argv) > 2 and sys.argv[2] == "dri" else {}
    data = big_baseline(**kw)
    for _ in range(n):
        decode_batch_on_device([data], device="cuda:0")
    torch.cuda.synchronize()
    print("ok", len(data))
