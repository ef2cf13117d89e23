"""Randomised differential test: a few hundred small Pillow-encoded images with
random sizes, subsamplings, qualities, restart intervals, progressive/optimised tables -- decoded in ONE
batch on the GPU and compared bit-exactly (coefficient planes and RGB) with the CPU oracle."""
import io

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _random_cases(n, seed):
    from PIL import Image, ImageFile
    ImageFile.MAXBLOCK = max(ImageFile.MAXBLOCK, 1 << 22)   # optimised progressive noise overflows the default
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        w = int(rng.integers(1, 200)) if rng.random() < 0.8 else int(rng.integers(200, 700))
        h = int(rng.integers(1, 200)) if rng.random() < 0.8 else int(rng.integers(200, 500))
        gray = rng.random() < 0.15
        kind = rng.integers(0, 4)
        if kind == 0:      # smooth
            y, x = np.mgrid[0:h, 0:w]
            img = np.stack([(x * 3 + y) % 256, (x + y * 2) % 256, (x * y) % 256], -1)
        elif kind == 1:    # noise
            img = rng.integers(0, 256, (h, w, 3))
        elif kind == 2:    # flat patches (rounding ties, saturated colours)
            img = np.zeros((h, w, 3), np.int64)
            for by in range(0, h, 8):
                for bx in range(0, w, 8):
                    img[by:by + 8, bx:bx + 8] = rng.choice([0, 3, 125, 128, 131, 253, 255], 3)
        else:              # mixture
            y, x = np.mgrid[0:h, 0:w]
            img = np.stack([128 + 100 * np.sin(x / 7 + y / 11), 128 + 100 * np.cos(x / 5 - y / 9),
                            128 + 100 * np.sin((x + y) / 13)], -1) + rng.normal(0, 20, (h, w, 3))
        img = np.clip(img, 0, 255).astype(np.uint8)
        kw = dict(quality=int(rng.integers(1, 101)))
        if gray:
            img = img[..., 0]
        else:
            kw["subsampling"] = int(rng.integers(0, 3))
        if rng.random() < 0.4:
            kw["progressive"] = True
        if rng.random() < 0.3:
            kw["optimize"] = True
        r = rng.random()
        if r < 0.25:
            kw["restart_marker_blocks"] = int(rng.integers(1, 40))
        elif r < 0.35:
            kw["restart_marker_rows"] = int(rng.integers(1, 4))
        b = io.BytesIO()
        try:
            Image.fromarray(img).save(b, "JPEG", **kw)
        except OSError:            # libjpeg "suspension not allowed": Pillow's buffer is too small for
            b = io.BytesIO()       # some optimised progressive files with many restart markers
            kw.pop("restart_marker_blocks", None)
            kw.pop("restart_marker_rows", None)
            Image.fromarray(img).save(b, "JPEG", **kw)
        out.append((b.getvalue(), (w, h, kw)))
    return out


@pytest.mark.parametrize("seed", [1, 2])
def test_random_images_one_batch(seed):
    from pyjpegdecoder_b200 import decode_batch
    cases = _random_cases(160, seed)
    decs = decode_batch([c[0] for c in cases], device="cuda:0")
    bad = []
    for d, (data, desc) in zip(decs, cases):
        ref = oracle.decode(data, want=("rgb", "coef"))
        ok = np.array_equal(d.image_tensor.cpu().numpy(), ref.rgb)
        okc = all(np.array_equal(a, b) for a, b in zip(d.coefficient_planes(), ref.coef))
        if not (ok and okc):
            bad.append((desc, ok, okc))
    assert not bad, bad[:5]
