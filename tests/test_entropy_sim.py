"""CPU checks of the entropy-stage device logic (host build of csrc/bj_entropy.cuh, orchestrated like
the kernels) against the reference-generated golden coefficient planes: bit-exact after every scan."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_case_names
from entropy_sim import decode_file


@pytest.mark.parametrize("name", golden_case_names())
@pytest.mark.parametrize("sub_bits", [1024, 128])
def test_entropy_logic_matches_reference(name, sub_bits):
    data = (GOLDEN / "cases" / f"{name}.jpg").read_bytes()
    z = np.load(GOLDEN / "cases" / f"{name}.npz")
    p, per_scan, _ = decode_file(data, sub_bits=sub_bits)
    if p.progressive:
        for k, grids in enumerate(per_scan, start=1):
            for c, g in enumerate(grids):
                assert np.array_equal(g, z[f"scan{k}_coef{c}"]), (k, c)
    for c, g in enumerate(per_scan[-1]):
        assert np.array_equal(g, z[f"coef{c}"]), c


def test_entropy_logic_base_image(golden_meta):
    import hashlib
    data = (GOLDEN / "base_image.jpg").read_bytes()
    p, per_scan, stats = decode_file(data)
    for k, grids in enumerate(per_scan):
        got = [hashlib.sha256(np.ascontiguousarray(g).tobytes()).hexdigest() for g in grids]
        assert got == golden_meta["base_image"]["scan_coef_sha256"][k], k


@pytest.mark.parametrize("seed", range(12))
def test_entropy_logic_random_images_against_oracle(seed):
    """Randomised differential check on the CPU: random small Pillow encodings (baseline / progressive, any
    subsampling, restart intervals, optimised tables) through the host build of the device decode logic, against
    the oracle's coefficient planes."""
    import io
    from PIL import Image
    import oracle
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(1, 150)), int(rng.integers(1, 120))
    if rng.random() < 0.5:
        img = rng.integers(0, 256, (h, w, 3))
    else:
        y, x = np.mgrid[0:h, 0:w]
        img = np.stack([128 + 100 * np.sin(x / 7 + y / 11), 128 + 100 * np.cos(x / 5 - y / 9),
                        128 + 100 * np.sin((x + y) / 13)], -1) + rng.normal(0, 15, (h, w, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    kw = dict(quality=int(rng.integers(5, 100)))
    if rng.random() < 0.2:
        img = img[..., 0]
    else:
        kw["subsampling"] = int(rng.integers(0, 3))
    if rng.random() < 0.6:
        kw["progressive"] = True
    if rng.random() < 0.3:
        kw["optimize"] = True
    if rng.random() < 0.4 and not kw.get("progressive"):
        kw["restart_marker_blocks"] = int(rng.integers(1, 20))
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    data = b.getvalue()
    ref = oracle.decode(data, want=("coef",))
    for sub_bits in (1024, 128):
        p, per_scan, _ = decode_file(data, sub_bits=sub_bits)
        for c, g in enumerate(per_scan[-1]):
            assert np.array_equal(g, np.asarray(ref.coef[c]).reshape(g.shape)), (kw, w, h, c, sub_bits)


def test_entropy_logic_survives_corrupted_scans():
    """Bit flips inside entropy-coded data: the device decode logic (host build) must either decode something or
    flag an error -- never hang or write outside its blocks (on the GPU that would poison the whole context)."""
    from pyjpegdecoder_b200.errors import JpegError
    from pyjpegdecoder_b200.parser import parse_jpeg
    rng = np.random.default_rng(0)
    names = golden_case_names()
    outcomes = {"decoded": 0, "flagged": 0}
    for _ in range(150):
        name = names[int(rng.integers(len(names)))]
        data = bytearray((GOLDEN / "cases" / f"{name}.jpg").read_bytes())
        p = parse_jpeg(bytes(data))
        sc = p.scans[int(rng.integers(len(p.scans)))]
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(sc.data_start, max(sc.data_start + 1, sc.data_end)))
            data[pos] ^= 1 << int(rng.integers(8))
        try:
            decode_file(bytes(data))
            outcomes["decoded"] += 1
        except (AssertionError, JpegError):
            outcomes["flagged"] += 1
    assert outcomes["decoded"] + outcomes["flagged"] == 150 and outcomes["flagged"] > 0
