"""GPU parity of the full decode path (parser -> un-stuff -> entropy -> pixels) through the public
entry point, against reference-generated fixtures.  Coefficient planes and RGB are bit-exact."""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, golden_case_names

pytestmark = pytest.mark.gpu


def _case(name):
    return (GOLDEN / "cases" / f"{name}.jpg").read_bytes(), np.load(GOLDEN / "cases" / f"{name}.npz")


def test_batch_all_golden_cases():
    """Every fixture (baseline + progressive, all subsamplings, DRI on/off) in ONE batch."""
    from pyjpegdecoder_b200 import decode_batch
    names = golden_case_names()
    decs = decode_batch([(GOLDEN / "cases" / f"{n}.jpg").read_bytes() for n in names], device="cuda:0")
    bad = []
    for name, d in zip(names, decs):
        z = np.load(GOLDEN / "cases" / f"{name}.npz")
        ok_rgb = d.image_array.shape == z["rgb"].shape and np.array_equal(d.image_array, z["rgb"])
        planes = d.coefficient_planes()
        ok_coef = all(np.array_equal(planes[c], z[f"coef{c}"]) for c in range(len(planes)))
        if not (ok_rgb and ok_coef):
            bad.append((name, ok_rgb, ok_coef))
    assert not bad, bad


@pytest.mark.parametrize("name", ["base_70x50_ss2", "base_gray_33x17_dri2", "prog_97x61_ss1_dri5", "base_1x1_ss2"])
def test_single_file_entry_point(name):
    from pyjpegdecoder_b200 import JpegDecoder
    data, z = _case(name)
    d = JpegDecoder(GOLDEN / "cases" / f"{name}.jpg", device="cuda:0")
    assert d.image_array.dtype == np.uint8
    assert np.array_equal(d.image_array, z["rgb"])
    assert d.image_tensor.is_cuda
    r = oracle.decode(data)
    assert (d.image_width, d.image_height) == (r.width, r.height)


@pytest.mark.parametrize("name", [n for n in golden_case_names() if n.startswith("prog_")][::3])
def test_progressive_coefficients_after_every_scan(name):
    """Bit-exact coefficient planes after each scan (bug-compatible AC refinement, jpeg_decoder.py:1114)."""
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import decode_batch_on_device
    data, z = _case(name)
    p = parse_jpeg(data)
    for k in range(1, len(p.scans) + 1):
        res = decode_batch_on_device([data], device="cuda:0", upto_wave=k)
        grids = res.coefficient_grids(0)
        for c, gr in enumerate(grids):
            assert np.array_equal(gr, z[f"scan{k}_coef{c}"]), (k, c)


def test_base_image_full(golden_meta):
    """The reference's own 4160x2340 progressive example with per-scan DRI: final RGB hash."""
    from pyjpegdecoder_b200 import JpegDecoder
    g = golden_meta["base_image"]
    d = JpegDecoder(GOLDEN / "base_image.jpg", device="cuda:0")
    assert d.image_array.shape == (g["width"], g["height"], 3)
    assert hashlib.sha256(np.ascontiguousarray(d.image_array).tobytes()).hexdigest() == g["rgb_sha256"]
    planes = d.coefficient_planes()
    assert [hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() for x in planes] == g["scan_coef_sha256"][-1]


@pytest.mark.parametrize("k", [1, 2])
def test_after_scan_renders(k, golden_meta):
    from pyjpegdecoder_b200 import JpegDecoder
    g = golden_meta["base_image"]
    data = (GOLDEN / "base_image.jpg").read_bytes()
    cut = data[: g[f"after_scan_{k}"]["truncate_at"]] + b"\xff\xd9"
    d = JpegDecoder(cut, device="cuda:0")
    hw3 = np.ascontiguousarray(np.swapaxes(d.image_array, 0, 1))
    assert hashlib.sha256(hw3.tobytes()).hexdigest() == g[f"after_scan_{k}"]["rgb_hw3_sha256"]


def test_corrupted_scans_never_poison_the_device():
    """Bit flips inside the entropy-coded data of many files: every batch either decodes or raises CorruptedJpeg,
    and a clean file still decodes bit-exactly afterwards (no illegal access, no hang)."""
    from pyjpegdecoder_b200 import JpegError, decode_batch
    from pyjpegdecoder_b200.parser import parse_jpeg
    rng = np.random.default_rng(5)
    names = golden_case_names()
    flagged = 0
    for rnd in range(6):
        batch = []
        for _ in range(24):
            name = names[int(rng.integers(len(names)))]
            data = bytearray((GOLDEN / "cases" / f"{name}.jpg").read_bytes())
            p = parse_jpeg(bytes(data))
            sc = p.scans[int(rng.integers(len(p.scans)))]
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(sc.data_start, max(sc.data_start + 1, sc.data_end)))
                data[pos] ^= 1 << int(rng.integers(8))
            batch.append(bytes(data))
        try:
            decode_batch(batch, device="cuda:0")
        except JpegError:
            flagged += 1
    assert flagged > 0
    name = names[0]
    data, z = _case(name)
    d = decode_batch([data] * 4, device="cuda:0")[0]
    assert np.array_equal(d.image_array, z["rgb"])


def test_output_consumers_dlpack_cuda_array_interface_and_save(tmp_path):
    """The step after the path: zero-copy hand-off and save() (jpeg_decoder.py:1485-1532)."""
    import torch
    from PIL import Image
    from pyjpegdecoder_b200 import JpegDecoder
    data, z = _case("base_120x88_ss2")
    src = tmp_path / "picture.jpg"
    src.write_bytes(data)
    d = JpegDecoder(src, device="cuda:0")
    t = torch.from_dlpack(d)
    assert t.is_cuda and t.data_ptr() == d.image_tensor.data_ptr()
    assert np.array_equal(np.swapaxes(t.cpu().numpy(), 0, 1), z["rgb"])
    cai = d.__cuda_array_interface__
    assert cai["shape"] == tuple(d.image_tensor.shape) and cai["typestr"] == "|u1" and cai["data"][0] == d.image_tensor.data_ptr()
    p1 = d.save()
    p2 = d.save()
    assert p1 == tmp_path / "picture.png" and p2 == tmp_path / "picture (1).png"
    assert np.array_equal(np.swapaxes(np.array(Image.open(p1)), 0, 1), z["rgb"])
    p3 = d.save(tmp_path / "out.unknownext")
    assert p3.suffix == ".png" and p3.exists()


def test_on_error_return_reports_each_bad_file_and_decodes_the_rest():
    """decode_batch(on_error="return"): bad files come back as exception instances in their positions (header
    problems found on the host, entropy-data problems from the per-image device error words); good files of the
    same batch are bit-exact."""
    from pyjpegdecoder_b200 import CorruptedJpeg, JpegDecoder, JpegError, NotJpeg, decode_batch
    from pyjpegdecoder_b200.parser import parse_jpeg
    names = golden_case_names()[:6]
    files, expect = [], []
    for name in names:
        data, z = _case(name)
        files.append(data)
        expect.append(z["rgb"])
    files.insert(1, b"not a jpeg at all")
    expect.insert(1, NotJpeg)
    good, _ = _case(names[0])
    files.insert(3, good[: len(good) // 2])                       # header intact, entropy data and EOI missing
    expect.insert(3, JpegError)
    p = parse_jpeg(good)
    sc = p.scans[0]
    broken = bytearray(good)
    mid = (sc.data_start + sc.data_end) // 2
    broken[mid:mid + 64] = bytes(64)                              # 512 zero bits: no such run of codes in this file
    files.append(bytes(broken))
    expect.append(None)                                           # decodes to something or is flagged, never raises
    res = decode_batch(files, device="cuda:0", on_error="return")
    assert len(res) == len(files)
    for r, e in zip(res, expect):
        if e is None:
            assert isinstance(r, (JpegDecoder, CorruptedJpeg))
        elif isinstance(e, type):
            assert isinstance(r, e), r
        else:
            assert isinstance(r, JpegDecoder)
            assert np.array_equal(r.image_array, e)
    with pytest.raises(JpegError):
        decode_batch(files, device="cuda:0")                      # the default still raises
