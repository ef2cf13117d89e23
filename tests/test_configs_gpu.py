"""GPU parity at the BASELINE.json configuration sizes.  The Python reference cannot run on the GPU box
(and would need minutes per image), so these compare the CUDA path with the CPU oracle (pinned to the
reference by tests/test_oracle.py) on the same seeded synthetic files, bit-exactly: coefficient planes
and RGB.  Generator = SURVEY.md 8(d) (bench.synth_image)."""
import io

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _encode(w, h, seed, gray=False, **kw):
    from PIL import Image
    import bench
    img = bench.synth_image(seed, w, h)
    if gray:
        img = img[..., 0]
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", quality=75, **kw)
    return b.getvalue()


def _check(datas):
    from pyjpegdecoder_b200 import decode_batch
    decs = decode_batch(datas, device="cuda:0")
    for d, data in zip(decs, datas):
        ref = oracle.decode(data, want=("rgb", "coef"))
        planes = d.coefficient_planes()
        for c in range(len(planes)):
            assert np.array_equal(planes[c], ref.coef[c]), f"coefficient plane {c} differs"
        got = d.image_tensor.cpu().numpy()
        assert got.shape == ref.rgb.shape
        assert np.array_equal(got, ref.rgb), f"max diff {np.abs(got.astype(int) - ref.rgb.astype(int)).max()}"
        assert d.image_array.shape[:2] == (ref.width, ref.height)   # reference layout is x-major (:626)


def test_config1_512x512_baseline_420():
    _check([_encode(512, 512, 0, subsampling=2)])


@pytest.mark.parametrize("kw", [dict(restart_marker_rows=1), dict(restart_marker_blocks=16)])
def test_config2_4k_baseline_420_with_restart_intervals(kw):
    _check([_encode(3840, 2160, 1, subsampling=2, **kw)])


def test_config3_progressive_10_scans_4160x2340():
    _check([_encode(4160, 2340, 2, subsampling=2, progressive=True)])


def test_config4_batch_of_1080p_baseline_420():
    datas = [_encode(1920, 1080, s, subsampling=2) for s in range(12)]
    _check(datas)
    # a batch decodes to the same pixels as its images one by one (sharding does not change results)
    from pyjpegdecoder_b200 import JpegDecoder, decode_batch
    one = JpegDecoder(datas[3], device="cuda:0").image_tensor
    many = decode_batch(datas, device="cuda:0")[3].image_tensor
    assert bool((one == many).all())


@pytest.mark.parametrize("kind", ["gray", "422", "444"])
def test_config5_8192x8192_baseline_mixed_subsampling(kind):
    kw = {"gray": dict(gray=True), "422": dict(subsampling=1), "444": dict(subsampling=0)}[kind]
    _check([_encode(8192, 8192, 5, **kw)])


@pytest.mark.parametrize("kind", ["gray", "422", "444"])
def test_config5_progressive_mixed_subsampling(kind):
    kw = {"gray": dict(gray=True), "422": dict(subsampling=1), "444": dict(subsampling=0)}[kind]
    _check([_encode(2048, 2048, 6, progressive=True, **kw)])


@pytest.mark.parametrize("kind", ["gray", "422", "444"])
def test_config5_8192x8192_progressive(kind):
    """BASELINE.json configs[4] at its full size in progressive mode (10 scans / 6 for grey, no restart markers)."""
    kw = {"gray": dict(gray=True), "422": dict(subsampling=1), "444": dict(subsampling=0)}[kind]
    _check([_encode(8192, 8192, 15, progressive=True, **kw)])


@pytest.mark.parametrize("name,layout", [("w440_96x80", 3), ("w440_33x41_dri3", 3), ("wgen_y22_cb21_cr11_48x48", 0),
                                         ("wgen_y22_cb12_cr11_50x37", 0), ("wgen_y21_cb11_cr21_64x24", 0),
                                         ("wgen_y22_cb22_cr11_32x32", 0), ("wwrap_16x16_ss0", 4)])
def test_layouts_pillow_cannot_encode(name, layout):
    """4:4:0 (its own kernel instance), chroma components with their own sampling factors (the generic kernel) and a
    dequantisation product that wraps int16 (:869): files from tests/jpeg_writer.py, decoded ALONE so that the kernel
    under test is the only one launched, against the fixtures the unmodified reference produced."""
    from conftest import GOLDEN
    from pyjpegdecoder_b200 import JpegDecoder
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.plan import layout_of
    data = (GOLDEN / "cases" / f"{name}.jpg").read_bytes()
    z = np.load(GOLDEN / "cases" / f"{name}.npz")
    assert layout_of(parse_jpeg(data)) == layout
    d = JpegDecoder(data, device="cuda:0")
    assert np.array_equal(d.image_array, z["rgb"])
    for c, plane in enumerate(d.coefficient_planes()):
        assert np.array_equal(plane, z[f"coef{c}"])


def test_config5_mixed_batch_one_launch_sequence():
    """grey + 4:2:2 + 4:4:4 + 4:2:0, baseline and progressive, odd sizes, in ONE batch."""
    datas = [_encode(1000, 700, 7, gray=True), _encode(1023, 769, 8, subsampling=1),
             _encode(801, 1201, 9, subsampling=0), _encode(640, 480, 10, subsampling=2, progressive=True),
             _encode(333, 777, 11, subsampling=1, progressive=True, restart_marker_blocks=11),
             _encode(1280, 720, 12, subsampling=2, restart_marker_rows=2)]
    _check(datas)


def test_truncated_entropy_data_raises_corrupted():
    from pyjpegdecoder_b200 import CorruptedJpeg, JpegDecoder
    data = _encode(512, 512, 0, subsampling=2)
    cut = data[: len(data) * 2 // 3] + b"\xff\xd9"
    with pytest.raises(CorruptedJpeg):
        JpegDecoder(cut, device="cuda:0")


def test_multi_gpu_dispatcher_single_device():
    """The dispatcher with the devices available here (one on the test box): same pixels, same order."""
    import torch
    from pyjpegdecoder_b200 import decode_files_multi_gpu
    datas = [_encode(320 + 16 * i, 200 + 8 * i, 20 + i, subsampling=2) for i in range(5)]
    devs = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    res = decode_files_multi_gpu(datas, devices=devs)
    for d, data in zip(res, datas):
        assert np.array_equal(d.image_tensor.cpu().numpy(), oracle.decode(data, want=("rgb",)).rgb)


def test_chain_step_without_the_stream_head_bitmap(monkeypatch):
    """Scans with more than 262144 subsequences do not fit the chain kernel's shared-memory bitmap of stream heads
    and look the heads up by binary search instead; BJ_PHASE_NO_BITMAP forces that path on a file with thousands of
    restart intervals (and on a progressive one), results unchanged."""
    from pyjpegdecoder_b200.pipeline import DevicePipeline
    monkeypatch.setattr(DevicePipeline, "EXTRA_PHASE_FLAGS", 8)
    _check([_encode(3840, 2160, 1, subsampling=2, restart_marker_blocks=16),
            _encode(1024, 768, 3, subsampling=2, progressive=True, restart_marker_rows=1),
            _encode(640, 480, 4, subsampling=1)])
