"""GPU parity of the pixel stages (bj_pixels through the C ABI) against the oracle / golden fixtures.

Inputs are the ORACLE's quantised coefficient planes, so these tests pin kernel I/C independently of
the entropy kernels.  Bit-exact: samples, canvas and RGB."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN, golden_case_names

pytestmark = pytest.mark.gpu


def _setup(names):
    import torch
    from pyjpegdecoder_b200 import _native, stages
    from pyjpegdecoder_b200.layout import grids_to_device
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.plan import BatchGeometry
    datas = [(GOLDEN / "cases" / f"{n}.jpg").read_bytes() for n in names]
    parsed = [parse_jpeg(d) for d in datas]
    orc = [oracle.decode(d) for d in datas]
    geom = BatchGeometry(parsed)
    coefs = np.concatenate([grids_to_device(p, o.coef) for p, o in zip(parsed, orc)])
    dev = torch.device("cuda:0")
    dg = stages.DeviceGeometry(geom, dev)
    coef_t = torch.from_numpy(coefs).to(dev)
    return torch, _native, stages, parsed, orc, geom, dg, coef_t


def test_pixels_all_golden_cases_one_batch():
    names = golden_case_names()
    torch, nat, stages, parsed, orc, geom, dg, coef_t = _setup(names)
    stats = torch.zeros(4, dtype=torch.int32, device=dg.device)
    rgb = stages.run_pixels(dg, coef_t, nat.IN_COEF, nat.OUT_RGB, stats=stats)
    canvas = stages.run_pixels(dg, coef_t, nat.IN_COEF, nat.OUT_CANVAS)
    samples = stages.run_pixels(dg, coef_t, nat.IN_COEF, nat.OUT_SAMPLES)
    rgb2 = stages.run_pixels(dg, samples, nat.IN_SAMPLES, nat.OUT_RGB)
    rgb3 = stages.run_pixels(dg, coef_t, nat.IN_COEF, nat.OUT_RGB, force_generic=True)   # generic kernel only
    torch.cuda.synchronize()
    for a, b in zip(stages.image_views(geom, rgb), stages.image_views(geom, rgb3)):
        assert torch.equal(a, b), "layout-specialised and generic pixel kernels disagree"
    rgb_v = stages.image_views(geom, rgb)
    rgb2_v = stages.image_views(geom, rgb2)
    can_v = stages.image_views(geom, canvas)
    bad = []
    for i, name in enumerate(names):
        z = np.load(GOLDEN / "cases" / f"{name}.npz")
        got = rgb_v[i].cpu().numpy()
        want = np.swapaxes(z["rgb"], 0, 1)
        p = parsed[i]
        want_canvas = np.swapaxes(z["canvas"], 0, 1)[:p.height, :p.width]
        got_canvas = can_v[i].cpu().numpy().reshape(p.height, p.width, -1)
        ok = (np.array_equal(got, want) and np.array_equal(got_canvas, want_canvas)
              and np.array_equal(rgb2_v[i].cpu().numpy(), want))
        if not ok:
            bad.append((name, int(np.abs(got.astype(int) - want.astype(int)).max()),
                        int(np.abs(got_canvas.astype(int) - want_canvas.astype(int)).max())))
    assert not bad, bad
    assert int(stats[0]) > 0  # the exact path was exercised (saturated / flat fixtures contain ties)


def test_pixels_samples_match_oracle_canvas_luma():
    names = ["base_120x88_ss2", "base_97x61_ss0", "base_gray_70x50"]
    torch, nat, stages, parsed, orc, geom, dg, coef_t = _setup(names)
    from pyjpegdecoder_b200.layout import samples_device_to_planes
    samples = stages.run_pixels(dg, coef_t, nat.IN_COEF, nat.OUT_SAMPLES).cpu().numpy()
    for i, p in enumerate(parsed):
        b0 = geom.block_offsets[i]
        nb = p.mcus_x * p.mcus_y * p.blocks_per_mcu
        planes = samples_device_to_planes(p, samples[b0:b0 + nb])
        # luma is never upsampled: the oracle canvas holds exactly these samples
        assert np.array_equal(planes[0], orc[i].canvas[:, :, 0])


def test_pixels_base_image_full_size(golden_meta):
    """4160x2340 4:2:0 (the reference's own example): oracle coefficients -> RGB must hash to the
    reference's output."""
    import hashlib
    import torch
    from pyjpegdecoder_b200 import _native as nat, stages
    from pyjpegdecoder_b200.layout import grids_to_device
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.plan import BatchGeometry
    data = (GOLDEN / "base_image.jpg").read_bytes()
    p = parse_jpeg(data)
    o = oracle.decode(data, want=("coef",))
    geom = BatchGeometry([p])
    dg = stages.DeviceGeometry(geom, torch.device("cuda:0"))
    coef_t = torch.from_numpy(grids_to_device(p, o.coef)).to(dg.device)
    rgb = stages.run_pixels(dg, coef_t, nat.IN_COEF, nat.OUT_RGB)
    img = stages.image_views(geom, rgb)[0].cpu().numpy()
    ref_order = np.ascontiguousarray(np.swapaxes(img, 0, 1))
    assert hashlib.sha256(ref_order.tobytes()).hexdigest() == golden_meta["base_image"]["rgb_sha256"]
