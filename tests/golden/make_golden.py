#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(tbpaolini/PyJpegDecoder, /root/reference/jpeg_decoder.py) in this container.

The reference cannot travel to the GPU box, so its outputs are committed here as small
fixtures together with this script (the only file that imports the reference).

Usage:  python tests/golden/make_golden.py [--big | --writer]     (needs /root/reference, Pillow, scipy)
        --writer: only the cases written by tests/jpeg_writer.py (sampling layouts Pillow cannot encode, int16 wrap)

What is recorded per case (tests/golden/cases/<name>.jpg + <name>.npz):
  rgb      uint8  (W,H,3) or (W,H)  -- JpegDecoder.image_array            (jpeg_decoder.py:1373-1386)
  canvas   int16  (AW,AH,nc)        -- image_array on entry to end_of_image (Y/Cb/Cr after IDCT+upsample)
  coef{c}  int16  (BH,BW,64)        -- quantised coefficients of component c, zig-zag order,
                                       over the padded block grid (after the last scan)
  scan{k}_coef{c}                   -- the same after progressive scan k (1-based), progressive only
Hooks follow SURVEY.md Appendix C: baseline coefficients are the arguments of undo_zigzag
(jpeg_decoder.py:869), progressive ones are image_array snapshots after every scan with
scan_amount forced high so the final IDCT stage (jpeg_decoder.py:1308) never fires early.
"""
import contextlib
import hashlib
import io
import json
import sys
from pathlib import Path

import numpy as np
from PIL import Image

REF_DIR = "/root/reference"
HERE = Path(__file__).resolve().parent
CASES = HERE / "cases"

sys.path.insert(0, REF_DIR)
import jpeg_decoder as jd  # noqa: E402  (the reference itself)

jd.JpegDecoder.show = lambda self: None  # reference opens a GUI at :1389

ZAGZIG = jd.zagzig


def synth(w, h, seed, channels=3, saturate=False):
    """Synthetic content of SURVEY.md section 8(d)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([128 + 100 * np.sin(x / 37 + y / 91),
                    128 + 100 * np.cos(x / 53 - y / 29),
                    128 + 100 * np.sin((x + y) / 71)], -1)
    img = img + rng.normal(0, 12, img.shape)
    if saturate:
        # hard saturated colour patches: exercises the "no clamp before colour" rule (:1698)
        img = np.zeros_like(img)
        pal = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [255, 0, 255],
                        [0, 255, 255], [255, 255, 255], [0, 0, 0]], np.float64)
        for by in range(0, h, 8):
            for bx in range(0, w, 8):
                img[by:by + 8, bx:bx + 8] = pal[rng.integers(0, len(pal))]
    img = np.clip(img, 0, 255).astype(np.uint8)
    if channels == 1:
        img = img[..., 0]
    return img


class Recorder:
    """Runs the reference on a file and records coefficient planes / canvas."""

    def __init__(self):
        self.blocks = []
        self.snaps = []
        self.canvas = None
        self.in_dqt = False
        self.cur = None

    def run(self, path, stop_after_scan=None):
        rec = self
        ouz = jd.undo_zigzag
        odq = jd.JpegDecoder.define_quantization_table
        oprog = jd.JpegDecoder.progressive_dct_scan
        oeoi = jd.JpegDecoder.end_of_image
        oidct = jd.InverseDCT

        def uz(b):
            if not rec.in_dqt:
                rec.blocks.append(np.array(b, dtype=np.int16))
            return ouz(b)

        def dq(self_, data):
            rec.in_dqt = True
            try:
                return odq(self_, data)
            finally:
                rec.in_dqt = False

        def prog(self_, *a, **k):
            rec.cur = self_
            n_before = len(rec.snaps)
            oprog(self_, *a, **k)
            if len(rec.snaps) == n_before:      # final stage did not fire inside this call
                rec.snaps.append(self_.image_array.copy())

        class SnapIDCT(oidct):
            # The final stage instantiates InverseDCT() (:1311) right after the last scan and
            # before touching image_array (:1317): snapshot the last scan's coefficients there.
            def __init__(self_i):
                if rec.cur is not None and rec.cur.scan_mode == "progressive_dct":
                    rec.snaps.append(rec.cur.image_array.copy())

        def eoi(self_, data):
            rec.canvas = self_.image_array.copy()
            return oeoi(self_, data)

        jd.undo_zigzag = uz
        jd.JpegDecoder.define_quantization_table = dq
        jd.JpegDecoder.progressive_dct_scan = prog
        jd.JpegDecoder.end_of_image = eoi
        jd.InverseDCT = SnapIDCT
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                dec = jd.JpegDecoder(Path(path))
        finally:
            jd.undo_zigzag = ouz
            jd.JpegDecoder.define_quantization_table = odq
            jd.JpegDecoder.progressive_dct_scan = oprog
            jd.JpegDecoder.end_of_image = oeoi
            jd.InverseDCT = oidct
        return dec


def comp_grids(dec):
    """Padded block grid (BW,BH) of each component in frame order."""
    comps = sorted(dec.color_components.values(), key=lambda c: c.order)
    out = []
    for c in comps:
        rh = dec.sample_shape[0] // c.shape[0]
        rv = dec.sample_shape[1] // c.shape[1]
        out.append((dec.array_width // rh // 8, dec.array_height // rv // 8))
    return comps, out


def planes_from_snapshot(dec, snap):
    comps, grids = comp_grids(dec)
    res = []
    zz = np.array(ZAGZIG)
    for c, (bw, bh) in zip(comps, grids):
        a = snap[:8 * bw, :8 * bh, c.order]  # [x, y]
        a = a.reshape(bw, 8, bh, 8)          # [bx, u, by, v]
        blk = a.transpose(2, 0, 1, 3)        # [by, bx, u, v]
        res.append(np.ascontiguousarray(blk[:, :, zz[:, 0], zz[:, 1]]).astype(np.int16))
    return res


def planes_from_blocks(dec, blocks):
    """Baseline: blocks in decode order -> per component grids. Single interleaved scan, or
    one scan per component (non-interleaved)."""
    comps, grids = comp_grids(dec)
    res = [np.zeros((bh, bw, 64), np.int16) for (bw, bh) in grids]
    nc = len(comps)
    if nc == 1:
        bw = -(-dec.image_width // 8)
        bh = -(-dec.image_height // 8)
        assert len(blocks) == bw * bh, (len(blocks), bw, bh)
        for k, b in enumerate(blocks):
            res[0][k // bw, k % bw] = b
        return res
    hmax = dec.sample_shape[0] // 8
    vmax = dec.sample_shape[1] // 8
    mx = -(-dec.image_width // (8 * hmax))
    my = -(-dec.image_height // (8 * vmax))
    per_mcu = sum(c.repeat for c in comps)
    assert len(blocks) == mx * my * per_mcu, (len(blocks), mx, my, per_mcu)
    k = 0
    for m in range(mx * my):
        mcy, mcx = divmod(m, mx)
        for ci, c in enumerate(comps):
            h, v = c.horizontal_sampling, c.vertical_sampling
            for r in range(c.repeat):
                by, bx = divmod(r, h)
                res[ci][mcy * v + by, mcx * h + bx] = blocks[k]
                k += 1
    return res


def record_case(name, jpeg_bytes, meta):
    path = CASES / f"{name}.jpg"
    path.write_bytes(jpeg_bytes)
    rec = Recorder()
    dec = rec.run(path)
    out = {"rgb": np.ascontiguousarray(dec.image_array), "canvas": rec.canvas}
    if dec.scan_mode == "baseline_dct":
        planes = planes_from_blocks(dec, rec.blocks)
    else:
        for k, snap in enumerate(rec.snaps, start=1):
            for ci, p in enumerate(planes_from_snapshot(dec, snap)):
                out[f"scan{k}_coef{ci}"] = p
        planes = planes_from_snapshot(dec, rec.snaps[-1])
    for ci, p in enumerate(planes):
        out[f"coef{ci}"] = p
    np.savez_compressed(CASES / f"{name}.npz", **out)
    meta[name] = {
        "width": int(dec.image_width), "height": int(dec.image_height),
        "mode": dec.scan_mode, "ncomp": len(dec.color_components),
        "scans": int(dec.scan_amount),
        "rgb_sha256": hashlib.sha256(np.ascontiguousarray(dec.image_array).tobytes()).hexdigest(),
    }
    print(name, meta[name]["mode"], dec.image_array.shape, flush=True)


def encode(img, **kw):
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def small_cases():
    cases = []
    # (name, w, h, seed, channels, save kwargs, saturate)
    sizes = [(1, 1), (8, 8), (17, 20), (33, 17), (64, 64), (70, 50), (75, 43), (97, 61), (120, 88)]
    i = 0
    for (w, h) in sizes:
        for ss in (0, 1, 2):
            i += 1
            cases.append((f"base_{w}x{h}_ss{ss}", w, h, i, 3, dict(quality=75, subsampling=ss), False))
    cases += [
        ("base_gray_70x50", 70, 50, 101, 1, dict(quality=75), False),
        ("base_gray_8x8", 8, 8, 102, 1, dict(quality=90), False),
        ("base_gray_33x17_dri2", 33, 17, 103, 1, dict(quality=75, restart_marker_blocks=2), False),
        ("base_70x50_ss2_dri1", 70, 50, 104, 3, dict(quality=75, subsampling=2, restart_marker_blocks=1), False),
        ("base_70x50_ss2_dri3", 70, 50, 105, 3, dict(quality=75, subsampling=2, restart_marker_blocks=3), False),
        ("base_97x61_ss1_dri7", 97, 61, 106, 3, dict(quality=75, subsampling=1, restart_marker_blocks=7), False),
        ("base_97x61_ss0_dri13", 97, 61, 107, 3, dict(quality=75, subsampling=0, restart_marker_blocks=13), False),
        ("base_120x88_ss2_row1", 120, 88, 108, 3, dict(quality=75, subsampling=2, restart_marker_rows=1), False),
        ("base_120x88_ss2_opt", 120, 88, 109, 3, dict(quality=75, subsampling=2, optimize=True), False),
        ("base_120x88_ss2_q30", 120, 88, 110, 3, dict(quality=30, subsampling=2), False),
        ("base_120x88_ss2_q95", 120, 88, 111, 3, dict(quality=95, subsampling=2), False),
        ("base_64x48_ss0_q100", 64, 48, 112, 3, dict(quality=100, subsampling=0), False),
        ("base_sat_96x64_ss2", 96, 64, 113, 3, dict(quality=90, subsampling=2), True),
        ("base_sat_96x64_ss0", 96, 64, 114, 3, dict(quality=90, subsampling=0), True),
        ("base_sat_80x48_ss1", 80, 48, 115, 3, dict(quality=75, subsampling=1), True),
        ("base_256x128_ss2_opt", 256, 128, 116, 3, dict(quality=75, subsampling=2, optimize=True), False),
    ]
    for (w, h) in [(1, 1), (8, 8), (17, 20), (33, 17), (70, 50), (97, 61), (120, 88)]:
        for ss in (0, 1, 2):
            i += 1
            cases.append((f"prog_{w}x{h}_ss{ss}", w, h, 200 + i, 3,
                          dict(quality=75, subsampling=ss, progressive=True), False))
    cases += [
        ("prog_gray_70x50", 70, 50, 301, 1, dict(quality=75, progressive=True), False),
        ("prog_gray_97x61_dri5", 97, 61, 302, 1, dict(quality=75, progressive=True, restart_marker_blocks=5), False),
        ("prog_70x50_ss2_dri2", 70, 50, 303, 3, dict(quality=75, subsampling=2, progressive=True, restart_marker_blocks=2), False),
        ("prog_97x61_ss1_dri5", 97, 61, 304, 3, dict(quality=75, subsampling=1, progressive=True, restart_marker_blocks=5), False),
        ("prog_120x88_ss0_dri30", 120, 88, 305, 3, dict(quality=75, subsampling=0, progressive=True, restart_marker_blocks=30), False),
        ("prog_120x88_ss2_row1", 120, 88, 306, 3, dict(quality=75, subsampling=2, progressive=True, restart_marker_rows=1), False),
        ("prog_120x88_ss2_q95", 120, 88, 307, 3, dict(quality=95, subsampling=2, progressive=True), False),
        ("prog_120x88_ss2_q30", 120, 88, 308, 3, dict(quality=30, subsampling=2, progressive=True), False),
        ("prog_sat_96x64_ss2", 96, 64, 309, 3, dict(quality=90, subsampling=2, progressive=True), True),
        ("prog_200x120_ss0", 200, 120, 310, 3, dict(quality=75, subsampling=0, progressive=True), False),
    ]
    return cases


def upsample_weights():
    """Unit-impulse weights of the reference's ResizeGrid (jpeg_decoder.py:1588-1626) for the three
    tile shapes the decoder uses; they pin scipy/Qhull's triangulation."""
    out = {}
    rz = jd.ResizeGrid()
    for (ow, oh, nw, nh) in [(8, 8, 16, 16), (8, 8, 16, 8), (8, 8, 8, 16), (16, 8, 16, 16), (8, 16, 16, 16)]:
        w = np.zeros((ow, oh, nw, nh), np.float64)
        for i in range(ow):
            for j in range(oh):
                blk = np.zeros((ow, oh), np.float64)
                blk[i, j] = 15.0
                # call griddata exactly as the reference does but keep the float result
                key = ((ow, oh), (nw, nh))
                rz(np.zeros((ow, oh), np.int16), (nw, nh))  # fills the caches
                new_xy = rz.mesh_cache[key]
                old_xy = rz.indices_cache[key[0]]
                w[i, j] = jd.griddata(old_xy, blk.ravel(), new_xy)
        wi = np.rint(w).astype(np.int16)
        assert np.abs(w - wi).max() < 1e-9, "weights are not integers /15"
        out[f"w_{ow}x{oh}_{nw}x{nh}"] = wi
    return out


def big_fixture(meta):
    """The reference's own example (progressive, 10 scans, DRI per scan)."""
    src = Path(REF_DIR) / "progressive scan example" / "base image.jpg"
    data = src.read_bytes()
    rec = Recorder()
    dec = rec.run(src)
    entry = {
        "file_sha256": hashlib.sha256(data).hexdigest(),
        "width": int(dec.image_width), "height": int(dec.image_height),
        "rgb_sha256": hashlib.sha256(np.ascontiguousarray(dec.image_array).tobytes()).hexdigest(),
        "scan_coef_sha256": [],
    }
    for snap in rec.snaps:
        planes = planes_from_snapshot(dec, snap)
        entry["scan_coef_sha256"].append([hashlib.sha256(p.tobytes()).hexdigest() for p in planes])
    # after-scan renders shipped with the reference (rows 0..2339 of the 2352-row canvas)
    for k, off in ((1, 0x2a740), (2, 0x7a53c)):
        png = np.array(Image.open(Path(REF_DIR) / "progressive scan example" / f"after scan 0{k}.png").convert("RGB"))
        rows = np.ascontiguousarray(png[:dec.image_height])
        entry[f"after_scan_{k}"] = {"truncate_at": off,
                                    "rgb_hw3_sha256": hashlib.sha256(rows.tobytes()).hexdigest()}
    meta["base_image"] = entry
    print("base_image", entry["rgb_sha256"], flush=True)


def writer_cases():
    """Files Pillow cannot encode, written by tests/jpeg_writer.py: 4:4:0, layouts whose chroma components have their
    own sampling factors (the generic pixel kernel), and quantised coefficients whose dequantisation product wraps
    the reference's int16 arithmetic (jpeg_decoder.py:869)."""
    sys.path.insert(0, str(HERE.parent))
    import jpeg_writer as jw
    qt = jw.std_qtables(75)
    out = []

    def from_image(name, w, h, seed, sampling, ri=0):
        img = synth(w, h, seed)
        comps = jw.image_to_components(img, sampling, qt)
        out.append((name, jw.write_baseline(w, h, comps, qt, restart_interval=ri)))

    from_image("w440_40x48", 40, 48, 401, [(1, 2), (1, 1), (1, 1)])
    from_image("w440_33x41_dri3", 33, 41, 402, [(1, 2), (1, 1), (1, 1)], ri=3)
    from_image("w440_96x80", 96, 80, 403, [(1, 2), (1, 1), (1, 1)])
    from_image("wgen_y22_cb21_cr11_48x48", 48, 48, 404, [(2, 2), (2, 1), (1, 1)])
    from_image("wgen_y22_cb12_cr11_50x37", 50, 37, 405, [(2, 2), (1, 2), (1, 1)])
    from_image("wgen_y21_cb11_cr21_64x24", 64, 24, 406, [(2, 1), (1, 1), (2, 1)])
    from_image("wgen_y22_cb22_cr11_32x32", 32, 32, 407, [(2, 2), (2, 2), (1, 1)])
    # int16 wrap of coefficient * Q: quantisation tables full of 255, DC and a few AC coefficients beyond 128
    rng = np.random.default_rng(408)
    q255 = {0: [255] * 64, 1: [255] * 64}
    comps = []
    for i in range(3):
        blocks = np.zeros((2, 2, 64), np.int32)
        blocks[..., 0] = rng.integers(-4, 5, (2, 2)) * 40 + np.array([[200, -150], [135, 129]])   # |DC * 255| > 32767
        blocks[..., 1] = rng.integers(-140, 141, (2, 2))
        blocks[..., 5] = rng.integers(-3, 4, (2, 2))
        comps.append({"h": 1, "v": 1, "tq": 0 if i == 0 else 1, "blocks": blocks})
    out.append(("wwrap_16x16_ss0", jw.write_baseline(16, 16, comps, q255)))
    return out


def main():
    CASES.mkdir(parents=True, exist_ok=True)
    meta_path = HERE / "golden.json"
    meta = json.loads(meta_path.read_text()) if meta_path.exists() else {}
    meta.setdefault("cases", {})
    meta["versions"] = {"numpy": np.__version__, "scipy": __import__("scipy").__version__,
                        "pillow": __import__("PIL").__version__}
    if "--big" in sys.argv:
        big_fixture(meta)
    elif "--writer" in sys.argv:
        for name, data in writer_cases():
            record_case(name, data, meta["cases"])
    else:
        for (name, w, h, seed, ch, kw, sat) in small_cases():
            img = synth(w, h, seed, ch, sat)
            record_case(name, encode(img, **kw), meta["cases"])
        np.savez_compressed(HERE / "upsample_weights.npz", **upsample_weights())
        np.save(HERE / "idct_table.npy", jd.InverseDCT.idct_table)
    meta_path.write_text(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
