"""Test-side baseline JPEG writer -- TEST INFRASTRUCTURE ONLY.

Pillow cannot produce every file the decode path has code for (4:4:0 and other unusual sampling factors, quantised
coefficients large enough to wrap the reference's int16 dequantisation product, jpeg_decoder.py:869).  This module
writes such files from scratch: baseline sequential DCT, one interleaved scan, arbitrary sampling factors per
component, Huffman tables = the Annex-K tables (taken from a Pillow-encoded file, so nothing is typed in by hand),
optional restart intervals.  Coefficients are either given directly (zig-zag order, already quantised) or computed
from an image with a float FDCT.  Used by tests/golden/make_golden.py to create fixtures that the unmodified reference
then decodes.
"""
from __future__ import annotations

import io
from typing import Dict, List, Sequence

import numpy as np

ZIGZAG_NAT = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
              28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
              54, 47, 55, 62, 63]   # zig-zag index -> natural index v*8+u


def _pillow_segments(quality: int = 75):
    """DHT and DQT payloads of a (non-optimised) Pillow file: {dest: (counts[16], values)} and {id: 64 zig-zag values}."""
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(np.zeros((16, 16, 3), np.uint8) + 100).save(b, "JPEG", quality=quality, subsampling=2)
    d = b.getvalue()
    huff, qt = {}, {}
    pos = 2
    while pos + 4 <= len(d):
        assert d[pos] == 0xFF
        m = d[pos + 1]
        size = (d[pos + 2] << 8) | d[pos + 3]
        seg = d[pos + 4:pos + 2 + size]
        if m == 0xC4:
            q = 0
            while q < len(seg):
                dest, counts = seg[q], list(seg[q + 1:q + 17])
                n = sum(counts)
                huff[dest] = (counts, list(seg[q + 17:q + 17 + n]))
                q += 17 + n
        elif m == 0xDB:
            q = 0
            while q < len(seg):
                qt[seg[q] & 15] = list(seg[q + 1:q + 65])
                q += 65
        elif m == 0xDA:
            break
        pos += 2 + size
    return huff, qt


def _codes(counts: Sequence[int], values: Sequence[int]) -> Dict[int, tuple]:
    """Canonical Huffman codes of a DHT table: {symbol: (code, length)} (T.81 Annex C)."""
    out, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            out[values[k]] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return out


class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, value: int, nbits: int) -> None:
        if nbits == 0:
            return
        self.acc = (self.acc << nbits) | (value & ((1 << nbits) - 1))
        self.n += nbits
        while self.n >= 8:
            byte = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(byte)
            if byte == 0xFF:
                self.out.append(0x00)     # byte stuffing
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self) -> None:
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)   # pad with ones


def _category(v: int) -> int:
    return int(abs(v)).bit_length()


def _value_bits(v: int, cat: int) -> int:
    return v if v >= 0 else v + (1 << cat) - 1


def _encode_block(bw: _BitWriter, blk: Sequence[int], pred: int, dc, ac) -> int:
    """Huffman-code one block of 64 quantised coefficients (zig-zag order).  Returns the new DC predictor."""
    diff = int(blk[0]) - pred
    cat = _category(diff)
    bw.put(*dc[cat])
    bw.put(_value_bits(diff, cat), cat)
    run = 0
    last = max((k for k in range(1, 64) if blk[k]), default=0)
    for k in range(1, last + 1):
        v = int(blk[k])
        if v == 0:
            run += 1
            continue
        while run > 15:
            bw.put(*ac[0xF0])
            run -= 16
        cat = _category(v)
        bw.put(*ac[(run << 4) | cat])
        bw.put(_value_bits(v, cat), cat)
        run = 0
    if last < 63:
        bw.put(*ac[0x00])
    return int(blk[0])


def write_baseline(width: int, height: int, comps: List[dict], qtables: Dict[int, Sequence[int]],
                   restart_interval: int = 0, quality_tables: int = 75) -> bytes:
    """Baseline JPEG with one interleaved scan.
    comps: per component {"h", "v", "tq", "blocks"}; blocks = int array [blocks_v][blocks_h][64] (zig-zag, quantised)
    over the padded MCU grid (mcus_y * v rows, mcus_x * h columns).  Luma-style tables (DC 0 / AC 0) for the first
    component, chroma-style (DC 1 / AC 1) for the others.  qtables: {id: 64 zig-zag values, each 1..255}."""
    huff, _ = _pillow_segments(quality_tables)
    hmax = max(c["h"] for c in comps)
    vmax = max(c["v"] for c in comps)
    mcus_x = -(-width // (8 * hmax))
    mcus_y = -(-height // (8 * vmax))
    out = bytearray(b"\xff\xd8")
    out += b"\xff\xe0" + (16).to_bytes(2, "big") + b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00"
    for tid, q in sorted(qtables.items()):
        assert len(q) == 64 and all(1 <= int(x) <= 255 for x in q)
        out += b"\xff\xdb" + (67).to_bytes(2, "big") + bytes([tid]) + bytes(int(x) for x in q)
    nc = len(comps)
    out += b"\xff\xc0" + (8 + 3 * nc).to_bytes(2, "big") + b"\x08" + height.to_bytes(2, "big") + width.to_bytes(2, "big") + bytes([nc])
    for i, c in enumerate(comps):
        out += bytes([i + 1, (c["h"] << 4) | c["v"], c["tq"]])
    for dest in (0x00, 0x10, 0x01, 0x11):
        counts, values = huff[dest]
        out += b"\xff\xc4" + (19 + len(values)).to_bytes(2, "big") + bytes([dest]) + bytes(counts) + bytes(values)
    if restart_interval:
        out += b"\xff\xdd" + (4).to_bytes(2, "big") + restart_interval.to_bytes(2, "big")
    out += b"\xff\xda" + (6 + 2 * nc).to_bytes(2, "big") + bytes([nc])
    for i in range(nc):
        out += bytes([i + 1, 0x00 if i == 0 else 0x11])
    out += b"\x00\x3f\x00"
    codes = {d: _codes(*huff[d]) for d in huff}
    bw = _BitWriter()
    pred = [0] * nc
    rst = 0
    n_mcu = mcus_x * mcus_y
    for m in range(n_mcu):
        my, mx = divmod(m, mcus_x)
        for i, c in enumerate(comps):
            dc, ac = (codes[0x00], codes[0x10]) if i == 0 else (codes[0x01], codes[0x11])
            for by in range(c["v"]):
                for bx in range(c["h"]):
                    pred[i] = _encode_block(bw, c["blocks"][my * c["v"] + by][mx * c["h"] + bx], pred[i], dc, ac)
        if restart_interval and (m + 1) % restart_interval == 0 and m + 1 < n_mcu:
            bw.flush()
            bw.out += bytes([0xFF, 0xD0 + (rst & 7)])
            rst += 1
            pred = [0] * nc
    bw.flush()
    out += bw.out + b"\xff\xd9"
    return bytes(out)


def blocks_from_plane(plane: np.ndarray, q: Sequence[int], blocks_v: int, blocks_h: int) -> np.ndarray:
    """Float FDCT + quantisation of a sample plane (values 0..255, edge-replicated to the padded grid):
    int32 [blocks_v][blocks_h][64] in zig-zag order."""
    from scipy.fft import dctn
    h, w = plane.shape
    pad = np.pad(plane.astype(np.float64), ((0, blocks_v * 8 - h), (0, blocks_h * 8 - w)), mode="edge") - 128.0
    blk = pad.reshape(blocks_v, 8, blocks_h, 8).transpose(0, 2, 1, 3)          # [by][bx][y][x]
    coef = dctn(blk, type=2, norm="ortho", axes=(2, 3))                            # [by][bx][v][u]
    nat = coef.reshape(blocks_v, blocks_h, 64)
    qn = np.zeros(64)
    qn[ZIGZAG_NAT] = np.asarray(q, dtype=np.float64)                               # natural order
    quant = np.rint(nat / qn).astype(np.int32)
    return quant[:, :, ZIGZAG_NAT]


def image_to_components(rgb: np.ndarray, sampling: Sequence[tuple], qtables: Dict[int, Sequence[int]]) -> List[dict]:
    """RGB image -> component dicts for write_baseline.  sampling: [(h, v)] per component (Y, Cb, Cr); chroma-style
    components are box-averaged down to their own resolution."""
    h_img, w_img = rgb.shape[:2]
    r, g, b = (rgb[..., k].astype(np.float64) for k in range(3))
    planes = [0.299 * r + 0.587 * g + 0.114 * b,
              128 - 0.168736 * r - 0.331264 * g + 0.5 * b,
              128 + 0.5 * r - 0.418688 * g - 0.081312 * b]
    hmax = max(s[0] for s in sampling)
    vmax = max(s[1] for s in sampling)
    mcus_x = -(-w_img // (8 * hmax))
    mcus_y = -(-h_img // (8 * vmax))
    comps = []
    for i, (hs, vs) in enumerate(sampling):
        fx, fy = hmax // hs, vmax // vs
        p = planes[i]
        p = np.pad(p, ((0, (-h_img) % fy), (0, (-w_img) % fx)), mode="edge")
        p = p.reshape(p.shape[0] // fy, fy, p.shape[1] // fx, fx).mean(axis=(1, 3))
        tq = 0 if i == 0 else 1
        comps.append({"h": hs, "v": vs, "tq": tq, "blocks": blocks_from_plane(np.clip(p, 0, 255), qtables[tq], mcus_y * vs, mcus_x * hs)})
    return comps


def std_qtables(quality: int = 75) -> Dict[int, List[int]]:
    return _pillow_segments(quality)[1]
