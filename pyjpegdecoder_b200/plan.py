"""Batch geometry for the device pipeline: bj_image records, quantisation-table buffer, buffer sizes."""
from __future__ import annotations

from math import cos, pi
from typing import List, Sequence

import numpy as np

from . import _native
from .errors import CorruptedJpeg
from .layout import slot0_of, total_blocks
from .parser import ParsedJpeg

_IDCT_TABLE_T = None


def idct_table_t() -> np.ndarray:
    """The reference's InverseDCT.idct_table (jpeg_decoder.py:1541-1553), evaluated with the very same
    Python expression (so the doubles are identical), transposed to [u][v][x][y] for the device."""
    global _IDCT_TABLE_T
    if _IDCT_TABLE_T is None:
        t = np.zeros((8, 8, 8, 8), dtype=np.float64)
        for x in range(8):
            for y in range(8):
                for u in range(8):
                    for v in range(8):
                        Cu = 2 ** (-0.5) if u == 0 else 1.0
                        Cv = 2 ** (-0.5) if v == 0 else 1.0
                        t[x, y, u, v] = 0.25 * Cu * Cv * cos((2 * x + 1) * pi * u / 16) * cos((2 * y + 1) * pi * v / 16)
        _IDCT_TABLE_T = np.ascontiguousarray(t.transpose(2, 3, 0, 1))
    return _IDCT_TABLE_T


def scan_levels(p: ParsedJpeg, serial: bool = False) -> List[int]:
    """Dependency level of every scan of an image: scans of the same level touch disjoint coefficients and can
    run in the same wave of launches.  A scan depends on every earlier scan that shares a component with it and
    whose spectral band overlaps its own (DC scans: band {0}); refinement scans additionally wait for every
    earlier scan of their components whatever the band (their parse looks at whole blocks).
    A progressive file from libjpeg's default script goes from 10 waves to 5: DC | 4 x AC first | Y refine |
    DC refine + Cr, Cb, Y refine.  serial=True: one level per scan (per-scan parity tests)."""
    if serial:
        return list(range(len(p.scans)))
    levels: List[int] = []
    for k, sc in enumerate(p.scans):
        lv = 0
        for j in range(k):
            sj = p.scans[j]
            if not set(sc.comps) & set(sj.comps):
                continue
            overlap = not (sc.se < sj.ss or sj.se < sc.ss)
            both_ac = sc.ss > 0 and sj.ss > 0
            if overlap or (both_ac and (sc.kind == "ac_refine" or sj.kind == "ac_refine")):
                lv = max(lv, levels[j] + 1)
        levels.append(lv)
    return levels


def layout_of(p: ParsedJpeg) -> int:
    """BJ_LAYOUT_* code: which specialised pixel kernel handles the image (0 = generic)."""
    if p.ncomp == 1:
        return _native.LAYOUT_GRAY
    c0, c1, c2 = p.components
    if (c1.h, c1.v, c2.h, c2.v) != (1, 1, 1, 1):
        return _native.LAYOUT_GENERIC
    return {(2, 2): _native.LAYOUT_420, (2, 1): _native.LAYOUT_422, (1, 2): _native.LAYOUT_440,
            (1, 1): _native.LAYOUT_444}.get((c0.h, c0.v), _native.LAYOUT_GENERIC)


# MCUs per CTA of the layout-specialised kernels: 4:2:0 -> 32 (csrc/bj_pixels_mma.cu), the others 6 warps x
# (32 // blocks_per_mcu) MCUs (Lay<...>::STRIP in csrc/bj_pixels_fast.cu); checked against bj_pixels_fast_strip() at load
FAST_STRIP = {_native.LAYOUT_420: 32, _native.LAYOUT_422: 48, _native.LAYOUT_440: 48, _native.LAYOUT_444: 60,
              _native.LAYOUT_GRAY: 192}


def choose_strip(mcus_x: int, blocks_per_mcu: int, layout: int = 0) -> int:
    """MCUs per CTA of the pixel kernels: fixed by the specialised kernel for its layouts, otherwise as
    equal as possible with at most 192 blocks."""
    if layout in FAST_STRIP:
        return FAST_STRIP[layout]
    max_m = max(1, _native.PIXEL_MAX_BLOCKS // blocks_per_mcu)
    n_strips = -(-mcus_x // max_m)
    return -(-mcus_x // n_strips)


class BatchGeometry:
    """Geometry of a batch of parsed images laid out back to back in the device buffers."""

    def __init__(self, parsed: Sequence[ParsedJpeg], channels_last_pitch_align: int = 1):
        n = len(parsed)
        self.parsed = list(parsed)
        self.images = np.zeros(n, dtype=_native.IMAGE_DTYPE)
        qt_rows: List[np.ndarray] = []
        qt_index = {}
        blk = 0
        out = 0
        self.out_offsets = []
        self.out_shapes = []
        self.block_offsets = []
        max_strips = 0
        self.layout_mask = 0
        for i, p in enumerate(parsed):
            rec = self.images[i]
            ch = 3 if p.ncomp == 3 else 1
            pitch = p.width * ch
            # keep every image start 16-byte aligned so the store path can use 128-bit stores
            out = (out + 15) & ~15
            rec["coef_block0"] = blk
            rec["out_offset"] = out
            rec["out_pitch"] = pitch
            rec["width"], rec["height"] = p.width, p.height
            rec["mcus_x"], rec["mcus_y"] = p.mcus_x, p.mcus_y
            rec["ncomp"] = p.ncomp
            rec["hmax"], rec["vmax"] = p.hmax, p.vmax
            rec["blocks_per_mcu"] = p.blocks_per_mcu
            s0 = slot0_of(p)
            for c in p.components:
                rec["hs"][c.order], rec["vs"][c.order] = c.h, c.v
                rec["slot0"][c.order] = s0[c.order]
                if c.tq not in p.qtables:
                    raise CorruptedJpeg("Component refers to a quantization table that the file does not define.")
                key = p.qtables[c.tq].tobytes()
                if key not in qt_index:
                    qt_index[key] = len(qt_rows)
                    qt_rows.append(p.qtables[c.tq])
                rec["qtab"][c.order] = qt_index[key]
            rec["layout"] = layout_of(p)
            self.layout_mask |= 1 << int(rec["layout"])
            strip = choose_strip(p.mcus_x, p.blocks_per_mcu, int(rec["layout"]))
            rec["strip_mcus"] = strip
            rec["strips_per_row"] = -(-p.mcus_x // strip)
            max_strips = max(max_strips, int(rec["strips_per_row"]) * p.mcus_y)
            self.block_offsets.append(blk)
            self.out_offsets.append(out)
            self.out_shapes.append((p.height, p.width, 3) if ch == 3 else (p.height, p.width))
            blk += total_blocks(p)
            out += p.height * pitch
        self.total_blocks = blk
        self.out_bytes = out
        self.max_strips = max_strips
        self.qtabs = np.ascontiguousarray(np.stack(qt_rows).astype(np.int16))
