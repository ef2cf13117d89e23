"""Coefficient-buffer layout helpers (host side, numpy).

Device layout (include/b200jpeg.h): per image, 128-byte blocks of 64 int16 in zig-zag order, stored
MCU-major over the padded MCU grid: block = mcu * blocks_per_mcu + slot0[c] + (by % v) * h + (bx % h),
mcu = (by // v) * mcus_x + bx // h.  The reference keeps quantised coefficients either nowhere
(baseline, jpeg_decoder.py:869) or packed in image_array[8*bx+u, 8*by+v, c] (progressive, :1029,
:1225); tests compare through per-component grids (BH, BW, 64) in zig-zag order.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .parser import ParsedJpeg

# zig-zag index -> (u, v) = (horizontal, vertical) frequency; same as the reference's zagzig (:1672-1681)
ZIGZAG_UV = [(0, 0), (1, 0), (0, 1), (0, 2), (1, 1), (2, 0), (3, 0), (2, 1), (1, 2), (0, 3), (0, 4), (1, 3),
             (2, 2), (3, 1), (4, 0), (5, 0), (4, 1), (3, 2), (2, 3), (1, 4), (0, 5), (0, 6), (1, 5), (2, 4),
             (3, 3), (4, 2), (5, 1), (6, 0), (7, 0), (6, 1), (5, 2), (4, 3), (3, 4), (2, 5), (1, 6), (0, 7),
             (1, 7), (2, 6), (3, 5), (4, 4), (5, 3), (6, 2), (7, 1), (7, 2), (6, 3), (5, 4), (4, 5), (3, 6),
             (2, 7), (3, 7), (4, 6), (5, 5), (6, 4), (7, 3), (7, 4), (6, 5), (5, 6), (4, 7), (5, 7), (6, 6),
             (7, 5), (7, 6), (6, 7), (7, 7)]


def slot0_of(p: ParsedJpeg) -> List[int]:
    out, s = [], 0
    for c in p.components:
        out.append(s)
        s += c.h * c.v
    return out


def total_blocks(p: ParsedJpeg) -> int:
    return p.mcus_x * p.mcus_y * p.blocks_per_mcu


def block_index_grid(p: ParsedJpeg, ci: int) -> np.ndarray:
    """(BH, BW) array: device block index (relative to the image) of every block of component ci."""
    c = p.components[ci]
    bw, bh = p.mcus_x * c.h, p.mcus_y * c.v
    by, bx = np.mgrid[0:bh, 0:bw]
    mcu = (by // c.v) * p.mcus_x + bx // c.h
    return mcu * p.blocks_per_mcu + slot0_of(p)[ci] + (by % c.v) * c.h + (bx % c.h)


def grids_to_device(p: ParsedJpeg, grids: List[np.ndarray]) -> np.ndarray:
    """Per-component (BH, BW, 64) grids -> (total_blocks, 64) int16 in device order."""
    buf = np.zeros((total_blocks(p), 64), np.int16)
    for ci, g in enumerate(grids):
        buf[block_index_grid(p, ci).ravel()] = g.reshape(-1, 64)
    return buf


def device_to_grids(p: ParsedJpeg, buf: np.ndarray) -> List[np.ndarray]:
    """(total_blocks, 64) device-order buffer -> per-component (BH, BW, 64) grids."""
    buf = np.asarray(buf).reshape(-1, 64)
    out = []
    for ci in range(p.ncomp):
        idx = block_index_grid(p, ci)
        out.append(buf[idx.ravel()].reshape(idx.shape[0], idx.shape[1], 64))
    return out


def samples_device_to_planes(p: ParsedJpeg, buf: np.ndarray) -> List[np.ndarray]:
    """Sample buffer (total_blocks, 64) [y][x] -> per-component (8*BH, 8*BW) int16 planes."""
    out = []
    for g in device_to_grids(p, buf):
        bh, bw, _ = g.shape
        out.append(g.reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(8 * bh, 8 * bw))
    return out
