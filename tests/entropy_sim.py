"""Drives tests/hostsim/entropy_hostsim.cpp (host build of csrc/bj_entropy.cuh) over whole files.
TEST-ONLY helper: sequential emulation of the GPU entropy pipeline."""
import ctypes

import numpy as np

from hostsim import build
from pyjpegdecoder_b200.huffman import build_scan_blob
from pyjpegdecoder_b200.parser import parse_jpeg

MODES = {"baseline": 0, "dc_first": 1, "dc_refine": 2, "ac_first": 3, "ac_refine": 4}
MAX_SLOTS = 10


class SimScan(ctypes.Structure):
    _fields_ = [("lut", ctypes.c_void_p),
                ("dc_tab", ctypes.c_uint16 * MAX_SLOTS), ("ac_tab", ctypes.c_uint16 * MAX_SLOTS),
                ("slot_comp", ctypes.c_uint8 * MAX_SLOTS),
                ("nslots", ctypes.c_int), ("ss", ctypes.c_int), ("se", ctypes.c_int), ("al", ctypes.c_int),
                ("mode", ctypes.c_int)]


_L = None


def lib():
    global _L
    if _L is None:
        _L = build("entropy_hostsim")
        _L.hs_unstuff.restype = ctypes.c_uint64
        _L.hs_unstuff.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                  ctypes.c_void_p, ctypes.c_uint32]
        _L.hs_decode_stream.restype = ctypes.c_uint32
        _L.hs_decode_stream.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                        ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int]
        _L.hs_acrefine_stream.restype = ctypes.c_uint32
        _L.hs_acrefine_stream.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                          ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        _L.hs_dcrefine_stream.restype = None
        _L.hs_dcrefine_stream.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int,
                                          ctypes.c_void_p]
    return _L


def scan_block_list(p, sc):
    """[(component, by, bx)] in the scan's decoding order (jpeg_decoder.py:774-805, :1125-1126)."""
    out = []
    if len(sc.comps) > 1:
        for m in range(sc.mcus_x * sc.mcus_y):
            my, mx = divmod(m, sc.mcus_x)
            for ci in sc.comps:
                c = p.components[ci]
                for r in range(c.h * c.v):
                    out.append((ci, my * c.v + r // c.h, mx * c.h + r % c.h))
    else:
        ci = sc.comps[0]
        for m in range(sc.mcus_x * sc.mcus_y):
            my, mx = divmod(m, sc.mcus_x)
            out.append((ci, my, mx))
    return out


def decode_file(data: bytes, sub_bits: int = 1024, upto_scan=None, warm: int = 1):
    """Returns (parsed, per-component grids after each scan, stats per stream)."""
    L = lib()
    p = parse_jpeg(data)
    grids = [np.zeros((p.mcus_y * c.v, p.mcus_x * c.h, 64), np.int16) for c in p.components]
    per_scan = []
    all_stats = []
    for si, sc in enumerate(p.scans):
        if upto_scan is not None and si >= upto_scan:
            break
        blob, dc_off, ac_off = build_scan_blob(sc.dc_specs, sc.ac_specs)
        blob = np.ascontiguousarray(blob)
        s = SimScan()
        s.lut = blob.ctypes.data
        interleaved = len(sc.comps) > 1
        slot = 0
        for k, ci in enumerate(sc.comps):
            c = p.components[ci]
            for _ in range(c.h * c.v if interleaved else 1):
                s.dc_tab[slot], s.ac_tab[slot], s.slot_comp[slot] = dc_off[k], ac_off[k], k
                slot += 1
        s.nslots, s.ss, s.se, s.al, s.mode = slot, sc.ss, sc.se, sc.al, MODES[sc.kind]
        raw = np.frombuffer(data, np.uint8)[sc.data_start:sc.data_end].copy()
        n_mcu = sc.mcus_x * sc.mcus_y
        ri = sc.ri if sc.ri > 0 else n_mcu
        n_streams = -(-n_mcu // ri)
        words = np.zeros(len(raw) // 4 + 80, np.uint32)
        starts = np.full(n_streams, 2 ** 64 - 1, np.uint64)
        nbytes = L.hs_unstuff(raw.ctypes.data, len(raw), words.ctypes.data, 0, starts.ctypes.data, n_streams)
        assert (starts != 2 ** 64 - 1).all(), "missing restart markers"
        ends = np.append(starts[1:], np.uint64(nbytes))
        blocks = scan_block_list(p, sc)
        coef = np.stack([grids[ci][by, bx] for (ci, by, bx) in blocks]).astype(np.int16)
        coef = np.ascontiguousarray(coef)
        for m in range(n_streams):
            b_lo = m * ri * s.nslots
            b_hi = min((m + 1) * ri, n_mcu) * s.nslots
            sub = coef[b_lo:b_hi]
            stats = np.zeros(4, np.uint32)
            if sc.kind in ("baseline", "dc_first", "ac_first"):
                err = L.hs_decode_stream(words.ctypes.data, len(words), int(starts[m]), int(ends[m]),
                                         ctypes.byref(s), b_hi - b_lo, sub_bits, sub.ctypes.data, stats.ctypes.data, warm)
                all_stats.append(stats)
            elif sc.kind == "dc_refine":
                L.hs_dcrefine_stream(words.ctypes.data, int(starts[m]), b_hi - b_lo, sc.al, sub.ctypes.data)
                err = 0
            else:
                err = L.hs_acrefine_stream(words.ctypes.data, len(words), int(starts[m]), int(ends[m]),
                                           ctypes.byref(s), b_hi - b_lo, sub.ctypes.data)
            assert err == 0, (si, sc.kind, m, err)
        for (ci, by, bx), blk in zip(blocks, coef):
            grids[ci][by, bx] = blk
        per_scan.append([g.copy() for g in grids])
    return p, per_scan, all_stats
