"""Device Huffman tables: canonical code construction (jpeg_decoder.py:366-377) flattened into the
two-level lookup tables that bj_entropy.cuh reads (see the entry format there).

The reference keeps one dict {bit-string: symbol} per table (:366-377); the device wants a LUT
indexed by the next 9 bits, with 128-entry second-level tables for longer codes.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .errors import CorruptedJpeg
from .parser import HuffSpec

L1_BITS = 9
L2_BITS = 16 - L1_BITS
# not a code: length 0, but consumes 1 bit and advances 1 so that a speculating decoder makes progress
INVALID_ENTRY = 1 | (1 << 8)


def canonical_codes(spec: HuffSpec) -> List[Tuple[int, int, int]]:
    """[(code, length, symbol)] in the order of the DHT segment (:368-374)."""
    out = []
    code = 0
    k = 0
    for length in range(1, 17):
        code <<= 1
        for _ in range(spec.counts[length - 1]):
            if k >= len(spec.values):
                raise CorruptedJpeg("Failed to parse Huffman tables.")
            out.append((code, length, spec.values[k]))
            code += 1
            k += 1
    return out


@lru_cache(maxsize=4096)
def build_table(spec: HuffSpec, is_dc: bool) -> np.ndarray:
    """uint32 LUT for one table: 512 first-level entries + 128 per second-level table."""
    l1 = np.full(1 << L1_BITS, INVALID_ENTRY, np.uint32)
    subs: List[np.ndarray] = []
    sub_of_prefix: Dict[int, int] = {}
    for code, length, sym in canonical_codes(spec):
        if code >> length:
            continue  # over-subscribed table: the reference's dict would hold it, no bit pattern reaches it
        if is_dc:
            if sym > 16:
                continue  # not decodable as a DC size
            total, adv = length + sym, 1
        else:
            total = length + (sym & 15)
            adv = 64 if sym == 0 else (16 if sym == 0xF0 else (sym >> 4) + 1)
        if total > 31:
            # cannot happen for length <= 16 and size <= 15
            continue
        entry = total | (adv << 8) | (length << 16) | (sym << 24)
        if length <= L1_BITS:
            lo = code << (L1_BITS - length)
            l1[lo:lo + (1 << (L1_BITS - length))] = entry
        else:
            code16 = code << (16 - length)
            prefix = code16 >> L2_BITS
            if prefix not in sub_of_prefix:
                sub_of_prefix[prefix] = len(subs)
                subs.append(np.full(1 << L2_BITS, INVALID_ENTRY, np.uint32))
            sub = subs[sub_of_prefix[prefix]]
            lo = code16 & ((1 << L2_BITS) - 1)
            sub[lo:lo + (1 << (16 - length))] = entry
    for prefix, si in sub_of_prefix.items():
        off = (1 << L1_BITS) + si * (1 << L2_BITS)
        if off > 0xFFFF:
            raise CorruptedJpeg("Huffman table too irregular for the device tables.")
        l1[prefix] = 0x80 | (off << 8)
    t = np.concatenate([l1] + subs) if subs else l1
    t.setflags(write=False)
    return t


_EMPTY = np.full(1 << L1_BITS, INVALID_ENTRY, np.uint32)


def build_scan_blob(dc_specs: Sequence, ac_specs: Sequence) -> Tuple[np.ndarray, List[int], List[int]]:
    """LUT blob for one scan: distinct tables back to back.  Returns (blob, dc offsets per scan
    component, ac offsets per scan component).  A missing table (e.g. no AC table in a DC scan)
    maps to an all-invalid table, so decoding with it reports BJ_ERR_BAD_CODE like the reference's
    KeyError/CorruptedJpeg would."""
    parts: List[np.ndarray] = []
    where: Dict[tuple, int] = {}
    size = 0

    def place(spec, is_dc):
        nonlocal size
        key = (spec, is_dc)
        if key not in where:
            t = _EMPTY if spec is None else build_table(spec, is_dc)
            where[key] = size
            parts.append(t)
            size += len(t)
        return where[key]

    dc_off = [place(s, True) for s in dc_specs]
    ac_off = [place(s, False) for s in ac_specs]
    if size > 0xFFFF:
        raise CorruptedJpeg("Huffman tables too large for the device tables.")
    return np.concatenate(parts), dc_off, ac_off
