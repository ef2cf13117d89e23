"""Host-side marker and segment parsing (stays in Python, as in the reference).

Mirrors the control flow of JpegDecoder.__init__ (jpeg_decoder.py:78-110) and the segment handlers
start_of_frame (:112-247), define_huffman_table (:249-390), define_quantization_table (:392-472),
define_restart_interval (:474-503) and start_of_scan (:505-652) -- same acceptance rules, same
exception classes -- but instead of decoding a scan when its SOS is met, it records a scan
descriptor (byte range of the entropy-coded segment, component/table selectors, Ss/Se/Ah/Al, the
restart interval and Huffman tables in force) that the device pipeline consumes later.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .errors import CorruptedJpeg, NotJpeg, UnsupportedJpeg

# first marker that ends an entropy-coded segment: 0xFF followed by anything but 0x00 / RSTn
# (the reference's main loop skips FF00 and RSTn, :93)
_SEGMENT_END = re.compile(rb"\xff[^\x00\xd0-\xd7]")
_SOS = b"\xff\xda"
_DNL = b"\xff\xdc"

_HOST = None


def _host_lib():
    """The C helpers of libb200jpeg.so (memchr-based scans), or False when the library is not built
    (the pure-Python scan is used then: parsing never needs a GPU)."""
    global _HOST
    if _HOST is None:
        try:
            import ctypes
            from .build import LIB
            L = ctypes.CDLL(str(LIB))
            L.bj_host_find_marker.restype = ctypes.c_uint64
            L.bj_host_find_marker.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64]
            L.bj_host_count_sos.restype = ctypes.c_uint32
            L.bj_host_count_sos.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64]
            _HOST = L
        except (OSError, AttributeError):
            _HOST = False
    return _HOST


def find_segment_end(data: bytes, pos: int) -> int:
    """End of the entropy-coded segment that starts at pos: the first 0xFF not followed by 0x00 / RSTn."""
    L = _host_lib()
    if L:
        return int(L.bj_host_find_marker(data, len(data), pos))
    mt = _SEGMENT_END.search(data, pos)
    return mt.start() if mt else len(data)


def count_sos(data: bytes, pos: int) -> int:
    L = _host_lib()
    if L:
        return int(L.bj_host_count_sos(data, len(data), pos))
    return data.count(_SOS, pos)


@dataclass(frozen=True)
class HuffSpec:
    """A DHT table exactly as transmitted: 16 code-length counts + values (:305-324)."""
    counts: bytes
    values: bytes


@dataclass
class Component:
    id: int
    order: int          # index on the last axis of the reference's image_array (:218)
    h: int
    v: int
    tq: int


@dataclass
class Scan:
    comps: Tuple[int, ...]          # component orders, in SOS order
    td: Tuple[int, ...]
    ta: Tuple[int, ...]
    ss: int
    se: int
    ah: int
    al: int
    ri: int                         # restart interval in force (:474-478)
    data_start: int                 # entropy-coded segment [data_start, data_end) in the file
    data_end: int
    dc_specs: Tuple[Optional[HuffSpec], ...]
    ac_specs: Tuple[Optional[HuffSpec], ...]
    mcus_x: int = 0                 # scan MCU grid (:609-619)
    mcus_y: int = 0
    kind: str = "baseline"          # baseline | dc_first | dc_refine | ac_first | ac_refine


@dataclass
class ParsedJpeg:
    width: int = 0
    height: int = 0
    progressive: bool = False
    components: List[Component] = field(default_factory=list)
    hmax: int = 1
    vmax: int = 1
    mcus_x: int = 0                 # interleaved (padded) MCU grid
    mcus_y: int = 0
    scans: List[Scan] = field(default_factory=list)
    qtables: Dict[int, np.ndarray] = field(default_factory=dict)   # key = raw Pq/Tq byte, 64 int16 zig-zag
    huff_specs: Dict[int, HuffSpec] = field(default_factory=dict)  # key = raw Tc/Th byte (final state)
    restart_interval: int = 0
    scan_amount: int = 0
    finished: bool = False          # EOI seen (:1388)
    file_size: int = 0

    @property
    def ncomp(self) -> int:
        return len(self.components)

    @property
    def blocks_per_mcu(self) -> int:
        return sum(c.h * c.v for c in self.components)

    @property
    def canvas_size(self) -> Tuple[int, int]:
        """(array_width, array_height) of the reference's padded image_array (:627-632)."""
        return self.mcus_x * 8 * self.hmax, self.mcus_y * 8 * self.vmax


def _be16(b: bytes, i: int) -> int:
    return (b[i] << 8) | b[i + 1]


def parse_jpeg(data: bytes) -> ParsedJpeg:
    """Walk the markers of one JPEG file image and return its descriptors."""
    if not data.startswith(b"\xff\xd8\xff"):                       # (:39-40)
        raise NotJpeg("File is not a JPEG image.")
    n = len(data)
    p = ParsedJpeg(file_size=n)
    mode: Optional[str] = None
    by_id: Dict[int, Component] = {}
    huff: Dict[int, HuffSpec] = {}
    ri = 0
    pos = 2
    find = data.find
    while True:                                                     # (:78-110)
        pos = find(b"\xff", pos)
        if pos < 0 or pos + 1 >= n:
            break                                                   # ran off the file (:79-83)
        m = data[pos + 1]
        if m == 0xFF:                                               # fill byte before a marker (T.81 B.1.1.2): the
            pos += 1                                                # reference would misread it as a segment; skip it
            continue
        pos += 2
        if m == 0x00 or 0xD0 <= m <= 0xD7:                          # (:93)
            continue
        if m == 0xD9:                                               # EOI (:1368)
            p.finished = True
            break
        if pos + 2 > n:
            break
        size = _be16(data, pos) - 2
        pos += 2
        seg = data[pos:pos + max(size, 0)]
        if m == 0xC0 or m == 0xC2:                                  # start_of_frame (:112-247)
            mode = "progressive_dct" if m == 0xC2 else "baseline_dct"
            p.progressive = m == 0xC2
            if len(seg) < 1 or seg[0] != 8:
                raise UnsupportedJpeg("Unsupported color depth. Only 8-bit greyscale and 24-bit RGB are supported.")
            if len(seg) < 6:
                raise CorruptedJpeg("Failed to parse the start of frame.")
            p.height = _be16(seg, 1)
            p.width = _be16(seg, 3)
            if p.width == 0:
                raise CorruptedJpeg("Image width cannot be zero.")
            nc = seg[5]
            if nc not in (1, 3):
                if nc == 4:
                    raise UnsupportedJpeg("CMYK color space is not supported. Only RGB and greyscale are supported.")
                raise UnsupportedJpeg("Unsupported color space. Only RGB and greyscale are supported.")
            if len(seg) < 6 + 3 * nc:
                raise CorruptedJpeg("Failed to parse the start of frame.")
            p.components = []
            by_id = {}
            for i in range(nc):
                cid, hv, tq = seg[6 + 3 * i], seg[7 + 3 * i], seg[8 + 3 * i]
                c = Component(id=cid, order=i, h=hv >> 4, v=hv & 15, tq=tq)
                p.components.append(c)
                by_id[cid] = c
            pos += size
        elif m == 0xC4:                                             # define_huffman_table (:249-390)
            q, ln = 0, len(seg)
            while q < ln:
                dest = seg[q]
                counts = seg[q + 1:q + 17]
                total = sum(counts)
                q += 17
                vals = seg[q:q + total]
                q += total
                if q > ln or len(counts) < 16:
                    raise CorruptedJpeg("Failed to parse Huffman tables.")
                huff[dest] = HuffSpec(bytes(counts), bytes(vals))
            pos += size
        elif m == 0xDB:                                             # define_quantization_table (:392-472)
            q, ln = 0, len(seg)
            while q < ln:
                dest = seg[q]
                if dest >> 4 == 1:                                  # Pq = 1: 16-bit entries (the reference reads 64
                    vals = seg[q + 1:q + 129]                       # bytes whatever Pq says, :443-454, and files the
                    if len(vals) != 128:                            # table under the wrong id; here it is parsed)
                        raise CorruptedJpeg("Failed to parse quantization tables.")
                    t = np.frombuffer(vals, dtype=">u2").astype(np.int32)
                    if (t > 32767).any():
                        raise UnsupportedJpeg("Quantization table entries above 32767 are not supported.")
                    p.qtables[dest & 15] = t.astype(np.int16)
                    q += 129
                    continue
                vals = seg[q + 1:q + 65]
                if len(vals) != 64:
                    raise CorruptedJpeg("Failed to parse quantization tables.")
                p.qtables[dest] = np.frombuffer(vals, dtype=np.uint8).astype(np.int16)
                q += 65
            pos += size
        elif m == 0xDD:                                             # define_restart_interval (:474-478)
            ri = _be16(seg, 0) if len(seg) >= 2 else 0
            pos += 2
        elif m == 0xDA:                                             # start_of_scan (:505-652)
            if mode is None:
                raise UnsupportedJpeg("Encoding mode not supported. Only 'Baseline DCT' and 'Progressive DCT' are supported.")
            if len(seg) < 1:
                raise CorruptedJpeg("Failed to parse the start of scan.")
            ns = seg[0]
            need = 1 + 2 * ns + (3 if p.progressive else 0)
            if ns < 1 or len(seg) < need:
                raise CorruptedJpeg("Failed to parse the start of scan.")
            comps, td, ta = [], [], []
            for i in range(ns):
                cid, tb = seg[1 + 2 * i], seg[2 + 2 * i]
                if cid not in by_id:
                    raise CorruptedJpeg("Scan refers to a colour component that the frame does not define.")
                comps.append(by_id[cid].order)
                td.append(tb >> 4)
                ta.append(tb & 15)
            ss, se, ah, al = 0, 63, 0, 0
            if p.progressive:
                t = 1 + 2 * ns
                ss, se, ah, al = seg[t], seg[t + 1], seg[t + 2] >> 4, seg[t + 2] & 15
            pos += size
            if pos > n:
                # a declared SOS length that points past the end of the file would give the scan a negative byte
                # range (the reference dies with an IndexError in its bit reader, :654-695)
                raise CorruptedJpeg("Start of scan segment extends past the end of the file.")
            if p.height == 0:                                       # DNL lookup (:575-581)
                d = find(_DNL, pos)
                if d < 0 or d + 6 > n:
                    raise CorruptedJpeg("Image height cannot be zero.")
                p.height = _be16(data, d + 4)
            if not p.scans:
                p.scan_amount = count_sos(data, pos) + 1            # (:635-637)
                _set_geometry(p)
            sc = Scan(comps=tuple(comps), td=tuple(td), ta=tuple(ta), ss=ss, se=se, ah=ah, al=al, ri=ri,
                      data_start=pos, data_end=n,
                      dc_specs=tuple(huff.get(t_) for t_ in td),
                      ac_specs=tuple(huff.get(0x10 | t_) for t_ in ta))
            _classify_scan(p, sc)
            sc.data_end = find_segment_end(data, pos)
            if sc.data_end < sc.data_start:
                raise CorruptedJpeg("Entropy-coded segment has a negative length.")
            pos = sc.data_end
            p.scans.append(sc)
        else:
            pos += max(size, 0)                                     # unknown segment skipped by length (:106)
    p.huff_specs = dict(huff)
    p.restart_interval = ri
    if not p.scans:
        raise CorruptedJpeg("No scan found in the file.")
    return p


def _set_geometry(p: ParsedJpeg) -> None:
    """MCU geometry (:583-632).  A single-component frame decodes as 8x8 MCUs whatever its sampling
    factors say (:595-598, :612-619)."""
    if p.ncomp == 1:
        p.components[0].h = p.components[0].v = 1
    for c in p.components:
        if c.h < 1 or c.v < 1:
            raise CorruptedJpeg("Sampling factors cannot be zero.")
    p.hmax = max(c.h for c in p.components)
    p.vmax = max(c.v for c in p.components)
    if sum(c.h * c.v for c in p.components) > 10:
        raise CorruptedJpeg("More than 10 blocks per MCU.")
    kinds = set()
    for c in p.components:
        if p.hmax % c.h or p.vmax % c.v or p.hmax // c.h > 2 or p.vmax // c.v > 2 or c.h > 2 or c.v > 2:
            raise UnsupportedJpeg(
                "Unsupported chroma subsampling: only sampling ratios of 1 or 2 per axis are supported "
                "(4:4:4, 4:2:2, 4:4:0, 4:2:0 and greyscale).")
        k = (p.hmax // c.h, p.vmax // c.v)
        if k != (1, 1):
            kinds.add(k)
    if len(kinds) > 2:
        raise UnsupportedJpeg("Unsupported chroma subsampling: more than two different upsampling ratios.")
    p.mcus_x = -(-p.width // (8 * p.hmax))
    p.mcus_y = -(-p.height // (8 * p.vmax))


def _classify_scan(p: ParsedJpeg, sc: Scan) -> None:
    ns = len(sc.comps)
    if ns > 1:
        sc.mcus_x, sc.mcus_y = p.mcus_x, p.mcus_y                    # (:609-611)
    else:
        c = p.components[sc.comps[0]]
        rh, rv = p.hmax // c.h, p.vmax // c.v
        sc.mcus_x = -(-(-(-p.width // rh)) // 8)                     # ceil(ceil(W/rh)/8) == ceil((W/rh)/8) (:612-619)
        sc.mcus_y = -(-(-(-p.height // rv)) // 8)
    if not p.progressive:
        sc.kind = "baseline"
        if ns == 1 and p.ncomp > 1:
            c = p.components[sc.comps[0]]
            if c.h * c.v > 1 or c.h != p.hmax or c.v != p.vmax:
                # the reference mis-stores such scans (:784 vs :889-891): nothing to be compatible with
                raise UnsupportedJpeg("Non-interleaved baseline scans of subsampled images are not supported.")
        return
    if sc.ss == 0 and sc.se == 0:
        is_dc = True
    elif sc.ss > 0 and sc.se >= sc.ss:
        is_dc = False
    else:
        raise CorruptedJpeg("Progressive JPEG images cannot contain both DC and AC values in the same scan.")  # (:922)
    if sc.ah == 0:
        refining = False
    elif sc.ah - sc.al == 1:
        refining = True
    else:
        raise CorruptedJpeg("Progressive JPEG images cannot contain more than 1 bit for each value on a refining scan.")  # (:934)
    if not is_dc and ns > 1:
        raise CorruptedJpeg("An AC progressive scan can only have a single color component.")  # (:967)
    if sc.se > 63:
        raise CorruptedJpeg("Spectral selection ends past coefficient 63.")
    if is_dc and ns == 1 and p.ncomp > 1:
        c = p.components[sc.comps[0]]
        if c.h * c.v > 1:
            # the reference positions these blocks with the interleaved stride (:993-994) and runs out of
            # its array: nothing to be compatible with
            raise UnsupportedJpeg("Non-interleaved DC scans of a subsampled component are not supported.")
    sc.kind = ("dc_" if is_dc else "ac_") + ("refine" if refining else "first")
