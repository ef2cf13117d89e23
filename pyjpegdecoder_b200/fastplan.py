"""Batch planning at GPU speed: the host side of `decode_batch` for large batches.

parser.py + pipeline.BatchPlan interpret every file one by one in Python (~0.3 ms per file), an order
of magnitude slower than the GPU decodes them.  This module produces the SAME plan (identical arrays,
checked in tests/test_fastplan.py) much faster:

  1. the byte-level marker walk of every file (where are the segments, where does each entropy-coded
     run end -- the only part that touches all the bytes) runs in C on several host threads
     (`bj_host_walk_batch`, csrc/bj_host.cu, mirroring jpeg_decoder.py:78-110);
  2. the segment payloads that determine the parse (SOF, DHT, DQT, DRI, SOS, DNL) form a key; files
     with the same key (same encoder settings and size: the normal case inside a batch) share one
     template that parser.py computes once -- the Python parser stays the single source of truth;
  3. the per-batch arrays (struct bj_scan / bj_image records, offsets, tiles, groups) are assembled
     with vectorised numpy from the templates and the per-file run offsets.
"""
from __future__ import annotations

import copy
import ctypes
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _native
from .errors import JpegError, NotJpeg
from .huffman import build_scan_blob
from .layout import slot0_of, total_blocks
from .parser import ParsedJpeg, parse_jpeg
from .pipeline import ENTROPY_THREADS, MAX_SLOTS, MODES, SCAN_DTYPE, SUBSEQ_BITS, UNSTUFF_TILE, ScanGroup
from .plan import choose_strip, layout_of, scan_levels

ENTRY_DTYPE = np.dtype([("start", "<u8"), ("end", "<u8"), ("marker", "<u4"), ("reserved", "<u4")])
MAX_ENTRIES = 256
ENTROPY_RUN = 0x100
# segments whose payload changes the parse (everything else -- APPn, COM, ... -- is skipped by length)
_KEY_MARKERS = frozenset(list(range(0xC0, 0xD0)) + [0xDB, 0xDD, 0xDA, 0xDC, 0xD9])

_BOUND = False


def _lib():
    global _BOUND
    L = _native.lib()
    if not _BOUND:
        L.bj_host_walk_batch.restype = None
        L.bj_host_walk_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.bj_host_walk_batch_keys.restype = None
        L.bj_host_walk_batch_keys.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _BOUND = True
    return L


def walk_batch(raw: np.ndarray, offsets: np.ndarray, sizes: np.ndarray, threads: Optional[int] = None):
    """Marker walk of every file of the packed buffer.  Returns (entries[n, MAX_ENTRIES], counts[n],
    key_hash[n, 2]): the 128-bit hash of each file's parse-relevant segments (csrc/bj_host.cu)."""
    n = len(offsets)
    entries = np.empty((n, MAX_ENTRIES), dtype=ENTRY_DTYPE)
    counts = np.empty(n, dtype=np.int32)
    hashes = np.empty((n, 2), dtype=np.uint64)
    if threads is None:
        from .pipeline import host_threads
        threads = host_threads()
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
    _lib().bj_host_walk_batch_keys(raw.ctypes.data, offsets.ctypes.data, sizes.ctypes.data, n, entries.ctypes.data,
                                   MAX_ENTRIES, counts.ctypes.data, hashes.ctypes.data, threads)
    return entries, counts, hashes


class _Template:
    """Everything about a file that depends only on its parse-relevant segments."""

    def __init__(self, p: ParsedJpeg):
        self.parsed = p
        self.nscan = len(p.scans)
        self.total_blocks = total_blocks(p)
        self.ch = 3 if p.ncomp == 3 else 1
        self.pitch = p.width * self.ch
        self.out_size = p.height * self.pitch
        self.out_shape = (p.height, p.width, 3) if self.ch == 3 else (p.height, p.width)
        rec = np.zeros(1, dtype=_native.IMAGE_DTYPE)[0]
        rec["out_pitch"] = self.pitch
        rec["width"], rec["height"] = p.width, p.height
        rec["mcus_x"], rec["mcus_y"] = p.mcus_x, p.mcus_y
        rec["ncomp"] = p.ncomp
        rec["hmax"], rec["vmax"] = p.hmax, p.vmax
        rec["blocks_per_mcu"] = p.blocks_per_mcu
        s0 = slot0_of(p)
        self.qtabs = []
        for c in p.components:
            rec["hs"][c.order], rec["vs"][c.order] = c.h, c.v
            rec["slot0"][c.order] = s0[c.order]
            if c.tq not in p.qtables:
                from .errors import CorruptedJpeg
                raise CorruptedJpeg("Component refers to a quantization table that the file does not define.")
            self.qtabs.append(p.qtables[c.tq])
        rec["layout"] = layout_of(p)
        strip = choose_strip(p.mcus_x, p.blocks_per_mcu, int(rec["layout"]))
        rec["strip_mcus"] = strip
        rec["strips_per_row"] = -(-p.mcus_x // strip)
        self.image_rec = rec
        self.strips = int(rec["strips_per_row"]) * p.mcus_y
        # static part of every scan record + its LUT blob
        self.scan_recs = np.zeros(self.nscan, dtype=SCAN_DTYPE)
        self.blobs = []
        self.modes = np.zeros(self.nscan, dtype=np.int64)
        self.levels = np.asarray(scan_levels(p), dtype=np.int64)
        self.data_start = np.zeros(self.nscan, dtype=np.int64)
        for k, sc in enumerate(p.scans):
            r = self.scan_recs[k]
            n_mcu = sc.mcus_x * sc.mcus_y
            ri = sc.ri if sc.ri > 0 else n_mcu
            blob, dc_off, ac_off = build_scan_blob(sc.dc_specs, sc.ac_specs)
            self.blobs.append(((sc.dc_specs, sc.ac_specs), blob))
            slot = 0
            for kk, ci in enumerate(sc.comps):
                c = p.components[ci]
                nb = c.h * c.v if len(sc.comps) > 1 else 1
                for rr in range(nb):
                    r["slot_frame"][slot] = s0[ci] + rr
                    r["slot_comp"][slot] = kk
                    r["slot_dc"][slot] = dc_off[kk]
                    r["slot_ac"][slot] = ac_off[kk]
                    slot += 1
            if slot > MAX_SLOTS:
                from .errors import CorruptedJpeg
                raise CorruptedJpeg("More than 10 blocks per MCU.")
            c0 = p.components[sc.comps[0]]
            r["n_streams"], r["ri"], r["n_mcu"], r["mcus_x"] = -(-n_mcu // ri), ri, n_mcu, sc.mcus_x
            r["lut_len"] = len(blob)
            r["frame_mcus_x"], r["frame_bpm"] = p.mcus_x, p.blocks_per_mcu
            r["nslots"], r["mode"] = slot, MODES[sc.kind]
            r["ss"], r["se"], r["ah"], r["al"] = sc.ss, sc.se, sc.ah, sc.al
            r["interleaved"] = 1 if (len(sc.comps) > 1 or p.ncomp == 1) else 0
            r["comp_h"], r["comp_v"], r["comp_slot0"] = c0.h, c0.v, s0[sc.comps[0]]
            r["ncomp_scan"] = len(sc.comps)
            self.modes[k] = MODES[sc.kind]
            self.data_start[k] = sc.data_start


class LazyParsed:
    """Sequence of ParsedJpeg: template copies with the file's own entropy-run offsets, built on demand."""

    def __init__(self, templates, tid, run_start, run_end, run_base, sizes):
        self._t, self._tid = templates, tid
        self._rs, self._re, self._rb, self._sz = run_start, run_end, run_base, sizes

    def __len__(self):
        return len(self._tid)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        t = self._t[int(self._tid[i])]
        p = copy.copy(t.parsed)
        p.file_size = int(self._sz[i])
        p.scans = []
        b = int(self._rb[i])
        for k, sc in enumerate(t.parsed.scans):
            s2 = copy.copy(sc)
            s2.data_start, s2.data_end = int(self._rs[b + k]), int(self._re[b + k])
            p.scans.append(s2)
        return p

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class FastGeometry:
    """Same attributes as plan.BatchGeometry."""


class FastPlan:
    """Same attributes as pipeline.BatchPlan, assembled with numpy from per-key templates."""

    _cache: Dict[bytes, object] = {}

    def __init__(self, raw: np.ndarray, offsets: Sequence[int], sizes: Sequence[int], threads: Optional[int] = None,
                 walked=None):
        """walked: (entries, counts, key_hash) if the marker walk was already done while the files were packed
        (pipeline.pack_files(walk=True))."""
        n = len(offsets)
        offsets = np.asarray(offsets, dtype=np.int64)
        sizes = np.asarray(sizes, dtype=np.int64)
        self.raw_bytes = int(raw.size)
        entries, counts, hashes = walked if walked is not None else walk_batch(raw, offsets, sizes, threads)
        if (counts == -1).any():
            raise NotJpeg(f"File {int(np.nonzero(counts == -1)[0][0])}: File is not a JPEG image.")
        if (counts < 0).any():
            raise _Fallback("marker walk overflow")
        # ---- templates: one per distinct key hash (no per-file Python work) -------------------------------
        raw_bytes_view = memoryview(raw)
        cmax = int(counts.max()) if n else 0
        valid = np.arange(cmax, dtype=np.int32)[None, :] < counts[:, None]
        is_run = (entries["marker"][:, :cmax] == ENTROPY_RUN) & valid
        nrun = is_run.sum(axis=1).astype(np.int64)
        run_start_l = entries["start"][:, :cmax][is_run]     # row-major: per file, in file order
        run_end_l = entries["end"][:, :cmax][is_run]
        hv = np.ascontiguousarray(hashes).view(np.dtype([("a", "<u8"), ("b", "<u8")])).reshape(n)
        uniq, first_idx, inv = np.unique(hv, return_index=True, return_inverse=True)
        templates: List[_Template] = []
        cache = FastPlan._cache
        offs_l, sizes_l = offsets.tolist(), sizes.tolist()
        for u, i0 in zip(uniq.tolist(), first_idx.tolist()):
            tpl = cache.get(u)
            if tpl is None:
                data = raw_bytes_view[offs_l[i0]:offs_l[i0] + sizes_l[i0]].tobytes()
                try:
                    tpl = _Template(parse_jpeg(data))
                except JpegError as e:
                    tpl = (type(e), str(e))          # cache class + message, not the instance (no traceback growth)
                if len(cache) > 4096:
                    cache.clear()
                cache[u] = tpl
            if isinstance(tpl, tuple):
                raise tpl[0](f"File {i0}: {tpl[1]}")
            templates.append(tpl)
        tid = inv.reshape(n).astype(np.int64)
        if (np.array([t.nscan for t in templates], dtype=np.int64)[tid] != nrun).any():
            raise _Fallback("scan count differs from the template")
        run_start = np.asarray(run_start_l, dtype=np.int64)
        run_end = np.asarray(run_end_l, dtype=np.int64)
        run_base = np.concatenate(([0], np.cumsum(nrun)[:-1])) if n else np.zeros(0, np.int64)
        self.parsed = LazyParsed(templates, tid, run_start, run_end, run_base, sizes)
        self.any_progressive = any(t.parsed.progressive for t in templates)
        from .pipeline import covers_all_components
        self.needs_zero = self.any_progressive or any(not covers_all_components(t.parsed) for t in templates)
        # ---- geometry ----------------------------------------------------------------------------------
        g = FastGeometry()
        g.parsed = self.parsed
        t_img = np.zeros(len(templates), dtype=_native.IMAGE_DTYPE)
        for k_, t in enumerate(templates):
            t_img[k_] = t.image_rec
        g.images = t_img[tid].copy()
        t_blocks = np.array([t.total_blocks for t in templates], dtype=np.int64)
        t_out = np.array([t.out_size for t in templates], dtype=np.int64)
        blocks = t_blocks[tid]
        block0 = np.concatenate(([0], np.cumsum(blocks)[:-1]))
        out_sz = t_out[tid]
        out_al = (out_sz + 15) & ~15
        out0 = np.concatenate(([0], np.cumsum(out_al)[:-1]))
        g.images["coef_block0"] = block0
        g.images["out_offset"] = out0
        qt_rows: List[np.ndarray] = []
        qt_index: Dict[bytes, int] = {}
        t_qidx = np.zeros((len(templates), 3), dtype=np.uint32)
        # same de-duplication order as BatchGeometry: first use by image order, so walk images once per template
        first_use = {}
        for i in range(n):
            k = int(tid[i])
            if k not in first_use:
                first_use[k] = i
                for ci, q in enumerate(templates[k].qtabs):
                    kb = q.tobytes()
                    if kb not in qt_index:
                        qt_index[kb] = len(qt_rows)
                        qt_rows.append(q)
                    t_qidx[k, ci] = qt_index[kb]
                if len(first_use) == len(templates):
                    break
        g.images["qtab"] = t_qidx[tid]
        g.qtabs = np.ascontiguousarray(np.stack(qt_rows).astype(np.int16))
        g.block_offsets = block0.tolist()
        g.out_offsets = out0.tolist()
        g.out_shapes = [templates[int(k)].out_shape for k in tid]
        g.total_blocks = int(blocks.sum())
        g.out_bytes = int(out0[-1] + out_sz[-1]) if n else 0
        g.max_strips = int(max(t.strips for t in templates))
        g.layout_mask = 0
        for t in templates:
            g.layout_mask |= 1 << int(t.image_rec["layout"])
        self.geom = g
        # ---- scans -----------------------------------------------------------------------------------------
        t_nscan = np.array([t.nscan for t in templates], dtype=np.int64)
        t_sbase = np.concatenate(([0], np.cumsum(t_nscan)[:-1]))
        t_recs = np.zeros(int(t_nscan.sum()), dtype=SCAN_DTYPE)   # (np.concatenate would re-pack the padded dtype)
        for t, b0_ in zip(templates, t_sbase):
            t_recs[int(b0_):int(b0_) + t.nscan] = t.scan_recs
        t_modes = np.concatenate([t.modes for t in templates])
        t_levels = np.concatenate([t.levels for t in templates])
        nsc = t_nscan[tid]
        img = np.repeat(np.arange(n, dtype=np.int64), nsc)
        kidx = np.arange(len(img), dtype=np.int64) - np.repeat(run_base, nsc)
        flat_t = t_sbase[tid][img] + kidx
        mode = t_modes[flat_t]
        level = t_levels[flat_t]
        order = np.lexsort((kidx, img, mode, level))      # wave (dependency level), then mode, then image, then scan
        img, kidx, flat_t, mode, level = img[order], kidx[order], flat_t[order], mode[order], level[order]
        rstart, rend = run_start[order], run_end[order]
        recs = t_recs[flat_t].copy()
        raw_off = offsets[img] + rstart
        raw_len = rend - rstart
        if len(recs) and ((raw_len < 0).any() or (raw_off < 0).any() or (raw_off + raw_len > self.raw_bytes).any()):
            from .errors import CorruptedJpeg
            bad = int(img[np.flatnonzero((raw_len < 0) | (raw_off < 0) | (raw_off + raw_len > self.raw_bytes))[0]])
            raise CorruptedJpeg(f"File {bad}: entropy-coded segment lies outside the file.")
        n_streams = recs["n_streams"].astype(np.int64)
        n_sub_max = -(-(raw_len * 8) // SUBSEQ_BITS) + n_streams
        n_tiles = np.maximum(1, -(-((raw_off & 15) + raw_len) // UNSTUFF_TILE))
        recs["raw_off"], recs["raw_len"] = raw_off, raw_len
        recs["coef_block0"] = block0[img]
        recs["image"] = img
        recs["stream0"] = np.concatenate(([0], np.cumsum(n_streams)[:-1]))
        recs["sub0"] = np.concatenate(([0], np.cumsum(n_sub_max)[:-1]))
        recs["n_sub_max"] = n_sub_max
        recs["tile0"] = np.concatenate(([0], np.cumsum(n_tiles)[:-1]))
        # LUT blobs of the batch, de-duplicated in plan order (first use), like BatchPlan does
        lut_parts: List[np.ndarray] = []
        lut_off_of: Dict[tuple, int] = {}
        lut_size = 0
        t_lutoff = np.zeros(len(t_recs), dtype=np.int64)
        uniq, first_pos = np.unique(flat_t, return_index=True)
        t_owner = np.repeat(np.arange(len(templates)), t_nscan)
        for ft in uniq[np.argsort(first_pos)]:
            k = int(t_owner[ft])
            bkey, blob = templates[k].blobs[int(ft - t_sbase[k])]
            if bkey not in lut_off_of:
                lut_off_of[bkey] = lut_size
                lut_parts.append(blob)
                lut_size += len(blob)
            t_lutoff[ft] = lut_off_of[bkey]
        recs["lut_off"] = t_lutoff[flat_t]
        assert recs.dtype == SCAN_DTYPE and recs.dtype.itemsize == 144
        self.scans = recs
        self.n_streams = int(n_streams.sum())
        self.n_sub = int(n_sub_max.sum())
        self.n_tiles = int(n_tiles.sum())
        self.tile_scan = np.repeat(np.arange(len(recs), dtype=np.uint32), n_tiles)
        self.lut = np.concatenate(lut_parts).astype(np.uint32)
        # groups: runs of equal (wave, mode)
        self.groups: List[ScanGroup] = []
        if len(recs):
            change = np.flatnonzero((np.diff(level) != 0) | (np.diff(mode) != 0)) + 1
            bounds = np.concatenate(([0], change, [len(recs)]))
            blocks_scan = recs["n_mcu"].astype(np.int64) * recs["nslots"].astype(np.int64)
            for a, b in zip(bounds[:-1], bounds[1:]):
                self.groups.append(ScanGroup(first=int(a), count=int(b - a), mode=int(mode[a]),
                                             max_sub=int(n_sub_max[a:b].max()), max_streams=int(n_streams[a:b].max()),
                                             max_blocks=int(blocks_scan[a:b].max()), max_lut=int(recs["lut_len"][a:b].max())))
        chains = [(-(-gp.max_sub // ENTROPY_THREADS)) * gp.count for gp in self.groups if gp.mode in (0, 1, 3)]
        self.max_chain = max(chains) if chains else 1


class _Fallback(Exception):
    """Internal: the fast planner cannot handle this batch; use the per-file Python path."""


def plan_batch(raw_host, offsets: Sequence[int], sizes: Sequence[int], threads: Optional[int] = None, walked=None):
    """Plan for a packed batch: FastPlan when possible, else the per-file BatchPlan (same attributes)."""
    raw_np = raw_host.numpy() if hasattr(raw_host, "numpy") else np.asarray(raw_host)
    try:
        return FastPlan(raw_np, offsets, sizes, threads, walked)
    except _Fallback:
        from .pipeline import BatchPlan
        mv = memoryview(raw_np)
        parsed = [parse_jpeg(mv[o:o + s].tobytes()) for o, s in zip(offsets, sizes)]
        return BatchPlan(parsed, list(offsets), int(raw_np.size))
