"""Thin Python wrappers over the C ABI stages.  torch tensors are device buffers only."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _native
from .errors import NativeLibraryError
from .plan import BatchGeometry, idct_table_t


def require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise NativeLibraryError("no CUDA device is visible: the B200 decode path has no CPU fallback")
    return torch.device(device if device is not None else "cuda")


def to_device(a: np.ndarray, device, non_blocking: bool = False) -> torch.Tensor:
    """Upload a numpy array (any dtype, incl. structured) as a uint8/typed device tensor."""
    if a.dtype.fields is not None or a.dtype == np.uint16 or a.dtype == np.uint32 or a.dtype == np.uint64:
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device, non_blocking=non_blocking)


_TABLE_CACHE = {}


def device_idct_table(device) -> torch.Tensor:
    key = str(device)
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = torch.from_numpy(idct_table_t().reshape(-1)).to(device)
    return _TABLE_CACHE[key]


class DeviceGeometry:
    """BatchGeometry uploaded to one device."""

    def __init__(self, geom: BatchGeometry, device, images: Optional[torch.Tensor] = None,
                 qtabs: Optional[torch.Tensor] = None):
        """images / qtabs: device copies that already exist (pipeline.upload_descriptors), else uploaded here."""
        self.geom = geom
        self.device = device
        self.images = images if images is not None else to_device(geom.images, device)
        self.qtabs = qtabs if qtabs is not None else to_device(geom.qtabs, device)
        self.table = device_idct_table(device)


def run_pixels(dg: DeviceGeometry, inp: torch.Tensor, in_kind: int, out_kind: int,
               out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
               stream: Optional[torch.cuda.Stream] = None, force_generic: bool = False) -> torch.Tensor:
    """bj_pixels(): coefficient (or sample) buffer -> RGB / samples / canvas."""
    g = dg.geom
    L = _native.lib()
    if out is None:
        if out_kind == _native.OUT_RGB:
            out = torch.empty(max(g.out_bytes, 16), dtype=torch.uint8, device=dg.device)
        elif out_kind == _native.OUT_SAMPLES:
            out = torch.empty((g.total_blocks, 64), dtype=torch.int16, device=dg.device)
        else:
            out = torch.zeros(max(g.out_bytes, 16), dtype=torch.int16, device=dg.device)
    s = stream if stream is not None else torch.cuda.current_stream(dg.device)
    with torch.cuda.device(dg.device):
        st = L.bj_pixels(dg.images.data_ptr(), len(g.parsed), g.max_strips, inp.data_ptr(), in_kind, g.total_blocks,
                         dg.qtabs.data_ptr(), dg.table.data_ptr(), out.data_ptr(), out_kind,
                         0 if force_generic else g.layout_mask,
                         stats.data_ptr() if stats is not None else None, s.cuda_stream)
    _native.check(st, "bj_pixels")
    return out


def image_views(g: BatchGeometry, out: torch.Tensor):
    """Per-image (H, W, 3) / (H, W) views of the flat output buffer."""
    views = []
    for off, shape in zip(g.out_offsets, g.out_shapes):
        n = int(np.prod(shape))
        views.append(out[off:off + n].view(*shape))
    return views
