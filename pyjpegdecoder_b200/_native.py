"""ctypes binding of libb200jpeg.so (see include/b200jpeg.h for the C ABI).

The product path has no CPU fallback: `lib()` raises NativeLibraryError when the CUDA library is
missing or cannot be loaded.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_void_p

import numpy as np

from .errors import NativeLibraryError

_LIB = None

# numpy mirror of struct bj_image (72 bytes)
IMAGE_DTYPE = np.dtype([
    ("coef_block0", "<u8"), ("out_offset", "<u8"), ("out_pitch", "<u4"),
    ("width", "<u4"), ("height", "<u4"), ("mcus_x", "<u4"), ("mcus_y", "<u4"),
    ("qtab", "<u4", (3,)),
    ("ncomp", "u1"), ("hs", "u1", (3,)), ("vs", "u1", (3,)), ("hmax", "u1"), ("vmax", "u1"),
    ("blocks_per_mcu", "u1"), ("slot0", "u1", (3,)), ("pad0", "u1"),
    ("strip_mcus", "<u2"), ("strips_per_row", "<u2"), ("pad1", "<u2"), ("layout", "<u4"),
])
assert IMAGE_DTYPE.itemsize == 72

OUT_RGB, OUT_SAMPLES, OUT_CANVAS = 0, 1, 2
IN_COEF, IN_SAMPLES = 0, 1
PIXEL_MAX_BLOCKS = 192
LAYOUT_GENERIC, LAYOUT_420, LAYOUT_422, LAYOUT_440, LAYOUT_444, LAYOUT_GRAY = range(6)

ERR_BAD_CODE, ERR_OVERRUN, ERR_RST_COUNT, ERR_SYNC, ERR_COEF_INDEX = 1, 2, 4, 8, 16


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    from .build import LIB
    if not LIB.exists():
        raise NativeLibraryError(
            f"{LIB} is missing: build it with `python -m pyjpegdecoder_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    try:
        L = ctypes.CDLL(str(LIB))
    except OSError as e:  # pragma: no cover
        raise NativeLibraryError(f"cannot load {LIB}: {e}") from e
    L.bj_version.restype = c_int
    L.bj_last_cuda_error.restype = ctypes.c_char_p
    L.bj_sizeof.restype = c_int
    L.bj_sizeof.argtypes = [c_int]
    L.bj_pixels.restype = c_int
    L.bj_pixels.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_int,
                            ctypes.c_uint32, c_void_p, c_void_p]
    if L.bj_sizeof(0) != IMAGE_DTYPE.itemsize:
        raise NativeLibraryError("struct bj_image layout mismatch between Python and libb200jpeg.so")
    L.bj_pixels_fast_strip.restype = c_int
    L.bj_pixels_fast_strip.argtypes = [c_int]
    from .plan import FAST_STRIP
    for lay, strip in FAST_STRIP.items():
        if L.bj_pixels_fast_strip(lay) != strip:
            raise NativeLibraryError("pixel-kernel strip sizes differ between Python and libb200jpeg.so")
    _LIB = L
    return L


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().bj_last_cuda_error().decode(errors="replace") if status == 2 else f"status {status}"
        raise NativeLibraryError(f"{what} failed: {msg}")
