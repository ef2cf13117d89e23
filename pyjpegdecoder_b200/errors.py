"""Decoder exceptions: same names and hierarchy as the reference (jpeg_decoder.py:1714-1725)."""


class JpegError(Exception):
    """Parent of all other exceptions of this decoder."""


class NotJpeg(JpegError):
    """File is not a JPEG image."""


class CorruptedJpeg(JpegError):
    """Failed to parse the file headers."""


class UnsupportedJpeg(JpegError):
    """JPEG image is encoded in a way that our decoder does not support."""


class NativeLibraryError(RuntimeError):
    """The CUDA library (libb200jpeg.so) is missing, failed to load, or a CUDA call failed.
    There is no CPU fallback: the decode path fails loudly instead."""
