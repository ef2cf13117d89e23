"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/b200jpeg.h declares; struct mirrors match; the host layer fails loudly without CUDA; the parser
raises the reference's exception classes."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    from pyjpegdecoder_b200 import _native
    L = _native.lib()
    header = (ROOT / "include" / "b200jpeg.h").read_text()
    names = set(re.findall(r"\b(bj_[a-z_0-9]+)\s*\(", header))
    names -= {"bj_status"}
    assert {"bj_version", "bj_pixels", "bj_unstuff", "bj_entropy_plan", "bj_entropy_decode"} <= names
    for n in sorted(names):
        assert hasattr(L, n), f"libb200jpeg.so does not export {n}"
    assert L.bj_version() >= 100


def test_struct_mirrors_match_the_library():
    from pyjpegdecoder_b200 import _native, pipeline
    L = pipeline._bind()
    assert L.bj_sizeof(0) == _native.IMAGE_DTYPE.itemsize == 72
    assert L.bj_sizeof_entropy(1) == pipeline.SCAN_DTYPE.itemsize == 144
    assert L.bj_sizeof_entropy(2) == ctypes.sizeof(pipeline.EntropyBuffers)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyjpegdecoder_b200 import JpegDecoder, NativeLibraryError
    with pytest.raises(NativeLibraryError):
        JpegDecoder(GOLDEN / "cases" / "base_8x8_ss2.jpg")


def test_product_never_imports_the_oracle():
    for f in (ROOT / "pyjpegdecoder_b200").rglob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f


def test_parser_exceptions_match_reference_classes():
    from pyjpegdecoder_b200 import CorruptedJpeg, JpegError, NotJpeg, UnsupportedJpeg, parse_jpeg
    data = (GOLDEN / "cases" / "base_70x50_ss2.jpg").read_bytes()
    with pytest.raises(NotJpeg):
        parse_jpeg(b"\x89PNG\r\n")
    i = data.find(b"\xff\xc0")
    bad = bytearray(data); bad[i + 4] = 12
    with pytest.raises(UnsupportedJpeg):
        parse_jpeg(bytes(bad))                      # precision (:155-156)
    bad = bytearray(data); bad[i + 9] = 4
    with pytest.raises(UnsupportedJpeg):
        parse_jpeg(bytes(bad))                      # CMYK (:179-180)
    bad = bytearray(data); bad[i + 7] = 0; bad[i + 8] = 0
    with pytest.raises(CorruptedJpeg):
        parse_jpeg(bytes(bad))                      # width 0 (:173-174)
    bad = bytearray(data); bad[i + 1] = 0xC1       # extended sequential: not in the handler table (:44-52)
    with pytest.raises(JpegError):
        parse_jpeg(bytes(bad))
    assert issubclass(NotJpeg, JpegError) and issubclass(CorruptedJpeg, JpegError) and issubclass(UnsupportedJpeg, JpegError)


def test_parser_matches_oracle_geometry():
    import oracle
    from pyjpegdecoder_b200 import parse_jpeg
    for f in sorted((GOLDEN / "cases").glob("*.jpg"))[::5]:
        data = f.read_bytes()
        p = parse_jpeg(data)
        r = oracle.decode(data, want=("coef",))
        assert (p.width, p.height, p.ncomp, p.progressive) == (r.width, r.height, r.ncomp, r.progressive)
        assert len(p.scans) == r.scan_count == p.scan_amount
        for c, g in zip(p.components, r.coef):
            assert g.shape == (p.mcus_y * c.v, p.mcus_x * c.h, 64)


def test_huffman_lut_decodes_every_code():
    """Device LUT (two-level) against the canonical code list of the DHT segment (:366-377)."""
    from pyjpegdecoder_b200 import parse_jpeg
    from pyjpegdecoder_b200.huffman import build_table, canonical_codes
    for name in ("base_120x88_ss2", "base_120x88_ss2_opt", "prog_120x88_ss2_q95"):
        p = parse_jpeg((GOLDEN / "cases" / f"{name}.jpg").read_bytes())
        for dest, spec in p.huff_specs.items():
            is_dc = (dest >> 4) == 0
            t = build_table(spec, is_dc)
            for code, length, sym in canonical_codes(spec):
                peek = (code << (16 - length)) | ((1 << (16 - length)) - 1)   # code followed by ones
                e = int(t[peek >> 7])
                if e & 0x80:
                    e = int(t[((e >> 8) & 0xFFFF) + (peek & 127)])
                assert (e >> 24, (e >> 16) & 255) == (sym, length), (name, dest, code, length)
