"""Host parser beyond what the reference accepts: fill bytes in front of markers and 16-bit
quantisation tables.  The reference misreads both (jpeg_decoder.py:93-106 treats FF FF as a segment, :443-454 reads 64
bytes whatever Pq says), so there is no reference output to compare with: the checks are structural on the CPU (the
rewritten file parses to the same tables and the same entropy-coded bytes as the original) and, on the GPU, that the
rewritten file decodes to the same pixels as the original."""
import numpy as np
import pytest

from conftest import GOLDEN

NAMES = ["base_120x88_ss2", "prog_97x61_ss1_dri5", "base_gray_33x17_dri2", "base_256x128_ss2_opt"]


def _segments(data: bytes):
    """(marker, start of the segment incl. the FF xx, end) for every marker segment up to the first SOS."""
    out, pos = [], 2
    while True:
        assert data[pos] == 0xFF
        m = data[pos + 1]
        size = (data[pos + 2] << 8) | data[pos + 3]
        out.append((m, pos, pos + 2 + size))
        pos += 2 + size
        if m == 0xDA:
            return out


def with_fill_bytes(data: bytes) -> bytes:
    out = bytearray(data[:2])
    for k, (m, a, b) in enumerate(_segments(data)):
        out += b"\xff" * (1 + k % 3) + data[a:b]
    return bytes(out) + data[_segments(data)[-1][2]:].replace(b"\xff\xd9", b"\xff\xff\xff\xd9")


def with_16bit_dqt(data: bytes) -> bytes:
    out = bytearray(data[:2])
    segs = _segments(data)
    for (m, a, b) in segs:
        if m != 0xDB:
            out += data[a:b]
            continue
        seg, q, body = data[a + 4:b], 0, bytearray()
        while q < len(seg):
            body += bytes([0x10 | seg[q]]) + b"".join(bytes([0, v]) for v in seg[q + 1:q + 65])
            q += 65
        out += b"\xff\xdb" + (len(body) + 2).to_bytes(2, "big") + body
    return bytes(out) + data[segs[-1][2]:]


def _same_parse(a: bytes, b: bytes):
    from pyjpegdecoder_b200.parser import parse_jpeg
    pa, pb = parse_jpeg(a), parse_jpeg(b)
    assert (pa.width, pa.height, pa.progressive, len(pa.scans)) == (pb.width, pb.height, pb.progressive, len(pb.scans))
    assert pa.qtables.keys() == pb.qtables.keys()
    for k in pa.qtables:
        assert np.array_equal(pa.qtables[k], pb.qtables[k])
    for sa, sb in zip(pa.scans, pb.scans):
        assert a[sa.data_start:sa.data_end] == b[sb.data_start:sb.data_end]
        assert (sa.comps, sa.td, sa.ta, sa.ss, sa.se, sa.ah, sa.al, sa.ri) == (sb.comps, sb.td, sb.ta, sb.ss, sb.se, sb.ah, sb.al, sb.ri)
        assert sa.dc_specs == sb.dc_specs and sa.ac_specs == sb.ac_specs


@pytest.mark.parametrize("name", NAMES)
def test_fill_bytes_and_16bit_tables_parse_like_the_original(name):
    data = (GOLDEN / "cases" / f"{name}.jpg").read_bytes()
    _same_parse(data, with_fill_bytes(data))
    _same_parse(data, with_16bit_dqt(data))
    _same_parse(data, with_fill_bytes(with_16bit_dqt(data)))


def test_fast_planner_agrees_on_rewritten_files():
    """The C marker walk (fastplan) skips fill bytes like parser.py does: same plan arrays as the per-file path."""
    from pyjpegdecoder_b200.fastplan import FastPlan
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import BatchPlan
    files = [with_fill_bytes((GOLDEN / "cases" / f"{n}.jpg").read_bytes()) for n in NAMES] * 2
    offs, total = [], 0
    for f in files:
        offs.append(total)
        total += (len(f) + 15) & ~15
    raw = np.zeros(total + 64, np.uint8)
    for f, o in zip(files, offs):
        raw[o:o + len(f)] = np.frombuffer(f, np.uint8)
    fast = FastPlan(raw, offs, [len(f) for f in files])
    slow = BatchPlan([parse_jpeg(f) for f in files], offs, int(raw.size))
    assert np.array_equal(fast.scans, slow.scans)
    assert np.array_equal(fast.geom.images, slow.geom.images)


def test_oversized_16bit_table_entry_is_unsupported():
    from pyjpegdecoder_b200.errors import UnsupportedJpeg
    from pyjpegdecoder_b200.parser import parse_jpeg
    data = bytearray(with_16bit_dqt((GOLDEN / "cases" / "base_120x88_ss2.jpg").read_bytes()))
    p = data.find(b"\xff\xdb")
    data[p + 5] = 0x90           # first entry = 0x90xx > 32767
    with pytest.raises(UnsupportedJpeg):
        parse_jpeg(bytes(data))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_rewritten_files_decode_to_the_same_pixels(name):
    from pyjpegdecoder_b200 import decode_batch
    data = (GOLDEN / "cases" / f"{name}.jpg").read_bytes()
    z = np.load(GOLDEN / "cases" / f"{name}.npz")
    variants = [with_fill_bytes(data), with_16bit_dqt(data), with_fill_bytes(with_16bit_dqt(data))]
    for d in decode_batch(variants + variants, device="cuda:0"):      # 6 files: the fast planner path
        assert np.array_equal(d.image_array, z["rgb"])
    assert np.array_equal(decode_batch(variants[:1], device="cuda:0")[0].image_array, z["rgb"])


def test_on_error_return_needs_no_device_for_files_that_fail_on_the_host():
    """decode_batch(on_error="return") triages headers on the host: a list of files that are all bad comes back as
    exception instances without touching the GPU (so this runs on the CPU box)."""
    from pyjpegdecoder_b200 import JpegError, NotJpeg, decode_batch
    res = decode_batch([b"", b"\x89PNG\r\n\x1a\n" + bytes(32), b"\xff\xd8\xff\xd9"], on_error="return")
    assert len(res) == 3 and all(isinstance(r, JpegError) for r in res)
    assert isinstance(res[1], NotJpeg)
    with pytest.raises(ValueError):
        decode_batch([b"x"], on_error="ignore")
