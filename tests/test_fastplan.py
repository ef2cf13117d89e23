"""The fast batch planner (C marker walk + per-key templates + numpy assembly) must produce exactly the
plan that the per-file Python path produces."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_case_names


def _both(datas):
    from pyjpegdecoder_b200.fastplan import FastPlan
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import BatchPlan, pack_files
    raw, offs = pack_files(datas, pin=False)
    slow = BatchPlan([parse_jpeg(d) for d in datas], offs, raw.numel())
    fast = FastPlan(raw.numpy(), offs, [len(d) for d in datas], threads=4)
    return slow, fast


def _assert_same(slow, fast):
    assert slow.scans.dtype == fast.scans.dtype and slow.geom.images.dtype == fast.geom.images.dtype
    # byte-level: these arrays are uploaded as raw struct bj_scan[] / bj_image[]
    assert np.array_equal(np.ascontiguousarray(slow.scans).view(np.uint8), np.ascontiguousarray(fast.scans).view(np.uint8))
    assert np.array_equal(np.ascontiguousarray(slow.geom.images).view(np.uint8),
                          np.ascontiguousarray(fast.geom.images).view(np.uint8))
    assert np.array_equal(slow.geom.qtabs, fast.geom.qtabs)
    assert np.array_equal(slow.tile_scan, fast.tile_scan)
    assert np.array_equal(slow.lut, fast.lut)
    for a in ("n_streams", "n_sub", "n_tiles", "max_chain", "any_progressive", "raw_bytes"):
        assert getattr(slow, a) == getattr(fast, a), a
    for a in ("total_blocks", "out_bytes", "max_strips", "layout_mask", "block_offsets", "out_offsets", "out_shapes"):
        assert getattr(slow.geom, a) == getattr(fast.geom, a), a
    assert [vars(g) for g in slow.groups] == [vars(g) for g in fast.groups]
    assert len(slow.parsed) == len(fast.parsed)
    for p, q in zip(slow.parsed, fast.parsed):
        assert (p.width, p.height, p.progressive, p.file_size) == (q.width, q.height, q.progressive, q.file_size)
        assert [(s.data_start, s.data_end, s.kind) for s in p.scans] == [(s.data_start, s.data_end, s.kind) for s in q.scans]


def test_fastplan_equals_batchplan_on_all_fixtures():
    datas = [(GOLDEN / "cases" / f"{n}.jpg").read_bytes() for n in golden_case_names()]
    _assert_same(*_both(datas))


def test_fastplan_with_repeated_and_interleaved_files():
    names = ["base_120x88_ss2", "prog_97x61_ss1_dri5", "base_gray_70x50", "base_120x88_ss2", "prog_97x61_ss1_dri5",
             "base_97x61_ss0_dri13", "base_120x88_ss2"]
    datas = [(GOLDEN / "cases" / f"{n}.jpg").read_bytes() for n in names]
    _assert_same(*_both(datas))


def test_fastplan_big_file():
    data = (GOLDEN / "base_image.jpg").read_bytes()
    _assert_same(*_both([data, data]))


def test_fastplan_raises_like_the_parser():
    from pyjpegdecoder_b200 import NotJpeg, UnsupportedJpeg
    from pyjpegdecoder_b200.fastplan import FastPlan
    from pyjpegdecoder_b200.pipeline import pack_files
    good = (GOLDEN / "cases" / "base_70x50_ss2.jpg").read_bytes()
    bad = bytearray(good)
    i = good.find(b"\xff\xc0")
    bad[i + 4] = 12
    raw, offs = pack_files([good, bytes(bad)], pin=False)
    with pytest.raises(UnsupportedJpeg):
        FastPlan(raw.numpy(), offs, [len(good), len(bad)])
    raw, offs = pack_files([good, b"\x89PNG-not-a-jpeg"], pin=False)
    with pytest.raises(NotJpeg):
        FastPlan(raw.numpy(), offs, [len(good), 15])


def test_scan_levels_of_the_default_progressive_script():
    """libjpeg's default progressive script: 10 scans collapse into 3 dependency levels (5 launch waves)."""
    import io
    import numpy as np
    from PIL import Image
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.plan import scan_levels
    rng = np.random.default_rng(0)
    b = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(b, "JPEG", progressive=True, subsampling=2)
    p = parse_jpeg(b.getvalue())
    kinds = [s.kind for s in p.scans]
    assert kinds == ["dc_first", "ac_first", "ac_first", "ac_first", "ac_first", "ac_refine", "dc_refine",
                     "ac_refine", "ac_refine", "ac_refine"]
    assert scan_levels(p) == [0, 0, 0, 0, 0, 1, 1, 1, 1, 2]
    assert scan_levels(p, serial=True) == list(range(10))


def test_pack_with_walk_gives_the_same_walk_and_plan():
    """pack_files(walk=True) (copy + marker walk + key hash in one pass) against the separate walk, and the plans
    built from either."""
    import io
    import numpy as np
    from PIL import Image
    from pyjpegdecoder_b200.fastplan import plan_batch, walk_batch
    from pyjpegdecoder_b200.pipeline import pack_files
    rng = np.random.default_rng(11)
    datas = []
    for i in range(12):
        b = io.BytesIO()
        Image.fromarray(rng.integers(0, 256, (24 + 8 * (i % 3), 40, 3), dtype=np.uint8)).save(
            b, "JPEG", quality=70, progressive=bool(i % 2), subsampling=i % 3)
        datas.append(b.getvalue())
    buf, offs = pack_files(datas, pin=False, walk=True)
    entries, counts, hashes = buf._bj_walk
    e2, c2, h2 = walk_batch(buf.numpy(), np.asarray(offs), np.asarray([len(d) for d in datas]))
    assert np.array_equal(counts, c2) and np.array_equal(hashes, h2)
    for i, c in enumerate(counts):
        assert entries[i, :c].tobytes() == e2[i, :c].tobytes()
    a = plan_batch(buf, offs, [len(d) for d in datas], walked=buf._bj_walk)
    b_ = plan_batch(buf, offs, [len(d) for d in datas])
    assert a.scans.tobytes() == b_.scans.tobytes() and np.array_equal(a.tile_scan, b_.tile_scan)
    assert np.array_equal(a.lut, b_.lut) and [vars(g) for g in a.groups] == [vars(g) for g in b_.groups]


# ---- malformed headers: wrapped scan lengths must never reach the device ------------------------
def _crafted_sos_past_eof():
    """A valid small baseline file whose SOS length field is patched to point far past EOF."""
    from pathlib import Path
    data = bytearray((Path(__file__).parent / "golden" / "cases" / "base_120x88_ss2.jpg").read_bytes())
    sos = data.find(b"\xff\xda")
    data[sos + 2:sos + 4] = (0xFFF0).to_bytes(2, "big")
    return bytes(data[:sos + 40])


def test_sos_length_past_eof_is_corrupted_jpeg_in_both_planners():
    import numpy as np
    import pytest
    from pyjpegdecoder_b200.errors import CorruptedJpeg
    from pyjpegdecoder_b200.fastplan import FastPlan
    from pyjpegdecoder_b200.parser import parse_jpeg
    bad = _crafted_sos_past_eof()
    with pytest.raises(CorruptedJpeg):
        parse_jpeg(bad)
    files = [bad] * 8
    offs, total = [], 0
    for f in files:
        offs.append(total)
        total += (len(f) + 15) & ~15
    raw = np.zeros(total + 64, np.uint8)
    for f, o in zip(files, offs):
        raw[o:o + len(f)] = np.frombuffer(f, np.uint8)
    with pytest.raises(CorruptedJpeg):
        FastPlan(raw, offs, [len(f) for f in files])
    with pytest.raises(CorruptedJpeg):          # cached (type, message), raised again as a fresh instance
        FastPlan(raw, offs, [len(f) for f in files])


def test_too_many_blocks_per_mcu_is_a_jpeg_error():
    import pytest
    from pathlib import Path
    from pyjpegdecoder_b200.errors import JpegError
    from pyjpegdecoder_b200.parser import parse_jpeg
    data = bytearray((Path(__file__).parent / "golden" / "cases" / "base_120x88_ss2.jpg").read_bytes())
    sof = data.find(b"\xff\xc0")
    for i in range(3):                       # 2x2 sampling on every component: 12 blocks per MCU
        data[sof + 4 + 6 + 3 * i + 1] = 0x22
    with pytest.raises(JpegError):
        parse_jpeg(bytes(data))


def test_unscanned_components_force_a_zeroed_coefficient_buffer():
    from pathlib import Path
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import covers_all_components
    p = parse_jpeg((Path(__file__).parent / "golden" / "cases" / "base_120x88_ss2.jpg").read_bytes())
    assert covers_all_components(p)
    p.scans[0].comps = (0,)
    assert not covers_all_components(p)


def test_header_fuzz_only_jpeg_errors_and_in_bounds_plans():
    """Random byte flips in the marker segments (SOF, DHT, DQT, DRI, SOS lengths, ...): the host side either builds a
    plan whose scan byte ranges lie inside the packed buffer, or raises a JpegError -- never another exception type,
    never a wrapped / negative length (what the device kernels would then index with)."""
    import numpy as np
    from pathlib import Path
    from pyjpegdecoder_b200.errors import JpegError
    from pyjpegdecoder_b200.fastplan import FastPlan, _Fallback
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import BatchPlan
    cases = Path(__file__).parent / "golden" / "cases"
    names = sorted(p.name for p in cases.glob("*.jpg"))
    rng = np.random.default_rng(11)
    n_ok = n_err = 0
    for it in range(400):
        data = bytearray((cases / names[int(rng.integers(len(names)))]).read_bytes())
        first_scan = parse_jpeg(bytes(data)).scans[0].data_start
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(2, first_scan))
            data[pos] = int(rng.integers(256)) if rng.random() < 0.5 else data[pos] ^ (1 << int(rng.integers(8)))
        bad = bytes(data)
        files = [bad] * 5
        offs, total = [], 0
        for f in files:
            offs.append(total)
            total += (len(f) + 15) & ~15
        raw = np.zeros(total + 64, np.uint8)
        for f, o in zip(files, offs):
            raw[o:o + len(f)] = np.frombuffer(f, np.uint8)
        plans = []
        try:
            plans.append(BatchPlan([parse_jpeg(bad)], [0], len(bad) + 64))
        except JpegError:
            n_err += 1
        try:
            plans.append(FastPlan(raw, offs, [len(f) for f in files]))
        except (JpegError, _Fallback):
            pass
        for plan in plans:
            n_ok += 1
            sc = plan.scans
            assert (sc["raw_off"].astype(np.int64) + sc["raw_len"].astype(np.int64) <= plan.raw_bytes).all()
            assert (sc["raw_len"].astype(np.int64) < 2 ** 31).all() and (sc["n_sub_max"].astype(np.int64) < 2 ** 31).all()
            assert plan.geom.total_blocks < 2 ** 31 and plan.geom.out_bytes < 2 ** 40
    assert n_ok > 50 and n_err > 50
