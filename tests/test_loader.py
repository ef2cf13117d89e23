"""Streaming front end (pyjpegdecoder_b200/loader.py): chunking logic on the CPU, equality with decode_batch on the GPU."""
import io

import numpy as np
import pytest


def _jpeg(w, h, seed, **kw):
    from PIL import Image
    rng = np.random.default_rng(seed)
    b = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(b, "JPEG", quality=80, **kw)
    return b.getvalue()


def test_chunks_keeps_order_and_sizes():
    from pyjpegdecoder_b200.loader import _chunks
    assert [len(c) for c in _chunks(range(10), 4)] == [4, 4, 2]
    assert [x for c in _chunks(iter(range(7)), 3) for x in c] == list(range(7))
    assert list(_chunks([], 5)) == []
    # large sub-batches ramp up from n/2 by a quarter per step (a short first sub-batch cuts the start-up latency)
    assert [len(c) for c in _chunks(iter(range(3000)), 512)] == [256, 320, 400, 500, 512, 512, 500]
    assert [len(c) for c in _chunks(range(2000), 512, ramp=False)] == [512, 512, 512, 464]
    assert [x for c in _chunks(iter(range(1000)), 256) for x in c] == list(range(1000))


def test_decode_stream_needs_a_gpu_and_never_falls_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pyjpegdecoder_b200 import NativeLibraryError, decode_stream
    with pytest.raises(NativeLibraryError):
        next(decode_stream([_jpeg(16, 16, 0)]))
    with pytest.raises(ValueError):
        next(decode_stream([_jpeg(16, 16, 0)], chunk=0))


def test_host_pack_matches_python_copy():
    """bj_host_pack (threaded gather in C) against a plain Python copy."""
    from pyjpegdecoder_b200.pipeline import pack_files
    rng = np.random.default_rng(3)
    datas = [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in rng.integers(1, 5000, 40)]
    buf, offs = pack_files(datas, pin=False)
    v = buf.numpy()
    for d, o in zip(datas, offs):
        assert o % 16 == 0 and bytes(v[o:o + len(d)]) == d


@pytest.mark.gpu
def test_decode_stream_equals_decode_batch(tmp_path):
    from pyjpegdecoder_b200 import decode_batch, decode_stream
    files = []
    for i in range(37):
        kw = {}
        if i % 5 == 1:
            kw["progressive"] = True
        if i % 3 == 0:
            kw["subsampling"] = i % 3
        data = _jpeg(40 + 13 * (i % 7), 24 + 9 * (i % 5), i, **kw)
        if i % 2:
            path = tmp_path / f"f{i}.jpg"
            path.write_bytes(data)
            files.append(path)
        else:
            files.append(data)
    ref = decode_batch(files, device="cuda:0")
    out = [d for chunk in decode_stream(iter(files), chunk=8, device="cuda:0") for d in chunk]
    assert len(out) == len(ref) == 37
    for a, b in zip(out, ref):
        assert np.array_equal(a.image_array, b.image_array)


@pytest.mark.gpu
def test_concurrent_decode_batch_calls_do_not_share_staging_buffers():
    """Two host threads decoding different batches at the same time (what the multi-GPU dispatcher does with one
    thread per GPU): each call must get its own pinned staging buffer."""
    import threading
    from pyjpegdecoder_b200 import decode_batch
    sets = [[_jpeg(64 + 8 * t, 48, 100 * t + i) for i in range(24)] for t in range(2)]
    ref = [decode_batch(s, device="cuda:0") for s in sets]
    for _ in range(4):
        out = [None, None]

        def work(t):
            out[t] = decode_batch(sets[t], device="cuda:0")
        th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
        [x.start() for x in th]
        [x.join() for x in th]
        for t in range(2):
            assert out[t] is not None
            for a, b in zip(out[t], ref[t]):
                assert np.array_equal(a.image_array, b.image_array)


def test_lazy_views_and_lazy_decoder_attributes_on_cpu():
    """_LazyViews / JpegDecoder's lazy attributes need no GPU: exercise them on a CPU tensor with a real parse."""
    import torch
    from pyjpegdecoder_b200.decoder import JpegDecoder
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import _LazyViews
    from pyjpegdecoder_b200.plan import BatchGeometry

    datas = [_jpeg(40, 24, 1), _jpeg(17, 33, 2, progressive=True)]
    parsed = [parse_jpeg(d) for d in datas]
    geom = BatchGeometry(parsed)
    out = torch.arange(geom.out_bytes, dtype=torch.int64).to(torch.uint8)
    views = _LazyViews(geom, out)
    assert len(views) == 2 and views[0].shape == (24, 40, 3) and views[-1].shape == (33, 17, 3)
    assert views[1].data_ptr() == out.data_ptr() + geom.out_offsets[1] and views[1] is views[1]
    assert [tuple(v.shape) for v in views] == [(24, 40, 3), (33, 17, 3)]

    class FakePlan:
        pass

    class FakeBatch:
        pass
    fb = FakeBatch()
    fb.plan = FakePlan()
    fb.plan.parsed = parsed
    fb.images = views
    d = JpegDecoder(datas[1], _batch=fb, _index=1)
    assert "image_width" not in d.__dict__              # nothing materialised yet
    assert (d.image_width, d.image_height, d.scan_mode) == (17, 33, "progressive_dct")
    assert d.scan_count == len(parsed[1].scans) and d.file_size == len(datas[1])
    assert d.image_array.shape == (17, 33, 3)           # the reference's x-major layout
    with pytest.raises(AttributeError):
        d.no_such_attribute


@pytest.mark.gpu
def test_streamed_sub_batches_release_their_work_buffers():
    """decode_stream keeps only the pixels of a finished sub-batch (the coefficient planes are as large again);
    keep_coefficients=True keeps the planes readable."""
    from pyjpegdecoder_b200 import decode_batch, decode_stream
    files = [_jpeg(64, 48, i) for i in range(12)]
    ref = decode_batch(files, device="cuda:0")
    got = [d for part in decode_stream(files, chunk=5, device="cuda:0") for d in part]
    assert got[0]._batch.coef is None and "_pipe" not in got[0]._batch.stats
    for a, b in zip(got, ref):
        assert np.array_equal(a.image_array, b.image_array)
    with pytest.raises(RuntimeError):
        got[0].coefficient_planes()
    kept = [d for part in decode_stream(files, chunk=5, device="cuda:0", keep_coefficients=True) for d in part]
    for a, b in zip(kept, ref):
        for pa, pb in zip(a.coefficient_planes(), b.coefficient_planes()):
            assert np.array_equal(pa, pb)


@pytest.mark.gpu
def test_to_host_gives_the_same_arrays_through_one_copy_per_sub_batch():
    """to_host=True: image_array is a view of the sub-batch's pinned host copy (the reference's return type, host
    memory) -- same values and layout as the per-image path, for the one-shot and the streamed call."""
    from pyjpegdecoder_b200 import decode_batch, decode_stream
    files = [_jpeg(40 + 8 * (i % 5), 32 + 8 * (i % 3), i, subsampling=2 if i % 2 else 0) for i in range(14)]
    ref = decode_batch(files, device="cuda:0")
    one = decode_batch(files, device="cuda:0", to_host=True)
    many = [d for part in decode_stream(files, chunk=4, device="cuda:0", to_host=True) for d in part]
    for r, a, b in zip(ref, one, many):
        assert a._batch.host_image(a._index) is not None and b._batch.host_image(b._index) is not None
        assert a.image_array.shape == r.image_array.shape == b.image_array.shape
        assert np.array_equal(a.image_array, r.image_array) and np.array_equal(b.image_array, r.image_array)


def test_read_files_packed_matches_the_files(tmp_path):
    """Path inputs are read straight into the packed (pinned on a GPU box) buffer: offsets 16-byte aligned, bytes equal."""
    import os
    from pyjpegdecoder_b200.pipeline import read_files_packed
    blobs = [os.urandom(n) for n in (1, 15, 16, 17, 4097, 70001, 0, 33)]
    paths = []
    for i, b in enumerate(blobs):
        p = tmp_path / f"f{i}.bin"
        p.write_bytes(b)
        paths.append(p if i % 2 else str(p))
    buf, offs, sizes = read_files_packed(paths, read_threads=4, walk=False)
    raw = buf.numpy()
    assert sizes == [len(b) for b in blobs] and all(o % 16 == 0 for o in offs)
    for b, o in zip(blobs, offs):
        assert bytes(raw[o:o + len(b)]) == b
    with pytest.raises(OSError):
        read_files_packed([tmp_path / "missing.bin"])


@pytest.mark.gpu
def test_stream_of_paths_reads_into_pinned_memory_and_matches_bytes(tmp_path):
    from pyjpegdecoder_b200 import NotJpeg, decode_batch, decode_stream
    datas = [_jpeg(48 + 8 * (i % 4), 40 + 8 * (i % 3), i, subsampling=2 if i % 2 else 0, progressive=(i % 5 == 0)) for i in range(19)]
    paths = []
    for i, d in enumerate(datas):
        p = tmp_path / f"img{i}.jpg"
        p.write_bytes(d)
        paths.append(p)
    ref = decode_batch(datas, device="cuda:0")
    got = [d for part in decode_stream(paths, chunk=6, device="cuda:0") for d in part]
    assert len(got) == len(ref)
    for a, b in zip(decode_batch(paths, device="cuda:0"), ref):       # the one-shot call reads paths the same way
        assert np.array_equal(a.image_array, b.image_array)
    for a, b, p in zip(got, ref, paths):
        assert np.array_equal(a.image_array, b.image_array)
        assert a.file_path == p
    bad = tmp_path / "bad.jpg"
    bad.write_bytes(b"definitely not a jpeg")
    with pytest.raises(NotJpeg):
        list(decode_stream(paths[:5] + [bad], chunk=6, device="cuda:0"))


def test_read_files_packed_walks_like_the_batch_walker():
    """The walk that rides along with the file reads (C threads, right after each read) equals the stand-alone walk of
    the packed buffer: same entries, counts and key hashes -- so plan_batch builds the same plan either way."""
    from conftest import GOLDEN
    from pyjpegdecoder_b200.fastplan import plan_batch, walk_batch
    from pyjpegdecoder_b200.pipeline import read_files_packed
    paths = sorted((GOLDEN / "cases").glob("*.jpg"))[:12]
    buf, offs, sizes = read_files_packed(paths, read_threads=3, walk=True)
    entries, counts, hashes = buf._bj_walk
    e2, c2, h2 = walk_batch(buf.numpy(), np.asarray(offs, dtype=np.uint64), np.asarray(sizes, dtype=np.uint64))
    assert np.array_equal(counts, c2) and np.array_equal(hashes, h2)
    for i in range(len(paths)):
        assert np.array_equal(entries[i, :counts[i]], e2[i, :c2[i]])
    a = plan_batch(buf, offs, sizes, walked=buf._bj_walk)
    b = plan_batch(buf, offs, sizes)
    assert len(a.parsed) == len(b.parsed) == len(paths)
    assert [p.width for p in a.parsed] == [p.width for p in b.parsed]
