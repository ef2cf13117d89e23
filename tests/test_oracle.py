"""Pins the CPU oracle (oracle/jpeg_oracle.c) against outputs of the unmodified reference.

Fixtures come from tests/golden/make_golden.py (run where /root/reference exists).  Everything here
is bit-exact: RGB, the int16 Y/Cb/Cr canvas, quantised coefficient planes after every scan.
"""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, golden_case_names


def _load(name):
    data = (GOLDEN / "cases" / f"{name}.jpg").read_bytes()
    z = np.load(GOLDEN / "cases" / f"{name}.npz")
    return data, z


def test_idct_table_matches_reference():
    # jpeg_decoder.py:1541-1553, dumped from the reference's class attribute
    assert np.array_equal(np.load(GOLDEN / "idct_table.npy"), oracle.idct_table())


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_matches_reference(name, golden_meta):
    data, z = _load(name)
    m = golden_meta["cases"][name]
    r = oracle.decode(data)
    assert (r.width, r.height) == (m["width"], m["height"])
    assert np.array_equal(r.image_array, z["rgb"])
    assert np.array_equal(np.swapaxes(r.canvas, 0, 1), z["canvas"])
    for c in range(r.ncomp):
        assert np.array_equal(r.coef[c], z[f"coef{c}"])
    assert hashlib.sha256(np.ascontiguousarray(r.image_array).tobytes()).hexdigest() == m["rgb_sha256"]


@pytest.mark.parametrize("name", [n for n in golden_case_names() if n.startswith("prog_")])
def test_oracle_progressive_per_scan(name, golden_meta):
    data, z = _load(name)
    m = golden_meta["cases"][name]
    for k in range(1, m["scans"] + 1):
        r = oracle.decode(data, stop_after_scan=k, want=("coef",))
        for c in range(r.ncomp):
            assert np.array_equal(r.coef[c], z[f"scan{k}_coef{c}"]), (k, c)


def test_oracle_base_image_full(golden_meta):
    """The reference's own example file: full decode hash and per-scan coefficient hashes."""
    g = golden_meta["base_image"]
    data = (GOLDEN / "base_image.jpg").read_bytes()
    assert hashlib.sha256(data).hexdigest() == g["file_sha256"]
    r = oracle.decode(data)
    assert r.image_array.shape == (g["width"], g["height"], 3)
    assert hashlib.sha256(np.ascontiguousarray(r.image_array).tobytes()).hexdigest() == g["rgb_sha256"]
    for k, hashes in enumerate(g["scan_coef_sha256"], start=1):
        rk = oracle.decode(data, stop_after_scan=k, want=("coef",))
        assert [hashlib.sha256(p.tobytes()).hexdigest() for p in rk.coef] == hashes, k


@pytest.mark.parametrize("k", [1, 2])
def test_oracle_after_scan_renders(k, golden_meta):
    """`after scan 0k.png` shipped with the reference = decode of the file cut after scan k + EOI."""
    g = golden_meta["base_image"]
    data = (GOLDEN / "base_image.jpg").read_bytes()
    cut = data[: g[f"after_scan_{k}"]["truncate_at"]] + b"\xff\xd9"
    r = oracle.decode(cut, want=("rgb",))
    assert hashlib.sha256(np.ascontiguousarray(r.rgb).tobytes()).hexdigest() == g[f"after_scan_{k}"]["rgb_hw3_sha256"]
    # the same through the stop_after_scan switch used by the per-scan parity tests
    r2 = oracle.decode(data, stop_after_scan=k, want=("rgb",))
    assert np.array_equal(r.rgb, r2.rgb)


def test_upsample_matches_griddata_weights():
    """ResizeGrid (:1588-1626): unit-impulse weights recorded from the reference (integers /15)."""
    w = np.load(GOLDEN / "upsample_weights.npz")
    rng = np.random.default_rng(0)
    for key, (rh, rv) in {"w_8x8_16x16": (2, 2), "w_8x8_16x8": (2, 1), "w_8x8_8x16": (1, 2)}.items():
        wt = w[key].astype(np.int64)  # [i, j, a, b]
        for _ in range(20):
            tile = rng.integers(-300, 600, (8, 8)).astype(np.int16)
            n = np.einsum("ij,ijab->ab", tile.astype(np.int64), wt)
            want = np.floor_divide(2 * n + 15, 30).astype(np.int16)
            assert np.array_equal(oracle.upsample_tile(tile, rh, rv), want), key


def test_upsample_matches_live_scipy():
    """Regenerate the triangulation's weights from the scipy installed here: if a
    different scipy/Qhull changes the diagonal map this test says so."""
    scipy_interp = pytest.importorskip("scipy.interpolate")
    xx, yy = np.indices((8, 8))
    new_x, new_y = np.mgrid[0:7:16j, 0:7:16j]
    rng = np.random.default_rng(1)
    for _ in range(5):
        tile = rng.integers(-300, 600, (8, 8)).astype(np.int16)
        ref = np.round(scipy_interp.griddata((xx.flatten(), yy.flatten()), tile.ravel(), (new_x, new_y))).astype(np.int16)
        assert np.array_equal(oracle.upsample_tile(tile, 2, 2), ref)


def test_diag_map_constant():
    assert oracle.diag_map() == sum(
        1 << (7 * i + j) for i, row in enumerate(["1010010", "1001001", "1100101", "0011010", "1010101", "0101010", "1010101"])
        for j, ch in enumerate(row) if ch == "1")


def test_oracle_errors():
    with pytest.raises(oracle.OracleError) as e:
        oracle.decode(b"\x89PNG....")
    assert e.value.kind == "NotJpeg"
    data, _ = _load("base_70x50_ss2")
    bad = bytearray(data)
    i = data.find(b"\xff\xc0")
    bad[i + 4] = 12  # precision
    with pytest.raises(oracle.OracleError) as e:
        oracle.decode(bytes(bad))
    assert e.value.kind == "UnsupportedJpeg"
