"""CPU checks of the pixel-stage arithmetic (host build of csrc/bj_pixel_math.cuh) against the oracle.

The CUDA kernel accepts an fp32 result only when it is farther from a rounding tie than the error
bound; these tests check that this rule never accepts a wrong value, and that the interpolation
weights equal the reference's griddata weights.
"""
import ctypes

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from hostsim import build

ZZ_NAT = None


@pytest.fixture(scope="module")
def hs():
    L = build("pixel_hostsim")
    L.hs_div15.restype = ctypes.c_float
    L.hs_div15.argtypes = [ctypes.c_float]
    return L


def _idct(hs, blocks_nat):
    n = blocks_nat.shape[0]
    b = np.ascontiguousarray(blocks_nat, dtype=np.int32)
    out = np.empty((n, 64), np.int16)
    flg = np.empty(n, np.uint8)
    dist = np.empty(n, np.float32)
    hs.hs_idct_blocks(b.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p),
                      flg.ctypes.data_as(ctypes.c_void_p), dist.ctypes.data_as(ctypes.c_void_p))
    return out, flg.astype(bool), dist


def _oracle_idct(blocks_nat):
    # oracle wants block[u, v]; natural order is [v, u]; oracle returns [x, y]; we want [y, x]
    out = np.empty((blocks_nat.shape[0], 64), np.int16)
    for i, b in enumerate(blocks_nat):
        o = oracle.idct_block(b.reshape(8, 8).T.astype(np.int16))
        out[i] = o.T.reshape(64)
    return out


def _random_blocks(rng, n, mode):
    b = np.zeros((n, 64), np.int32)
    if mode == "dense":
        b[:] = rng.integers(-600, 600, (n, 64))
    elif mode == "sparse":
        mask = rng.random((n, 64)) < 0.15
        b[mask] = rng.integers(-300, 300, mask.sum())
        b[:, 0] = rng.integers(-1024, 1024, n)
    elif mode == "dc_only":
        b[:, 0] = rng.integers(-1024, 1024, n) * rng.integers(1, 20, n)
    elif mode == "ties":
        # DC = 8k+4 -> every sample is exactly k + 0.5: the reference decides by fp64 noise
        b[:, 0] = 8 * rng.integers(-120, 120, n) + 4
        m = rng.random(n) < 0.5
        b[m, 32] = 8 * rng.integers(-5, 5, m.sum())   # (v=4,u=0): keeps many samples on ties
    elif mode == "large":
        b[:] = rng.integers(-32768, 32767, (n, 64))
    return b


@pytest.mark.parametrize("mode", ["dense", "sparse", "dc_only", "ties", "large"])
def test_idct_fast_path_never_accepts_a_wrong_sample(hs, mode):
    rng = np.random.default_rng(hash(mode) & 0xFFFF)
    b = _random_blocks(rng, 1500, mode)
    fast, flagged, _ = _idct(hs, b)
    want = _oracle_idct(b)
    ok = ~flagged
    assert np.array_equal(fast[ok], want[ok])
    if mode == "ties":
        assert flagged.all()          # every such block must go to the exact path
    if mode == "sparse":
        assert flagged.mean() < 0.2   # and the exact path stays rare on ordinary content


def _idct_packed(hs, blocks_nat, lo4=False):
    n = blocks_nat.shape[0]
    b = np.ascontiguousarray(blocks_nat, dtype=np.int32)
    out = np.empty((n, 64), np.int16)
    flg = np.empty(n, np.uint8)
    dist = np.empty(n, np.float32)
    hs.hs_idct_blocks_packed(b.ctypes.data_as(ctypes.c_void_p), n, int(lo4), out.ctypes.data_as(ctypes.c_void_p),
                             flg.ctypes.data_as(ctypes.c_void_p), dist.ctypes.data_as(ctypes.c_void_p), None, None)
    return out, flg.astype(bool), dist


@pytest.mark.parametrize("mode", ["dense", "sparse", "dc_only", "ties", "large"])
def test_packed_idct_never_accepts_a_wrong_sample(hs, mode):
    """The FFMA2-shaped fast path of csrc/bj_pixels_mma.cu (DC peeling + packed IDCT + magic-add rounding)."""
    rng = np.random.default_rng((hash(mode) & 0xFFFF) + 1)
    b = _random_blocks(rng, 1500, mode)
    fast, flagged, _ = _idct_packed(hs, b)
    want = _oracle_idct(b)
    ok = ~flagged
    assert np.array_equal(fast[ok], want[ok])
    if mode == "ties":
        assert flagged.all()
    if mode == "sparse":
        assert flagged.mean() < 0.2
    if mode == "large":
        assert flagged.all()          # products beyond int16 (:869 wraps) always go to the exact path


def test_packed_idct_low_frequency_variant(hs):
    """Blocks confined to the 4x4 low-frequency corner: the LO4 variant must agree with the oracle (where accepted)
    and with the dense variant's decisions on the same blocks."""
    rng = np.random.default_rng(77)
    n = 3000
    b = np.zeros((n, 8, 8), np.int32)
    b[:, :4, :4] = rng.integers(-200, 200, (n, 4, 4)) * (rng.random((n, 4, 4)) < 0.5)
    b[:, 0, 0] = rng.integers(-1024, 1024, n)
    b = b.reshape(n, 64)
    lo, flo, _ = _idct_packed(hs, b, lo4=True)
    want = _oracle_idct(b)
    assert np.array_equal(lo[~flo], want[~flo])
    de, fde, _ = _idct_packed(hs, b, lo4=False)
    assert np.array_equal(de[~fde], want[~fde])
    assert flo.mean() < 0.1


def test_dc_peeling_lowers_the_exact_path_rate(hs):
    """Same blocks through the scalar fast path (tie threshold grows with |DC|) and the packed one (DC peeled)."""
    rng = np.random.default_rng(9)
    n = 20000
    b = np.zeros((n, 64), np.int32)
    mask = rng.random((n, 64)) < 0.2
    b[mask] = rng.integers(-60, 60, mask.sum())
    b[:, 0] = rng.integers(-1000, 1000, n)
    _, f_old, _ = _idct(hs, b)
    _, f_new, _ = _idct_packed(hs, b)
    assert f_new.mean() < 0.8 * f_old.mean()


def test_idct_fast_path_on_real_coefficients(hs):
    """Dequantised blocks of real files: fast path result == oracle wherever it is accepted."""
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.layout import ZIGZAG_UV
    nat_of_zz = np.array([v * 8 + u for (u, v) in ZIGZAG_UV])
    for name in ["base_120x88_ss2_q95", "base_256x128_ss2_opt", "base_sat_96x64_ss0", "prog_200x120_ss0"]:
        data = (GOLDEN / "cases" / f"{name}.jpg").read_bytes()
        p = parse_jpeg(data)
        r = oracle.decode(data)
        for ci, c in enumerate(p.components):
            q = p.qtables[c.tq].astype(np.int32)
            zz = r.coef[ci].reshape(-1, 64).astype(np.int32)
            deq = (zz * q).astype(np.int16).astype(np.int32)
            nat = np.zeros_like(deq)
            nat[:, nat_of_zz] = deq
            fast, flagged, _ = _idct(hs, nat)
            want = _oracle_idct(nat)
            assert np.array_equal(fast[~flagged], want[~flagged]), name
            fast2, flagged2, _ = _idct_packed(hs, nat)
            assert np.array_equal(fast2[~flagged2], want[~flagged2]), name


@pytest.mark.parametrize("kind,key", [((2, 2), "w_8x8_16x16"), ((2, 1), "w_8x8_16x8"), ((1, 2), "w_8x8_8x16")])
def test_interpolation_weights_match_griddata(hs, kind, key):
    rh, rv = kind
    w = np.zeros((256, 4), np.int32)
    cell = np.zeros((256, 2), np.int32)
    hs.hs_weights(rh, rv, w.ctypes.data_as(ctypes.c_void_p), cell.ctypes.data_as(ctypes.c_void_p))
    gold = np.load(GOLDEN / "upsample_weights.npz")[key]   # [i, j, a, b] weights * 15
    for b in range(8 * rv):
        for a in range(8 * rh):
            i, j = cell[b * 16 + a]
            dense = np.zeros((8, 8), np.int64)
            for (di, dj, k) in ((0, 0, 0), (1, 0, 1), (0, 1, 2), (1, 1, 3)):
                if w[b * 16 + a, k]:
                    dense[min(i + di, 7), min(j + dj, 7)] += w[b * 16 + a, k]
            assert np.array_equal(dense, gold[:, :, a, b]), (a, b)
            assert w[b * 16 + a].sum() == 15


def test_div15_round_exact(hs):
    n = np.arange(-15 * 4000, 15 * 4000 + 1)
    want = np.floor_divide(2 * n + 15, 30)
    got = np.array([hs.hs_div15(float(x)) for x in n[::7]])
    assert np.array_equal(got, want[::7])


def test_colour_integer_aware_path_never_accepts_a_wrong_pixel(hs):
    """The specialised kernel's colour path: exhaustive over Cb, Cr in [-130, 390] and a few Y."""
    cb, cr = np.meshgrid(np.arange(-130, 391), np.arange(-130, 391), indexing="ij")
    for Y in (-7, 0, 1, 100, 128, 255, 301):
        ycc = np.stack([np.full(cb.size, Y), cb.ravel(), cr.ravel()], -1).astype(np.int16)
        n = ycc.shape[0]
        rgb = np.empty((n, 3), np.uint8)
        slow = np.empty(n, np.uint8)
        hs.hs_color2(ycc.ctypes.data_as(ctypes.c_void_p), n, rgb.ctypes.data_as(ctypes.c_void_p),
                     slow.ctypes.data_as(ctypes.c_void_p))
        Yf, Cb, Cr = (ycc[:, k].astype(np.float64) for k in range(3))
        R = Yf + 1.402 * (Cr - 128.0)
        G = Yf - 0.34414 * (Cb - 128.0) - 0.71414 * (Cr - 128.0)
        B = Yf + 1.772 * (Cb - 128.0)
        want = np.round(np.clip(np.stack((R, G, B), -1), 0.0, 255.0)).astype(np.uint8)
        ok = slow == 0
        assert np.array_equal(rgb[ok], want[ok]), Y
        inside = (np.abs(Cb - 128) < 128) & (np.abs(Cr - 128) < 128)
        assert slow[inside].mean() < 0.02


def test_colour_fast_path_never_accepts_a_wrong_pixel(hs):
    rng = np.random.default_rng(5)
    n = 400000
    ycc = np.empty((n, 3), np.int16)
    ycc[:, 0] = rng.integers(-40, 300, n)
    ycc[:, 1] = rng.integers(-60, 320, n)
    ycc[:, 2] = rng.integers(-60, 320, n)
    # exact decimal ties of the colour matrix: Cr-128 = +-250, Cb-128 = +-125
    ycc[:2000, 2] = 128 + 250
    ycc[2000:4000, 1] = 128 + 125
    ycc[4000:6000, 1] = 3
    rgb = np.empty((n, 3), np.uint8)
    tie = np.empty(n, np.uint8)
    hs.hs_color(ycc.ctypes.data_as(ctypes.c_void_p), n, rgb.ctypes.data_as(ctypes.c_void_p), tie.ctypes.data_as(ctypes.c_void_p))
    Y, Cb, Cr = (ycc[:, k].astype(np.float64) for k in range(3))
    R = Y + 1.402 * (Cr - 128.0)
    G = Y - 0.34414 * (Cb - 128.0) - 0.71414 * (Cr - 128.0)
    B = Y + 1.772 * (Cb - 128.0)
    want = np.round(np.clip(np.stack((R, G, B), -1), 0.0, 255.0)).astype(np.uint8)   # jpeg_decoder.py:1693-1700
    ok = tie == 0
    assert np.array_equal(rgb[ok], want[ok])
    assert tie.mean() < 0.05
