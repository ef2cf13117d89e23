#!/usr/bin/env python3
"""Error probe of the fp32 fast IDCT (csrc/bj_pixel_math.cuh, idct8x8_packed + DC peeling) against its bound.

Runs the host build of the device header (tests/hostsim, bit-identical arithmetic) over many blocks of several
families -- random dense / sparse, single coefficients, energy piled on the worst-weighted positions, sign
patterns chosen to align the rounding errors -- evaluates the exact value of every sample in float64 with the
reference's own table (jpeg_decoder.py:1541-1553) and prints the largest observed |fp32 - exact| / T, where
T = BJ_IDCT_ERR_U * sum w|x| + BJ_IDCT_ERR_ABS is the tie threshold the kernel uses.  The bit-exactness argument of
the fast path needs this ratio to stay below 1.   usage: python tools/errprobe.py [blocks per family]
"""
import ctypes
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def families(rng, n):
    z = np.zeros((n, 8, 8), np.int32)
    out = {}
    a = rng.integers(-600, 600, (n, 8, 8)); out["dense"] = a
    a = z.copy(); m = rng.random((n, 8, 8)) < 0.2; a[m] = rng.integers(-300, 300, m.sum()); a[:, 0, 0] = rng.integers(-1024, 1024, n); out["sparse"] = a
    a = z.copy(); v, u = rng.integers(0, 8, n), rng.integers(0, 8, n); a[np.arange(n), v, u] = rng.integers(-2000, 2000, n); out["single"] = a
    a = z.copy(); a[:, 1, 1] = rng.integers(-3000, 3000, n); a[:, 1, 3] = rng.integers(-3000, 3000, n); a[:, 3, 1] = rng.integers(-3000, 3000, n); out["odd_low"] = a
    a = rng.integers(0, 2, (n, 8, 8)) * 2 - 1; a = a * rng.integers(100, 120, (n, 8, 8)); out["signs"] = a
    a = z.copy(); a[:, :4, :4] = rng.integers(-500, 500, (n, 4, 4)); out["lo4"] = a
    a = rng.integers(-4000, 4000, (n, 8, 8)) * (rng.random((n, 8, 8)) < 0.5); out["large"] = a
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    from hostsim import build
    import oracle
    hs = build("pixel_hostsim")
    tab = oracle.idct_table()                      # [x][y][u][v]
    rng = np.random.default_rng(1234)
    worst = 0.0
    for name, blk in families(rng, n).items():      # blk[n, v, u]
        for lo4 in ((False, True) if name == "lo4" else (False,)):
            b = np.ascontiguousarray(blk.reshape(n, 64), dtype=np.int32)
            out = np.empty((n, 64), np.int16); flg = np.empty(n, np.uint8); dist = np.empty(n, np.float32)
            thr = np.empty(n, np.float32); raw = np.empty((n, 64), np.float64)
            hs.hs_idct_blocks_packed(b.ctypes.data_as(ctypes.c_void_p), n, int(lo4), out.ctypes.data_as(ctypes.c_void_p),
                                     flg.ctypes.data_as(ctypes.c_void_p), dist.ctypes.data_as(ctypes.c_void_p),
                                     thr.ctypes.data_as(ctypes.c_void_p), raw.ctypes.data_as(ctypes.c_void_p))
            # exact[y, x] = sum_{u,v} blk[v, u] * tab[x, y, u, v]
            exact = np.einsum("nvu,xyuv->nyx", blk.astype(np.float64), tab).reshape(n, 64)
            err = np.abs(raw - exact).max(axis=1)
            ratio = err / thr
            worst = max(worst, float(ratio.max()))
            s = np.abs(blk.reshape(n, 64)[:, 1:]).sum(axis=1)
            print(f"{name:8s} lo4={int(lo4)}  max err/T {ratio.max():.3f}  mean {ratio.mean():.3f}  "
                  f"max err/sum|AC| {np.max(err / np.maximum(s, 1)):.2e}  flagged {flg.mean() * 100:.2f} %")
    print(f"worst err/T over all families: {worst:.3f}  ({'OK' if worst < 1 else 'BOUND VIOLATED'})")
    return 0 if worst < 1 else 1


if __name__ == "__main__":
    sys.exit(main())
