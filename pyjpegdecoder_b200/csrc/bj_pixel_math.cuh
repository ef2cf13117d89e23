// bj_pixel_math.cuh -- per-thread arithmetic of the pixel stages (IDCT fast path, chroma
// interpolation weights, colour conversion).  Every function is __host__ __device__ and uses only
// explicitly fused/unfused operations (fmaf, plain + and *).  The CUDA build passes -fmad=false so
// nvcc never contracts a*b+c on its own; the host-side unit tests in tests/hostsim (g++
// -ffp-contract=off) therefore exercise bit-identical arithmetic.
//
// Reference being replaced (tbpaolini/PyJpegDecoder, jpeg_decoder.py):
//   InverseDCT            :1535-1573   out[x,y] = sum_uv block[u,v] * 0.25*Cu*Cv*cos(..)*cos(..) in fp64,
//                                      np.round (half-even), +128, no clamp
//   ResizeGrid            :1580-1626   griddata linear interpolation 8 -> 16 (align corners)
//   YCbCr_to_RGB          :1683-1700   fp64 matrix, clip, np.round
// The reference rounds fp64 values; the fast paths below work in fp32 and report how close each
// value is to a rounding tie so the caller can fall back to the exact fp64 evaluation.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BJ_HD __host__ __device__ __forceinline__
#define BJ_HDM __host__ __device__ __forceinline__  // member functions
#else
#define BJ_HD static inline
#define BJ_HDM inline
#endif

namespace bj {

// zig-zag index -> natural index v*8+u (v = vertical frequency).  Same permutation as the
// reference's zagzig table (:1672-1681), which lists (x, y) = (u, v).
#define BJ_ZZ_NATURAL                                                                               \
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, \
        13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58,   \
        59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63

// cos(k*pi/16) / 2
#define BJ_C1 0.49039264020161522f
#define BJ_C2 0.46193976625564337f
#define BJ_C3 0.41573480615127262f
#define BJ_C4 0.35355339059327379f
#define BJ_C5 0.27778511650980114f
#define BJ_C6 0.19134171618254489f
#define BJ_C7 0.09754516100806417f

// 1.5 * 2^23: adding it to |x| < 2^22 rounds x to the nearest integer (ties to even) in fp32 and
// leaves that integer in the low mantissa bits.
#define BJ_MAGIC 12582912.0f
#define BJ_MAGIC_BITS 0x4B400000

// One 8-point inverse DCT (even/odd decomposition, 34 flops), in place over 8 values with stride 1.
// out[n] = sum_k a_k X[k] cos((2n+1) k pi / 16),  a_0 = 1/(2 sqrt 2), a_k = 1/2.
BJ_HD void idct8(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6, float& x7) {
    float s04 = x0 + x4, d04 = x0 - x4;
    float f0 = fmaf(BJ_C6, x6, BJ_C2 * x2);
    float f1 = fmaf(-BJ_C2, x6, BJ_C6 * x2);
    float a0 = fmaf(s04, BJ_C4, f0), a3 = fmaf(s04, BJ_C4, -f0);
    float a1 = fmaf(d04, BJ_C4, f1), a2 = fmaf(d04, BJ_C4, -f1);
    float o0 = fmaf(BJ_C7, x7, fmaf(BJ_C5, x5, fmaf(BJ_C3, x3, BJ_C1 * x1)));
    float o1 = fmaf(-BJ_C5, x7, fmaf(-BJ_C1, x5, fmaf(-BJ_C7, x3, BJ_C3 * x1)));
    float o2 = fmaf(BJ_C3, x7, fmaf(BJ_C7, x5, fmaf(-BJ_C1, x3, BJ_C5 * x1)));
    float o3 = fmaf(-BJ_C1, x7, fmaf(BJ_C3, x5, fmaf(-BJ_C5, x3, BJ_C7 * x1)));
    x0 = a0 + o0; x7 = a0 - o0;
    x1 = a1 + o1; x6 = a1 - o1;
    x2 = a2 + o2; x5 = a2 - o2;
    x3 = a3 + o3; x4 = a3 - o3;
}

// fp32 error model of idct8x8_fast:
//   |fast - exact| <= BJ_IDCT_ERR_REL * sum|AC coef| + BJ_IDCT_ERR_DC * |DC| + BJ_IDCT_ERR_ABS.
// Worst-case analysis of the 34-flop butterfly: at most 7 roundings of magnitude <= 0.5 * sum|x| per
// 1-D pass (3.5 * 2^-24 per unit of sum|x|), two passes -> 2.1e-7 * sum|coef|; the DC term only
// passes through 3 roundings per pass at gain 0.354 -> 7 * 2^-24 * 0.125 |DC| = 5.2e-8 |DC|.
// Measured worst case over 1.1e9 samples of random/sparse/DC-heavy blocks: 5.3e-8 * sum|coef|
// (tests/hostsim, tools/errprobe).
#define BJ_IDCT_ERR_REL 2.4e-7f
#define BJ_IDCT_ERR_DC 5.5e-8f
#define BJ_IDCT_ERR_ABS 1.0e-6f

// 2-D IDCT of one block held as f[v*8+u] (natural order); result f[y*8+x].
BJ_HD void idct8x8_fast(float* f) {
#pragma unroll
    for (int u = 0; u < 8; u++)
        idct8(f[u], f[8 + u], f[16 + u], f[24 + u], f[32 + u], f[40 + u], f[48 + u], f[56 + u]);
#pragma unroll
    for (int y = 0; y < 8; y++)
        idct8(f[8 * y], f[8 * y + 1], f[8 * y + 2], f[8 * y + 3], f[8 * y + 4], f[8 * y + 5], f[8 * y + 6], f[8 * y + 7]);
}

// Round v (|v| < 2^22) to nearest-even; returns the rounded value as float and the distance of v
// from the nearest rounding tie (0 = exactly on a tie, 0.5 = exactly an integer).
BJ_HD float round_tie(float v, float& tie_dist) {
    float w = v + BJ_MAGIC;
    float r = w - BJ_MAGIC;
    tie_dist = 0.5f - fabsf(v - r);
    return r;
}

// ---- chroma interpolation (ResizeGrid, :1588-1626) ---------------------------------------------
// Output sample a (0..15) of a 16-long axis reads source cell i = floor(7a/15) with fraction
// s/15, s = 7a mod 15; a = 15 is the right edge of cell 6 (s = 15).
constexpr BJ_HD void up_cell(int a, int& i, int& s) {
    if (a == 15) { i = 6; s = 15; }
    else { i = (7 * a) / 15; s = (7 * a) % 15; }
}

// Diagonal map of Qhull's Delaunay triangulation of the 8x8 grid as scipy.interpolate.griddata
// produces it (bit 7*i+j set: cell (i,j) split along (i,j)-(i+1,j+1)); scipy 1.18.1, pinned by
// tests/test_oracle.py::test_upsample_matches_live_scipy.
#define BJ_DIAG_MAP 0x155555594e4a5ull

// Integer weights (fifteenths) of the four corners P(i,j), P(i+1,j), P(i,j+1), P(i+1,j+1) for
// fractions (s,t); exactly one of them is zero (three-tap barycentric interpolation).
constexpr BJ_HD void up_weights_2d(int i, int j, int s, int t, int& w00, int& w10, int& w01, int& w11) {
    bool diag = (BJ_DIAG_MAP >> (7 * i + j)) & 1ull;
    if (diag) {
        if (s >= t) { w00 = 15 - s; w10 = s - t; w01 = 0; w11 = t; }
        else        { w00 = 15 - t; w10 = 0; w01 = t - s; w11 = s; }
    } else {
        if (s + t <= 15) { w00 = 15 - s - t; w10 = s; w01 = t; w11 = 0; }
        else             { w00 = 0; w10 = 15 - t; w01 = 15 - s; w11 = s + t - 15; }
    }
}

// N/15 rounded to nearest for integer N held exactly in a float (|N| < 2^22).  N/15 is never
// within 1/30 of a tie (15 is odd) and the fp32 error is < 2^-9 for |N/15| < 2^15, so the magic-add
// rounding is exact: equals floor((2N+15)/30).
BJ_HD float div15_round(float n) {
    float w = fmaf(n, 1.0f / 15.0f, BJ_MAGIC);
    return w - BJ_MAGIC;
}

// ---- colour, integer-aware form (used by the layout-specialised kernel) ---------------------------
// Y, Cb, Cr are integers (:1573, :1626), so R - Y = 1.402 (Cr-128), B - Y = 1.772 (Cb-128) and
// G - Y = -0.34414 (Cb-128) - 0.71414 (Cr-128) take values on the grids k/500, k/250 and k/50000:
//   * R - Y is a rounding tie only for Cr-128 = 250 (mod 500), otherwise at least 0.002 away;
//   * B - Y is a tie only for Cb-128 = 125 (mod 250), otherwise at least 0.004 away;
//   * G - Y can come within 2e-5 of a tie, or hit it exactly.
// With |Cb-128|, |Cr-128| < 250 the fp32 offsets are within 4e-5 (R, B) and 3e-5 (G) of the exact
// values, so R and B round correctly unless |Cb-128| == 125, and G is safe when it is farther than
// BJ_G_ERR from a tie.  Everything else goes to the fp64 evaluation.  clip-then-round (:1698-1700)
// equals clamp(Y + round(offset), 0, 255) away from ties.
#define BJ_CHROMA_GUARD 250.0f
#define BJ_G_ERR 3.0e-5f

// ---- colour (YCbCr_to_RGB, :1683-1700) ---------------------------------------------------------
// fp32 fast path.  err bound: each channel is at most two fused multiply-adds of magnitudes below
// |Y| + |cb| + |cr| plus the representation error of the constants.
#define BJ_COLOR_ERR_REL 3.0e-7f
#define BJ_COLOR_ERR_ABS 1.0e-6f

BJ_HD void ycc_to_rgb_fast(float Y, float Cb, float Cr, float& R, float& G, float& B, float& err) {
    float cb = Cb - 128.0f, cr = Cr - 128.0f;
    R = fmaf(1.402f, cr, Y);
    G = fmaf(-0.71414f, cr, fmaf(-0.34414f, cb, Y));
    B = fmaf(1.772f, cb, Y);
    err = fmaf(fabsf(Y) + fabsf(cb) + fabsf(cr), BJ_COLOR_ERR_REL, BJ_COLOR_ERR_ABS);
}

}  // namespace bj
