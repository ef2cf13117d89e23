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

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2 = PTX fma/add/mul.rn.f32x2, sm_100+) -------------
// Two independent IEEE fp32 operations per instruction; every lane result is bit-identical to the scalar
// fmaf / + / * of the host build, so tests/hostsim exercises the very same arithmetic.
#if defined(__CUDACC__)
typedef float2 F2;
#else
struct F2 { float x, y; };
#endif
BJ_HD F2 f2(float a, float b) { F2 r; r.x = a; r.y = b; return r; }
BJ_HD F2 f2s(float c) { return f2(c, c); }
BJ_HD F2 f2fma(F2 a, F2 b, F2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
BJ_HD F2 f2add(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return f2(a.x + b.x, a.y + b.y);
#endif
}
BJ_HD F2 f2mul(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return f2(a.x * b.x, a.y * b.y);
#endif
}
BJ_HD F2 f2sub(F2 a, F2 b) { return f2fma(b, f2s(-1.0f), a); }  // a - b with one rounding

// Even (a0..a3) and odd (o0..o3) halves of the 8-point inverse DCT of TWO data sets at once;
// out[n] = a_n + o_n, out[7-n] = a_n - o_n.  At most 6 roundings on any path (the error model above allows 7).
BJ_HD void idct8_halves2(F2 x0, F2 x1, F2 x2, F2 x3, F2 x4, F2 x5, F2 x6, F2 x7, F2 (&a)[4], F2 (&o)[4]) {
    const F2 s04 = f2add(x0, x4), d04 = f2sub(x0, x4);
    const F2 e0 = f2fma(f2s(BJ_C6), x6, f2mul(f2s(BJ_C2), x2));
    const F2 e1 = f2fma(f2s(-BJ_C2), x6, f2mul(f2s(BJ_C6), x2));
    const F2 t0 = f2mul(s04, f2s(BJ_C4)), t1 = f2mul(d04, f2s(BJ_C4));
    a[0] = f2add(t0, e0); a[3] = f2sub(t0, e0);
    a[1] = f2add(t1, e1); a[2] = f2sub(t1, e1);
    o[0] = f2fma(f2s(BJ_C7), x7, f2fma(f2s(BJ_C5), x5, f2fma(f2s(BJ_C3), x3, f2mul(f2s(BJ_C1), x1))));
    o[1] = f2fma(f2s(-BJ_C5), x7, f2fma(f2s(-BJ_C1), x5, f2fma(f2s(-BJ_C7), x3, f2mul(f2s(BJ_C3), x1))));
    o[2] = f2fma(f2s(BJ_C3), x7, f2fma(f2s(BJ_C7), x5, f2fma(f2s(-BJ_C1), x3, f2mul(f2s(BJ_C5), x1))));
    o[3] = f2fma(f2s(-BJ_C1), x7, f2fma(f2s(BJ_C3), x5, f2fma(f2s(-BJ_C5), x3, f2mul(f2s(BJ_C7), x1))));
}
// The same with x4..x7 == 0 (blocks whose non-zero coefficients lie in the 4x4 low-frequency corner).
BJ_HD void idct8_halves2_lo4(F2 x0, F2 x1, F2 x2, F2 x3, F2 (&a)[4], F2 (&o)[4]) {
    const F2 t = f2mul(f2s(BJ_C4), x0);
    a[0] = f2fma(f2s(BJ_C2), x2, t); a[3] = f2fma(f2s(-BJ_C2), x2, t);
    a[1] = f2fma(f2s(BJ_C6), x2, t); a[2] = f2fma(f2s(-BJ_C6), x2, t);
    o[0] = f2fma(f2s(BJ_C3), x3, f2mul(f2s(BJ_C1), x1));
    o[1] = f2fma(f2s(-BJ_C7), x3, f2mul(f2s(BJ_C3), x1));
    o[2] = f2fma(f2s(-BJ_C1), x3, f2mul(f2s(BJ_C5), x1));
    o[3] = f2fma(f2s(-BJ_C5), x3, f2mul(f2s(BJ_C7), x1));
}

// 2-D IDCT of one block in packed form.  In: P[v][h] = (f[v][2h], f[v][2h+1]) (natural order, v = vertical
// frequency).  Out: R[yp][x] = (out[2yp][x], out[2yp+1][x]).  The column pass runs on column pairs; its last
// butterfly stage is scalar so that it can write the row-pair layout the row pass wants (no register moves).
// LO4: only P[0..3][0..1] are read (4x4 low-frequency corner, the rest is zero).
template <bool LO4>
BJ_HD void idct8x8_packed(const F2 (&P)[8][4], F2 (&R)[4][8]) {
    F2 Q[4][8];  // Q[yp][u] = (g[2yp][u], g[2yp+1][u])
#pragma unroll
    for (int h = 0; h < (LO4 ? 2 : 4); h++) {
        F2 a[4], o[4];
        if (LO4) idct8_halves2_lo4(P[0][h], P[1][h], P[2][h], P[3][h], a, o);
        else idct8_halves2(P[0][h], P[1][h], P[2][h], P[3][h], P[4][h], P[5][h], P[6][h], P[7][h], a, o);
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const float lo0 = a[n].x + o[n].x, hi0 = a[n].x - o[n].x;  // column 2h: rows n and 7-n
            const float lo1 = a[n].y + o[n].y, hi1 = a[n].y - o[n].y;  // column 2h+1
            if (n & 1) { Q[n >> 1][2 * h].y = lo0; Q[n >> 1][2 * h + 1].y = lo1; }
            else       { Q[n >> 1][2 * h].x = lo0; Q[n >> 1][2 * h + 1].x = lo1; }
            if ((7 - n) & 1) { Q[(7 - n) >> 1][2 * h].y = hi0; Q[(7 - n) >> 1][2 * h + 1].y = hi1; }
            else             { Q[(7 - n) >> 1][2 * h].x = hi0; Q[(7 - n) >> 1][2 * h + 1].x = hi1; }
        }
    }
#pragma unroll
    for (int yp = 0; yp < 4; yp++) {
        F2 a[4], o[4];
        if (LO4) idct8_halves2_lo4(Q[yp][0], Q[yp][1], Q[yp][2], Q[yp][3], a, o);
        else idct8_halves2(Q[yp][0], Q[yp][1], Q[yp][2], Q[yp][3], Q[yp][4], Q[yp][5], Q[yp][6], Q[yp][7], a, o);
#pragma unroll
        for (int n = 0; n < 4; n++) {
            R[yp][n] = f2add(a[n], o[n]);
            R[yp][7 - n] = f2sub(a[n], o[n]);
        }
    }
}

// ---- fp32 error model of idct8x8_packed (first-order rounding analysis, u = 2^-24) --------------------------
// One 1-D pass (idct8_halves2 + the final a +- o), inputs x_k with incoming absolute errors e_k:
//     |err(out_n)| <= u * sum_k G_k |x_k| + sum_k B_k e_k
// B_k = largest basis magnitude seen by input k (C4 for k = 0, 4; C2 for k = 2, 6; C1 for odd k).  G_k counts, along
// the worst path from input k to any output, every rounding (each at most u times the magnitude bound of its
// result) and the representation error of every fp32 constant (relative u):
//   k = 0, 4:  t = (x0 +- x4) * C4 [constant + rounding: 2, plus 1 for the sum in the second pass, whose inputs
//              are not integers], a = t +- e [1], out = a +- o [1]                      -> C4 * (4 + r) = 1.415 / 1.768
//   k = 2:     C2*x2 [2], fma with C6*x6 [1], a0 [1], out [1] (x C2)                     -> 5 * C2      = 2.310
//   k = 6:     -C2*x6 inside the fma of e1 [2], a1 [1], out [1] (x C2)                    -> 4 * C2      = 1.848
//   k = 1,3,5,7: chain of 1 product + 3 fmas: input k sees (5, 4, 3, 2) roundings/constants, then out [1]:
//              6*C1 = 2.943, 5*C1 = 2.452, 4*C1 = 1.962, 3*C1 = 1.472 (worst output row each)
// Two passes (columns, then rows): |fast - exact| <= u * sum_{v,u} |x_vu| * (G2_u B_v + B_u G1_v).  The low-frequency
// variant (idct8_halves2_lo4) has fewer operations on every path, so the same weights bound it.  The kernel
// accumulates S_w = sum w_vu |x_vu| with one FFMA per coefficient (the weight is an immediate) and uses
// T = BJ_IDCT_ERR_U * S_w + BJ_IDCT_ERR_ABS; tests/test_pixel_math.py and tools/errprobe.py measure the actual error
// against this bound.  Every weight is >= BJ_IDCT_W_MIN, so S_w / BJ_IDCT_W_MIN bounds the plain sum |x|.
#define BJ_IDCT_ERR_U 6.1e-8f   /* 2^-24 * 1.023: first-order bound + second-order slack */
#define BJ_IDCT_W_MIN 1.12f
constexpr BJ_HD float idct_err_weight(int v, int u) {
    constexpr float G1[8] = {1.416f, 2.945f, 2.312f, 2.454f, 1.416f, 1.964f, 1.850f, 1.473f};
    constexpr float G2[8] = {1.770f, 2.945f, 2.312f, 2.454f, 1.770f, 1.964f, 1.850f, 1.473f};
    constexpr float B[8] = {0.3536f, 0.4904f, 0.4620f, 0.4904f, 0.3536f, 0.4904f, 0.4620f, 0.4904f};
    return G2[u] * B[v] + B[u] * G1[v];
}

// DC peeling: the DC coefficient adds the constant DC/8 to every sample, and DC/8 = I + r/8 with I = floor(DC/8),
// r = DC mod 8.  The fast path transforms r instead of DC and adds I in the integer domain (it rides on the
// rounding constant), so the fp32 error of the DC path scales with r < 8 instead of |DC| (hundreds): the tie
// threshold -- and with it the number of blocks that need the exact recompute -- drops by about 40 %.
BJ_HD float dc_peel(float& dc) {
    const float i = floorf(dc * 0.125f);  // exact: dc is an integer below 2^24
    dc = fmaf(-8.0f, i, dc);              // exact, in [0, 8)
    return i;
}

// Packed IDCT + rounding + tie distance of one block.  P: dequantised coefficients (DC already peeled),
// shift = 1.5 * 2^23 + 128 + I.  W[yp][x] = R + shift: the low 16 bits of each float are the int16 samples
// (rint(v) + 128 + I) of rows 2yp (.x) and 2yp+1 (.y); maxd = max |v - rint(v)| over the block.
template <bool LO4>
BJ_HD void idct8x8_round_packed(const F2 (&P)[8][4], float shift, F2 (&W)[4][8], float& maxd) {
    F2 R[4][8];
    idct8x8_packed<LO4>(P, R);
    maxd = 0.f;
#pragma unroll
    for (int yp = 0; yp < 4; yp++) {
#pragma unroll
        for (int x = 0; x < 8; x++) {
            W[yp][x] = f2add(R[yp][x], f2s(shift));
            const F2 nr = f2fma(W[yp][x], f2s(-1.0f), f2s(shift));  // -(rint(v)), exact
            const F2 d = f2add(R[yp][x], nr);
            maxd = fmaxf(fmaxf(maxd, fabsf(d.x)), fabsf(d.y));
        }
    }
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
