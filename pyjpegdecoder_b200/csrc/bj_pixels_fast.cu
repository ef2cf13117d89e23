// bj_pixels_fast.cu -- layout-specialised fused pixel kernel (sm_100a): de-zigzag + dequantise +
// 8x8 IDCT + level shift + chroma upsampling + YCbCr->RGB + clamp, coefficients in, RGB bytes out.
//
// Same arithmetic contract as the generic kernel in bj_pixels.cu (bit-exact with the reference's
// fp64 path, jpeg_decoder.py:869-891, :1306-1366, :1368-1386, :1683-1700) but the sampling layout is
// a template parameter (4:2:0, 4:2:2, 4:4:0, 4:4:4, greyscale), so every index computation, division
// and branch on the geometry folds at compile time.  What changes versus the generic kernel:
//   * luma samples stay int16 in shared memory (128 B per block); only chroma is kept as fp32
//     because it is interpolated.  R/G/B = clamp(Y + round(offset(Cb,Cr))): the rounding tie test
//     only involves the chroma offsets (Y is an integer), see bj_pixel_math.cuh;
//   * add + clamp is one VIADDMNMX (__viaddmin_s32_relu);
//   * the exact fp64 recompute of a flagged block iterates over its non-zero coefficients only
//     (ballot masks), in numpy's pairwise order;
//   * 60 KB of shared memory per CTA -> 3 CTAs (18 warps) per SM.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200jpeg.h"
#include "bj_pixel_math.cuh"

extern "C" bj_status bj_set_cuda_error(cudaError_t e, const char* where);

namespace {

constexpr int kThreads = 192;
constexpr int kWarps = kThreads / 32;
constexpr int kTileABytes = 192 * 128 + 512;

template <int HMAX_, int VMAX_, int NCOMP_>
struct Lay {
    static constexpr int HMAX = HMAX_, VMAX = VMAX_, NCOMP = NCOMP_;
    static constexpr int NY = HMAX * VMAX;
    static constexpr int BPM = NY + (NCOMP == 3 ? 2 : 0);
    static constexpr int MCU_W = 8 * HMAX, MCU_H = 8 * VMAX;
    static constexpr bool UPS = (NCOMP == 3) && (NY > 1);
    static constexpr int MAXM = 192 / BPM;
    static constexpr int MCU_B = NY * 128 + (NCOMP == 3 ? 512 : 0);  // tile B bytes per MCU
    static constexpr int CH = NCOMP == 3 ? 3 : 1;
    static constexpr int TILE_B = MAXM * MCU_B;
    static constexpr int W_BYTES = UPS ? 256 * 16 : 0;
    static constexpr int SMEM = kTileABytes + TILE_B + W_BYTES + 3 * 128 + 64 + 64;
};

__constant__ uint8_t c_zz_nat2[64] = {BJ_ZZ_NATURAL};

__device__ __forceinline__ int swzA(int blk, int chunk) { return blk * 128 + ((chunk ^ (blk & 7)) << 4); }
__device__ __forceinline__ float int_to_float_magic(int v) { return __int_as_float(v + BJ_MAGIC_BITS) - BJ_MAGIC; }

template <class L>
struct Tiles {
    unsigned char* a;   // coefficients / RGB staging
    unsigned char* b;   // samples
    float4* w;          // interpolation weights [b*16 + a]
    int16_t* qt;        // [3][64]
    uint8_t* nat_zz;    // u*8+v -> zig-zag index
    // luma block ys of MCU m, row y: 8 int16
    __device__ __forceinline__ unsigned char* yrow(int m, int ys, int y) const {
        return b + m * L::MCU_B + ys * 128 + ((y ^ ((m + ys) & 7)) << 4);
    }
    // chroma block k (0 = Cb, 1 = Cr) of MCU m, 16-byte chunk c (row y = chunks 2y, 2y+1): 4 floats
    __device__ __forceinline__ float* cchunk(int m, int k, int c) const {
        return reinterpret_cast<float*>(b + m * L::MCU_B + L::NY * 128 + k * 256 + (((c ^ (m + L::NY + k)) & 15) << 4));
    }
};

// ---- exact recompute of one block by a whole warp (InverseDCT.__call__, :1561-1573) ---------------
// Lane l owns output samples s = l and l + 32 (s = x*8 + y).  Accumulator j = v collects the products
// of u = 0..7 in order, then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)): numpy's pairwise sum of the 64
// products in C order [u][v]; zero coefficients only add +-0.0 and are skipped.
template <class L>
__device__ void recompute_block_exact(const Tiles<L>& t, int blk, int m, int slot, const double* __restrict__ tabT, int lane) {
    const int comp = slot < L::NY ? 0 : slot - L::NY + 1;
    const int16_t* qt = t.qt + comp * 64;
    // natural-order non-zero masks: lane l looks at n = l and n = l + 32
    int k0 = t.nat_zz[lane], k1 = t.nat_zz[lane + 32];
    int c0 = *reinterpret_cast<const int16_t*>(t.a + swzA(blk, k0 >> 3) + ((k0 & 7) << 1));
    int c1 = *reinterpret_cast<const int16_t*>(t.a + swzA(blk, k1 >> 3) + ((k1 & 7) << 1));
    int p0 = (int16_t)(c0 * (int)qt[k0]), p1 = (int16_t)(c1 * (int)qt[k1]);  // int16 product wraps (:869)
    const unsigned nz_lo = __ballot_sync(0xffffffffu, p0 != 0), nz_hi = __ballot_sync(0xffffffffu, p1 != 0);
    double r0[8], r1[8];
#pragma unroll
    for (int v = 0; v < 8; v++) r0[v] = r1[v] = 0.0;
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int v = 0; v < 8; v++) {
            const int n = u * 8 + v;
            const bool nz = (((n < 32) ? nz_lo : nz_hi) >> (n & 31)) & 1u;  // warp-uniform
            if (nz) {
                int prod = __shfl_sync(0xffffffffu, (n < 32) ? p0 : p1, n & 31);
                double p = (double)prod;
                const double* tt = tabT + n * 64;
                r0[v] = __dadd_rn(r0[v], __dmul_rn(p, tt[lane]));
                r1[v] = __dadd_rn(r1[v], __dmul_rn(p, tt[lane + 32]));
            }
        }
    }
    double s0 = __dadd_rn(__dadd_rn(__dadd_rn(r0[0], r0[1]), __dadd_rn(r0[2], r0[3])),
                          __dadd_rn(__dadd_rn(r0[4], r0[5]), __dadd_rn(r0[6], r0[7])));
    double s1 = __dadd_rn(__dadd_rn(__dadd_rn(r1[0], r1[1]), __dadd_rn(r1[2], r1[3])),
                          __dadd_rn(__dadd_rn(r1[4], r1[5]), __dadd_rn(r1[6], r1[7])));
    const int x0 = lane >> 3, y0 = lane & 7, x1 = x0 + 4;
    const int v0 = (int16_t)(__double2int_rn(s0)) + 128, v1 = (int16_t)(__double2int_rn(s1)) + 128;
    if (comp == 0) {
        int16_t* row = reinterpret_cast<int16_t*>(t.yrow(m, slot, y0));
        row[x0] = (int16_t)v0;
        row[x1] = (int16_t)v1;
    } else {
        t.cchunk(m, comp - 1, 2 * y0 + (x0 >> 2))[x0 & 3] = (float)v0;
        t.cchunk(m, comp - 1, 2 * y0 + (x1 >> 2))[x1 & 3] = (float)v1;
    }
}

// exact colour conversion of one pixel, fp64, evaluation order of :1693-1695, clip (:1698), round (:1700)
__device__ __noinline__ uint32_t ycc_to_rgb_exact_packed(int Yi, float Cbf, float Crf) {
    double Y = (double)Yi, cb = __dsub_rn((double)Cbf, 128.0), cr = __dsub_rn((double)Crf, 128.0);
    double r = __dadd_rn(Y, __dmul_rn(1.402, cr));
    double g = __dsub_rn(__dsub_rn(Y, __dmul_rn(0.34414, cb)), __dmul_rn(0.71414, cr));
    double b = __dadd_rn(Y, __dmul_rn(1.772, cb));
    r = fmin(fmax(r, 0.0), 255.0);
    g = fmin(fmax(g, 0.0), 255.0);
    b = fmin(fmax(b, 0.0), 255.0);
    return (uint32_t)__double2int_rn(r) | ((uint32_t)__double2int_rn(g) << 8) | ((uint32_t)__double2int_rn(b) << 16);
}

template <int A> struct Cell { static constexpr int i = (A == 15) ? 6 : (7 * A) / 15; };

template <class L, int HX>
__device__ __forceinline__ void pixel_run(const Tiles<L>& t, int m, int r, int row_stride_s, uint32_t* stats) {
    // luma: 8 int16 of row r
    const int ys = (r >> 3) * L::HMAX + HX, yy = r & 7;
    const uint4 yv = *reinterpret_cast<const uint4*>(t.yrow(m, ys, yy));
    const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
    int Y[8];
#pragma unroll
    for (int p = 0; p < 8; p++) Y[p] = (int)(int16_t)(yw[p >> 1] >> (16 * (p & 1)));
    const int px0 = m * L::MCU_W + 8 * HX;
    if (L::NCOMP == 1) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            lo |= (uint32_t)min(max(Y[p], 0), 255) << (8 * p);  // (:1385-1386)
            hi |= (uint32_t)min(max(Y[p + 4], 0), 255) << (8 * p);
        }
        *reinterpret_cast<uint2*>(t.a + r * row_stride_s + px0) = make_uint2(lo, hi);
        return;
    }
    float cb[8], cr[8];
    if (!L::UPS) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            float4 lo = *reinterpret_cast<const float4*>(t.cchunk(m, k, 2 * yy));
            float4 hi = *reinterpret_cast<const float4*>(t.cchunk(m, k, 2 * yy + 1));
            float* o = k ? cr : cb;
            o[0] = lo.x; o[1] = lo.y; o[2] = lo.z; o[3] = lo.w; o[4] = hi.x; o[5] = hi.y; o[6] = hi.z; o[7] = hi.w;
        }
    } else {
        int j = (L::VMAX == 2) ? ((r == 15) ? 6 : (7 * r) / 15) : r;
        int j2 = j < 7 ? j + 1 : 7;
        const float4* w = t.w + r * 16 + 8 * HX;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            float p0[8], p1[8];
            {
                float4 a0 = *reinterpret_cast<const float4*>(t.cchunk(m, k, 2 * j));
                float4 a1 = *reinterpret_cast<const float4*>(t.cchunk(m, k, 2 * j + 1));
                float4 b0 = *reinterpret_cast<const float4*>(t.cchunk(m, k, 2 * j2));
                float4 b1 = *reinterpret_cast<const float4*>(t.cchunk(m, k, 2 * j2 + 1));
                p0[0] = a0.x; p0[1] = a0.y; p0[2] = a0.z; p0[3] = a0.w; p0[4] = a1.x; p0[5] = a1.y; p0[6] = a1.z; p0[7] = a1.w;
                p1[0] = b0.x; p1[1] = b0.y; p1[2] = b0.z; p1[3] = b0.w; p1[4] = b1.x; p1[5] = b1.y; p1[6] = b1.z; p1[7] = b1.w;
            }
            float* o = k ? cr : cb;
#define BJ_PIX(P)                                                                          \
    {                                                                                      \
        constexpr int i = (L::HMAX == 2) ? Cell<8 * HX + P>::i : P;                        \
        constexpr int i2 = i < 7 ? i + 1 : 7;                                              \
        float4 ww = w[P];                                                                  \
        float n = fmaf(ww.w, p1[i2], fmaf(ww.z, p1[i], fmaf(ww.y, p0[i2], ww.x * p0[i]))); \
        o[P] = bj::div15_round(n);                                                         \
    }
            BJ_PIX(0) BJ_PIX(1) BJ_PIX(2) BJ_PIX(3) BJ_PIX(4) BJ_PIX(5) BJ_PIX(6) BJ_PIX(7)
#undef BJ_PIX
        }
    }
    // colour: offsets from chroma in fp32, integer add + clamp; see bj_pixel_math.cuh for the tie rules
    uint32_t rgb[8];
    float guard = 0.f, dgmax = 0.f;
    bool btie = false;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        float cbm = cb[p] - 128.0f, crm = cr[p] - 128.0f;
        float rC = 1.402f * crm, gC = fmaf(-0.71414f, crm, -0.34414f * cbm), bC = 1.772f * cbm;
        float wr = rC + BJ_MAGIC, wg = gC + BJ_MAGIC, wb = bC + BJ_MAGIC;
        dgmax = fmaxf(dgmax, fabsf(gC - (wg - BJ_MAGIC)));
        guard = fmaxf(guard, fmaxf(fabsf(cbm), fabsf(crm)));
        btie = btie || (fabsf(cbm) == 125.0f);
        const int yb = Y[p] - BJ_MAGIC_BITS;
        uint32_t R = (uint32_t)__viaddmin_s32_relu(__float_as_int(wr), yb, 255);
        uint32_t G = (uint32_t)__viaddmin_s32_relu(__float_as_int(wg), yb, 255);
        uint32_t B = (uint32_t)__viaddmin_s32_relu(__float_as_int(wb), yb, 255);
        rgb[p] = R | (G << 8) | (B << 16);
    }
    if (btie || guard >= BJ_CHROMA_GUARD || dgmax > 0.5f - BJ_G_ERR) {
#pragma unroll 1
        for (int p = 0; p < 8; p++) rgb[p] = ycc_to_rgb_exact_packed(Y[p], cb[p], cr[p]);
        if (stats) atomicAdd(&stats[1], 8u);
    }
    // 8 pixels x 3 bytes = 6 words
    uint32_t o[6];
    o[0] = rgb[0] | (rgb[1] << 24);
    o[1] = (rgb[1] >> 8) | (rgb[2] << 16);
    o[2] = (rgb[2] >> 16) | (rgb[3] << 8);
    o[3] = rgb[4] | (rgb[5] << 24);
    o[4] = (rgb[5] >> 8) | (rgb[6] << 16);
    o[5] = (rgb[6] >> 16) | (rgb[7] << 8);
    uint2* s2 = reinterpret_cast<uint2*>(t.a + r * row_stride_s + px0 * 3);
    s2[0] = make_uint2(o[0], o[1]);
    s2[1] = make_uint2(o[2], o[3]);
    s2[2] = make_uint2(o[4], o[5]);
}

template <int HMAX, int VMAX, int NCOMP, int LAYOUT>
__global__ void __launch_bounds__(kThreads, 3)
bj_pixels_fast_kernel(const bj_image* __restrict__ images, const int16_t* __restrict__ coef,
                      const int16_t* __restrict__ qtabs, const double* __restrict__ tabT,
                      uint8_t* __restrict__ out, uint32_t* __restrict__ stats) {
    using L = Lay<HMAX, VMAX, NCOMP>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ bj_image im;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&images[blockIdx.y]);
        if (tid < (int)(sizeof(bj_image) / 4)) reinterpret_cast<uint32_t*>(&im)[tid] = src[tid];
    }
    __syncthreads();
    if ((int)im.layout != LAYOUT) return;
    const int strips_total = im.mcus_y * im.strips_per_row;
    if ((int)blockIdx.x >= strips_total) return;
    Tiles<L> t;
    t.a = smem;
    t.b = smem + kTileABytes;
    t.w = reinterpret_cast<float4*>(t.b + L::TILE_B);
    t.qt = reinterpret_cast<int16_t*>(reinterpret_cast<unsigned char*>(t.w) + L::W_BYTES);
    t.nat_zz = reinterpret_cast<uint8_t*>(t.qt + 3 * 64);

    const int my = blockIdx.x / im.strips_per_row;
    const int m0 = (blockIdx.x % im.strips_per_row) * im.strip_mcus;
    const int M = min(min((int)im.strip_mcus, (int)im.mcus_x - m0), L::MAXM);
    const int nblk = M * L::BPM;
    const int64_t gblk0 = (int64_t)im.coef_block0 + ((int64_t)my * im.mcus_x + m0) * L::BPM;

    // ---- tables + coefficient load ----------------------------------------------------------------
    if (tid < 64) {
        int nat = c_zz_nat2[tid];
        t.nat_zz[(nat & 7) * 8 + (nat >> 3)] = (uint8_t)tid;
    }
    for (int i = tid; i < NCOMP * 64; i += kThreads) t.qt[i] = qtabs[(size_t)im.qtab[i >> 6] * 64 + (i & 63)];
    if (L::UPS) {
        for (int i = tid; i < 256; i += kThreads) {
            int b = i >> 4, a = i & 15;
            int ii = a & 7, s = 0, jj = b & 7, tt = 0;
            if (HMAX == 2) bj::up_cell(a, ii, s);
            if (VMAX == 2) bj::up_cell(b, jj, tt);
            int w00, w10, w01, w11;
            bj::up_weights_2d(ii, jj, s, tt, w00, w10, w01, w11);
            t.w[i] = make_float4((float)w00, (float)w10, (float)w01, (float)w11);
        }
    }
    {
        const uint4* g = reinterpret_cast<const uint4*>(coef + gblk0 * 64);
        for (int i = tid; i < nblk * 8; i += kThreads) *reinterpret_cast<uint4*>(t.a + swzA(i >> 3, i & 7)) = __ldg(g + i);
    }
    __syncthreads();

    // ---- phase A: one thread per block --------------------------------------------------------------
    {
        const int blk = tid;
        const int m = blk / L::BPM, slot = blk - m * L::BPM;
        bool flagged = false;
        if (blk < nblk) {
            const int comp = slot < L::NY ? 0 : slot - L::NY + 1;
            float f[64];
            float S = 0.f;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint4 v = *reinterpret_cast<const uint4*>(t.a + swzA(blk, c));
                uint4 q = *reinterpret_cast<const uint4*>(t.qt + comp * 64 + c * 8);
                const uint32_t vw[4] = {v.x, v.y, v.z, v.w}, qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    int cf = (int16_t)(vw[e >> 1] >> (16 * (e & 1)));
                    int qq = (int16_t)(qw[e >> 1] >> (16 * (e & 1)));
                    int prod = (int16_t)(cf * qq);  // int16 * int16 -> int16 wraps (:869, :1348)
                    float x = int_to_float_magic(prod);
                    constexpr uint8_t zz[64] = {BJ_ZZ_NATURAL};
                    f[zz[c * 8 + e]] = x;
                    if (c * 8 + e) S += fabsf(x);
                }
            }
            const float T = fmaf(S, BJ_IDCT_ERR_REL, fmaf(fabsf(f[0]), BJ_IDCT_ERR_DC, BJ_IDCT_ERR_ABS));
            bj::idct8x8_fast(f);
            float maxd = 0.f;
            if (comp == 0) {
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    uint32_t wv[8];
#pragma unroll
                    for (int x = 0; x < 8; x++) {
                        float v = f[y * 8 + x];
                        float w = v + BJ_MAGIC;
                        maxd = fmaxf(maxd, fabsf(v - (w - BJ_MAGIC)));
                        wv[x] = (uint32_t)(__float_as_int(w) - BJ_MAGIC_BITS + 128);
                    }
                    uint4 o = make_uint4(__byte_perm(wv[0], wv[1], 0x5410), __byte_perm(wv[2], wv[3], 0x5410),
                                         __byte_perm(wv[4], wv[5], 0x5410), __byte_perm(wv[6], wv[7], 0x5410));
                    *reinterpret_cast<uint4*>(t.yrow(m, slot, y)) = o;
                }
            } else {
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    float o[8];
#pragma unroll
                    for (int x = 0; x < 8; x++) {
                        float v = f[y * 8 + x];
                        float r = (v + BJ_MAGIC) - BJ_MAGIC;
                        maxd = fmaxf(maxd, fabsf(v - r));
                        o[x] = r + 128.0f;
                    }
                    *reinterpret_cast<float4*>(t.cchunk(m, comp - 1, 2 * y)) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(t.cchunk(m, comp - 1, 2 * y + 1)) = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
            flagged = maxd > 0.5f - T;
        }
        unsigned mask = __ballot_sync(0xffffffffu, flagged);
        if (mask) {
            __syncwarp();
            if (stats && lane == 0) atomicAdd(&stats[0], (uint32_t)__popc(mask));
            while (mask) {
                int src = __ffs(mask) - 1;
                mask &= mask - 1;
                int b2 = (warp << 5) + src;
                int m2 = b2 / L::BPM;
                recompute_block_exact<L>(t, b2, m2, b2 - m2 * L::BPM, tabT, lane);
            }
        }
    }
    __syncthreads();

    // ---- phase B: lanes = MCUs, warps walk (pixel row, 8-pixel run) pairs ----------------------------
    const int x0 = m0 * L::MCU_W, y0 = my * L::MCU_H;
    const int cols = min(M * L::MCU_W, (int)im.width - x0);
    const int rows = min(L::MCU_H, (int)im.height - y0);
    const int row_stride_s = ((M * L::MCU_W * L::CH) + 15) & ~15;
    constexpr int pairs = L::MCU_H * L::HMAX;
    const int chunks = (M + 31) >> 5;
    for (int it = warp; it < pairs * chunks; it += kWarps) {
        const int pair = it % pairs, chunk = it / pairs;
        const int r = pair / L::HMAX, hx = pair % L::HMAX;
        const int m = (chunk << 5) + lane;
        if (r >= rows || m >= M) continue;
        if (L::HMAX == 2 && hx) pixel_run<L, (L::HMAX == 2 ? 1 : 0)>(t, m, r, row_stride_s, stats);
        else pixel_run<L, 0>(t, m, r, row_stride_s, stats);
    }
    __syncthreads();

    // ---- coalesced copy-out ----------------------------------------------------------------------------
    uint8_t* gout = out + (int64_t)im.out_offset + (int64_t)y0 * im.out_pitch + (int64_t)x0 * L::CH;
    const int nbytes = cols * L::CH;
    const bool aligned = ((reinterpret_cast<uintptr_t>(gout) & 15) == 0) && ((im.out_pitch & 15) == 0);
    if (aligned) {
        const int nvec = nbytes >> 4;
        for (int i = tid; i < rows * nvec; i += kThreads) {
            int r = i / nvec, v = i - r * nvec;
            uint4 val = *reinterpret_cast<const uint4*>(t.a + r * row_stride_s + (v << 4));
            __stcs(reinterpret_cast<uint4*>(gout + (int64_t)r * im.out_pitch + (v << 4)), val);
        }
        const int tail = nbytes & 15;
        if (tail) {
            for (int i = tid; i < rows * tail; i += kThreads) {
                int r = i / tail, b = (nvec << 4) + (i - r * tail);
                gout[(int64_t)r * im.out_pitch + b] = t.a[r * row_stride_s + b];
            }
        }
    } else {
        for (int i = tid; i < rows * nbytes; i += kThreads) {
            int r = i / nbytes, b = i - r * nbytes;
            gout[(int64_t)r * im.out_pitch + b] = t.a[r * row_stride_s + b];
        }
    }
}

template <int HMAX, int VMAX, int NCOMP, int LAYOUT>
cudaError_t launch(const bj_image* images, int n_images, int max_strips, const int16_t* coef, const int16_t* qtabs,
                   const double* tabT, uint8_t* out, uint32_t* stats, cudaStream_t st) {
    using L = Lay<HMAX, VMAX, NCOMP>;
    auto k = bj_pixels_fast_kernel<HMAX, VMAX, NCOMP, LAYOUT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
    if (e != cudaSuccess) return e;
    k<<<dim3((unsigned)max_strips, (unsigned)n_images), kThreads, L::SMEM, st>>>(images, coef, qtabs, tabT, out, stats);
    return cudaGetLastError();
}

}  // namespace

// Launch the specialised kernels for every layout present in layout_mask (bits BJ_LAYOUT_420..GRAY).
extern "C" bj_status bj_pixels_fast_launch(const bj_image* images, int n_images, int max_strips, const int16_t* coef,
                                           const int16_t* qtabs, const double* tabT, uint8_t* out, uint32_t layout_mask,
                                           uint32_t* stats, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess && (layout_mask & (1u << BJ_LAYOUT_420)))
        e = launch<2, 2, 3, BJ_LAYOUT_420>(images, n_images, max_strips, coef, qtabs, tabT, out, stats, st);
    if (e == cudaSuccess && (layout_mask & (1u << BJ_LAYOUT_422)))
        e = launch<2, 1, 3, BJ_LAYOUT_422>(images, n_images, max_strips, coef, qtabs, tabT, out, stats, st);
    if (e == cudaSuccess && (layout_mask & (1u << BJ_LAYOUT_440)))
        e = launch<1, 2, 3, BJ_LAYOUT_440>(images, n_images, max_strips, coef, qtabs, tabT, out, stats, st);
    if (e == cudaSuccess && (layout_mask & (1u << BJ_LAYOUT_444)))
        e = launch<1, 1, 3, BJ_LAYOUT_444>(images, n_images, max_strips, coef, qtabs, tabT, out, stats, st);
    if (e == cudaSuccess && (layout_mask & (1u << BJ_LAYOUT_GRAY)))
        e = launch<1, 1, 1, BJ_LAYOUT_GRAY>(images, n_images, max_strips, coef, qtabs, tabT, out, stats, st);
    if (e != cudaSuccess) return bj_set_cuda_error(e, "bj_pixels/fast");
    return BJ_OK;
}
