// Host build of the pixel-stage arithmetic in pyjpegdecoder_b200/csrc/bj_pixel_math.cuh.
// TEST-ONLY: lets the CPU test-suite check the exact fp32 operation sequence the CUDA kernel runs
// (the header is __host__ __device__ and uses explicit fmaf only).  Never used by the product.
#include <cstdint>
#include <cmath>
#include <cstring>
#include "../../pyjpegdecoder_b200/csrc/bj_pixel_math.cuh"

extern "C" {

// blocks: n x 64 dequantised coefficients in natural order [v*8+u] (as int32).
// out: n x 64 rounded samples (+128) [y*8+x]; flagged: n bytes, 1 if the kernel would recompute the block exactly.
void hs_idct_blocks(const int32_t* blocks, int n, int16_t* out, uint8_t* flagged, float* min_tie_dist) {
    for (int b = 0; b < n; b++) {
        float f[64];
        float S = 0.f;
        for (int k = 0; k < 64; k++) {
            f[k] = (float)blocks[b * 64 + k];
            if (k) S += fabsf(f[k]);
        }
        const float T = fmaf(S, BJ_IDCT_ERR_REL, fmaf(fabsf(f[0]), BJ_IDCT_ERR_DC, BJ_IDCT_ERR_ABS));
        bj::idct8x8_fast(f);
        float mind = 1.0f;
        for (int k = 0; k < 64; k++) {
            float td;
            float r = bj::round_tie(f[k], td);
            mind = fminf(mind, td);
            out[b * 64 + k] = (int16_t)((int)r + 128);
        }
        flagged[b] = mind < T;
        if (min_tie_dist) min_tie_dist[b] = mind;
    }
}

// The packed fast path of csrc/bj_pixels_mma.cu (block_idct): DC peeling, packed IDCT (the host build runs the
// same fmaf / + / * sequence lane by lane), magic-add rounding, tie test.  blocks: n x 64 dequantised coefficients
// in natural order [v*8+u]; lo4 != 0 runs the 4x4 low-frequency variant (coefficients outside the corner must be 0).
void hs_idct_blocks_packed(const int32_t* blocks, int n, int lo4, int16_t* out, uint8_t* flagged, float* max_dist,
                           float* thresholds, double* raw) {
    using bj::F2;
    for (int b = 0; b < n; b++) {
        F2 P[8][4];
        float Sw = 0.f;
        for (int k = 0; k < 64; k++) {
            const float x = (float)blocks[b * 64 + k];
            const int v = k >> 3, u = k & 7;
            if (u & 1) P[v][u >> 1].y = x; else P[v][u >> 1].x = x;
            if (k) Sw = fmaf(fabsf(x), bj::idct_err_weight(v, u), Sw);
        }
        const float dc_abs = fabsf(P[0][0].x);
        const float S = Sw * (1.0f / BJ_IDCT_W_MIN);
        const float dc_int = bj::dc_peel(P[0][0].x);
        const float T = fmaf(fmaf(P[0][0].x, bj::idct_err_weight(0, 0), Sw), BJ_IDCT_ERR_U, BJ_IDCT_ERR_ABS);
        const float shift = (BJ_MAGIC + 128.0f) + dc_int;
        F2 W[4][8];
        float maxd;
        if (lo4) bj::idct8x8_round_packed<true>(P, shift, W, maxd);
        else bj::idct8x8_round_packed<false>(P, shift, W, maxd);
        for (int yp = 0; yp < 4; yp++)
            for (int x = 0; x < 8; x++) {
                uint32_t b0, b1;
                memcpy(&b0, &W[yp][x].x, 4);
                memcpy(&b1, &W[yp][x].y, 4);
                out[b * 64 + (2 * yp) * 8 + x] = (int16_t)(b0 & 0xffffu);
                out[b * 64 + (2 * yp + 1) * 8 + x] = (int16_t)(b1 & 0xffffu);
            }
        flagged[b] = (maxd > 0.5f - T) || (S + dc_abs > 32767.0f);
        if (max_dist) max_dist[b] = maxd;
        if (thresholds) thresholds[b] = T;
        if (raw) {   // the fp32 values before rounding (error probe): v = (W - shift) + (v - rint(v)) is not recoverable
                     // exactly from W, so run the transform again without the rounding step
            F2 R[4][8];
            if (lo4) bj::idct8x8_packed<true>(P, R); else bj::idct8x8_packed<false>(P, R);
            for (int yp = 0; yp < 4; yp++)
                for (int x = 0; x < 8; x++) {
                    raw[b * 64 + (2 * yp) * 8 + x] = (double)R[yp][x].x + (double)dc_int;
                    raw[b * 64 + (2 * yp + 1) * 8 + x] = (double)R[yp][x].y + (double)dc_int;
                }
        }
    }
}

// weights of kind (rh, rv) for MCU pixel (b, a): w[b*16+a][4] = w00,w10,w01,w11 and cell i,j
void hs_weights(int rh, int rv, int32_t* w, int32_t* cell) {
    for (int b = 0; b < 16; b++)
        for (int a = 0; a < 16; a++) {
            int ii = a & 7, s = 0, jj = b & 7, t = 0;
            if (rh == 2) bj::up_cell(a, ii, s);
            if (rv == 2) bj::up_cell(b, jj, t);
            int w00, w10, w01, w11;
            bj::up_weights_2d(ii, jj, s, t, w00, w10, w01, w11);
            int32_t* o = w + (b * 16 + a) * 4;
            o[0] = w00; o[1] = w10; o[2] = w01; o[3] = w11;
            cell[(b * 16 + a) * 2] = ii;
            cell[(b * 16 + a) * 2 + 1] = jj;
        }
}

float hs_div15(float n) { return bj::div15_round(n); }

// integer-aware colour path of the specialised kernel: returns rgb and the "needs fp64" flag per pixel
void hs_color2(const int16_t* ycc, int n, uint8_t* rgb, uint8_t* slow) {
    for (int i = 0; i < n; i++) {
        int Y = ycc[3 * i];
        float cbm = (float)ycc[3 * i + 1] - 128.0f, crm = (float)ycc[3 * i + 2] - 128.0f;
        float rC = 1.402f * crm, gC = fmaf(-0.71414f, crm, -0.34414f * cbm), bC = 1.772f * cbm;
        float wr = rC + BJ_MAGIC, wg = gC + BJ_MAGIC, wb = bC + BJ_MAGIC;
        float dg = fabsf(gC - (wg - BJ_MAGIC));
        bool s = fmaxf(fabsf(cbm), fabsf(crm)) >= BJ_CHROMA_GUARD || fabsf(cbm) == 125.0f || dg > 0.5f - BJ_G_ERR;
        int32_t ir, ig, ib;
        memcpy(&ir, &wr, 4); memcpy(&ig, &wg, 4); memcpy(&ib, &wb, 4);
        int v[3] = {Y + ir - BJ_MAGIC_BITS, Y + ig - BJ_MAGIC_BITS, Y + ib - BJ_MAGIC_BITS};
        for (int c = 0; c < 3; c++) rgb[3 * i + c] = (uint8_t)(v[c] < 0 ? 0 : (v[c] > 255 ? 255 : v[c]));
        slow[i] = s;
    }
}

// colour fast path for n pixels; tie[i] = 1 when the kernel would take the fp64 path
void hs_color(const int16_t* ycc, int n, uint8_t* rgb, uint8_t* tie) {
    for (int i = 0; i < n; i++) {
        float R, G, B, err;
        bj::ycc_to_rgb_fast((float)ycc[3 * i], (float)ycc[3 * i + 1], (float)ycc[3 * i + 2], R, G, B, err);
        float v[3] = {R, G, B};
        bool t = false;
        for (int c = 0; c < 3; c++) {
            float w = v[c] + BJ_MAGIC;
            float r = w - BJ_MAGIC;
            t = t || ((0.5f - fabsf(v[c] - r)) < err && v[c] > -1.0f && v[c] < 256.0f);
            int iv = (int)r;
            rgb[3 * i + c] = (uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
        }
        tie[i] = t;
    }
}
}
