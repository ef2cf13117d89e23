/*
 * oracle/jpeg_oracle.c -- CPU restatement of the reference decoder's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file restates, in plain scalar C, the algorithm of
 * tbpaolini/PyJpegDecoder (jpeg_decoder.py) so that the CUDA path can be checked against it on
 * the GPU box, where the Python reference does not exist.  It is imported only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
 * (pyjpegdecoder_b200) never links, loads or calls it.
 *
 * Parity is PINNED: tests/test_oracle.py checks this oracle bit-for-bit against fixtures produced
 * by the unmodified reference (tests/golden/make_golden.py): RGB output, the int16 Y/Cb/Cr canvas,
 * the quantised coefficient planes after every scan (54 baseline/progressive files), the full
 * decode of the reference's own "progressive scan example/base image.jpg" (sha256) and the two
 * after-scan renders shipped with it.
 *
 * Every function cites the reference lines (jpeg_decoder.py:LINE) it follows.  Where the reference
 * is non-conformant the oracle is bug-compatible:
 *   - AC successive-approximation correction is `coef |= bit << Al` on a two's-complement int16
 *     (:1114), not the T.81 sign-magnitude rule;
 *   - the byte after ANY 0xFF inside entropy data is dropped (:676-677);
 *   - restart handling counts MCUs and never looks at the marker (:667-669, :898);
 *   - no clamp between IDCT and colour conversion (:1573, :1626, :1698);
 *   - chroma upsampling is scipy.interpolate.griddata's Delaunay piece-wise linear interpolation
 *     of the MCU tile (:1588-1626), restated as exact integer arithmetic with the triangulation's
 *     diagonal map (checked against live scipy in tests/test_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, the IDCT sum must be
 * evaluated exactly like numpy does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_NOT_JPEG 1     /* NotJpeg        (:40)  */
#define ORC_UNSUPPORTED 2  /* UnsupportedJpeg (:150,156,180,182) */
#define ORC_CORRUPT 3      /* CorruptedJpeg  (:174,234,329,458,581,719,922,934,958,967) */
#define ORC_TRUNCATED 4    /* IndexError in the reference: entropy data ran past the file end */
#define ORC_NOMEM 5

#define ORC_FLAG_CONFORMANT_REFINE 1u /* T.81 refinement instead of the reference's OR (:1114) */

/* zagzig (:1672-1681): zig-zag index -> (x, y) = (horizontal, vertical) frequency */
static const uint8_t ZZ_X[64] = {
    0, 1, 0, 0, 1, 2, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4,
    3, 2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 5, 6, 7, 7, 6, 7};
static const uint8_t ZZ_Y[64] = {
    0, 0, 1, 2, 1, 0, 0, 1, 2, 3, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3,
    4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 4, 5, 6, 7, 7, 6, 5, 6, 7, 7};

/* Diagonal map of scipy/Qhull's Delaunay triangulation of the 8x8 integer grid, as used by
 * griddata at :1624 (scipy 1.18.1).  Bit (7*i + j) set: cell (x-cell i, y-cell j) is split by the
 * diagonal joining (i,j)-(i+1,j+1); clear: (i+1,j)-(i,j+1).  Row-major listing MSB first
 * = 0x14a4e53555555; tests regenerate it from live scipy. */
static const char* DIAG_ROWS[7] = {"1010010", "1001001", "1100101", "0011010",
                                   "1010101", "0101010", "1010101"};

typedef struct {
    int id, h, v, tq;
    int bw, bh;    /* padded block grid */
    int16_t* coef; /* [bh][bw][64], zig-zag order, quantised */
} orc_comp;

typedef struct {
    uint16_t count[17];
    uint16_t first[17]; /* first code of each length */
    uint16_t offset[17];
    uint8_t vals[256];
    int present;
} orc_huff;

typedef struct {
    /* outputs */
    int width, height, ncomp, progressive;
    int hmax, vmax, mcus_x, mcus_y;
    int canvas_w, canvas_h;
    int nscans;
    int restart_interval_last;
    orc_comp comp[3];
    int16_t* canvas; /* [canvas_h][canvas_w][ncomp]  (the reference's image_array, y-major here) */
    uint8_t* rgb;    /* [height][width][ncomp==3 ? 3 : 1] */
    /* decoder state */
    const uint8_t* f;
    size_t n, pos;
    orc_huff huff[2][16];
    int16_t qt[256][64]; /* zig-zag order as stored in the file */
    uint8_t qt_present[256];
    int ri;
    uint32_t flags;
    int stop_after_scan;
    int scan_count;
    int err;
    /* bit reader (:654-695) */
    uint64_t acc;
    int nbits;
} orc_dec;

/* ---- bit reader: bits_generator/get_bits (:654-695) ------------------------------------------ */
static void br_reset(orc_dec* d) {
    d->acc = 0;
    d->nbits = 0;
}
static void br_restart(orc_dec* d) { /* restart=True (:667-669): drop bits, jump the 2 marker bytes */
    br_reset(d);
    d->pos += 2;
}
static int br_bit(orc_dec* d) {
    if (d->nbits == 0) {
        if (d->pos >= d->n) {
            d->err = ORC_TRUNCATED;
            return 0;
        }
        uint8_t b = d->f[d->pos++];
        if (b == 0xFF) d->pos++; /* byte after any 0xFF is skipped (:676-677) */
        d->acc = b;
        d->nbits = 8;
    }
    d->nbits--;
    return (int)((d->acc >> d->nbits) & 1u);
}
static uint32_t br_bits(orc_dec* d, int n) {
    uint32_t v = 0;
    if (n > 16) {
        d->err = ORC_CORRUPT; /* no valid stream asks for more than 16 raw bits at once */
        return 0;
    }
    for (int i = 0; i < n && !d->err; i++) v = (v << 1) | (uint32_t)br_bit(d);
    return v;
}
/* bin_twos_complement (:1636-1646): JPEG EXTEND */
static int extend(uint32_t v, int n) {
    if (n <= 0 || n > 16) return 0;
    if (v >> (n - 1)) return (int)v;
    return (int)v - ((1 << n) - 1);
}
/* next_huffval (:712-722, :951-961); canonical code construction (:366-377) */
static int huff_decode(orc_dec* d, const orc_huff* h) {
    uint32_t code = 0;
    for (int len = 1; len <= 16; len++) {
        code = (code << 1) | (uint32_t)br_bit(d);
        if (d->err) return 0;
        if (h->count[len] && code >= h->first[len] && code - h->first[len] < h->count[len])
            return h->vals[h->offset[len] + (code - h->first[len])];
    }
    d->err = ORC_CORRUPT; /* >16 bits (:718-719) */
    return 0;
}

static uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

/* ---- segment parsers ------------------------------------------------------------------------- */
/* start_of_frame (:112-247) */
static void parse_sof(orc_dec* d, const uint8_t* s, size_t len, int marker) {
    if (marker == 0xC0)
        d->progressive = 0;
    else if (marker == 0xC2)
        d->progressive = 1;
    if (len < 6) {
        d->err = ORC_CORRUPT;
        return;
    }
    if (s[0] != 8) {
        d->err = ORC_UNSUPPORTED;
        return;
    }
    d->height = (int)be16(s + 1);
    d->width = (int)be16(s + 3);
    if (d->width == 0) {
        d->err = ORC_CORRUPT;
        return;
    }
    int nc = s[5];
    if (nc != 1 && nc != 3) {
        d->err = ORC_UNSUPPORTED;
        return;
    }
    if (len < (size_t)(6 + 3 * nc)) {
        d->err = ORC_CORRUPT;
        return;
    }
    d->ncomp = nc;
    d->hmax = d->vmax = 1;
    for (int i = 0; i < nc; i++) {
        d->comp[i].id = s[6 + 3 * i];
        d->comp[i].h = s[7 + 3 * i] >> 4;
        d->comp[i].v = s[7 + 3 * i] & 15;
        d->comp[i].tq = s[8 + 3 * i];
        if (d->comp[i].h > d->hmax) d->hmax = d->comp[i].h;
        if (d->comp[i].v > d->vmax) d->vmax = d->comp[i].v;
    }
}
/* define_huffman_table (:249-390) */
static void parse_dht(orc_dec* d, const uint8_t* s, size_t len) {
    size_t p = 0;
    while (p < len) {
        int dest = s[p++];
        if (p + 16 > len) {
            d->err = ORC_CORRUPT;
            return;
        }
        orc_huff* h = &d->huff[(dest >> 4) & 1][dest & 15];
        memset(h, 0, sizeof *h);
        int total = 0;
        for (int i = 1; i <= 16; i++) {
            h->count[i] = s[p + i - 1];
            total += h->count[i];
        }
        p += 16;
        if (p + (size_t)total > len || total > 256) {
            d->err = ORC_CORRUPT; /* (:327-329) */
            return;
        }
        memcpy(h->vals, s + p, (size_t)total);
        p += (size_t)total;
        uint32_t code = 0;
        int off = 0;
        for (int i = 1; i <= 16; i++) { /* (:368-374) */
            code <<= 1;
            h->first[i] = (uint16_t)code;
            h->offset[i] = (uint16_t)off;
            code += h->count[i];
            off += h->count[i];
        }
        h->present = 1;
    }
}
/* define_quantization_table (:392-472): 8-bit entries only, key = raw Pq/Tq byte */
static void parse_dqt(orc_dec* d, const uint8_t* s, size_t len) {
    size_t p = 0;
    while (p < len) {
        int dest = s[p++];
        if (p + 64 > len) {
            d->err = ORC_CORRUPT; /* (:457-458) */
            return;
        }
        for (int i = 0; i < 64; i++) d->qt[dest][i] = s[p + i];
        d->qt_present[dest] = 1;
        p += 64;
    }
}

/* ---- geometry (:583-632) --------------------------------------------------------------------- */
static int alloc_planes(orc_dec* d) {
    if (d->canvas) return 0;
    if (d->ncomp == 1) { /* single component: 8x8 MCUs, sampling factors are irrelevant */
        d->comp[0].h = d->comp[0].v = 1;
        d->hmax = d->vmax = 1;
    }
    d->mcus_x = (d->width + 8 * d->hmax - 1) / (8 * d->hmax);
    d->mcus_y = (d->height + 8 * d->vmax - 1) / (8 * d->vmax);
    d->canvas_w = d->mcus_x * 8 * d->hmax;
    d->canvas_h = d->mcus_y * 8 * d->vmax;
    for (int c = 0; c < d->ncomp; c++) {
        orc_comp* k = &d->comp[c];
        if (k->h < 1 || k->v < 1 || d->hmax % k->h || d->vmax % k->v) return ORC_UNSUPPORTED;
        int rh = d->hmax / k->h, rv = d->vmax / k->v;
        if (rh > 2 || rv > 2) return ORC_UNSUPPORTED; /* only the 8->16 griddata map is restated */
        k->bw = d->mcus_x * k->h;
        k->bh = d->mcus_y * k->v;
        k->coef = (int16_t*)calloc((size_t)k->bw * k->bh * 64, sizeof(int16_t));
        if (!k->coef) return ORC_NOMEM;
    }
    d->canvas = (int16_t*)calloc((size_t)d->canvas_w * d->canvas_h * d->ncomp, sizeof(int16_t));
    if (!d->canvas) return ORC_NOMEM;
    return 0;
}

typedef struct {
    int ci;     /* component index in frame order */
    int td, ta; /* table selectors */
} scan_comp;

/* ---- baseline scan (:697-906): entropy part only, coefficients kept quantised ---------------- */
static void baseline_scan(orc_dec* d, const scan_comp* sc, int ns) {
    int mcu_w, mcu_h, n_mcu;
    if (ns > 1) { /* (:609-611) */
        mcu_w = d->mcus_x;
        mcu_h = d->mcus_y;
    } else { /* (:612-619) */
        const orc_comp* k = &d->comp[sc[0].ci];
        int rh = d->hmax / k->h, rv = d->vmax / k->v;
        mcu_w = ((d->width + rh - 1) / rh + 7) / 8;
        mcu_h = ((d->height + rv - 1) / rv + 7) / 8;
    }
    n_mcu = mcu_w * mcu_h;
    int pred[3] = {0, 0, 0};
    br_reset(d);
    for (int m = 0; m < n_mcu && !d->err; m++) {
        int my = m / mcu_w, mx = m % mcu_w;
        for (int s = 0; s < ns; s++) {
            orc_comp* k = &d->comp[sc[s].ci];
            int h = ns > 1 ? k->h : 1, v = ns > 1 ? k->v : 1;
            const orc_huff* hd = &d->huff[0][sc[s].td];
            const orc_huff* ha = &d->huff[1][sc[s].ta];
            for (int r = 0; r < h * v; r++) {
                int bx = mx * h + r % h, by = my * v + r / h; /* (:875) */
                int16_t* blk = k->coef + ((size_t)by * k->bw + bx) * 64;
                memset(blk, 0, 128);
                int t = huff_decode(d, hd);
                int diff = extend(br_bits(d, t), t);
                pred[s] = (int16_t)(pred[s] + diff); /* previous_dc is int16 (:735, :818-819) */
                blk[0] = (int16_t)pred[s];
                int idx = 1;
                while (idx < 64 && !d->err) { /* (:834-866) */
                    int rs = huff_decode(d, ha);
                    if (rs == 0) break;
                    idx += rs >> 4;
                    if (idx >= 64) break;
                    int sz = rs & 15;
                    if (sz) blk[idx] = (int16_t)extend(br_bits(d, sz), sz);
                    idx++;
                }
                if (d->err) return;
            }
        }
        if (d->ri > 0 && (m + 1) % d->ri == 0 && (m + 1) != n_mcu) { /* (:898-900) */
            br_restart(d);
            pred[0] = pred[1] = pred[2] = 0;
        }
    }
}

/* ---- progressive scan (:908-1304) ------------------------------------------------------------ */
typedef struct {
    int16_t** p;
    size_t n, cap;
} refine_q;
static int rq_push(refine_q* q, int16_t* c) {
    if (q->n == q->cap) {
        size_t nc = q->cap ? q->cap * 2 : 1024;
        int16_t** np_ = (int16_t**)realloc(q->p, nc * sizeof *np_);
        if (!np_) return -1;
        q->p = np_;
        q->cap = nc;
    }
    q->p[q->n++] = c;
    return 0;
}
/* refine_ac (:1100-1115) */
static void refine_flush(orc_dec* d, refine_q* q, int al) {
    for (size_t i = 0; i < q->n && !d->err; i++) {
        int bit = br_bit(d);
        int16_t* c = q->p[i];
        if (d->flags & ORC_FLAG_CONFORMANT_REFINE) {
            if (bit && !((*c >> al) & 1)) *c = (int16_t)(*c >= 0 ? *c + (1 << al) : *c - (1 << al));
        } else {
            *c = (int16_t)(*c | (bit << al)); /* (:1114) two's-complement OR */
        }
    }
    q->n = 0;
}

static void progressive_scan(orc_dec* d, const scan_comp* sc, int ns, int ss, int se, int ah, int al) {
    int is_dc;
    if (ss == 0 && se == 0)
        is_dc = 1;
    else if (ss > 0 && se >= ss)
        is_dc = 0;
    else {
        d->err = ORC_CORRUPT; /* (:922) */
        return;
    }
    int refining;
    if (ah == 0)
        refining = 0;
    else if (ah - al == 1)
        refining = 1;
    else {
        d->err = ORC_CORRUPT; /* (:934) */
        return;
    }
    if (!is_dc && ns > 1) {
        d->err = ORC_CORRUPT; /* (:967) */
        return;
    }
    int mcu_w, mcu_h;
    if (ns > 1) {
        mcu_w = d->mcus_x;
        mcu_h = d->mcus_y;
    } else {
        const orc_comp* k = &d->comp[sc[0].ci];
        int rh = d->hmax / k->h, rv = d->vmax / k->v;
        mcu_w = ((d->width + rh - 1) / rh + 7) / 8;
        mcu_h = ((d->height + rv - 1) / rv + 7) / 8;
    }
    int n_mcu = mcu_w * mcu_h;
    br_reset(d);

    if (is_dc) { /* (:974-1057) */
        int pred[3] = {0, 0, 0};
        for (int m = 0; m < n_mcu && !d->err; m++) {
            for (int s = 0; s < ns; s++) {
                orc_comp* k = &d->comp[sc[s].ci];
                int h = ns > 1 ? k->h : 1, v = ns > 1 ? k->v : 1;
                /* (:993-994): the reference positions blocks with the component's own MCU shape even
                 * in a single-component scan; that is only right when h = v = 1 there. */
                int ox = (m % mcu_w) * k->h, oy = (m / mcu_w) * k->v;
                for (int r = 0; r < h * v; r++) {
                    int bx = ox + r % k->h, by = oy + r / k->h;
                    if (bx >= k->bw || by >= k->bh) {
                        d->err = ORC_UNSUPPORTED;
                        return;
                    }
                    int16_t* c0 = k->coef + ((size_t)by * k->bw + bx) * 64;
                    if (!refining) {
                        int t = huff_decode(d, &d->huff[0][sc[s].td]);
                        int diff = extend(br_bits(d, t), t);
                        pred[s] = (int16_t)(pred[s] + diff);
                        c0[0] = (int16_t)((uint32_t)(int32_t)pred[s] << al); /* (:1029) */
                    } else {
                        c0[0] = (int16_t)(c0[0] | (br_bit(d) << al)); /* (:1037-1038) */
                    }
                }
            }
            if (d->ri > 0 && (m + 1) % d->ri == 0 && (m + 1) != n_mcu) { /* (:1050-1053) */
                br_restart(d);
                pred[0] = pred[1] = pred[2] = 0;
            }
        }
        return;
    }

    /* AC scan (:1060-1302) */
    orc_comp* k = &d->comp[sc[0].ci];
    const orc_huff* ha = &d->huff[1][sc[0].ta];
    refine_q q = {0, 0, 0};
    int eob_run = 0, zero_run = 0;
    int cur = 0;
#define COEF(mcu, zi) (k->coef + ((size_t)((mcu) / mcu_w) * k->bw + (size_t)((mcu) % mcu_w)) * 64 + (zi))
    while (cur < n_mcu && !d->err) {
        int blk = cur; /* x, y of the block stay fixed until recomputed (:1125-1126, :1244-1245) */
        int idx = ss;
        while (idx <= se && !d->err) {
            int rs = huff_decode(d, ha);
            if (d->err) break;
            int run = rs >> 4, sz = rs & 15;
            if (rs == 0) { /* (:1138-1141) */
                eob_run = 1;
                break;
            } else if (rs == 0xF0) {
                zero_run = 16;
            } else if (sz == 0) { /* (:1144-1149) */
                eob_run = (1 << run) + (int)br_bits(d, run);
                break;
            } else {
                zero_run = run;
            }
            if (!refining && zero_run) { /* (:1177-1179) */
                idx += zero_run;
                zero_run = 0;
            } else {
                while (zero_run > 0) { /* (:1184-1193) */
                    if (idx > 63 || blk >= n_mcu) {
                        d->err = ORC_CORRUPT;
                        break;
                    }
                    int16_t* c = COEF(blk, idx);
                    if (*c == 0)
                        zero_run--;
                    else if (rq_push(&q, c)) {
                        d->err = ORC_NOMEM;
                        break;
                    }
                    idx++;
                }
            }
            if (d->err) break;
            if (sz > 0) { /* (:1201-1228) */
                int val = extend(br_bits(d, sz), sz);
                if (idx > 63 || blk >= n_mcu) {
                    d->err = ORC_CORRUPT;
                    break;
                }
                if (refining) {
                    while (*COEF(blk, idx) != 0) { /* (:1211-1215) */
                        if (rq_push(&q, COEF(blk, idx))) {
                            d->err = ORC_NOMEM;
                            break;
                        }
                        idx++;
                        if (idx > 63) {
                            d->err = ORC_CORRUPT;
                            break;
                        }
                    }
                    if (d->err) break;
                }
                *COEF(blk, idx) = (int16_t)((uint32_t)val << al); /* (:1225) */
                idx++;
            }
            if (refining) refine_flush(d, &q, al); /* (:1231-1232) */
        }
        if (d->err) break;
        if (idx > se) { /* (:1240-1245) */
            cur++;
            blk = cur;
        }
        if (!refining) { /* (:1248-1250) */
            cur += eob_run;
            eob_run = 0;
        } else { /* (:1258-1276) */
            while (eob_run > 0) {
                if (blk >= n_mcu || idx > 63) {
                    d->err = ORC_CORRUPT;
                    break;
                }
                int16_t* c = COEF(blk, idx);
                if (*c != 0 && rq_push(&q, c)) {
                    d->err = ORC_NOMEM;
                    break;
                }
                idx++;
                if (idx > se) {
                    eob_run--;
                    cur++;
                    idx = ss;
                    blk = cur;
                }
            }
            refine_flush(d, &q, al); /* (:1285-1286) */
        }
        if (d->ri > 0 && cur % d->ri == 0 && cur != n_mcu) br_restart(d); /* (:1297-1298) */
    }
#undef COEF
    free(q.p);
}

/* ---- pixel reconstruction -------------------------------------------------------------------- */
/* InverseDCT.__call__ (:1561-1573).  `tab` is the reference's idct_table[x][y][u][v] (:1541-1553).
 * in[u*8+v] = dequantised coefficient with horizontal frequency u (int16, :869/:1348).
 * np.sum over the 64 products runs numpy's pairwise routine: 8 accumulators r[j] += p[8i+j],
 * then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).  np.round is round-half-even (rint). */
static void idct_block(const double* tab, const int16_t* in, int16_t* out /* [x*8+y] */) {
    for (int xy = 0; xy < 64; xy++) {
        const double* t = tab + (size_t)xy * 64;
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = (double)in[j] * t[j];
        for (int i = 8; i < 64; i += 8)
            for (int j = 0; j < 8; j++) r[j] += (double)in[i + j] * t[i + j];
        double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        out[xy] = (int16_t)((int16_t)(long long)rint(s) + 128);
    }
}

static int diag_bit(int i, int j) { return DIAG_ROWS[i][j] == '1'; }

/* floor((2N+15)/30): N/15 rounded to nearest; never a tie because 15 is odd. */
static int16_t div15_round(int n) {
    int t = 2 * n + 15;
    int q = t / 30;
    if (t % 30 < 0) q--;
    return (int16_t)q;
}

/* ResizeGrid.__call__ (:1588-1626) for one 8x8 source tile P[x][y] (stride sx, sy in int16 units),
 * scaled by (rh, rv) in {1,2}: output tile O[(8*rh) x (8*rv)].  Align-corners mesh
 * np.mgrid[0:7:16j] (:1602-1606): output a -> source coordinate 7a/15. */
static void upsample_tile(const int16_t* P, int sx, int sy, int rh, int rv, int16_t* O, int ox, int oy) {
#define PP(i, j) ((int)P[(i) * sx + (j) * sy])
    int W = 8 * rh, H = 8 * rv;
    for (int a = 0; a < W; a++) {
        for (int b = 0; b < H; b++) {
            int i = a, s = 0, j = b, t = 0;
            if (rh == 2) {
                i = 7 * a / 15;
                s = 7 * a % 15;
                if (a == 15) {
                    i = 6;
                    s = 15;
                }
            }
            if (rv == 2) {
                j = 7 * b / 15;
                t = 7 * b % 15;
                if (b == 15) {
                    j = 6;
                    t = 15;
                }
            }
            int16_t val;
            if (rh == 2 && rv == 2) {
                int n;
                if (diag_bit(i, j)) {
                    if (s >= t)
                        n = (15 - s) * PP(i, j) + (s - t) * PP(i + 1, j) + t * PP(i + 1, j + 1);
                    else
                        n = (15 - t) * PP(i, j) + (t - s) * PP(i, j + 1) + s * PP(i + 1, j + 1);
                } else {
                    if (s + t <= 15)
                        n = (15 - s - t) * PP(i, j) + s * PP(i + 1, j) + t * PP(i, j + 1);
                    else
                        n = (s + t - 15) * PP(i + 1, j + 1) + (15 - t) * PP(i + 1, j) + (15 - s) * PP(i, j + 1);
                }
                val = div15_round(n);
            } else if (rh == 2) {
                val = s ? div15_round((15 - s) * PP(i, j) + s * PP(i + 1, j)) : (int16_t)PP(i, j);
            } else if (rv == 2) {
                val = t ? div15_round((15 - t) * PP(i, j) + t * PP(i, j + 1)) : (int16_t)PP(i, j);
            } else {
                val = (int16_t)PP(i, j);
            }
            O[(size_t)a * ox + (size_t)b * oy] = val;
        }
    }
#undef PP
}

/* De-zigzag + dequantise (:869, :1347-1348; int16 product wraps), IDCT, upsample, store: the pixel
 * half of baseline_dct_scan (:868-891) and the progressive final stage (:1306-1366).  Both give the
 * same canvas because the upsampler never looks outside one 8x8 source block (for the supported
 * ratios) -- checked against the reference's canvas in tests. */
static int reconstruct(orc_dec* d, const double* tab) {
    for (int c = 0; c < d->ncomp; c++) {
        orc_comp* k = &d->comp[c];
        if (!d->qt_present[k->tq]) return ORC_CORRUPT; /* KeyError in the reference */
        const int16_t* q = d->qt[k->tq];
        int rh = d->hmax / k->h, rv = d->vmax / k->v;
        for (int by = 0; by < k->bh; by++)
            for (int bx = 0; bx < k->bw; bx++) {
                const int16_t* z = k->coef + ((size_t)by * k->bw + bx) * 64;
                int16_t in[64], px[64];
                for (int i = 0; i < 64; i++)
                    in[ZZ_X[i] * 8 + ZZ_Y[i]] = (int16_t)((int)z[i] * (int)q[i]); /* wraps like int16*int16 */
                idct_block(tab, in, px);
                int16_t* o = d->canvas + ((size_t)(by * 8 * rv) * d->canvas_w + (size_t)bx * 8 * rh) * d->ncomp + c;
                /* px[x*8+y]; canvas is [y][x][c] */
                upsample_tile(px, 8, 1, rh, rv, o, d->ncomp, d->canvas_w * d->ncomp);
            }
    }
    return 0;
}

static double clip255(double x) { return x < 0.0 ? 0.0 : (x > 255.0 ? 255.0 : x); }

/* end_of_image (:1368-1390) + YCbCr_to_RGB (:1683-1700) */
static int finish(orc_dec* d) {
    int nc = d->ncomp == 3 ? 3 : 1;
    d->rgb = (uint8_t*)malloc((size_t)d->width * d->height * nc);
    if (!d->rgb) return ORC_NOMEM;
    for (int y = 0; y < d->height; y++)
        for (int x = 0; x < d->width; x++) {
            const int16_t* p = d->canvas + ((size_t)y * d->canvas_w + x) * d->ncomp;
            uint8_t* o = d->rgb + ((size_t)y * d->width + x) * nc;
            if (d->ncomp == 3) {
                double Y = p[0], Cb = p[1], Cr = p[2];
                double R = Y + 1.402 * (Cr - 128.0);
                double G = Y - 0.34414 * (Cb - 128.0) - 0.71414 * (Cr - 128.0);
                double B = Y + 1.772 * (Cb - 128.0);
                o[0] = (uint8_t)rint(clip255(R));
                o[1] = (uint8_t)rint(clip255(G));
                o[2] = (uint8_t)rint(clip255(B));
            } else {
                int v = p[0];
                o[0] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); /* (:1385-1386) */
            }
        }
    return 0;
}

/* ---- marker walk: JpegDecoder.__init__ (:29-110) + start_of_scan (:505-652) ------------------ */
static int count_sos(const uint8_t* f, size_t from, size_t n) { /* bytes.count(SOS) (:636) */
    int c = 0;
    for (size_t i = from; i + 1 < n; i++)
        if (f[i] == 0xFF && f[i + 1] == 0xDA) {
            c++;
            i++;
        }
    return c;
}

void orc_free(orc_dec* d) {
    if (!d) return;
    for (int c = 0; c < 3; c++) free(d->comp[c].coef);
    free(d->canvas);
    free(d->rgb);
    free(d);
}

/* Decode a whole file.  idct_table: 4096 doubles [x][y][u][v] computed by the caller with the
 * reference's expression (:1544-1553).  stop_after_scan > 0: stop the entropy stage after that
 * many scans (the pixel stages still run on what has been decoded, like a file truncated there
 * and closed with EOI).  Returns a handle (query with the accessors below) or NULL. */
orc_dec* orc_decode(const uint8_t* file, size_t n, const double* idct_table, uint32_t flags,
                    int stop_after_scan, int* err_out) {
    orc_dec* d = (orc_dec*)calloc(1, sizeof *d);
    if (!d) {
        if (err_out) *err_out = ORC_NOMEM;
        return NULL;
    }
    d->f = file;
    d->n = n;
    d->flags = flags;
    d->stop_after_scan = stop_after_scan;
    int finished = 0;
    if (n < 3 || file[0] != 0xFF || file[1] != 0xD8 || file[2] != 0xFF) { /* (:39-40) */
        d->err = ORC_NOT_JPEG;
        goto done;
    }
    d->pos = 2;
    while (!finished && !d->err) { /* (:78-110) */
        if (d->pos >= n) break;    /* IndexError -> loop ends silently (:79-83) */
        if (file[d->pos] != 0xFF) {
            d->pos++;
            continue;
        }
        if (d->pos + 1 >= n) break;
        int m = file[d->pos + 1];
        d->pos += 2;
        if (m == 0x00 || (m >= 0xD0 && m <= 0xD7)) continue; /* (:93) */
        if (m == 0xD9) { /* end_of_image (:1368): no length is needed */
            finished = 1;
            break;
        }
        if (d->pos + 2 > n) break;
        size_t seglen = be16(file + d->pos);
        d->pos += 2;
        size_t size = seglen >= 2 ? seglen - 2 : 0;
        const uint8_t* seg = file + d->pos;
        size_t avail = d->pos + size <= n ? size : n - d->pos;
        switch (m) {
            case 0xC0:
            case 0xC2:
                parse_sof(d, seg, avail, m);
                d->pos += size;
                break;
            case 0xC4:
                parse_dht(d, seg, avail);
                d->pos += size;
                break;
            case 0xDB:
                parse_dqt(d, seg, avail);
                d->pos += size;
                break;
            case 0xDD:
                if (avail >= 2) d->ri = (int)be16(seg);
                d->pos += 2; /* (:476-477) */
                break;
            case 0xDA: { /* start_of_scan (:505-652) */
                if (avail < 1 || d->ncomp == 0) {
                    d->err = ORC_CORRUPT;
                    break;
                }
                int ns = seg[0];
                scan_comp sc[4];
                if (ns < 1 || ns > 3 || avail < (size_t)(1 + 2 * ns + (d->progressive ? 3 : 0))) {
                    d->err = ORC_CORRUPT;
                    break;
                }
                for (int i = 0; i < ns && !d->err; i++) {
                    int id = seg[1 + 2 * i], tb = seg[2 + 2 * i];
                    sc[i].ci = -1;
                    for (int c = 0; c < d->ncomp; c++)
                        if (d->comp[c].id == id) sc[i].ci = c;
                    if (sc[i].ci < 0) d->err = ORC_CORRUPT; /* KeyError in the reference (:555) */
                    sc[i].td = tb >> 4;
                    sc[i].ta = tb & 15;
                }
                if (d->err) break;
                int ss = 0, se = 63, ah = 0, al = 0;
                if (d->progressive) {
                    const uint8_t* t = seg + 1 + 2 * ns;
                    ss = t[0];
                    se = t[1];
                    ah = t[2] >> 4;
                    al = t[2] & 15;
                }
                d->pos += size;
                if (d->height == 0) { /* DNL lookup (:575-581) */
                    size_t i;
                    int found = 0;
                    for (i = d->pos; i + 5 < n; i++)
                        if (file[i] == 0xFF && file[i + 1] == 0xDC) {
                            d->height = (int)be16(file + i + 4);
                            found = 1;
                            break;
                        }
                    if (!found) {
                        d->err = ORC_CORRUPT;
                        break;
                    }
                }
                if ((d->err = alloc_planes(d))) break;
                if (d->scan_count == 0) d->nscans = count_sos(file, d->pos, n) + 1; /* (:635-637) */
                if (d->stop_after_scan > 0 && d->scan_count >= d->stop_after_scan) {
                    /* behave like a file cut before this SOS and closed with EOI */
                    finished = 1;
                    break;
                }
                if (d->progressive)
                    progressive_scan(d, sc, ns, ss, se, ah, al);
                else
                    baseline_scan(d, sc, ns);
                d->scan_count++;
                d->restart_interval_last = d->ri;
                break;
            }
            default:
                d->pos += size; /* unknown segment skipped by length (:106) */
        }
    }
    if (!d->err && d->canvas) {
        d->err = reconstruct(d, idct_table);
        if (!d->err) d->err = finish(d);
    } else if (!d->err && !d->canvas) {
        d->err = ORC_CORRUPT;
    }
done:
    if (err_out) *err_out = d->err;
    return d;
}

/* ---- accessors for the ctypes wrapper -------------------------------------------------------- */
int orc_info(const orc_dec* d, int* out /* 16 ints */) {
    out[0] = d->width;
    out[1] = d->height;
    out[2] = d->ncomp;
    out[3] = d->progressive;
    out[4] = d->canvas_w;
    out[5] = d->canvas_h;
    out[6] = d->scan_count;
    out[7] = d->nscans;
    out[8] = d->hmax;
    out[9] = d->vmax;
    for (int c = 0; c < 3; c++) {
        out[10 + 2 * c] = d->comp[c].bw;
        out[11 + 2 * c] = d->comp[c].bh;
    }
    return d->err;
}
const int16_t* orc_coef(const orc_dec* d, int c) { return d->comp[c].coef; }
const int16_t* orc_canvas(const orc_dec* d) { return d->canvas; }
const uint8_t* orc_rgb(const orc_dec* d) { return d->rgb; }

/* Stand-alone pixel stages for kernel tests: coefficient planes -> canvas/RGB. */
void orc_idct_block(const double* tab, const int16_t* in_xy, int16_t* out_xy) { idct_block(tab, in_xy, out_xy); }
void orc_upsample_tile(const int16_t* P, int rh, int rv, int16_t* O /* [(8rh)][(8rv)] x-major */) {
    upsample_tile(P, 8, 1, rh, rv, O, 8 * rv, 1);
}
uint64_t orc_diag_map(void) {
    uint64_t m = 0;
    for (int i = 0; i < 7; i++)
        for (int j = 0; j < 7; j++)
            if (diag_bit(i, j)) m |= 1ull << (7 * i + j);
    return m;
}
