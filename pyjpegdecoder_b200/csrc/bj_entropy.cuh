// bj_entropy.cuh -- per-thread logic of the entropy stage (Huffman / run-length / progressive
// bookkeeping), written as __host__ __device__ templates over a bit source so that the CPU test-suite
// (tests/hostsim) can drive exactly the code the CUDA kernels in bj_entropy.cu run.
//
// Reference being replaced (tbpaolini/PyJpegDecoder, jpeg_decoder.py):
//   bits_generator / get_bits   :654-695    bit reader (un-stuffing is done by bj_unstuff.cu)
//   next_huffval                :712-722    canonical Huffman decode, <= 16 bits
//   bin_twos_complement         :1636-1646  EXTEND
//   baseline_dct_scan           :805-866    DC difference + AC run/size loop (EOB 0x00, ZRL 0xF0)
//   progressive_dct_scan        :983-1057   DC first / DC refine
//                               :1120-1256  AC first with EOB runs
//                               :1100-1115, :1183-1198, :1209-1232, :1258-1292  AC refinement
//
// Parallel decomposition (new design, nothing like it exists in the reference): a stream (restart
// interval, or a whole scan) is cut into subsequences of BJ_SUBSEQ_BITS bits.  The decoder state at
// a subsequence boundary is (bit position, zig-zag index, slot in the MCU).  Huffman codes
// self-synchronise: a decoder started at a wrong state falls into step with the true one after a few
// symbols, so a speculative pass that starts one subsequence early yields, for almost every
// subsequence, its true entry state; a fix-up pass re-decodes the few that disagree with their
// predecessor's exit state until nothing changes (then all states are true by induction from the
// known state at the stream start).  Prefix sums of "blocks started per subsequence" and of the DC
// differences then give every subsequence its first block index and DC predictors, and a final
// pass writes whole 128-byte blocks.
#pragma once
#include <stdint.h>

#include "../../include/b200jpeg.h"
#include "bj_pixel_math.cuh"  // BJ_HD

namespace bj {

// ---- packed decoder state ------------------------------------------------------------------------
// bits 0..43 bit position in the un-stuffed buffer, 44..50 zig-zag index, 51..54 slot.
BJ_HD uint64_t pack_state(uint64_t pos, int z, int slot) {
    return pos | ((uint64_t)(uint32_t)z << 44) | ((uint64_t)(uint32_t)slot << 51);
}
BJ_HD uint64_t state_pos(uint64_t s) { return s & ((1ull << 44) - 1); }
BJ_HD int state_z(uint64_t s) { return (int)((s >> 44) & 127); }
BJ_HD int state_slot(uint64_t s) { return (int)((s >> 51) & 15); }

// ---- Huffman LUT entry (built by pyjpegdecoder_b200/huffman.py) ----------------------------------
// Fields are byte aligned so that each one is a single LOP/PRMT on the device:
//   direct:   byte 0 = L + (symbol & 15) (bits consumed by code + value; bit 7 clear)
//             byte 1 = zig-zag advance for baseline AC (run + 1; 64 for EOB; 16 for ZRL; 1 for DC)
//             byte 2 = code length L (0 = no such code; such entries have byte 0 = byte 1 = 1 so that a
//                      speculating decoder still makes progress)
//             byte 3 = symbol
//   indirect: bit 7 set, bits 8..23 = offset of a 128-entry second-level table (relative to the table)
// first level: 512 entries indexed by the next 9 bits; second level by the following 7 bits.
BJ_HD uint32_t lut_lookup(const uint32_t* tab, uint32_t peek16) {
    uint32_t e = tab[peek16 >> 7];
    if (e & 0x80u) e = tab[((e >> 8) & 0xFFFFu) + (peek16 & 127u)];
    return e;
}
BJ_HD int ent_total(uint32_t e) { return (int)(e & 0xFFu); }
BJ_HD int ent_adv(uint32_t e) { return (int)((e >> 8) & 0xFFu); }
BJ_HD int ent_len(uint32_t e) { return (int)((e >> 16) & 0xFFu); }
BJ_HD int ent_sym(uint32_t e) { return (int)(e >> 24); }

// EXTEND (bin_twos_complement, :1636-1646)
BJ_HD int extend(uint32_t v, int n) { return (n == 0) ? 0 : ((v >> (n - 1)) ? (int)v : (int)v - ((1 << n) - 1)); }

// ---- bit reader over big-endian 32-bit words -----------------------------------------------------
// Two 32-bit words (w0 = current, w1 = next) and the number of bits of w0 already consumed: the next
// 32 bits are one funnel shift away, consuming bits is an add, and a refill is one word fetch about
// every third symbol.  One symbol never needs more than 16 code bits + 15 value bits = 31 bits.
BJ_HD uint32_t funnel_left(uint32_t hi, uint32_t lo, int sh) {  // upper 32 bits of (hi:lo) << sh, 0 <= sh < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, sh);
#else
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
#endif
}

template <class Src>
struct BitReader {
    const Src* src;
    uint32_t w0, w1;  // current and next word
    int o;            // bits of w0 consumed (0..31)
    uint32_t next;    // next word to fetch
    uint32_t rel;     // bit position relative to `base` (a stream is far below 2^32 bits)
    uint64_t base;    // absolute bit position of the stream start

    BJ_HDM void seek(const Src* s, uint64_t base_bit, uint32_t rel_bit) {
        src = s;
        base = base_bit;
        rel = rel_bit;
        const uint64_t p = base_bit + rel_bit;
        uint32_t w = (uint32_t)(p >> 5);
        o = (int)(p & 31);
        w0 = src->word(w);
        w1 = src->word(w + 1);
        next = w + 2;
    }
    BJ_HDM uint64_t abs_pos() const { return base + rel; }
    BJ_HDM uint32_t peek32() const { return funnel_left(w0, w1, o); }
    BJ_HDM uint32_t peek16() const { return peek32() >> 16; }
    // n bits (1..16) that follow the first `skipn` bits (skipn + n <= 32)
    BJ_HDM uint32_t bits_after(int skipn, int n) const { return (peek32() << skipn) >> (32 - n); }
    BJ_HDM void skip(int n) {  // n <= 32
        o += n;
        rel += (uint32_t)n;
        if (o >= 32) {
            o -= 32;
            w0 = w1;
            w1 = src->word(next++);
        }
    }
};

// value bits of a symbol taken from an already fetched 32-bit look-ahead
BJ_HD uint32_t take_bits(uint32_t pk, int skipn, int n) { return (pk << skipn) >> (32 - n); }

// ---- per-scan context (shared memory on the device) ----------------------------------------------
struct ScanCtx {
    uint16_t dc_tab[BJ_MAX_SLOTS];    // table offsets inside the blob
    uint16_t ac_tab[BJ_MAX_SLOTS];
    uint8_t slot_comp[BJ_MAX_SLOTS];  // slot -> DC predictor index
    int nslots;
    int ss, se, al;
};

struct SubCount {
    uint32_t blocks;  // blocks started (baseline / DC) or block advance (AC first) in the subsequence
    int32_t dc[3];    // sum of DC differences per scan component
};

// True when the bits from the reader's position to the end of the stream (end_rel, same base) are only
// the 1-padding of the last byte (fewer than 8 bits, all ones): no Huffman code consists of ones only,
// so real data never looks like this.
template <class Src>
BJ_HD bool at_padding(const BitReader<Src>& rd, uint32_t end_rel) {
    if (rd.rel >= end_rel) return true;
    uint32_t left = end_rel - rd.rel;
    if (left >= 8) return false;
    uint32_t ones = (1u << left) - 1u;
    return rd.bits_after(0, (int)left) == ones;
}

#define BJ_M_BASE 0
#define BJ_M_DCFIRST 1

// All positions below are 32-bit and relative to the stream start (rd.base): own_rel = first bit of
// the subsequence (0xFFFFFFFF: count nothing), stop_rel = its end, end_rel = end of the stream.
// `lut` is the scan's LUT blob; it is a separate argument (not a field of ScanCtx) so that the CUDA
// compiler can see when it points to shared memory and emit LDS instead of generic loads.

// ---- baseline / DC-first: counting pass ----------------------------------------------------------
// Decode from the reader's position with state (z, slot) until rd.rel >= stop_rel.  Blocks whose DC
// symbol starts at a position >= own_rel are counted and their DC differences summed.  Nothing is
// written.  FLAT loop: every iteration decodes exactly one symbol (DC or AC, chosen by a table
// select), so the 32 lanes of a warp stay converged although their blocks have different numbers of
// symbols; only the short DC / end-of-block bodies diverge.
template <int MODE, class Src>
BJ_HD void sync_run(BitReader<Src>& rd, int& z, int& slot, const ScanCtx& c, const uint32_t* lut, uint32_t own_rel,
                    uint32_t stop_rel, uint32_t end_rel, SubCount& cnt) {
    const int nslots = c.nslots;
    uint32_t dct = c.dc_tab[slot], act = c.ac_tab[slot];
    while (rd.rel < stop_rel) {
        const bool is_dc = (z == 0);
        if (is_dc && end_rel - rd.rel < 8 && at_padding(rd, end_rel)) {
            rd.rel = end_rel;
            break;
        }
        const uint32_t pk = rd.peek32();
        const uint32_t e = lut_lookup(lut + (is_dc ? dct : act), pk >> 16);
        int adv = ent_adv(e);
        if (is_dc) {
            if (rd.rel >= own_rel) {
                const int L = ent_len(e);
                const int t = L ? ent_sym(e) : 0;
                cnt.blocks++;
                int diff = extend(t ? take_bits(pk, L, t) : 0u, t);
                int k = c.slot_comp[slot];
                if (k == 0) cnt.dc[0] += diff;
                else if (k == 1) cnt.dc[1] += diff;
                else cnt.dc[2] += diff;
            }
            adv = (MODE == BJ_M_DCFIRST) ? 64 : 1;
        }
        rd.skip(ent_total(e));  // entries that are not a code consume 1 bit: any deterministic step will do
        z += adv;
        if (z >= 64) {
            z = 0;
            slot = (slot + 1 == nslots) ? 0 : slot + 1;
            dct = c.dc_tab[slot];
            act = c.ac_tab[slot];
        }
    }
}

// ---- baseline: writing pass ----------------------------------------------------------------------
// Sink: begin(), put(zigzag_index, value), commit(block_in_stream, slot) -- one whole block at a time.
// Returns BJ_ERR_* bits.  `blk` is the index (within the stream) of the first block this thread
// owns; on return it is one past the last block written.  If z != 0 on entry the open block belongs
// to the previous subsequence: it is decoded without being written.
// This loop is nested (per block: DC, AC symbols, commit) on purpose: the 24-instruction block flush
// then runs with all lanes of the warp converged, which measured faster on B200 than the flat form
// (2.8 vs 3.9 ms per 512 images) even though lanes wait for the block with the most symbols.
template <class Src, class Sink>
BJ_HD uint32_t base_write_run(BitReader<Src>& rd, int z, int slot, const ScanCtx& c, const uint32_t* lut, uint32_t stop_rel,
                              uint32_t end_rel, uint32_t& blk, uint32_t nblk_stream, int pred[3], Sink& sink) {
    // Single-exit loops only (errors are carried in `err`, never returned from inside a loop): the
    // compiler then reconverges the warp after every inner loop, which this nested form relies on.
    const int nslots = c.nslots;
    uint32_t err = 0;
    if (z != 0) {
        const uint32_t* const tab = lut + c.ac_tab[slot];
        while (z < 64) {
            uint32_t e = lut_lookup(tab, rd.peek16());
            if (ent_len(e) == 0) err |= BJ_ERR_BAD_CODE;
            if (rd.rel >= end_rel + 64) err |= BJ_ERR_OVERRUN;
            rd.skip(ent_total(e));
            z += err ? 64 : ent_adv(e);
        }
        slot = (slot + 1 == nslots) ? 0 : slot + 1;
    }
    while (err == 0 && rd.rel < stop_rel && blk < nblk_stream) {
        sink.begin();
        {
            const uint32_t pk = rd.peek32();
            uint32_t e = lut_lookup(lut + c.dc_tab[slot], pk >> 16);
            int L = ent_len(e), t = ent_sym(e);
            if (L == 0) { err |= BJ_ERR_BAD_CODE; t = 0; }
            int diff = extend(t ? take_bits(pk, L, t) : 0u, t);
            int k = c.slot_comp[slot];
            int pv;
            if (k == 0) pv = (pred[0] += diff);
            else if (k == 1) pv = (pred[1] += diff);
            else pv = (pred[2] += diff);
            sink.put(0, (int16_t)pv);  // previous_dc is int16 (:735, :818-820)
            rd.skip(ent_total(e));
        }
        const uint32_t* const tab = lut + c.ac_tab[slot];
        int zz = 1;
        while (zz < 64) {
            const uint32_t pk = rd.peek32();
            uint32_t e = lut_lookup(tab, pk >> 16);
            int L = ent_len(e), tot = ent_total(e);
            int s = tot - L;
            zz += ent_adv(e) - 1;  // zero run (EOB: jumps past 63, ZRL: 15)
            if (L == 0) { err |= BJ_ERR_BAD_CODE; zz = 64; }
            if (s && zz < 64) sink.put(zz, (int16_t)extend(take_bits(pk, L, s), s));
            rd.skip(tot);
            zz++;
        }
        if (rd.rel > end_rel + 7) err |= BJ_ERR_OVERRUN;
        if (err == 0) {
            sink.commit(blk, slot);
            blk++;
        }
        slot = (slot + 1 == nslots) ? 0 : slot + 1;
    }
    return err;
}

// ---- DC first: writing pass (one 16-bit store per block) -----------------------------------------
template <class Src, class Sink>
BJ_HD uint32_t dcfirst_write_run(BitReader<Src>& rd, int slot, const ScanCtx& c, const uint32_t* lut, uint32_t stop_rel,
                                 uint32_t end_rel, uint32_t& blk, uint32_t nblk_stream, int pred[3], Sink& sink) {
    while (rd.rel < stop_rel && blk < nblk_stream) {
        const uint32_t pk = rd.peek32();
        uint32_t e = lut_lookup(lut + c.dc_tab[slot], pk >> 16);
        int L = ent_len(e), t = ent_sym(e);
        if (L == 0) return BJ_ERR_BAD_CODE;
        int diff = extend(t ? take_bits(pk, L, t) : 0u, t);
        int k = c.slot_comp[slot];
        int pv;
        if (k == 0) pv = (pred[0] += diff);
        else if (k == 1) pv = (pred[1] += diff);
        else pv = (pred[2] += diff);
        rd.skip(ent_total(e));
        if (rd.rel > end_rel + 7) return BJ_ERR_OVERRUN;
        sink.store_dc(blk, slot, (int16_t)((uint32_t)(int32_t)(int16_t)pv << c.al));  // (:1029)
        blk++;
        slot = (slot + 1 == c.nslots) ? 0 : slot + 1;
    }
    return 0;
}

// ---- AC first (:1120-1256) -----------------------------------------------------------------------
// State: zig-zag index z in [ss, se].  A symbol either places a coefficient after a zero run, skips
// 16 zeros (ZRL), or ends the band of this block and of the next EOBRUN-1 blocks.  WRITE = false:
// count block advance only; WRITE = true: also store coefficients of symbols that start at or
// after own_rel (symbol-level ownership; the planes were zeroed before the first scan).
template <bool WRITE, class Src, class Sink>
BJ_HD uint32_t acfirst_run(BitReader<Src>& rd, int& z, const ScanCtx& c, const uint32_t* lut, uint32_t own_rel,
                           uint32_t stop_rel, uint32_t end_rel, uint32_t& blk, uint32_t nblk_stream, uint32_t& advance,
                           Sink& sink) {
    const uint32_t* tab = lut + c.ac_tab[0];
    while (rd.rel < stop_rel) {
        if (WRITE && blk >= nblk_stream) break;
        if (end_rel - rd.rel < 8 && at_padding(rd, end_rel)) {
            rd.rel = end_rel;
            break;
        }
        const uint32_t pk = rd.peek32();
        uint32_t e = lut_lookup(tab, pk >> 16);
        int L = ent_len(e), rs = ent_sym(e);
        bool own = rd.rel >= own_rel;
        if (L == 0) {
            if (WRITE) return BJ_ERR_BAD_CODE;
            rd.skip(1);
            continue;
        }
        int r = rs >> 4, s = rs & 15;
        uint32_t adv = 0;
        if (s) {
            z += r;
            if (WRITE && own) {
                if (z > 63) return BJ_ERR_COEF_INDEX;
                sink.store(blk, z, (int16_t)((uint32_t)extend(take_bits(pk, L, s), s) << c.al));  // (:1225)
            }
            z++;
            rd.skip(L + s);
        } else if (r == 15) {
            z += 16;  // (:1142-1143)
            rd.skip(L);
        } else {
            uint32_t run = (1u << r) + (r ? take_bits(pk, L, r) : 0u);  // (:1144-1149)
            rd.skip(L + r);
            adv = run;
            z = c.ss;
        }
        if (z > c.se) {
            adv = 1;
            z = c.ss;
        }
        blk += adv;
        if (own) advance += adv;
    }
    return 0;
}

// ---- AC refinement (:1100-1115, :1183-1198, :1209-1232, :1258-1292) -------------------------------
// Sequential over one stream: the number of correction bits after a symbol depends on which
// coefficients of the block are already non-zero.  Coef: at(block_in_stream, zigzag) -> int16_t&.
// Correction is the reference's `coef |= bit << Al` on the two's-complement value (:1114), which is
// NOT the T.81 rule for negative coefficients; bit-exact parity with the reference requires it.
template <class Src, class Coef>
BJ_HD uint32_t acrefine_stream(BitReader<Src>& rd, const ScanCtx& c, const uint32_t* lut, uint32_t end_rel,
                               uint32_t nblk_stream, Coef& coef) {
    const uint32_t* tab = lut + c.ac_tab[0];
    const int ss = c.ss, se = c.se, al = c.al;
    uint32_t blk = 0;
    while (blk < nblk_stream) {
        int z = ss;
        uint32_t eob_run = 0;
        while (z <= se) {
            if (rd.rel > end_rel + 7) return BJ_ERR_OVERRUN;
            const uint32_t pk = rd.peek32();
            uint32_t e = lut_lookup(tab, pk >> 16);
            int L = ent_len(e), rs = ent_sym(e);
            if (L == 0) return BJ_ERR_BAD_CODE;
            int r = rs >> 4, s = rs & 15;
            int zero_run;
            int newval = 0;
            if (rs == 0) {
                rd.skip(L);
                eob_run = 1;
                break;
            } else if (rs == 0xF0) {
                rd.skip(L);
                zero_run = 16;
            } else if (s == 0) {
                eob_run = (1u << r) + (r ? take_bits(pk, L, r) : 0u);
                rd.skip(L + r);
                break;
            } else {
                zero_run = r;
                newval = extend(take_bits(pk, L, s), s);  // value bits come right after the code (:1202)
                rd.skip(L + s);
            }
            // skip `zero_run` zero-history coefficients, refining the non-zero ones passed (:1184-1193);
            // their correction bits follow in the same order (:1107-1115)
            while (zero_run > 0) {
                if (z > 63) return BJ_ERR_COEF_INDEX;
                int16_t& cf = coef.at(blk, z);
                if (cf == 0) zero_run--;
                else {
                    cf = (int16_t)(cf | (int16_t)(rd.bits_after(0, 1) << al));
                    rd.skip(1);
                }
                z++;
            }
            if (s) {
                // a new coefficient lands on the next zero-history position (:1211-1215)
                for (;;) {
                    if (z > 63) return BJ_ERR_COEF_INDEX;
                    int16_t& cf = coef.at(blk, z);
                    if (cf == 0) break;
                    cf = (int16_t)(cf | (int16_t)(rd.bits_after(0, 1) << al));
                    rd.skip(1);
                    z++;
                }
                coef.at(blk, z) = (int16_t)((uint32_t)newval << al);  // (:1225)
                z++;
            }
        }
        if (z > se) {
            blk++;
            continue;
        }
        // end-of-band run: the rest of this band and the bands of the next eob_run-1 blocks only
        // carry correction bits for their non-zero coefficients (:1258-1292)
        while (eob_run > 0 && blk < nblk_stream) {
            for (; z <= se; z++) {
                int16_t& cf = coef.at(blk, z);
                if (cf != 0) {
                    if (rd.rel > end_rel + 7) return BJ_ERR_OVERRUN;
                    cf = (int16_t)(cf | (int16_t)(rd.bits_after(0, 1) << al));
                    rd.skip(1);
                }
            }
            eob_run--;
            blk++;
            z = ss;
        }
    }
    return 0;
}

}  // namespace bj
