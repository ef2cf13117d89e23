// Host build of the entropy-stage logic in pyjpegdecoder_b200/csrc/bj_entropy.cuh, driven by a
// sequential emulation of the kernel orchestration in bj_entropy.cu (speculate -> fix-up rounds ->
// prefix sums -> write).  TEST-ONLY: lets the CPU suite check the decode logic, the convergence of
// the self-synchronising scheme and its bookkeeping against the oracle.  Never used by the product.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../pyjpegdecoder_b200/csrc/bj_entropy.cuh"

using namespace bj;

struct HostSrc {
    const uint32_t* w;
    uint64_t n;
    uint32_t word(uint32_t i) const { return i < n ? w[i] : 0xFFFFFFFFu; }
    void start(uint32_t) const {}
};

extern "C" {

// Sequential model of bj_unstuff: drops the byte after 0xFF when it is 0x00, removes RSTn markers
// and records where each restart interval starts.  Output: big-endian words.
// Returns number of compacted bytes; starts[k] = byte offset of stream k (k < n_streams).
uint64_t hs_unstuff(const uint8_t* raw, uint64_t n, uint32_t* words, uint64_t out_base, uint64_t* starts,
                    uint32_t n_streams) {
    uint64_t o = out_base;
    uint32_t k = 0;
    if (n_streams) starts[0] = o;
    auto put = [&](uint8_t b) {
        uint32_t& w = words[o >> 2];
        int sh = 24 - 8 * (int)(o & 3);
        w = (w & ~(0xFFu << sh)) | ((uint32_t)b << sh);
        o++;
    };
    for (uint64_t i = 0; i < n; i++) {
        uint8_t b = raw[i];
        uint8_t prev = i ? raw[i - 1] : 0;
        uint8_t next = (i + 1 < n) ? raw[i + 1] : 0;
        if (b == 0x00 && prev == 0xFF) continue;
        if (prev == 0xFF && (b & 0xF8) == 0xD0) {
            k++;
            if (k < n_streams) starts[k] = o;
            continue;
        }
        if (b == 0xFF && (next & 0xF8) == 0xD0) continue;
        put(b);
    }
    return o - out_base;
}

struct SimScan {
    const uint32_t* lut;
    uint16_t dc_tab[BJ_MAX_SLOTS], ac_tab[BJ_MAX_SLOTS];
    uint8_t slot_comp[BJ_MAX_SLOTS];
    int nslots, ss, se, al, mode;
};

static ScanCtx make_ctx(const SimScan& s) {
    ScanCtx c;
    for (int i = 0; i < BJ_MAX_SLOTS; i++) {
        c.dc_tab[i] = s.dc_tab[i];
        c.ac_tab[i] = s.ac_tab[i];
        c.slot_comp[i] = s.slot_comp[i];
    }
    ctx_finish(c);
    c.nslots = s.nslots;
    c.ss = s.ss;
    c.se = s.se;
    c.al = s.al;
    return c;
}

struct BlockSink {  // baseline: a whole block at a time
    int16_t cur[64];
    int16_t* out;  // [nblk_stream][64] in scan order
    void begin() { memset(cur, 0, sizeof cur); }
    void put(int z, int16_t v) { cur[z] = v; }
    void commit(uint32_t blk, int) { memcpy(out + (size_t)blk * 64, cur, 128); }
    void store_dc(uint32_t blk, int, int16_t v) { out[(size_t)blk * 64] = v; }
    void store(uint32_t blk, int z, int16_t v) { out[(size_t)blk * 64 + z] = v; }
};

// Decode one stream with the parallel scheme, emulated sequentially.
//   coef: [nblk_stream][64] blocks in scan order (for AC first: band coefficients stored in place)
//   stats[0] = subsequences, [1] = entry states wrong after the speculative pass,
//   [2] = fix-up rounds needed, [3] = total re-decodes in fix-up
// Returns error bits.
uint32_t hs_decode_stream(const uint32_t* words, uint64_t n_words, uint64_t start_byte, uint64_t end_byte,
                          const SimScan* ss_, uint32_t nblk_stream, int sub_bits, int16_t* coef, uint32_t* stats, int warm) {
    const SimScan& sc = *ss_;
    ScanCtx c = make_ctx(sc);
    HostSrc src{words, n_words};
    const uint64_t b0 = start_byte * 8, b1 = end_byte * 8;
    const uint64_t S = (uint64_t)sub_bits;
    const uint32_t nsub = (uint32_t)((b1 - b0 + S - 1) / S);
    const int z0 = (sc.mode == BJ_MODE_AC_FIRST) ? sc.ss : 0;
    std::vector<uint64_t> entry(nsub), exitst(nsub);
    std::vector<SubCount> cnt(nsub);
    BlockSink dummy{};
    const uint32_t nbits = (uint32_t)(b1 - b0);
    auto run_sub = [&](uint32_t l, uint64_t st) {  // decode subsequence l from entry state st -> exit state, counts
        uint32_t own = l * (uint32_t)S, stop = own + (uint32_t)S < nbits ? own + (uint32_t)S : nbits;
        BitReader<HostSrc> rd;
        rd.seek(&src, b0, (uint32_t)(state_pos(st) - b0));
        int z = state_z(st), slot = state_slot(st);
        SubCount k{0, {0, 0, 0}};
        if (sc.mode == BJ_MODE_BASELINE) sync_run<BJ_M_BASE>(rd, z, slot, c, sc.lut, own, stop, nbits, k);
        else if (sc.mode == BJ_MODE_DC_FIRST) sync_run<BJ_M_DCFIRST>(rd, z, slot, c, sc.lut, own, stop, nbits, k);
        else {
            uint32_t blk = 0, adv = 0;
            acfirst_run<false>(rd, z, c, sc.lut, own, stop, nbits, blk, 0xFFFFFFFFu, adv, dummy);
            k.blocks = adv;
        }
        cnt[l] = k;
        exitst[l] = pack_state(rd.abs_pos(), z, slot);
    };
    // pass 1: speculate (thread l starts `warm` subsequences early, at an assumed block start)
    for (uint32_t l = 0; l < nsub; l++) {
        uint64_t st;
        if (l == 0) st = pack_state(b0, z0, 0);
        else {
            uint32_t own = l * (uint32_t)S;
            BitReader<HostSrc> rd;
            rd.seek(&src, b0, own - (uint32_t)S * ((uint32_t)warm < l ? (uint32_t)warm : l));
            int z = z0, slot = 0;
            SubCount k{0, {0, 0, 0}};
            if (sc.mode == BJ_MODE_BASELINE) sync_run<BJ_M_BASE>(rd, z, slot, c, sc.lut, 0xFFFFFFFFu, own, nbits, k);
            else if (sc.mode == BJ_MODE_DC_FIRST) sync_run<BJ_M_DCFIRST>(rd, z, slot, c, sc.lut, 0xFFFFFFFFu, own, nbits, k);
            else {
                uint32_t blk = 0, adv = 0;
                acfirst_run<false>(rd, z, c, sc.lut, 0xFFFFFFFFu, own, nbits, blk, 0xFFFFFFFFu, adv, dummy);
            }
            st = pack_state(rd.abs_pos(), z, slot);
        }
        entry[l] = st;
        run_sub(l, st);
    }
    // pass 2: fix-up rounds (Jacobi style, like one kernel launch per round)
    uint32_t wrong0 = 0, rounds = 0, redo = 0;
    for (;;) {
        std::vector<uint32_t> todo;
        for (uint32_t l = 1; l < nsub; l++)
            if (entry[l] != exitst[l - 1]) todo.push_back(l);
        if (rounds == 0) wrong0 = (uint32_t)todo.size();
        if (todo.empty()) break;
        std::vector<uint64_t> newentry;
        for (uint32_t l : todo) newentry.push_back(exitst[l - 1]);
        for (size_t i = 0; i < todo.size(); i++) {
            entry[todo[i]] = newentry[i];
            run_sub(todo[i], newentry[i]);
        }
        redo += (uint32_t)todo.size();
        rounds++;
        if (rounds > nsub + 2) return BJ_ERR_SYNC;
    }
    if (stats) {
        stats[0] = nsub;
        stats[1] = wrong0;
        stats[2] = rounds;
        stats[3] = redo;
    }
    // pass 3: exclusive prefix sums
    std::vector<SubCount> pre(nsub);
    SubCount acc{0, {0, 0, 0}};
    for (uint32_t l = 0; l < nsub; l++) {
        pre[l] = acc;
        acc.blocks += cnt[l].blocks;
        for (int k = 0; k < 3; k++) acc.dc[k] += cnt[l].dc[k];
    }
    // pass 4: write
    uint32_t err = 0;
    BlockSink sink{};
    sink.out = coef;
    for (uint32_t l = 0; l < nsub; l++) {
        uint32_t own = l * (uint32_t)S, stop = own + (uint32_t)S < nbits ? own + (uint32_t)S : nbits;
        BitReader<HostSrc> rd;
        rd.seek(&src, b0, (uint32_t)(state_pos(entry[l]) - b0));
        int z = state_z(entry[l]), slot = state_slot(entry[l]);
        uint32_t blk = pre[l].blocks;
        int pred[3] = {pre[l].dc[0], pre[l].dc[1], pre[l].dc[2]};
        if (sc.mode == BJ_MODE_BASELINE) {
            err |= base_write_run(rd, z, slot, c, sc.lut, stop, nbits, blk, nblk_stream, pred, sink);
        } else if (sc.mode == BJ_MODE_DC_FIRST) {
            err |= dcfirst_write_run(rd, slot, c, sc.lut, stop, nbits, blk, nblk_stream, pred, sink);
        } else {
            uint32_t adv = 0;
            err |= acfirst_run<true>(rd, z, c, sc.lut, own, stop, nbits, blk, nblk_stream, adv, sink);
        }
    }
    if (acc.blocks < nblk_stream) err |= BJ_ERR_OVERRUN;
    return err;
}

// AC refinement the way the device runs it: a sequential parse over chunks of 32 blocks that only
// sees the blocks' non-zero masks and records per-block start positions, then an independent
// re-decode of every block from its recorded position (acrefine_parse_kernel / acrefine_apply_kernel).
static uint64_t nonzero_mask(const int16_t* b) {
    uint64_t m = 0;
    for (int z = 0; z < 64; z++)
        if (b[z]) m |= 1ull << z;
    return m;
}

uint32_t hs_acrefine_stream(const uint32_t* words, uint64_t n_words, uint64_t start_byte, uint64_t end_byte,
                            const SimScan* ss_, uint32_t nblk_stream, int16_t* coef) {
    ScanCtx c = make_ctx(*ss_);
    HostSrc src{words, n_words};
    const uint32_t* tab = ss_->lut + c.ac_tab[0];
    const uint32_t end_rel = (uint32_t)((end_byte - start_byte) * 8);
    std::vector<uint32_t> pos(nblk_stream);
    {   // the device's orchestration: pre-decoded window, rebuilt whenever the next block could run out of it
        std::vector<uint32_t> pre(BJ_ACR_WIN_BITS);
        uint32_t win_base = 0, pos_bits = 0, eob_run = 0;
        auto build = [&](uint32_t base) {
            win_base = base;
            for (uint32_t k = 0; k < BJ_ACR_WIN_BITS; k++) pre[k] = acrefine_predecode(src, start_byte * 8 + base + k, tab);
        };
        build(0);
        for (uint32_t cb = 0; cb < nblk_stream; cb += 32) {
            int nb = (int)std::min<uint32_t>(32, nblk_stream - cb);
            uint8_t tabs[32 * BJ_ACR_TAB_STRIDE];
            for (int i = 0; i < nb; i++)
                acrefine_build_table(nonzero_mask(coef + (size_t)(cb + i) * 64), c.ss, c.se, tabs + i * BJ_ACR_TAB_STRIDE);
            int i = 0;
            while (i < nb) {
                if (pos_bits > win_base + BJ_ACR_WIN_BITS - BJ_ACR_WIN_SLACK) build(pos_bits);
                uint32_t err = 0;
                i = acrefine_parse_window(pre.data(), win_base, win_base + BJ_ACR_WIN_BITS - BJ_ACR_WIN_SLACK, c, end_rel, tabs, i,
                                          nb, pos_bits, eob_run, pos.data() + cb, err);
                if (err) return err;
            }
        }
    }
    uint32_t err = 0;
    for (uint32_t b = nblk_stream; b-- > 0;) {  // any order: blocks are independent now
        BitReader<HostSrc> rd;
        rd.seek(&src, start_byte * 8, pos[b] & ~BJ_ACR_IN_EOBRUN);
        uint32_t eob_run = (pos[b] & BJ_ACR_IN_EOBRUN) ? 1u : 0u;
        int16_t* p = coef + (size_t)b * 64;
        err |= acrefine_block<true>(rd, c, tab, end_rel, nonzero_mask(p), eob_run, p);
    }
    return err;
}

// DC refinement (:1036-1043): bit b of the stream belongs to block b.
void hs_dcrefine_stream(const uint32_t* words, uint64_t start_byte, uint32_t nblk_stream, int al, int16_t* coef) {
    for (uint32_t b = 0; b < nblk_stream; b++) {
        uint64_t bit = start_byte * 8 + b;
        uint32_t v = (words[bit >> 5] >> (31 - (bit & 31))) & 1u;
        coef[(size_t)b * 64] = (int16_t)(coef[(size_t)b * 64] | (int16_t)(v << al));
    }
}
// EXPERIMENT SUPPORT (test-only): how far does a decoder that starts at a subsequence boundary with the
// default state (z0, slot 0) have to go before it falls into step with the true decode?  hist[k] counts
// subsequences whose distance is in [k*64, (k+1)*64) bits (last bin: never within `limit` bits).
void hs_sync_distance(const uint32_t* words, uint64_t n_words, uint64_t start_byte, uint64_t end_byte,
                      const SimScan* ss_, int sub_bits, uint32_t limit, uint32_t* hist, int nbins) {
    const SimScan& sc = *ss_;
    if (sc.mode != BJ_MODE_BASELINE) return;
    ScanCtx c = make_ctx(sc);
    HostSrc src{words, n_words};
    const uint64_t b0 = start_byte * 8, b1 = end_byte * 8;
    const uint32_t nbits = (uint32_t)(b1 - b0);
    std::vector<uint16_t> truth(nbits + 64, 0xFFFF);  // per bit position: z | slot << 7 at a true symbol start
    {
        BitReader<HostSrc> rd;
        rd.seek(&src, b0, 0);
        int z = 0, slot = 0;
        SubCount k{0, {0, 0, 0}};
        while (rd.rel < nbits) {
            truth[rd.rel] = (uint16_t)(z | (slot << 7));
            uint32_t before = rd.rel;
            sync_run<BJ_M_BASE>(rd, z, slot, c, sc.lut, 0xFFFFFFFFu, rd.rel + 1, nbits, k);
            if (rd.rel == before) break;
        }
    }
    const uint32_t nsub = (nbits + sub_bits - 1) / sub_bits;
    for (uint32_t l = 1; l < nsub; l++) {
        const uint32_t own = l * (uint32_t)sub_bits;
        BitReader<HostSrc> rd;
        rd.seek(&src, b0, own);
        int z = 0, slot = 0;
        SubCount k{0, {0, 0, 0}};
        uint32_t d = limit;
        while (rd.rel < nbits && rd.rel - own < limit) {
            if (truth[rd.rel] == (uint16_t)(z | (slot << 7))) { d = rd.rel - own; break; }
            uint32_t before = rd.rel;
            sync_run<BJ_M_BASE>(rd, z, slot, c, sc.lut, 0xFFFFFFFFu, rd.rel + 1, nbits, k);
            if (rd.rel == before) break;
        }
        int bin = (int)(d / 64);
        if (bin >= nbins) bin = nbins - 1;
        hist[bin]++;
    }
}
}
