// bj_unstuff.cu -- device byte un-stuffing and restart-marker removal for all scans of a batch.
//
// Replaces the reader side of bits_generator/get_bits (jpeg_decoder.py:654-695): inside entropy-coded
// data the byte after 0xFF is dropped (:676-677) and restart markers are stepped over (:667-669).
// Doing this once, as a stream compaction, lets every later decode pass read the bitstream with
// plain 32-bit loads and funnel shifts.
//
// Three launches: per-tile counts (kept bytes, restart markers) -> exclusive scan of the tile sums
// -> scatter.  Output is a byte stream stored as BIG-ENDIAN 32-bit words (stream byte k lives at byte
// address k ^ 3), so that a native uint32 load yields the next 32 bits MSB first.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200jpeg.h"

extern "C" bj_status bj_set_cuda_error(cudaError_t e, const char* where);

namespace {

constexpr int kTile = BJ_UNSTUFF_TILE;  // raw bytes per CTA
constexpr int kThreads = kTile / 16;    // 16 bytes per thread
constexpr uint64_t kByteMask = (1ull << 40) - 1;

// Per-byte flags of a thread's 16 bytes, kept in the bit order the SWAR classification produces them in: byte i
// (word k = i >> 2, byte j = i & 3 of that word) is bit 8 j + k.  Nothing ever needs them in natural order: counts are
// popcounts, and the scatter loop tests a compile-time bit per byte.
#define BJ_FLAG_BIT(i) (8 * ((i) & 3) + ((i) >> 2))
struct Flags {
    uint32_t keep;    // byte i is kept
    uint32_t marker;  // byte i is the second byte of a restart marker
};

// 0x80 in every byte of w that is zero (exact, no carries between bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t w) { return ~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w | 0x7F7F7F7Fu); }
// the four per-word byte masks (0x80 per flagged byte) of a 16-byte chunk -> one word with byte i at bit BJ_FLAG_BIT(i)
__device__ __forceinline__ uint32_t gather16(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    return (m0 >> 7) | (m1 >> 6) | (m2 >> 5) | (m3 >> 4);
}
// flags of byte i moved to byte i + 1 (the flag of byte 15 falls out)
__device__ __forceinline__ uint32_t next_byte(uint32_t f) { return ((f << 8) & 0x0F0F0F00u) | ((f >> 23) & 0x0Eu); }
// flags of byte i moved to byte i - 1 (the flag of byte 0 falls out)
__device__ __forceinline__ uint32_t prev_byte(uint32_t f) { return ((f >> 8) & 0x000F0F0Fu) | ((f & 0x0Eu) << 23); }
// bytes [lo, hi) of the chunk, 0 <= lo <= hi <= 16
__device__ __forceinline__ uint32_t range16(uint32_t lo, uint32_t hi) {
    uint32_t f = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        if ((uint32_t)i >= lo && (uint32_t)i < hi) f |= 1u << BJ_FLAG_BIT(i);
    return f;
}

// Classify the 16 bytes starting at offset `off` of the 16-byte aligned region `seg`; the scan's bytes are
// [lead, len) of that region (lead = raw_off & 15), everything else is dropped.  Word-parallel: the three byte
// classes that matter (0xFF, 0x00, RSTn's second byte 0xD0..0xD7) are found with carry-free byte tests on the four
// words, everything else is mask arithmetic.  A kept byte is one that is inside the scan and is neither a stuffed
// zero (previous byte 0xFF), nor the 0xFF or the second byte of a restart marker (jpeg_decoder.py:667-669, :676-677).
// prev / next: the byte before / after the chunk (they come from the neighbouring lanes where possible).
__device__ __forceinline__ Flags classify(uint4 data, uint32_t off, uint32_t lead, uint32_t len, uint32_t prev, uint32_t next) {
    Flags f{0u, 0u};
    if (off >= len) return f;
    const uint32_t w[4] = {data.x, data.y, data.z, data.w};
    const uint32_t lo = lead > off ? lead - off : 0u, hi = len - off < 16u ? len - off : 16u;
    const uint32_t inside = (lo == 0u && hi == 16u) ? 0x0F0F0F0Fu : range16(lo, hi);
    uint32_t ffm[4], zm[4], rm[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ffm[k] = zero_bytes(~w[k]);
        zm[k] = zero_bytes(w[k]);
        rm[k] = zero_bytes((w[k] & 0xF8F8F8F8u) ^ 0xD0D0D0D0u);
    }
    const uint32_t ff = gather16(ffm[0], ffm[1], ffm[2], ffm[3]) & inside;
    const uint32_t zero = gather16(zm[0], zm[1], zm[2], zm[3]) & inside;
    const uint32_t rst = gather16(rm[0], rm[1], rm[2], rm[3]) & inside;
    const bool prev_ff = (off > lead) && prev == 0xFFu;
    const bool next_rst = (off + 16u < len) && (next & 0xF8u) == 0xD0u;
    const uint32_t after_ff = next_byte(ff) | (prev_ff ? 1u : 0u);                       // bytes that follow a 0xFF
    const uint32_t before_rst = prev_byte(rst) | (next_rst ? (1u << BJ_FLAG_BIT(15)) : 0u);  // bytes followed by D0..D7
    const uint32_t stuffed = after_ff & zero;
    const uint32_t rst2 = after_ff & rst;
    const uint32_t rst1 = ff & before_rst;
    f.keep = inside & ~(stuffed | rst2 | rst1);
    f.marker = rst2;
    return f;
}

// The chunk of a thread plus the bytes around it: neighbours inside the warp hand them over with shuffles, the warp's
// first and last lanes read them from memory.
__device__ __forceinline__ Flags load_classify(const uint8_t* __restrict__ seg, uint32_t off, uint32_t lead, uint32_t len, uint4& data) {
    const int lane = threadIdx.x & 31;
    data = make_uint4(0, 0, 0, 0);
    if (off < len) data = __ldg(reinterpret_cast<const uint4*>(seg + off));
    uint32_t prev = __shfl_up_sync(0xffffffffu, data.w >> 24, 1);
    uint32_t next = __shfl_down_sync(0xffffffffu, data.x & 0xFFu, 1);
    if (lane == 0) prev = (off > lead && off < len) ? (uint32_t)__ldg(seg + off - 1) : 0u;
    if (lane == 31) next = (off + 16u < len) ? (uint32_t)__ldg(seg + off + 16) : 0u;
    return classify(data, off, lead, len, prev, next);
}

__global__ void __launch_bounds__(kThreads) unstuff_count_kernel(const uint8_t* __restrict__ raw,
                                                                const bj_scan* __restrict__ scans,
                                                                const uint32_t* __restrict__ tile_scan,
                                                                uint64_t* __restrict__ tile_sum) {
    const uint32_t tile = blockIdx.x;
    const bj_scan& sc = scans[tile_scan[tile]];
    const uint32_t off = (tile - sc.tile0) * kTile + threadIdx.x * 16;
    const uint32_t lead = (uint32_t)(sc.raw_off & 15);
    uint4 d;
    Flags f = load_classify(raw + (sc.raw_off & ~15ull), off, lead, lead + sc.raw_len, d);
    uint64_t v = (uint64_t)__popc(f.keep) | ((uint64_t)__popc(f.marker) << 40);
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ uint64_t ws[kThreads / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < kThreads / 32; i++) s += ws[i];
        tile_sum[tile] = s;
    }
}

// Exclusive scan of n 64-bit values by one CTA (n is a few hundred thousand at most).  tile_sum has
// n + 1 entries: on return entry i holds the sum of the first i inputs.
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint64_t* __restrict__ v, int n) {
    __shared__ uint64_t part[1024];
    const int t = threadIdx.x;
    const int chunk = (n + 1023) / 1024;
    const int lo = min(t * chunk, n), hi = min(lo + chunk, n);
    uint64_t s = 0;
    for (int i = lo; i < hi; i++) s += v[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        uint64_t a = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += a;
        __syncthreads();
    }
    uint64_t run = part[t] - s;
    for (int i = lo; i < hi; i++) {
        uint64_t x = v[i];
        v[i] = run;
        run += x;
    }
    if (t == 1023) v[n] = part[1023];
}

__global__ void __launch_bounds__(kThreads) unstuff_scatter_kernel(const uint8_t* __restrict__ raw,
                                                                  const bj_scan* __restrict__ scans,
                                                                  const uint32_t* __restrict__ tile_scan,
                                                                  const uint64_t* __restrict__ tile_sum,
                                                                  uint32_t* __restrict__ words,
                                                                  uint64_t* __restrict__ stream_start) {
    __shared__ uint8_t sbuf[kTile + 8];
    __shared__ uint32_t wk[kThreads / 32], wm[kThreads / 32];
    const uint32_t tile = blockIdx.x;
    const uint32_t sidx = tile_scan[tile];
    const bj_scan& sc = scans[sidx];
    const uint32_t off = (tile - sc.tile0) * kTile + threadIdx.x * 16;
    const uint32_t lead = (uint32_t)(sc.raw_off & 15);
    uint4 d;
    Flags f = load_classify(raw + (sc.raw_off & ~15ull), off, lead, lead + sc.raw_len, d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // exclusive prefix of (kept, markers) over the CTA
    uint32_t nk = __popc(f.keep), nm = __popc(f.marker);
    uint32_t pk = nk, pm = nm;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(0xffffffffu, pk, o), b = __shfl_up_sync(0xffffffffu, pm, o);
        if (lane >= o) { pk += a; pm += b; }
    }
    if (lane == 31) { wk[warp] = pk; wm[warp] = pm; }
    __syncthreads();
    uint32_t bk = 0, bm = 0;
    for (int i = 0; i < warp; i++) { bk += wk[i]; bm += wm[i]; }
    uint32_t total = 0;
    for (int i = 0; i < kThreads / 32; i++) total += wk[i];
    const uint32_t my_k = bk + pk - nk, my_m = bm + pm - nm;

    const uint64_t tsum = tile_sum[tile];
    const uint64_t obase = tsum & kByteMask;                                      // output byte of the tile's first kept byte
    const uint64_t mbase = (tsum >> 40) - (tile_sum[sc.tile0] >> 40);            // restart markers of this scan before the tile
    if (off == 0) stream_start[sc.stream0] = obase;                               // stream 0 starts with the scan (off is region-relative)
    const uint32_t shift = (uint32_t)(obase & 3);                                 // keep smem aligned with the output words
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&d);
    uint32_t k = my_k, m = my_m;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (f.keep & (1u << BJ_FLAG_BIT(i))) sbuf[shift + k++] = b[i];
        if (f.marker & (1u << BJ_FLAG_BIT(i))) {
            uint64_t ord = mbase + m++ + 1;  // stream that starts right after this marker
            if (ord < sc.n_streams) stream_start[sc.stream0 + ord] = obase + k;
        }
    }
    __syncthreads();
    // write out: whole words with 32-bit stores, the ragged ends byte by byte (address ^ 3)
    const uint64_t o0 = obase, o1 = obase + total;
    const uint64_t w_lo = (o0 + 3) >> 2, w_hi = o1 >> 2;  // words [w_lo, w_hi) are entirely ours
    const uint64_t abase = o0 & ~3ull;                    // output byte that sbuf[0] corresponds to
    if (w_hi > w_lo) {
        for (uint64_t w = w_lo + threadIdx.x; w < w_hi; w += kThreads) {
            const uint8_t* s = sbuf + (w * 4 - abase);
            words[w] = ((uint32_t)s[0] << 24) | ((uint32_t)s[1] << 16) | ((uint32_t)s[2] << 8) | (uint32_t)s[3];
        }
    }
    uint8_t* wb = reinterpret_cast<uint8_t*>(words);
    const uint64_t head_end = (w_lo << 2) < o1 ? (w_lo << 2) : o1;  // head bytes [o0, head_end)
    if (threadIdx.x < 4) {
        uint64_t o = o0 + threadIdx.x;
        if (o < head_end) wb[o ^ 3] = sbuf[o - abase];
    } else if (threadIdx.x < 8 && w_hi >= w_lo) {                   // tail bytes [4*w_hi, o1)
        uint64_t o = (w_hi << 2) + (threadIdx.x - 4);
        if (o < o1) wb[o ^ 3] = sbuf[o - abase];
    }
}

}  // namespace

extern "C" bj_status bj_unstuff(const uint8_t* raw, const bj_scan* scans, int n_scans, const uint32_t* tile_scan,
                                int n_tiles, uint64_t* tile_sum, uint32_t* words_out, uint64_t* stream_start,
                                uint64_t* stream_end, int n_streams_total, void* stream) {
    if (!raw || !scans || n_scans <= 0 || !tile_scan || n_tiles <= 0 || !tile_sum || !words_out || !stream_start ||
        !stream_end || n_streams_total <= 0)
        return BJ_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stream_start, 0xFF, sizeof(uint64_t) * (size_t)n_streams_total, st);
    if (e != cudaSuccess) return bj_set_cuda_error(e, "bj_unstuff/memset");
    unstuff_count_kernel<<<n_tiles, kThreads, 0, st>>>(raw, scans, tile_scan, tile_sum);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tile_sum, n_tiles);
    unstuff_scatter_kernel<<<n_tiles, kThreads, 0, st>>>(raw, scans, tile_scan, tile_sum, words_out, stream_start);
    e = cudaGetLastError();
    if (e != cudaSuccess) return bj_set_cuda_error(e, "bj_unstuff/launch");
    return BJ_OK;
}
