// bj_host.cu -- host-side helpers of the C ABI (no device code).
//
// The marker walk stays in Python (pyjpegdecoder_b200/parser.py, mirroring jpeg_decoder.py:78-110), but
// finding the end of an entropy-coded segment means scanning hundreds of kilobytes for the first 0xFF
// that is not followed by 0x00 or RSTn (the bytes the reference's main loop skips, :93).  That scan is
// done here with memchr instead of a Python regular expression (~15x faster).
#include <stdint.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/b200jpeg.h"

extern "C" {

// First position p >= pos with data[p] == 0xFF and data[p+1] not in {0x00, 0xD0..0xD7}; n if none.
uint64_t bj_host_find_marker(const uint8_t* data, uint64_t n, uint64_t pos) {
    while (pos + 1 < n) {
        const uint8_t* q = (const uint8_t*)memchr(data + pos, 0xFF, (size_t)(n - 1 - pos));
        if (!q) return n;
        uint64_t p = (uint64_t)(q - data);
        uint8_t m = data[p + 1];
        if (m != 0x00 && (m < 0xD0 || m > 0xD7)) return p;
        pos = p + 2;
    }
    return n;
}

// Number of SOS markers (0xFF 0xDA byte pairs, non-overlapping) in data[pos:n) -- bytes.count(SOS) of
// jpeg_decoder.py:635-637.
uint32_t bj_host_count_sos(const uint8_t* data, uint64_t n, uint64_t pos) {
    uint32_t c = 0;
    while (pos + 1 < n) {
        const uint8_t* q = (const uint8_t*)memchr(data + pos, 0xFF, (size_t)(n - 1 - pos));
        if (!q) break;
        uint64_t p = (uint64_t)(q - data);
        if (data[p + 1] == 0xDA) {
            c++;
            pos = p + 2;
        } else {
            pos = p + 1;
        }
    }
    return c;
}

}  // extern "C"

// ---- batch marker walk ---------------------------------------------------------------------------------
// Mirrors the control flow of JpegDecoder.__init__ (jpeg_decoder.py:78-110) at the byte level only: it
// reports where every segment and every entropy-coded run lies; interpreting the segment payloads stays
// in Python (parser.py).  Entries: marker 0x100 = entropy-coded run [start, end); otherwise a marker
// segment whose payload is [start, end) (start points after the 2-byte length).  Returns the number of
// entries, or -1 if the file does not start with FFD8FF, or -2 if max_entries is too small.
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <thread>
#include <vector>

static int walk_one(const uint8_t* d, uint64_t n, bj_host_entry* out, int max_entries) {
    if (n < 3 || d[0] != 0xFF || d[1] != 0xD8 || d[2] != 0xFF) return -1;
    int k = 0;
    uint64_t pos = 2;
    for (;;) {
        if (pos >= n) break;
        const uint8_t* q = (const uint8_t*)memchr(d + pos, 0xFF, (size_t)(n - pos));
        if (!q) break;
        pos = (uint64_t)(q - d);
        if (pos + 1 >= n) break;
        uint32_t m = d[pos + 1];
        if (m == 0xFF) {  // fill byte in front of a marker (parser.py does the same)
            pos += 1;
            continue;
        }
        pos += 2;
        if (m == 0x00 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (k >= max_entries) return -2;
        if (m == 0xD9) {
            out[k++] = bj_host_entry{pos, pos, m, 0};
            break;
        }
        if (pos + 2 > n) break;
        uint64_t size = ((uint64_t)d[pos] << 8 | d[pos + 1]);
        size = size >= 2 ? size - 2 : 0;
        pos += 2;
        uint64_t end = pos + size < n ? pos + size : n;
        out[k++] = bj_host_entry{pos, end, m, 0};
        if (m == 0xDD) {
            pos += 2;  // the reference advances by 2, not by the segment length (:476-477)
        } else if (m == 0xDA) {
            pos += size;
            if (pos > n) pos = n;  // declared SOS length past EOF: empty run; parser.py raises CorruptedJpeg for it
            if (k >= max_entries) return -2;
            uint64_t e = bj_host_find_marker(d, n, pos);
            out[k++] = bj_host_entry{pos, e, 0x100, 0};
            pos = e;
        } else {
            pos += size;
        }
    }
    return k;
}

extern "C" int bj_host_walk(const uint8_t* data, uint64_t n, bj_host_entry* entries, int max_entries) {
    return walk_one(data, n, entries, max_entries);
}

// Walk n_files files stored in one buffer (file i = raw[off[i], off[i] + size[i])) with n_threads host
// threads.  entries: [n_files][max_entries]; counts[i] = result of bj_host_walk for file i.
extern "C" void bj_host_walk_batch(const uint8_t* raw, const uint64_t* off, const uint64_t* size, int n_files,
                                   bj_host_entry* entries, int max_entries, int32_t* counts, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
    auto work = [&](int t) {
        for (int i = t; i < n_files; i += n_threads)
            counts[i] = walk_one(raw + off[i], size[i], entries + (size_t)i * max_entries, max_entries);
    };
    if (n_threads == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
}

// 128-bit hash of what determines a file's parse: the sequence of markers, the payload of every segment that
// changes the parse (SOFn, DHT, DQT, DRI, SOS, DNL, EOI) and one tag per entropy-coded run.  Files with equal
// hashes share one parsed template on the host (fastplan.py); two independent 64-bit multiply-mix streams make an
// accidental collision a 2^-128 event.
static inline uint64_t mix64(uint64_t h, uint64_t v, uint64_t k) {
    h ^= v;
    h *= k;
    h ^= h >> 29;
    return h;
}
static void key_hash_one(const uint8_t* d, const bj_host_entry* e, int count, uint64_t out[2]) {
    uint64_t a = 0x9E3779B97F4A7C15ull, b = 0xC2B2AE3D27D4EB4Full;
    for (int i = 0; i < count; i++) {
        const uint32_t m = e[i].marker;
        if (m == 0x100) {
            a = mix64(a, 0xE0E0E0E0ull, 0xFF51AFD7ED558CCDull);
            b = mix64(b, 0x0E0E0E0Eull, 0xC4CEB9FE1A85EC53ull);
            continue;
        }
        const bool key = (m >= 0xC0 && m <= 0xCF) || m == 0xDB || m == 0xDD || m == 0xDA || m == 0xDC || m == 0xD9;
        if (!key) continue;
        const uint64_t len = e[i].end - e[i].start;
        a = mix64(a, ((uint64_t)m << 32) | (len & 0xFFFFFFFFull), 0xFF51AFD7ED558CCDull);
        b = mix64(b, ((uint64_t)len << 8) | m, 0xC4CEB9FE1A85EC53ull);
        const uint8_t* p = d + e[i].start;
        uint64_t k = 0;
        for (; k + 8 <= len; k += 8) {
            uint64_t v;
            memcpy(&v, p + k, 8);
            a = mix64(a, v, 0xFF51AFD7ED558CCDull);
            b = mix64(b, v, 0xC4CEB9FE1A85EC53ull);
        }
        uint64_t v = 0;
        for (int s = 0; k < len; k++, s += 8) v |= (uint64_t)p[k] << s;
        a = mix64(a, v, 0xFF51AFD7ED558CCDull);
        b = mix64(b, v ^ 0xA5A5A5A5A5A5A5A5ull, 0xC4CEB9FE1A85EC53ull);
    }
    out[0] = a;
    out[1] = b;
}

// bj_host_walk_batch + the key hash of every file (key_hash: [n_files][2]; files whose walk failed get 0, 0).
extern "C" void bj_host_walk_batch_keys(const uint8_t* raw, const uint64_t* off, const uint64_t* size, int n_files,
                                        bj_host_entry* entries, int max_entries, int32_t* counts, uint64_t* key_hash,
                                        int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
    auto work = [&](int t) {
        for (int i = t; i < n_files; i += n_threads) {
            bj_host_entry* e = entries + (size_t)i * max_entries;
            counts[i] = walk_one(raw + off[i], size[i], e, max_entries);
            key_hash[2 * i] = key_hash[2 * i + 1] = 0;
            if (counts[i] > 0) key_hash_one(raw + off[i], e, counts[i], key_hash + 2 * i);
        }
    };
    if (n_threads == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
}

// Copy with streaming (non-temporal) stores.  The packed buffer is only ever read by the GPU's copy engine, so
// caching it on the host is useless, and regular stores would first READ every destination line (write-allocate):
// the gather of the streaming front end is bound by host memory bandwidth, and this takes a third of its traffic
// away.  dst must be 16-byte aligned (the packed buffer is page aligned and every file starts at a multiple of 16).
static void copy_streaming(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__SSE2__)
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
        }
        if (i < n) memcpy(dst + i, src + i, n - i);
        _mm_sfence();
        return;
    }
#endif
    memcpy(dst, src, n);
}

// Copy n_files separate host buffers into one packed buffer (dst + off[i]) with n_threads host threads:
// a single Python-level copy loop tops out near 5 GB/s, far below what the H2D copy that follows can take.
extern "C" void bj_host_pack(const uint8_t* const* src, const uint64_t* size, const uint64_t* off, int n_files, uint8_t* dst,
                             int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
    auto work = [&](int t) {
        for (int i = t; i < n_files; i += n_threads) copy_streaming(dst + off[i], src[i], size[i]);
    };
    if (n_threads == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
}

// bj_host_pack and bj_host_walk_batch_keys in one pass: every thread copies a file into the packed buffer and walks
// the SOURCE right away, while the copy's reads still sit in that core's cache -- the file bytes are read from
// memory once on the host instead of twice (and the streaming stores of the copy never come back into the cache).
extern "C" void bj_host_pack_walk_keys(const uint8_t* const* src, const uint64_t* size, const uint64_t* off, int n_files,
                                       uint8_t* dst, bj_host_entry* entries, int max_entries, int32_t* counts,
                                       uint64_t* key_hash, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
    auto work = [&](int t) {
        for (int i = t; i < n_files; i += n_threads) {
            copy_streaming(dst + off[i], src[i], size[i]);
            bj_host_entry* e = entries + (size_t)i * max_entries;
            counts[i] = walk_one(src[i], size[i], e, max_entries);
            key_hash[2 * i] = key_hash[2 * i + 1] = 0;
            if (counts[i] > 0) key_hash_one(src[i], e, counts[i], key_hash + 2 * i);
        }
    };
    if (n_threads == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
}



// ---- files straight from disk (page cache) into the packed buffer ----------------------------------------
// Python-level open/read/close costs ~45 us per file even from 16 threads (the interpreter lock); these run the same
// system calls from plain host threads.

// size[i] = size of paths[i] in bytes, or -errno.
extern "C" void bj_host_stat_files(const char* const* paths, int n_files, int64_t* size, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
    auto work = [&](int t) {
        for (int i = t; i < n_files; i += n_threads) {
            struct stat st;
            size[i] = (stat(paths[i], &st) == 0) ? (int64_t)st.st_size : -(int64_t)errno;
        }
    };
    if (n_threads == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
}

static int read_whole(const char* path, uint8_t* dst, uint64_t n) {
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return errno ? errno : EIO;
    uint64_t got = 0;
    int err = 0;
    while (got < n) {
        const ssize_t k = read(fd, dst + got, n - got);
        if (k < 0) {
            if (errno == EINTR) continue;
            err = errno ? errno : EIO;
            break;
        }
        if (k == 0) {  // the file shrank after it was measured
            err = EIO;
            break;
        }
        got += (uint64_t)k;
    }
    close(fd);
    return err;
}

// Read every file into dst + off[i] (size[i] bytes, as measured by bj_host_stat_files) and, when entries != NULL,
// walk and hash it right away (bj_host_walk_batch_keys), while the kernel's copy is still in that core's cache.
// status[i] = 0 or the errno of the failed open/read.
extern "C" void bj_host_read_files(const char* const* paths, const uint64_t* size, const uint64_t* off, int n_files, uint8_t* dst,
                                   int32_t* status, bj_host_entry* entries, int max_entries, int32_t* counts, uint64_t* key_hash,
                                   int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
    auto work = [&](int t) {
        for (int i = t; i < n_files; i += n_threads) {
            uint8_t* d = dst + off[i];
            status[i] = read_whole(paths[i], d, size[i]);
            if (!entries) continue;
            bj_host_entry* e = entries + (size_t)i * max_entries;
            counts[i] = status[i] ? 0 : walk_one(d, size[i], e, max_entries);
            key_hash[2 * i] = key_hash[2 * i + 1] = 0;
            if (counts[i] > 0) key_hash_one(d, e, counts[i], key_hash + 2 * i);
        }
    };
    if (n_threads == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
}
