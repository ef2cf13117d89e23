// bj_host.cu -- host-side helpers of the C ABI (no device code).
//
// The marker walk stays in Python (pyjpegdecoder_b200/parser.py, mirroring jpeg_decoder.py:78-110), but
// finding the end of an entropy-coded segment means scanning hundreds of kilobytes for the first 0xFF
// that is not followed by 0x00 or RSTn (the bytes the reference's main loop skips, :93).  That scan is
// done here with memchr instead of a Python regular expression (~15x faster).
#include <stdint.h>
#include <string.h>

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
