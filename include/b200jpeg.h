/*
 * b200jpeg.h -- C ABI of the B200-native JPEG decode hot path (libb200jpeg.so).
 *
 * This is the drop-in boundary for the hot path of tbpaolini/PyJpegDecoder (jpeg_decoder.py):
 * everything between "scan descriptors parsed on the host" and "uint8 RGB pixels".  The reference has
 * no FFI of its own (it is one pure-Python file); each entry point below names the reference
 * function(s) whose work it replaces.  The Python host layer (pyjpegdecoder_b200/) binds these with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch/CUDA-runtime types in signatures
 *     (a CUDA stream is passed as void*; NULL = the legacy default stream).
 *   - All buffers are DEVICE pointers owned by the caller unless a name ends in _host.
 *     The library allocates nothing per call; work space sizes come from bj_*_workspace_bytes().
 *   - Every call is asynchronous on the given stream and returns a bj_status (0 = ok) that only
 *     reports launch/argument errors.  Data errors (corrupt entropy data, ...) are written to the
 *     per-image device error words (BJ_ERR_* bits) and mapped by the host to the reference's
 *     exception classes (jpeg_decoder.py:1714-1725).
 *   - Re-entrant per (device, stream); no global mutable state.
 *
 * Data layout in HBM
 *   coefficient buffer  int16, 128-byte blocks of 64 coefficients in ZIG-ZAG order, quantised.
 *                       Per image the blocks are stored MCU-major over the padded MCU grid:
 *                       block index = coef_block0 + mcu * blocks_per_mcu + slot, slots in frame
 *                       component order, h*v blocks per component, row-major inside the MCU
 *                       (the interleaved scan order of jpeg_decoder.py:774-805/:875).
 *   sample buffer       int16, same block indexing, each block 8x8 row-major [y][x] after
 *                       de-zigzag * Q, IDCT, round, +128 (jpeg_decoder.py:869-872).
 *   output              uint8 row-major (H, W, 3) or (H, W); the reference's (W, H, 3) array
 *                       (jpeg_decoder.py:626, :1373-1386) is the transposed VIEW of it.
 */
#ifndef B200JPEG_H
#define B200JPEG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BJ_VERSION 100

typedef int bj_status;
#define BJ_OK 0
#define BJ_E_ARG 1      /* bad argument */
#define BJ_E_CUDA 2     /* CUDA runtime error, see bj_last_cuda_error() */
#define BJ_E_NOGPU 3    /* no CUDA device */

/* per-image device error word bits */
#define BJ_ERR_BAD_CODE 1u      /* no Huffman code within 16 bits   -> CorruptedJpeg (:718-719, :957-958) */
#define BJ_ERR_OVERRUN 2u       /* entropy data ended early          -> IndexError in the reference */
#define BJ_ERR_RST_COUNT 4u     /* fewer restart markers than the MCU count requires */
#define BJ_ERR_SYNC 8u          /* internal: speculative decode did not converge (host retries) */
#define BJ_ERR_COEF_INDEX 16u   /* coefficient index ran past 63 in a progressive scan */

#define BJ_MAX_COMP 3
#define BJ_MAX_SLOTS 10 /* blocks per MCU (T.81 limit) */

/* ---------------------------------------------------------------------------------------------
 * Image geometry for the pixel stages.  Filled by the host from SOF/SOS (jpeg_decoder.py:112-247,
 * :583-632).  72 bytes, 8-byte aligned.
 * ------------------------------------------------------------------------------------------- */
typedef struct bj_image {
    uint64_t coef_block0;  /* first block of this image in the coefficient / sample buffers */
    uint64_t out_offset;   /* offset of pixel (0,0) in the output buffer, in output ELEMENTS (uint8 or int16) */
    uint32_t out_pitch;    /* output elements per row (>= width * channels) */
    uint32_t width, height;
    uint32_t mcus_x, mcus_y;      /* padded MCU grid (interleaved geometry, :609-611) */
    uint32_t qtab[BJ_MAX_COMP];   /* index of each component's table in the qtab buffer (64 int16, zig-zag order) */
    uint8_t ncomp;                /* 1 (greyscale) or 3 (YCbCr) */
    uint8_t hs[BJ_MAX_COMP];      /* sampling factors, 1 or 2 (forced to 1 when ncomp == 1); hmax/hs and
                                     vmax/vs in {1,2}; at most two distinct upsampling kinds per image */
    uint8_t vs[BJ_MAX_COMP];
    uint8_t hmax, vmax;
    uint8_t blocks_per_mcu;
    uint8_t slot0[BJ_MAX_COMP];   /* first block slot of each component inside an MCU */
    uint16_t strip_mcus;          /* MCUs handled by one CTA of the pixel kernels (host-chosen, <= 192 / blocks_per_mcu) */
    uint16_t strips_per_row;      /* ceil(mcus_x / strip_mcus) */
    uint32_t reserved;
} bj_image;

/* Output selector of bj_pixels(). */
#define BJ_OUT_RGB 0     /* uint8: fused IDCT + upsample + YCbCr->RGB (+ clamp)            */
#define BJ_OUT_SAMPLES 1 /* int16 sample buffer: de-zigzag, dequantise, IDCT, +128 only    */
#define BJ_OUT_CANVAS 2  /* int16 (H, W, ncomp): Y/Cb/Cr after upsampling, before colour   */
/* Input selector */
#define BJ_IN_COEF 0     /* coefficient buffer (zig-zag, quantised) */
#define BJ_IN_SAMPLES 1  /* sample buffer produced by BJ_OUT_SAMPLES */

int bj_version(void);
int bj_sizeof(int what); /* 0: sizeof(bj_image) -- lets a binding verify its struct mirror */
const char* bj_last_cuda_error(void);

/*
 * Pixel stages.  Replaces, for a whole batch of images in one launch:
 *   undo_zigzag * Q              jpeg_decoder.py:1648-1662, :869, :1347-1348   (int16 product wraps)
 *   InverseDCT.__call__          :1561-1573   (fp64 sum in numpy's pairwise order, round-half-even, +128)
 *   ResizeGrid.__call__          :1588-1626   (griddata piece-wise linear 8->16 on Qhull's triangulation)
 *   YCbCr_to_RGB + clip + crop   :1683-1700, :1372-1386
 * and the loops that drive them (:868-891 baseline, :1306-1366 progressive final stage).
 *
 *   images      device array of n_images bj_image
 *   in          coefficient buffer (BJ_IN_COEF) or sample buffer (BJ_IN_SAMPLES)
 *   qtabs       int16 quantisation tables, 64 entries each, zig-zag order as in the DQT segment
 *   idct_table_t 4096 doubles [u][v][x][y]: the reference's InverseDCT.idct_table (:1541-1553,
 *               indexed [x][y][u][v] there) TRANSPOSED so that the 64 output samples of one (u,v)
 *               are contiguous; computed by the host with the reference's expression; used only to
 *               resolve samples that the fp32 fast path cannot round safely.
 *   out         BJ_OUT_RGB: uint8 (H, W, 3|1) per image at out_offset/out_pitch;
 *               BJ_OUT_SAMPLES: int16 sample buffer (block indexing as the coefficient buffer);
 *               BJ_OUT_CANVAS: int16 (H, W, ncomp) per image at out_offset/out_pitch (int16 elements)
 *   max_strips  max over images of mcus_y * strips_per_row (grid x size)
 *   stats       optional device uint32[4]: [0] blocks recomputed exactly, [1] pixels recomputed exactly
 */
bj_status bj_pixels(const bj_image* images, int n_images, int max_strips, const void* in, int in_kind,
                    const int16_t* qtabs, const double* idct_table_t, void* out, int out_kind,
                    uint32_t* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200JPEG_H */
