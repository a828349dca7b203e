/* hiccup_b200.h -- C ABI of the B200-native hiccup encode/decode hot path.
 *
 * The reference (nhomble/hiccup) has no FFI: its boundary is the Python function API of
 * hiccup/compression.py and hiccup/codec.py.  Each entry point below names the reference code it
 * replaces (file:line under /root/reference); hiccup_b200/compression.py and hiccup_b200/codec.py
 * compose them back into the reference's eight functions, and INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative hic_status on failure;
 *     hic_last_error() returns the message of the calling thread's last failure;
 *   - all pointers are caller-owned; `d_` arguments are device pointers on the current CUDA device,
 *     `h_` arguments are host pointers; `stream` is a cudaStream_t passed as void* (NULL = default);
 *   - calls are asynchronous on `stream` unless the comment says otherwise; no hidden global state
 *     besides read-only constant tables;
 *   - images are H x W x 3 uint8, C-contiguous, channel 0 treated as R (the reference feeds BGR
 *     from cv2.imread through the same arithmetic, run.py:19);
 *   - batches are n images of one shape, contiguous.
 *
 * Coefficient layout ("zigzag blocks"): per image, luminance blocks, then Cr blocks, then Cb
 * blocks; blocks in raster block order of their (zero-padded) plane; each block is 64 int16 in the
 * reference's scan order (transform._zigzag_indices, transform.py:106-124), i.e. one 128-byte line.
 * Coefficients that fall outside the unpadded plane are zero, as they are after the reference's
 * crop (transform.py:63) and re-pad (codec.py:288,294).
 */
#ifndef HICCUP_B200_H
#define HICCUP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hic_status {
    HIC_OK = 0,
    HIC_ERR_INVALID = -1,     /* bad argument */
    HIC_ERR_CUDA = -2,        /* CUDA runtime error; message has the details */
    HIC_ERR_CAPACITY = -3,    /* an output buffer was too small; nothing past its end was written */
    HIC_ERR_CORRUPT = -4      /* a bit stream did not decode to the expected symbol count */
} hic_status;

/* ---- runtime plumbing ------------------------------------------------------------------------ */
int hic_version(void);
const char* hic_last_error(void);
int hic_device_count(int* count);
int hic_set_device(int device);
int hic_device_name(char* buf, size_t buflen);
int hic_malloc(void** d_ptr, size_t bytes);
int hic_free(void* d_ptr);
int hic_host_alloc(void** h_ptr, size_t bytes);          /* pinned */
int hic_host_free(void* h_ptr);
int hic_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
int hic_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
int hic_memset(void* d_ptr, int value, size_t bytes, void* stream);
int hic_stream_create(void** stream);
int hic_stream_destroy(void* stream);
int hic_stream_sync(void* stream);                       /* blocks the host */

/* ---- DCT-mode geometry ----------------------------------------------------------------------- */
typedef struct hic_dct_geometry {
    int32_t h, w;             /* luminance plane = image */
    int32_t hc, wc;           /* chroma planes after pyrDown: h/2, w/2 (transform.py:160-166) */
    int32_t nby_l, nbx_l;     /* luminance blocks: ceil(h/8), ceil(w/8) (transform.py:17-42) */
    int32_t nby_c, nbx_c;     /* chroma blocks */
    int64_t nb_l, nb_c;       /* blocks per plane */
    int64_t blocks_per_image; /* nb_l + 2 nb_c */
    int32_t out_h, out_w;     /* decoded image: 2 hc, 2 wc (transform.force_merge, :269-277) */
} hic_dct_geometry;
int hic_dct_geometry_of(int32_t h, int32_t w, hic_dct_geometry* out);

/* One rounding-tie candidate: a block whose float32 quantised value landed inside the error band
 * around a half-integer at one or more scan positions (bit k of mask = scan position k). */
typedef struct hic_tie_record {
    uint32_t block;           /* global block index: image * blocks_per_image + block in image */
    uint32_t reserved;
    uint64_t mask;
} hic_tie_record;

/* d_stats[0] candidates flagged (blocks), [1] coefficients re-evaluated in float64,
 * [2] coefficients whose float32 rounding was changed by the re-evaluation, [3] records dropped
 * because tie_capacity was too small (if non-zero the coefficients are NOT guaranteed exact). */
#define HIC_TIE_STATS 4

/* K1 -- fused encode transform.  Replaces compression.jpeg_compression (compression.py:16-39):
 * cv2.cvtColor RGB->YCrCb, cv2.pyrDown of both chroma planes, x-128, zero pad, 8x8 DCT-II
 * (scipy.fftpack.dct rows then columns), division by the Annex-K table, np.round, crop; plus the
 * zigzag of transform.ac_components (transform.py:260-266).  float32 butterflies; every value
 * within the float32 error band of a rounding tie is then re-evaluated by a second kernel in
 * float64 with scipy's exact operation order, so the output equals the reference's bit for bit.
 * d_ties needs room for tie_capacity records (blocks_per_image * n is always enough);
 * d_stats is HIC_TIE_STATS uint32, zeroed by this call. */
int hic_dct_forward(const uint8_t* d_rgb, int32_t n, int32_t h, int32_t w, int16_t* d_coef,
                    hic_tie_record* d_ties, uint32_t tie_capacity, uint32_t* d_stats, void* stream);

/* Zigzag blocks <-> the reference's CompressedImage planes (model.py:38-74): int32, cropped to the
 * unpadded plane shape, block (by,bx) coefficient (u,v) at [8by+u, 8bx+v] (transform.py:45-64). */
int hic_blocks_to_planes(const int16_t* d_coef, int32_t n, int32_t h, int32_t w, int32_t* d_lum,
                         int32_t* d_cr, int32_t* d_cb, void* stream);
int hic_planes_to_blocks(const int32_t* d_lum, const int32_t* d_cr, const int32_t* d_cb, int32_t n,
                         int32_t h, int32_t w, int16_t* d_coef, void* stream);

/* K7+K8 -- decode transform.  Replaces compression.jpeg_decompression (compression.py:42-56):
 * coefficient * table, idct rows then columns, /256, +128, astype(uint8) (truncate, wrap),
 * cv2.pyrUp of both chroma planes, crop of the luminance plane, cv2.cvtColor YCrCb->RGB.
 * d_y (n*h*w), d_cr, d_cb (n*hc*wc each) are scratch planes; d_rgb_out is n*out_h*out_w*3.
 * float32 butterflies: pixels are within +-1 LSB of the reference's (or +-255 where its uint8 cast
 * wraps); tests count both. */
int hic_dct_inverse(const int16_t* d_coef, int32_t n, int32_t h, int32_t w, uint8_t* d_y,
                    uint8_t* d_cr, uint8_t* d_cb, uint8_t* d_rgb_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HICCUP_B200_H */
