/*
 * felics_b200.h -- C ABI of the B200-native FELICS codec engine.
 *
 * This is the drop-in boundary for the hot path of visanalexandru/felics: the
 * per-channel encode/decode loop behind the crate's compress/decompress API.
 * Every entry point names the reference interface it replaces (paths relative
 * to /root/reference/src).  Plain pointers and sizes only; no C++ or torch types.
 * The implementation is hand-written sm_100a CUDA (felics_b200/csrc); there is
 * no CPU fallback: every compute entry point fails with FELICS_ERR_CUDA when no
 * usable device exists.
 *
 * Pixel layout (both directions): row-major, host-endian samples, u8 for
 * pixel_depth 0 and u16 for pixel_depth 1; RGB is interleaved R,G,B per pixel
 * exactly like image::ImageBuffer::as_raw() (compression.rs:276, :346-352).
 * The compressed form is the reference's .fel container: 14-byte header
 * (format.rs:51-61) followed by one MSB-first bit stream (compression.rs:270-280).
 */
#ifndef FELICS_B200_H
#define FELICS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Return codes.  -1..-7 mirror `enum DecompressionError` (compression/error.rs:5-19)
 * in declaration order; the rest belong to this ABI. */
#define FELICS_OK 0
#define FELICS_ERR_IO (-1)                  /* IoError: short / truncated input (read_exact, BitRead EOF) */
#define FELICS_ERR_INVALID_VALUE (-2)       /* InvalidValue: decoded sample does not fit the pixel type */
#define FELICS_ERR_VALUE_OVERFLOW (-3)      /* ValueOverflow */
#define FELICS_ERR_INVALID_DIMENSIONS (-4)  /* InvalidDimensions */
#define FELICS_ERR_INVALID_COLOR_TYPE (-5)  /* InvalidColorType */
#define FELICS_ERR_INVALID_PIXEL_DEPTH (-6) /* InvalidPixelDepth */
#define FELICS_ERR_INVALID_SIGNATURE (-7)   /* InvalidSignature */
#define FELICS_ERR_BUFFER_TOO_SMALL (-8)    /* output capacity too small; required size returned */
#define FELICS_ERR_CUDA (-9)                /* CUDA failure; text in felics_last_error() */
#define FELICS_ERR_UNSUPPORTED (-10)        /* valid in the reference, not built yet */
#define FELICS_ERR_CORRUPT (-11)            /* the reference would panic on this stream (assert / unwrap) */
#define FELICS_ERR_INVALID_ARGUMENT (-12)

/* Header of a .fel file: `struct Header` (compression/format.rs:44-49).
 * color_type: 0 Gray, 1 Rgb (format.rs:8-11).  pixel_depth: 0 Eight, 1 Sixteen (format.rs:27-30). */
typedef struct felics_header {
    uint8_t color_type;
    uint8_t pixel_depth;
    uint32_t width;
    uint32_t height;
} felics_header;

#define FELICS_HEADER_BYTES 14

/* One context owns a device, a stream and its scratch memory.  Calls on one
 * context are serialised by the caller; distinct contexts are independent
 * (the reference is pure and re-entrant: SURVEY.md 8b "Threading"). */
typedef struct felics_ctx felics_ctx;

/* device < 0 selects the current CUDA device. */
int felics_ctx_create(int device, felics_ctx **out);
void felics_ctx_destroy(felics_ctx *ctx);
/* Run on an external stream (a cudaStream_t, e.g. torch's current stream); NULL restores the context's own. */
int felics_ctx_set_stream(felics_ctx *ctx, void *cuda_stream);
/* Thread-local text of the last failure of any call on this thread. */
const char *felics_last_error(void);

/* read_header / write_header (compression/format.rs:51-84).  Host only. */
int felics_read_header(const uint8_t *buf, size_t len, felics_header *out);
int felics_write_header(const felics_header *hdr, uint8_t *out14);
/* Number of pixel bytes an image with this header occupies. */
size_t felics_pixel_bytes(const felics_header *hdr);
/* A capacity that always suffices for felics_compress (worst-case code lengths). */
size_t felics_compress_bound(const felics_header *hdr);

/* CompressDecompress::compress / compress_image (compression/traits.rs:48-50,
 * compression.rs:255-282 Luma, :322-371 Rgb, :412-418): pixels in HOST memory,
 * whole .fel file (header included) to HOST memory.  On FELICS_ERR_BUFFER_TOO_SMALL
 * *out_len holds the required size. */
int felics_compress(felics_ctx *ctx, const void *pixels, const felics_header *hdr,
                    uint8_t *out, size_t cap, size_t *out_len);

/* decompress_image / CompressDecompress::decompress (compression.rs:420-441,
 * traits.rs:57-64): .fel bytes in HOST memory, pixels to HOST memory.
 * hdr_out is filled as soon as the header parses (also on later errors). */
int felics_decompress(felics_ctx *ctx, const uint8_t *fel, size_t len,
                      void *pixels_out, size_t cap, felics_header *hdr_out);

/* Same two calls with DEVICE-resident buffers (no host<->device copy of the
 * payload; used to time the kernels alone).  d_out and d_fel need 4-byte alignment, and the decoder loads whole
 * 32-bit words: the allocation behind d_fel must be readable up to the next multiple of four bytes after `len`. */
int felics_compress_device(felics_ctx *ctx, const void *d_pixels, const felics_header *hdr,
                           uint8_t *d_out, size_t cap, size_t *out_len);
int felics_decompress_device(felics_ctx *ctx, const uint8_t *d_fel, size_t len,
                             void *d_pixels_out, size_t cap, felics_header *hdr_out);

/* Batch of n images that share one header (the tile batches of BASELINE.json
 * configs 4/5; each image is an independent compress_image call in the
 * reference, tests/compress.rs:15-35).  Image i's pixels start at
 * pixels + i*felics_pixel_bytes(hdr); its .fel occupies
 * arena[offsets[i] .. offsets[i+1]).  `offsets` has n+1 entries (host memory).
 * *_device variants take device pointers for pixels/arena (the arena 4-byte aligned, except for batches of 8-bit gray
 * images, which accept any alignment). */
int felics_compress_batch(felics_ctx *ctx, size_t n, const void *pixels, const felics_header *hdr,
                          uint8_t *arena, size_t arena_cap, uint64_t *offsets);
int felics_compress_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header *hdr,
                                 uint8_t *d_arena, size_t arena_cap, uint64_t *offsets);
/* Decode n .fel files laid out in one arena (offsets as above, n+1 entries) that
 * all carry the header *hdr; image i goes to pixels_out + i*felics_pixel_bytes(hdr).
 * status[i] receives the per-image return code; the call returns the first
 * non-zero status (or FELICS_OK). */
int felics_decompress_batch(felics_ctx *ctx, size_t n, const uint8_t *arena, const uint64_t *offsets,
                            const felics_header *hdr, void *pixels_out, int *status);
int felics_decompress_batch_device(felics_ctx *ctx, size_t n, const uint8_t *d_arena, const uint64_t *offsets,
                                   const felics_header *hdr, void *d_pixels_out, int *status);

/* Batches of images of DIFFERENT shapes and pixel types (BASELINE.json configs[4]; the reference's own loop over the
 * image suite compresses 146 differently sized files one after the other, tests/compress.rs:74-103): image i has header
 * hdrs[i] and its pixels at pixels[i] (HOST memory); its .fel occupies arena[offsets[i] .. offsets[i+1]) (HOST memory,
 * n+1 offsets).  Images that share a header are encoded together as one batch on the device; the streams are byte-identical
 * to n separate felics_compress calls.  On FELICS_ERR_BUFFER_TOO_SMALL offsets[n] holds the size the arena needs. */
int felics_compress_batch_v(felics_ctx *ctx, size_t n, const void *const *pixels, const felics_header *hdrs,
                            uint8_t *arena, size_t arena_cap, uint64_t *offsets);
/* The inverse: n .fel files of any shapes in one arena; image i is decoded into pixels_out[i], which holds caps[i] bytes.
 * hdrs_out[i] receives the file's header whenever it parses; status[i] the per-image return code
 * (FELICS_ERR_BUFFER_TOO_SMALL when caps[i] is too small); the call returns the first non-zero status. */
int felics_decompress_batch_v(felics_ctx *ctx, size_t n, const uint8_t *arena, const uint64_t *offsets,
                              void *const *pixels_out, const size_t *caps, felics_header *hdrs_out, int *status);

/* ---- band sidecar (opt-in; NOT part of the reference .fel format) ----
 * A .fel file is one serial bit chain per plane (compression.rs:151-248), so one big image decodes on one warp.
 * felics_sidecar_build writes, for the 8-bit image compressed by the LAST felics_compress / felics_compress_device call
 * on this context (call it right after, before the input pixels are released), a side file that records the decoder
 * state at the start of every band of `band_rows` rows (0 = about 64 bands per plane; otherwise a multiple of
 * 4096 / gcd(width, 4096)).  felics_decompress_sidecar decodes the bands in parallel; the .fel bytes are unchanged and
 * still decode with felics_decompress.  A sidecar that does not belong to the file yields FELICS_ERR_CORRUPT. */
int felics_sidecar_build(felics_ctx *ctx, uint32_t band_rows, uint8_t *sidecar_out, size_t cap, size_t *out_len);
int felics_decompress_sidecar(felics_ctx *ctx, const uint8_t *fel, size_t len, const uint8_t *sidecar, size_t sidecar_len,
                              void *pixels_out, size_t cap, felics_header *hdr_out);

/* Instrumentation (not part of the reference API): per-stage device times from
 * CUDA events on the context's stream, and the number of kernel launches. */
int felics_profile_enable(felics_ctx *ctx, int on);
int felics_profile_reset(felics_ctx *ctx);
int felics_profile_stage_count(void);
const char *felics_profile_stage_name(int stage);
/* accumulated milliseconds / number of launches of `stage` since the last reset */
double felics_profile_stage_ms(felics_ctx *ctx, int stage);
uint64_t felics_profile_stage_launches(felics_ctx *ctx, int stage);
uint64_t felics_profile_total_launches(felics_ctx *ctx);

const char *felics_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FELICS_B200_H */
