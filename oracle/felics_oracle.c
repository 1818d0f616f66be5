/*
 * felics_oracle.c -- CPU restatement of the visanalexandru/felics channel codec.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the B200 path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (felics_b200/) never links, imports
 * or calls anything in oracle/.
 *
 * Parity status: the reference is Rust and no Rust toolchain exists in this
 * image, so the reference itself cannot be compiled or run here.  This
 * restatement is pinned against
 *   - every unit vector of the reference's own tests (tests/test_oracle_*.py),
 *   - the compressed-size tables published in DOC.md:385-396 and :469-477
 *     (reproduced to the byte from the image-suite TIFFs),
 * which pins the predictor, classes, code lengths, k tie-break, count halving,
 * colour transform, 64-bit channel preamble, 14-byte header and final padding.
 * The BYTE ORDER of the bit stream rests on the documented BigEndian contract of
 * the third-party crate bitstream-io 2.4.2 (Cargo.lock:264-267; not vendored):
 * most significant bit first, write_unary0(v) = v ones then a zero,
 * write_signed(32) = two's complement big-endian, byte_align pads with zeros.
 * No reference test pins file bytes, so byte-level parity against a running
 * Rust `cfelics` is "parity unpinned" (see DESIGN.md).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference/src).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* error codes mirror compression/error.rs:5-19 in declaration order */
enum {
    FO_OK = 0,
    FO_IO_ERROR = -1,
    FO_INVALID_VALUE = -2,
    FO_VALUE_OVERFLOW = -3,
    FO_INVALID_DIMENSIONS = -4,
    FO_INVALID_COLOR_TYPE = -5,
    FO_INVALID_PIXEL_DEPTH = -6,
    FO_INVALID_SIGNATURE = -7,
    FO_BUFFER_TOO_SMALL = -8,
    FO_PANIC = -11 /* the reference would panic (assert / unwrap) */
};

/* ------------------------------------------------------------------------ */
/* Bit I/O: bitstream-io BitWriter/BitReader<_, BigEndian> semantics          */
/* (call sites compression.rs:270,279-280,296; rice_coding.rs:34-36,46-47;    */
/*  phase_in_coding.rs:69,80-81,95,106).  `mock` reproduces the reference's   */
/*  test sink coding/bitwrite_mock.rs:30-41, whose write() is LSB-first.      */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint8_t *buf;
    size_t cap;      /* bytes */
    size_t nbytes;   /* whole bytes flushed to buf */
    uint64_t acc;    /* pending bits, right-aligned */
    uint32_t nacc;   /* number of pending bits (< 8 between calls) */
    uint64_t nbits;  /* bits written so far */
    int overflow;
    int mock;
} bitwriter;

static void bw_init_grow(bitwriter *w) {
    w->cap = 1 << 16;
    w->buf = (uint8_t *)malloc(w->cap);
    w->nbytes = 0; w->acc = 0; w->nacc = 0; w->nbits = 0;
    w->overflow = 0; w->mock = 0;
}

static int bw_grow(bitwriter *w) {
    size_t ncap = w->cap * 2;
    uint8_t *nb = (uint8_t *)realloc(w->buf, ncap);
    if (!nb) { w->overflow = 1; return 0; }
    w->buf = nb; w->cap = ncap;
    return 1;
}

/* append `bits` (<= 32) bits, most significant first */
static inline void bw_put(bitwriter *w, uint32_t bits, uint32_t value) {
    if (bits == 0) return;
    if (bits < 32) value &= (1u << bits) - 1u;
    w->acc = (w->acc << bits) | value;
    w->nacc += bits;
    w->nbits += bits;
    while (w->nacc >= 8) {
        if (w->nbytes >= w->cap && !bw_grow(w)) { w->nacc -= 8; continue; }
        w->buf[w->nbytes++] = (uint8_t)(w->acc >> (w->nacc - 8));
        w->nacc -= 8;
    }
}

/* BitWrite::write_bit -- MSB of each byte is written first (DOC.md:303). */
static inline void bw_write_bit(bitwriter *w, int bit) { bw_put(w, 1, bit ? 1u : 0u); }

/* BitWrite::write(bits, value): most significant of the `bits` bits first.
 * In mock mode: least significant first (bitwrite_mock.rs:34-39). */
static inline void bw_write(bitwriter *w, uint32_t bits, uint32_t value) {
    if (w->mock) {
        for (uint32_t i = 0; i < bits; i++) { bw_write_bit(w, value & 1u); value >>= 1; }
        return;
    }
    bw_put(w, bits, value);
}

/* BitWrite::write_signed(32, v): two's complement, big-endian bit order. */
static void bw_write_signed32(bitwriter *w, int32_t v) { bw_write(w, 32, (uint32_t)v); }

/* BitWrite::write_unary0(v): v one-bits then a zero (rice_coding.rs:76-77
 * pins the polarity: Rice k=0 of 12 is "1111111111110"). */
static inline void bw_write_unary0(bitwriter *w, uint32_t v) {
    while (v >= 32) { bw_put(w, 32, 0xFFFFFFFFu); v -= 32; }
    bw_put(w, v + 1, ((1u << v) - 1u) << 1); /* v ones then a zero; v+1 <= 32 */
}

/* BitWrite::byte_align: pad with zero bits. */
static void bw_byte_align(bitwriter *w) { if (w->nbits & 7) bw_put(w, 8 - (uint32_t)(w->nbits & 7), 0); }

static inline int bw_bit_at(const bitwriter *w, uint64_t i) {
    if ((i >> 3) < w->nbytes) return (w->buf[i >> 3] >> (7 - (i & 7))) & 1;
    uint32_t j = (uint32_t)(i - (uint64_t)w->nbytes * 8); /* index into pending bits */
    return (int)((w->acc >> (w->nacc - 1 - j)) & 1);
}

typedef struct {
    const uint8_t *buf;
    uint64_t nbits_total;
    uint64_t pos;
    int eof;
} bitreader;

/* BitRead::read(bits): most significant first; a short read is an io::Error. */
static inline uint32_t br_read(bitreader *r, uint32_t bits) {
    if (bits == 0) return 0;
    if (r->pos + bits > r->nbits_total) { r->eof = 1; r->pos = r->nbits_total; return 0; }
    size_t byte = (size_t)(r->pos >> 3);
    uint32_t off = (uint32_t)(r->pos & 7);
    uint32_t need = (off + bits + 7) / 8;
    uint64_t v = 0;
    for (uint32_t i = 0; i < need; i++) v = (v << 8) | r->buf[byte + i];
    v >>= (need * 8 - off - bits);
    if (bits < 32) v &= (1ull << bits) - 1ull; else v &= 0xFFFFFFFFull;
    r->pos += bits;
    return (uint32_t)v;
}
static inline int br_read_bit(bitreader *r) {
    if (r->pos >= r->nbits_total) { r->eof = 1; return 0; }
    int b = (r->buf[r->pos >> 3] >> (7 - (r->pos & 7))) & 1;
    r->pos++;
    return b;
}
static int32_t br_read_signed32(bitreader *r) { return (int32_t)br_read(r, 32); }
/* BitRead::read_unary0: count one-bits up to the terminating zero. */
static inline uint32_t br_read_unary0(bitreader *r) {
    uint32_t n = 0;
    for (;;) {
        int b = br_read_bit(r);
        if (r->eof) return n;
        if (!b) return n;
        n++;
    }
}

/* ------------------------------------------------------------------------ */
/* RiceCoder  (coding/rice_coding.rs:19-58)                                   */
/* ------------------------------------------------------------------------ */
/* rice_coding.rs:56-58 */
static inline uint32_t rice_code_length(uint32_t k, uint32_t number) { return (number >> k) + 1 + k; }

/* rice_coding.rs:26-39 */
static void rice_encode(bitwriter *w, uint32_t k, uint32_t number) {
    uint32_t quotient = number >> k;
    uint32_t remainder = number & ((1u << k) - 1u);
    bw_write_unary0(w, quotient);
    bw_write(w, k, remainder);
}

/* rice_coding.rs:42-51; returns 0 and sets *panic when checked_mul overflows */
static uint32_t rice_decode(bitreader *r, uint32_t k, int *panic) {
    uint32_t quotient = br_read_unary0(r);
    uint32_t remainder = br_read(r, k);
    uint64_t prod = (uint64_t)quotient << k;
    if (prod > 0xFFFFFFFFull) { *panic = 1; return 0; }
    uint64_t res = prod + remainder;
    if (res > 0xFFFFFFFFull) { *panic = 1; return 0; }
    return (uint32_t)res;
}

/* ------------------------------------------------------------------------ */
/* PhaseInCoder  (coding/phase_in_coding.rs:23-112)                           */
/* ------------------------------------------------------------------------ */
typedef struct { uint32_t n, m, left_p, right_p; } phase_in;

/* phase_in_coding.rs:27-41; returns 0 if the reference would panic */
static int phase_in_new(phase_in *c, uint32_t n) {
    if (n == 0) return 0;                       /* "n is 0!" */
    uint32_t m = 31u - (uint32_t)__builtin_clz(n);
    if (m + 1 >= 32) return 0;                  /* "n is too big!" */
    uint32_t lpw = 1u << m, rpw = 1u << (m + 1);
    c->n = n; c->m = m; c->left_p = n - lpw; c->right_p = rpw - n;
    return 1;
}
/* phase_in_coding.rs:50-52 / :55-57 */
static inline uint32_t pi_rotate_right(const phase_in *c, uint32_t x) { return (x + c->n - c->left_p) % c->n; }
static inline uint32_t pi_rotate_left(const phase_in *c, uint32_t x) { return (x + c->left_p) % c->n; }

/* phase_in_coding.rs:64-84 */
static void phase_in_encode(const phase_in *c, bitwriter *w, uint32_t number) {
    number = pi_rotate_right(c, number);
    if (number < c->right_p) {
        bw_write(w, c->m, number);
    } else {
        uint32_t pair = (number - c->right_p) / 2;
        uint32_t last_bit = (number - c->right_p) % 2;
        bw_write(w, c->m, pair + c->right_p);
        bw_write_bit(w, last_bit == 1);
    }
}
/* phase_in_coding.rs:90-112 */
static uint32_t phase_in_decode(const phase_in *c, bitreader *r) {
    uint32_t first_m = br_read(r, c->m);
    if (first_m < c->right_p) return pi_rotate_left(c, first_m);
    uint32_t pair = first_m - c->right_p;
    uint32_t number = pair * 2 + c->right_p;
    if (br_read_bit(r)) number += 1;
    return pi_rotate_left(c, number);
}

/* ------------------------------------------------------------------------ */
/* KEstimator  (compression/parameter_selection.rs:24-85)                     */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint32_t max_context;
    const uint8_t *k_values;
    uint32_t nk;
    uint32_t *map;        /* [(max_context+1) * nk], zero-initialised (:29-33) */
    int has_halve;
    uint32_t halve_at;
} kestimator;

static int kest_new(kestimator *e, uint32_t max_context, const uint8_t *k_values, uint32_t nk,
                    int has_halve, uint32_t halve_at) {
    if (nk == 0) return 0; /* parameter_selection.rs:25-27 panics */
    e->max_context = max_context; e->k_values = k_values; e->nk = nk;
    e->has_halve = has_halve; e->halve_at = halve_at;
    e->map = (uint32_t *)calloc((size_t)(max_context + 1) * nk, sizeof(uint32_t));
    return e->map != NULL;
}
static void kest_free(kestimator *e) { free(e->map); e->map = NULL; }

/* parameter_selection.rs:49-65 */
static int kest_update(kestimator *e, uint32_t context, uint32_t encoded) {
    if (context > e->max_context) return 0; /* assert! */
    uint32_t *row = e->map + (size_t)context * e->nk;
    for (uint32_t ki = 0; ki < e->nk; ki++) row[ki] += rice_code_length(e->k_values[ki], encoded);
    if (e->has_halve) {
        uint32_t mn = row[0];
        for (uint32_t ki = 1; ki < e->nk; ki++) if (row[ki] < mn) mn = row[ki];
        if (mn > e->halve_at) for (uint32_t ki = 0; ki < e->nk; ki++) row[ki] /= 2;
    }
    return 1;
}
/* parameter_selection.rs:71-85: `<=` so ties go to the LAST index */
static int kest_get_k(const kestimator *e, uint32_t context) {
    if (context > e->max_context) return -1; /* assert! */
    const uint32_t *row = e->map + (size_t)context * e->nk;
    uint32_t smallest = 0xFFFFFFFFu, best = 0;
    for (uint32_t i = 0; i < e->nk; i++) if (row[i] <= smallest) { best = i; smallest = row[i]; }
    return e->k_values[best];
}

/* ------------------------------------------------------------------------ */
/* nearest_neighbours  (compression/misc.rs:6-24)                             */
/* returns 0 for None                                                         */
/* ------------------------------------------------------------------------ */
static int nearest_neighbours(size_t i, size_t width, size_t *a, size_t *b) {
    size_t x = i % width, y = i / width;
    if (x > 0 && y > 0) { *a = i - 1; *b = i - width; return 1; }
    if (y == 0) {
        if (x >= 2) { *a = i - 1; *b = i - 2; return 1; }
        return 0;
    }
    if (y >= 2) { *a = i - width; *b = i - 2 * width; return 1; }
    if (x + 1 < width) { *a = i - width; *b = i - width + 1; return 1; }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* colour transform  (compression/color_transform.rs:11-26); C `/` on int32   */
/* truncates toward zero exactly like Rust's.                                 */
/* ------------------------------------------------------------------------ */
static inline void rgb_to_ycocg(int32_t r, int32_t g, int32_t b, int32_t *y, int32_t *co, int32_t *cg) {
    int32_t c_o = r - b;
    int32_t t = b + c_o / 2;
    int32_t c_g = g - t;
    *y = t + c_g / 2; *co = c_o; *cg = c_g;
}
static inline void ycocg_to_rgb(int32_t y, int32_t co, int32_t cg, int32_t *r, int32_t *g, int32_t *b) {
    int32_t t = y - cg / 2;
    *g = cg + t;
    *b = t - co / 2;
    *r = *b + co;
}

/* ------------------------------------------------------------------------ */
/* CodingOptions (compression.rs:63-68) and Intensity consts (traits.rs:25-43)*/
/* ------------------------------------------------------------------------ */
typedef struct {
    uint32_t max_context;
    const uint8_t *k_values;
    uint32_t nk;
    int has_scaling;
    uint32_t scaling;
} coding_options;

static const uint8_t K_U8[6] = {0, 1, 2, 3, 4, 5};
static const uint8_t K_U16[15] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14};

static coding_options options_for_depth(int depth) {
    coding_options o;
    if (depth == 0) { o.max_context = 255u * 2; o.k_values = K_U8; o.nk = 6; }
    else { o.max_context = 65535u * 2; o.k_values = K_U16; o.nk = 15; }
    o.has_scaling = 1; o.scaling = 1024;
    return o;
}

/* per-pixel trace (debug aid for the GPU parity tests; not in the reference) */
typedef struct {
    uint8_t *cls;    /* 0 in-range, 1 above, 2 below; 255 for the two raw pixels */
    uint8_t *k;      /* k from get_k at this pixel */
    uint32_t *ctx;   /* context */
    uint32_t *len;   /* code length in bits including the class marker */
} channel_trace;

/* intensity marker, compression.rs:29-45: InRange '1', Above '01', Below '00' */
static void encode_intensity(bitwriter *w, int cls) {
    if (cls == 0) { bw_write_bit(w, 1); }
    else if (cls == 1) { bw_write_bit(w, 0); bw_write_bit(w, 1); }
    else { bw_write_bit(w, 0); bw_write_bit(w, 0); }
}
/* compression.rs:48-61 */
static int decode_intensity(bitreader *r) {
    if (br_read_bit(r)) return 0;
    if (br_read_bit(r)) return 1;
    return 2;
}

/* compress_channel, compression.rs:76-148 */
static int compress_channel(const int32_t *channel, uint32_t width, uint32_t height,
                            coding_options opt, bitwriter *w, channel_trace *tr) {
    uint64_t total64 = (uint64_t)width * height;
    if (total64 > 0xFFFFFFFFull) return FO_PANIC; /* checked_mul().unwrap(), :86 */
    size_t total = (size_t)total64;

    if (width == 0 || height == 0) {                 /* :94-98 */
        bw_write_signed32(w, 0); bw_write_signed32(w, 0);
        return FO_OK;
    }
    if (width == 1 && height == 1) {                 /* :99-103 */
        bw_write_signed32(w, channel[0]); bw_write_signed32(w, 0);
        if (tr) { tr->cls[0] = 255; }
        return FO_OK;
    }
    bw_write_signed32(w, channel[0]);                /* :104-107 */
    bw_write_signed32(w, channel[1]);
    if (tr) { tr->cls[0] = 255; tr->cls[1] = 255; }

    kestimator est;
    if (!kest_new(&est, opt.max_context, opt.k_values, opt.nk, opt.has_scaling, opt.scaling)) return FO_PANIC;

    int rc = FO_OK;
    for (size_t i = 2; i < total; i++) {             /* :117-146 */
        size_t a, b;
        if (!nearest_neighbours(i, width, &a, &b)) { rc = FO_PANIC; break; }
        int32_t p = channel[i], v1 = channel[a], v2 = channel[b];
        int32_t h = v1 > v2 ? v1 : v2;
        int32_t l = v1 < v2 ? v1 : v2;
        uint32_t context = (uint32_t)(h - l);
        int k = kest_get_k(&est, context);
        if (k < 0) { rc = FO_PANIC; break; }
        uint64_t before = w->nbits;
        int cls;
        if (p >= l && p <= h) {
            cls = 0;
            encode_intensity(w, 0);
            phase_in pc;
            if (!phase_in_new(&pc, context + 1)) { rc = FO_PANIC; break; }
            phase_in_encode(&pc, w, (uint32_t)(p - l));
        } else if (p < l) {
            cls = 2;
            encode_intensity(w, 2);
            uint32_t e = (uint32_t)(l - p - 1);
            rice_encode(w, (uint32_t)k, e);
            kest_update(&est, context, e);
        } else {
            cls = 1;
            encode_intensity(w, 1);
            uint32_t e = (uint32_t)(p - h - 1);
            rice_encode(w, (uint32_t)k, e);
            kest_update(&est, context, e);
        }
        if (tr) { tr->cls[i] = (uint8_t)cls; tr->k[i] = (uint8_t)k; tr->ctx[i] = context; tr->len[i] = (uint32_t)(w->nbits - before); }
    }
    kest_free(&est);
    if (w->overflow) return FO_BUFFER_TOO_SMALL;
    return rc;
}

/* decompress_channel, compression.rs:151-248.  *out is malloc'ed (total i32). */
static int decompress_channel(uint32_t width, uint32_t height, coding_options opt, bitreader *r,
                              int32_t **out, size_t *out_n) {
    *out = NULL; *out_n = 0;
    int32_t pixel1 = br_read_signed32(r);            /* :161-162 */
    int32_t pixel2 = br_read_signed32(r);
    if (r->eof) return FO_IO_ERROR;

    if (width == 0 || height == 0) return FO_OK;     /* :166-168 */
    if (width == 1 && height == 1) {                 /* :169-171 */
        *out = (int32_t *)malloc(sizeof(int32_t)); (*out)[0] = pixel1; *out_n = 1;
        return FO_OK;
    }
    uint64_t total64 = (uint64_t)width * height;     /* :176-180 */
    if (total64 > 0xFFFFFFFFull) return FO_INVALID_DIMENSIONS;
    size_t total = (size_t)total64;
    int32_t *buf = (int32_t *)calloc(total, sizeof(int32_t));
    if (!buf) return FO_INVALID_DIMENSIONS;
    buf[0] = pixel1; buf[1] = pixel2;

    kestimator est;
    if (!kest_new(&est, opt.max_context, opt.k_values, opt.nk, opt.has_scaling, opt.scaling)) { free(buf); return FO_PANIC; }

    int rc = FO_OK;
    for (size_t i = 2; i < total; i++) {             /* :193-246 */
        size_t a, b;
        if (!nearest_neighbours(i, width, &a, &b)) { rc = FO_PANIC; break; }
        int32_t v1 = buf[a], v2 = buf[b];
        int32_t h = v1 > v2 ? v1 : v2;
        int32_t l = v1 < v2 ? v1 : v2;
        int64_t ctx64 = (int64_t)h - (int64_t)l;
        if (ctx64 > 0x7FFFFFFFll) { rc = FO_PANIC; break; }   /* debug overflow on h - l */
        uint32_t context = (uint32_t)ctx64;
        int k = kest_get_k(&est, context);
        if (k < 0) { rc = FO_PANIC; break; }                  /* assert!(context <= max_context) */
        int cls = decode_intensity(r);
        if (r->eof) { rc = FO_IO_ERROR; break; }
        int64_t value;
        if (cls == 0) {
            phase_in pc;
            if (!phase_in_new(&pc, context + 1)) { rc = FO_PANIC; break; }
            uint32_t v = phase_in_decode(&pc, r);
            if (r->eof) { rc = FO_IO_ERROR; break; }
            if (v > 0x7FFFFFFFu) { rc = FO_INVALID_VALUE; break; }
            value = (int64_t)v + l;
        } else {
            int panic = 0;
            uint32_t e = rice_decode(r, (uint32_t)k, &panic);
            if (r->eof) { rc = FO_IO_ERROR; break; }
            if (panic) { rc = FO_PANIC; break; }
            kest_update(&est, context, e);
            if (e > 0x7FFFFFFFu) { rc = FO_INVALID_VALUE; break; }
            value = (cls == 2) ? (int64_t)l - e - 1 : (int64_t)e + h + 1;
        }
        if (value > 0x7FFFFFFFll || value < -0x80000000ll) { rc = FO_VALUE_OVERFLOW; break; }
        buf[i] = (int32_t)value;
    }
    kest_free(&est);
    if (rc != FO_OK) { free(buf); return rc; }
    *out = buf; *out_n = total;
    return FO_OK;
}

/* ------------------------------------------------------------------------ */
/* container  (compression/format.rs:51-84) + image impls                     */
/* (compression.rs:255-282 Luma, :322-371 Rgb, :284-315, :373-410, :420-441)  */
/* ------------------------------------------------------------------------ */
ORACLE_API int felics_oracle_read_header(const uint8_t *buf, size_t len, uint8_t *color, uint8_t *depth,
                                         uint32_t *width, uint32_t *height) {
    if (len < 4) return FO_IO_ERROR;                 /* read_exact, format.rs:67-68 */
    if (memcmp(buf, "FLCS", 4) != 0) return FO_INVALID_SIGNATURE;
    if (len < 5) return FO_IO_ERROR;
    if (buf[4] > 1) return FO_INVALID_COLOR_TYPE;    /* format.rs:14-23 */
    if (len < 6) return FO_IO_ERROR;
    if (buf[5] > 1) return FO_INVALID_PIXEL_DEPTH;   /* format.rs:32-41 */
    if (len < 14) return FO_IO_ERROR;
    *color = buf[4]; *depth = buf[5];
    *width = ((uint32_t)buf[6] << 24) | ((uint32_t)buf[7] << 16) | ((uint32_t)buf[8] << 8) | buf[9];
    *height = ((uint32_t)buf[10] << 24) | ((uint32_t)buf[11] << 16) | ((uint32_t)buf[12] << 8) | buf[13];
    return FO_OK;
}

static void write_header(uint8_t *h, int color, int depth, uint32_t width, uint32_t height) {
    memcpy(h, "FLCS", 4);
    h[4] = (uint8_t)color; h[5] = (uint8_t)depth;
    h[6] = (uint8_t)(width >> 24); h[7] = (uint8_t)(width >> 16); h[8] = (uint8_t)(width >> 8); h[9] = (uint8_t)width;
    h[10] = (uint8_t)(height >> 24); h[11] = (uint8_t)(height >> 16); h[12] = (uint8_t)(height >> 8); h[13] = (uint8_t)height;
}

static inline int32_t sample_at(const void *pixels, int depth, size_t idx) {
    return depth == 0 ? (int32_t)((const uint8_t *)pixels)[idx] : (int32_t)((const uint16_t *)pixels)[idx];
}

/* Generalised compress with explicit options (lets tests reproduce the K-set
 * table DOC.md:385-389 and the "without colour transform" column :469-477).
 * color: 0 gray, 1 rgb.  depth: 0 = u8, 1 = u16 (host-endian samples).
 * use_transform: 1 is the reference behaviour for rgb.
 * Returns FO_OK and *out_len, or FO_BUFFER_TOO_SMALL with *out_len = needed. */
ORACLE_API int felics_oracle_compress_opts(const void *pixels, int color, int depth, uint32_t width, uint32_t height,
                                           const uint8_t *k_values, uint32_t nk, int has_scaling, uint32_t scaling,
                                           int use_transform, uint8_t *out, size_t cap, size_t *out_len,
                                           uint64_t *payload_bits) {
    coding_options opt = options_for_depth(depth);
    if (k_values) { opt.k_values = k_values; opt.nk = nk; opt.has_scaling = has_scaling; opt.scaling = scaling; }
    bitwriter w; bw_init_grow(&w);
    size_t npix = (size_t)width * height;
    int rc = FO_OK;
    if (color == 0) {
        int32_t *ch = (int32_t *)malloc((npix ? npix : 1) * sizeof(int32_t));   /* compression.rs:276 */
        for (size_t i = 0; i < npix; i++) ch[i] = sample_at(pixels, depth, i);
        rc = compress_channel(ch, width, height, opt, &w, NULL);
        free(ch);
    } else {
        int32_t *y = (int32_t *)malloc((npix ? npix : 1) * sizeof(int32_t));
        int32_t *co = (int32_t *)malloc((npix ? npix : 1) * sizeof(int32_t));
        int32_t *cg = (int32_t *)malloc((npix ? npix : 1) * sizeof(int32_t));
        for (size_t i = 0; i < npix; i++) {                                      /* compression.rs:346-356 */
            int32_t r = sample_at(pixels, depth, 3 * i), g = sample_at(pixels, depth, 3 * i + 1), b = sample_at(pixels, depth, 3 * i + 2);
            if (use_transform) rgb_to_ycocg(r, g, b, &y[i], &co[i], &cg[i]);
            else { y[i] = r; co[i] = g; cg[i] = b; }
        }
        rc = compress_channel(y, width, height, opt, &w, NULL);                  /* :365-367 */
        if (rc == FO_OK) rc = compress_channel(co, width, height, opt, &w, NULL);
        if (rc == FO_OK) rc = compress_channel(cg, width, height, opt, &w, NULL);
        free(y); free(co); free(cg);
    }
    if (payload_bits) *payload_bits = w.nbits;
    bw_byte_align(&w);                                                           /* :279 / :368 */
    size_t total = 14 + w.nbytes;
    if (out_len) *out_len = total;
    if (rc == FO_OK) {
        if (w.overflow || total > cap) rc = FO_BUFFER_TOO_SMALL;
        else { write_header(out, color, depth, width, height); memcpy(out + 14, w.buf, total - 14); }
    }
    free(w.buf);
    return rc;
}

/* compress_image / CompressDecompress::compress with the reference's options */
ORACLE_API int felics_oracle_compress(const void *pixels, int color, int depth, uint32_t width, uint32_t height,
                                      uint8_t *out, size_t cap, size_t *out_len) {
    return felics_oracle_compress_opts(pixels, color, depth, width, height, NULL, 0, 0, 0, 1, out, cap, out_len, NULL);
}

/* upper bound for the output of felics_oracle_compress (for callers sizing a buffer) */
ORACLE_API size_t felics_oracle_compress_bound(int color, int depth, uint32_t width, uint32_t height) {
    size_t npix = (size_t)width * height;
    size_t ch = color ? 3 : 1;
    /* marker 2 + unary up to max_context + 1 + k bits; generous */
    size_t per = depth == 0 ? 520 : 131100;
    return 14 + ch * 8 + (npix * ch * per + 7) / 8 + 8;
}

/* decompress_image (compression.rs:420-441).  pixels_out receives u8/u16
 * samples (interleaved RGB for colour). */
ORACLE_API int felics_oracle_decompress(const uint8_t *fel, size_t len, void *pixels_out, size_t cap_bytes,
                                        uint8_t *color_out, uint8_t *depth_out, uint32_t *width_out, uint32_t *height_out) {
    uint8_t color, depth; uint32_t width, height;
    int rc = felics_oracle_read_header(fel, len, &color, &depth, &width, &height);
    if (rc != FO_OK) return rc;
    if (color_out) *color_out = color;
    if (depth_out) *depth_out = depth;
    if (width_out) *width_out = width;
    if (height_out) *height_out = height;
    coding_options opt = options_for_depth(depth);
    bitreader r; r.buf = fel + 14; r.nbits_total = (uint64_t)(len - 14) * 8; r.pos = 0; r.eof = 0;
    int nch = color ? 3 : 1;
    int32_t *ch[3] = {NULL, NULL, NULL};
    size_t n[3] = {0, 0, 0};
    for (int c = 0; c < nch; c++) {                                              /* :302 / :392-394 */
        rc = decompress_channel(width, height, opt, &r, &ch[c], &n[c]);
        if (rc != FO_OK) { for (int j = 0; j < c; j++) free(ch[j]); return rc; }
    }
    size_t npix = n[0];
    int32_t maxv = depth == 0 ? 255 : 65535;
    size_t bps = depth == 0 ? 1 : 2;
    if (npix * (size_t)nch * bps > cap_bytes) { for (int j = 0; j < nch; j++) free(ch[j]); return FO_BUFFER_TOO_SMALL; }
    rc = FO_OK;
    for (size_t i = 0; i < npix && rc == FO_OK; i++) {
        int32_t v[3];
        if (nch == 1) v[0] = ch[0][i];
        else ycocg_to_rgb(ch[0][i], ch[1][i], ch[2][i], &v[0], &v[1], &v[2]);   /* :402-407 */
        for (int c = 0; c < nch; c++) {
            if (v[c] < 0 || v[c] > maxv) { rc = FO_INVALID_VALUE; break; }       /* try_into, :305-310 */
            if (depth == 0) ((uint8_t *)pixels_out)[i * nch + c] = (uint8_t)v[c];
            else ((uint16_t *)pixels_out)[i * nch + c] = (uint16_t)v[c];
        }
    }
    for (int j = 0; j < nch; j++) free(ch[j]);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* unit-level entry points for the golden-vector tests                        */
/* ------------------------------------------------------------------------ */
static size_t bits_to_string(const bitwriter *w, char *out, size_t cap) {
    size_t n = (size_t)w->nbits;
    if (n + 1 > cap) n = cap - 1;
    for (size_t i = 0; i < n; i++) out[i] = bw_bit_at(w, i) ? '1' : '0';
    out[n] = 0;
    return n;
}

/* mock=1 renders like BitWriterMock (rice_coding.rs:70-82 golden strings). */
ORACLE_API int felics_oracle_rice_bits(uint32_t k, uint32_t number, int mock, char *out, size_t cap) {
    if (k >= 32) return FO_PANIC; /* rice_coding.rs:20 "k is too big!" */
    bitwriter w; bw_init_grow(&w); w.mock = mock;
    rice_encode(&w, k, number);
    bits_to_string(&w, out, cap);
    free(w.buf);
    return FO_OK;
}
ORACLE_API uint32_t felics_oracle_rice_code_length(uint32_t k, uint32_t number) { return rice_code_length(k, number); }

ORACLE_API int felics_oracle_phase_in_params(uint32_t n, uint32_t *m, uint32_t *left_p, uint32_t *right_p) {
    phase_in c;
    if (!phase_in_new(&c, n)) return FO_PANIC;
    *m = c.m; *left_p = c.left_p; *right_p = c.right_p;
    return FO_OK;
}
ORACLE_API int felics_oracle_phase_in_bits(uint32_t n, uint32_t number, int mock, char *out, size_t cap) {
    phase_in c;
    if (!phase_in_new(&c, n)) return FO_PANIC;
    if (number >= n) return FO_PANIC; /* phase_in_coding.rs:63 assert */
    bitwriter w; bw_init_grow(&w); w.mock = mock;
    phase_in_encode(&c, &w, number);
    bits_to_string(&w, out, cap);
    free(w.buf);
    return FO_OK;
}

/* Encode a list of Rice (kind 0: a=k, b=value) / phased-in (kind 1: a=n, b=value)
 * codes through the real MSB-first writer, byte-align, decode them back
 * (rice_coding.rs:91-107, :111-135; phase_in_coding.rs:231-252). */
ORACLE_API int felics_oracle_codes_roundtrip(const uint32_t *kind, const uint32_t *a, const uint32_t *b, size_t n,
                                             uint32_t *decoded, uint8_t *bytes_out, size_t cap, size_t *nbytes) {
    bitwriter w; bw_init_grow(&w);
    for (size_t i = 0; i < n; i++) {
        if (kind[i] == 0) rice_encode(&w, a[i], b[i]);
        else { phase_in c; if (!phase_in_new(&c, a[i])) { free(w.buf); return FO_PANIC; } phase_in_encode(&c, &w, b[i]); }
    }
    bw_byte_align(&w);
    size_t nb = w.nbytes;
    if (nbytes) *nbytes = nb;
    if (bytes_out) memcpy(bytes_out, w.buf, nb < cap ? nb : cap);
    bitreader r; r.buf = w.buf; r.nbits_total = w.nbits; r.pos = 0; r.eof = 0;
    for (size_t i = 0; i < n; i++) {
        if (kind[i] == 0) { int panic = 0; decoded[i] = rice_decode(&r, a[i], &panic); }
        else { phase_in c; phase_in_new(&c, a[i]); decoded[i] = phase_in_decode(&c, &r); }
    }
    int eof = r.eof;
    free(w.buf);
    return eof ? FO_IO_ERROR : FO_OK;
}

/* KEstimator as an opaque handle (parameter_selection.rs tests :96-183) */
ORACLE_API void *felics_oracle_kest_new(uint32_t max_context, const uint8_t *k_values, uint32_t nk, int has_halve, uint32_t halve_at) {
    kestimator *e = (kestimator *)malloc(sizeof(kestimator));
    uint8_t *kv = (uint8_t *)malloc(nk ? nk : 1);
    memcpy(kv, k_values, nk);
    if (!kest_new(e, max_context, kv, nk, has_halve, halve_at)) { free(kv); free(e); return NULL; }
    return e;
}
ORACLE_API int felics_oracle_kest_update(void *h, uint32_t context, uint32_t encoded) { return kest_update((kestimator *)h, context, encoded) ? FO_OK : FO_PANIC; }
ORACLE_API int felics_oracle_kest_get_k(void *h, uint32_t context) { return kest_get_k((kestimator *)h, context); }
ORACLE_API uint32_t felics_oracle_kest_entry(void *h, uint32_t context, uint32_t ki) { kestimator *e = (kestimator *)h; return e->map[(size_t)context * e->nk + ki]; }
ORACLE_API void felics_oracle_kest_free(void *h) { kestimator *e = (kestimator *)h; free((void *)e->k_values); kest_free(e); free(e); }

ORACLE_API int felics_oracle_nearest_neighbours(uint64_t i, uint64_t width, uint64_t *a, uint64_t *b) {
    size_t aa = 0, bb = 0;
    int ok = nearest_neighbours((size_t)i, (size_t)width, &aa, &bb);
    *a = aa; *b = bb;
    return ok;
}
ORACLE_API void felics_oracle_rgb_to_ycocg(int32_t r, int32_t g, int32_t b, int32_t *out3) { rgb_to_ycocg(r, g, b, &out3[0], &out3[1], &out3[2]); }
ORACLE_API void felics_oracle_ycocg_to_rgb(int32_t y, int32_t co, int32_t cg, int32_t *out3) { ycocg_to_rgb(y, co, cg, &out3[0], &out3[1], &out3[2]); }

/* exhaustive u8 reversibility + range check, color_transform.rs:35-73.
 * returns 0 on success; ranges[6] = min/max of y, co, cg */
ORACLE_API int felics_oracle_color_transform8_exhaustive(int32_t *ranges) {
    int32_t mn[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, mx[3] = {-0x7FFFFFFF, -0x7FFFFFFF, -0x7FFFFFFF};
    for (int r = 0; r < 256; r++) for (int g = 0; g < 256; g++) for (int b = 0; b < 256; b++) {
        int32_t v[3], rn, gn, bn;
        rgb_to_ycocg(r, g, b, &v[0], &v[1], &v[2]);
        ycocg_to_rgb(v[0], v[1], v[2], &rn, &gn, &bn);
        if (rn != r || gn != g || bn != b) return 1;
        for (int c = 0; c < 3; c++) { if (v[c] < mn[c]) mn[c] = v[c]; if (v[c] > mx[c]) mx[c] = v[c]; }
    }
    for (int c = 0; c < 3; c++) { ranges[2 * c] = mn[c]; ranges[2 * c + 1] = mx[c]; }
    return 0;
}

/* Per-pixel trace of one channel (i32 samples): class, k, context, code length.
 * Debug aid for GPU parity tests; follows compress_channel exactly. */
ORACLE_API int felics_oracle_trace_channel(const int32_t *channel, uint32_t width, uint32_t height, int depth,
                                           uint8_t *cls, uint8_t *k, uint32_t *ctx, uint32_t *len, uint64_t *total_bits) {
    coding_options opt = options_for_depth(depth);
    bitwriter w; bw_init_grow(&w);
    channel_trace tr = {cls, k, ctx, len};
    int rc = compress_channel(channel, width, height, opt, &w, &tr);
    if (total_bits) *total_bits = w.nbits;
    free(w.buf);
    return rc;
}

/* Batch helpers for the CPU baseline: encode/decode `n` equally sized images,
 * images [first, first+count) handled by the calling thread.  Returns total
 * compressed bytes through *bytes.  (Timing harness only.) */
ORACLE_API int felics_oracle_compress_many(const uint8_t *pixels, size_t image_stride, int color, int depth,
                                           uint32_t width, uint32_t height, size_t count, uint8_t *scratch, size_t scratch_cap,
                                           uint64_t *bytes) {
    uint64_t tot = 0;
    for (size_t i = 0; i < count; i++) {
        size_t len = 0;
        int rc = felics_oracle_compress(pixels + i * image_stride, color, depth, width, height, scratch, scratch_cap, &len);
        if (rc != FO_OK) return rc;
        tot += len;
    }
    if (bytes) *bytes = tot;
    return FO_OK;
}
