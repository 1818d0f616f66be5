// FELICS decode for sm_100a.
//
// Replaces decompress_channel (/root/reference/src/compression.rs:151-248).  A plane's
// bit stream is strictly serial (every code length depends on the previously decoded
// samples and on the adaptive k), and the three planes of an RGB file share one
// unaligned stream with no recorded lengths (compression.rs:392-394), so one decoder
// walks one file.  Parallelism comes from the batch: one warp per file, the estimator
// table (511 contexts x 6 counters, parameter_selection.rs:29-33) in shared memory.
#include "ctx.h"
#include "device_common.cuh"
#include "serial16.cuh"
#include "sidecar.cuh"

#include <algorithm>
#include <cstring>

namespace felics {

struct DecArgs {
    const uint32_t *words;     // the arena viewed as 32-bit words (4-byte aligned)
    uint64_t arena_words;      // number of whole words that may be read
    const uint64_t *offsets;   // device copy, n+1 entries (bytes)
    int16_t *planes;           // [n*nch][pstride]: padded, see plane_stride8
    size_t pstride;
    int *status;               // [n]
    uint32_t w, h, npix, nch;
    uint8_t color, depth;
};

struct BitReader {
    const uint32_t *words;
    uint64_t nwords;
    uint64_t pos, end;   // absolute bit positions in the arena
    uint64_t cwi;        // cached word index
    uint32_t c0, c1;
    bool eof;

    __device__ __forceinline__ void init(const uint32_t *w, uint64_t nw, uint64_t p, uint64_t e) {
        words = w; nwords = nw; pos = p; end = e; cwi = ~0ull; c0 = c1 = 0; eof = false;
    }
    __device__ __forceinline__ uint32_t load(uint64_t wi) const { return wi < nwords ? bswap32(__ldg(words + wi)) : 0u; }
    // next 32 bits of the stream, MSB first (bits past `end` are whatever follows; callers bound by `end`)
    __device__ __forceinline__ uint32_t peek32() {
        uint64_t wi = pos >> 5;
        if (wi != cwi) {
            c0 = (wi == cwi + 1) ? c1 : load(wi);
            c1 = load(wi + 1);
            cwi = wi;
        }
        uint32_t sh = (uint32_t)(pos & 31);
        return sh ? ((c0 << sh) | (c1 >> (32 - sh))) : c0;
    }
    // BitRead::read(n), n in 0..32
    __device__ __forceinline__ uint32_t read(uint32_t n) {
        if (n == 0) return 0;
        if (pos + n > end) { eof = true; pos = end; return 0; }
        uint32_t v = peek32() >> (32 - n);
        pos += n;
        return v;
    }
    // BitRead::read_unary0: ones up to the terminating zero
    __device__ __forceinline__ uint32_t read_unary0() {
        uint32_t q = 0;
        for (;;) {
            if (pos >= end) { eof = true; return q; }
            uint64_t avail64 = end - pos;
            uint32_t avail = avail64 > 32 ? 32u : (uint32_t)avail64;
            uint32_t ones = __clz(~peek32());   // 32 when all ones
            if (ones >= avail) { q += avail; pos += avail; if (avail < 32) { eof = true; return q; } continue; }
            q += ones;
            pos += ones + 1;
            return q;
        }
    }
};

// Internal status of the fast decoders: the stream did something no valid file does (a sample outside the 16-bit planes,
// a unary run of tens of thousands of ones).  The reference keeps such values as i32 and goes on, and what it finally reports
// depends on what comes next (IoError at the end of the input, the `context <= max_context` assertion, InvalidValue at the
// final try_into, compression.rs:208-243, :305-310), so the file is decoded again by k_decode_exact, which keeps i32 planes
// and follows the reference check by check.  Never returned to the caller.
constexpr int FELICS_NEED_EXACT = 1;

struct ExactArgs {
    const uint32_t *words;
    uint64_t arena_words, off0, off1;
    int32_t *planes;           // [nch][npix]
    int *status;
    uint32_t w, h, npix, nch;
};

// decompress_channel as written (compression.rs:151-248), one lane, i32 samples: only for files the fast decoders gave up on
__global__ void __launch_bounds__(32) k_decode_exact(ExactArgs a) {
    __shared__ uint32_t tab[(NBIN - 1) * NK];
    const uint32_t lane = threadIdx.x;
    int st = FELICS_OK;
    BitReader br;
    br.init(a.words, a.arena_words, 8 * (a.off0 + FELICS_HEADER_BYTES), 8 * a.off1);
    const uint32_t w = a.w;
    for (uint32_t ch = 0; ch < a.nch && st == FELICS_OK; ch++) {
        for (uint32_t j = lane; j < (NBIN - 1) * NK; j += 32) tab[j] = 0;
        __syncwarp();
        if (lane == 0) {
            int32_t *pl = a.planes + (size_t)ch * a.npix;
            const int32_t p1 = (int32_t)br.read(32), p2 = (int32_t)br.read(32);   // :161-162
            if (br.eof) st = FELICS_ERR_IO;
            else {
                if (a.npix >= 1) pl[0] = p1;
                if (a.npix >= 2) pl[1] = p2;
            }
            uint32_t x = 0, y = 0;
            if (a.npix >= 3) { x = 2 % w; y = 2 / w; }
            for (uint32_t i = 2; i < a.npix && st == FELICS_OK; i++) {
                uint32_t ia, ib;
                if (x > 0 && y > 0) { ia = i - 1; ib = i - w; }
                else if (y == 0) { ia = i - 1; ib = i - 2; }
                else if (y >= 2) { ia = i - w; ib = i - 2 * w; }
                else { ia = i - w; ib = i - w + 1; }
                const long long v1 = pl[ia], v2 = pl[ib];
                const long long hi = max(v1, v2), lo = min(v1, v2);
                if (hi - lo > 510ll) { st = FELICS_ERR_CORRUPT; break; }   // h - l overflow / assert!(context <= max_context), parameter_selection.rs:72
                const uint32_t ctx = (uint32_t)(hi - lo);
                uint32_t *row = tab + ctx * NK;
                uint32_t rr[NK];
#pragma unroll
                for (int k = 0; k < NK; k++) rr[k] = row[k];
                const int k = argmin_last(rr);                                // get_k before the marker is read (:202)
                long long value;
                const uint32_t in_range = br.read(1);                         // decode_intensity (:48-61)
                uint32_t above = 0;
                if (!in_range) above = br.read(1);
                if (br.eof) { st = FELICS_ERR_IO; break; }
                if (in_range) {
                    const uint32_t nn = ctx + 1;
                    const int m = 31 - __clz(nn);
                    const uint32_t left_p = nn - (1u << m), right_p = (2u << m) - nn;
                    uint32_t xx = br.read((uint32_t)m);
                    if (!br.eof && xx >= right_p) xx = (xx - right_p) * 2 + right_p + br.read(1);   // phase_in_coding.rs:102-109
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    xx = (xx + left_p) % nn;                                  // rotate_left (:55-57)
                    value = lo + (long long)xx;
                } else {
                    const uint32_t q = br.read_unary0();
                    const uint32_t rem = br.eof ? 0u : br.read((uint32_t)k);
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    if (((unsigned long long)q << k) > 0xFFFFFFFFull) { st = FELICS_ERR_CORRUPT; break; }   // checked_mul(..).unwrap(), rice_coding.rs:50
                    const uint32_t e = (q << k) + rem;
                    uint32_t mn = 0xffffffffu;
#pragma unroll
                    for (int kk = 0; kk < NK; kk++) {                         // update (parameter_selection.rs:49-65)
                        rr[kk] += (e >> kk) + 1u + (uint32_t)kk;
                        mn = min(mn, rr[kk]);
                    }
                    if (mn > HALVE_AT) {
#pragma unroll
                        for (int kk = 0; kk < NK; kk++) rr[kk] >>= 1;
                    }
#pragma unroll
                    for (int kk = 0; kk < NK; kk++) row[kk] = rr[kk];
                    if (e > 0x7FFFFFFFu) { st = FELICS_ERR_INVALID_VALUE; break; }   // u32 -> i32 try_into (:222-224, :234-236)
                    value = above ? hi + (long long)e + 1 : lo - (long long)e - 1;
                }
                if (value < -2147483648ll || value > 2147483647ll) { st = FELICS_ERR_VALUE_OVERFLOW; break; }   // checked_add / checked_sub
                pl[i] = (int32_t)value;
                if (++x == w) { x = 0; y++; }
            }
        }
        st = __shfl_sync(0xffffffffu, st, 0);
    }
    if (lane == 0) *a.status = st;
}

// i32 planes of one image -> pixels with the try_into range checks (compression.rs:305-310, :402-407; color_transform.rs:20-26)
__global__ void k_unplane_exact8(const int32_t *__restrict__ planes, uint8_t *__restrict__ px, uint32_t npix, uint32_t nch, int *__restrict__ status) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        if (nch == 1) {
            const int v = planes[i];
            if (v < 0 || v > 255) { atomicCAS(status, FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
            px[i] = (uint8_t)v;
        } else {
            const long long y = planes[i], co = planes[(size_t)npix + i], cg = planes[2 * (size_t)npix + i];
            const long long t = y - cg / 2, g = cg + t, b = t - co / 2, r = b + co;
            if (r < 0 || r > 255 || g < 0 || g > 255 || b < 0 || b > 255) { atomicCAS(status, FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
            px[3 * i] = (uint8_t)r; px[3 * i + 1] = (uint8_t)g; px[3 * i + 2] = (uint8_t)b;
        }
    }
}

__global__ void __launch_bounds__(32) k_decode_wide(DecArgs a, uint32_t n) {
    __shared__ uint32_t tab[(NBIN - 1) * NK];
    const uint32_t img = blockIdx.x;
    if (img >= n) return;
    const uint32_t lane = threadIdx.x;
    const uint64_t off0 = a.offsets[img], off1 = a.offsets[img + 1];
    const uint64_t len = off1 - off0;
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(a.words);
    int st = FELICS_OK;
    // read_header order (format.rs:63-84), then decompress_with_header's checks (compression.rs:289-294)
    if (lane == 0) {
        const uint8_t *hb = bytes + off0;
        if (len < 4) st = FELICS_ERR_IO;
        else if (hb[0] != 'F' || hb[1] != 'L' || hb[2] != 'C' || hb[3] != 'S') st = FELICS_ERR_INVALID_SIGNATURE;
        else if (len < 5) st = FELICS_ERR_IO;
        else if (hb[4] > 1) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (len < 6) st = FELICS_ERR_IO;
        else if (hb[5] > 1) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else if (len < FELICS_HEADER_BYTES) st = FELICS_ERR_IO;
        else if (hb[4] != a.color) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (hb[5] != a.depth) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else {
            uint32_t w = ((uint32_t)hb[6] << 24) | ((uint32_t)hb[7] << 16) | ((uint32_t)hb[8] << 8) | hb[9];
            uint32_t h = ((uint32_t)hb[10] << 24) | ((uint32_t)hb[11] << 16) | ((uint32_t)hb[12] << 8) | hb[13];
            if (w != a.w || h != a.h) st = FELICS_ERR_INVALID_DIMENSIONS;
        }
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    if (st != FELICS_OK) { if (lane == 0) a.status[img] = st; return; }

    BitReader br;
    br.init(a.words, a.arena_words, 8 * (off0 + FELICS_HEADER_BYTES), 8 * off1);
    const uint32_t w = a.w;

    for (uint32_t ch = 0; ch < a.nch && st == FELICS_OK; ch++) {
        for (uint32_t j = lane; j < (NBIN - 1) * NK; j += 32) tab[j] = 0;   // fresh estimator per channel (:186-190)
        __syncwarp();
        if (lane == 0) {
            int16_t *pl = a.planes + ((size_t)img * a.nch + ch) * a.pstride;
            int32_t p1 = (int32_t)br.read(32);   // read_signed(32) twice (:161-162)
            int32_t p2 = (int32_t)br.read(32);
            if (br.eof) st = FELICS_ERR_IO;
            else if (a.npix >= 1) {
                if (p1 < -32768 || p1 > 32767 || (a.npix >= 2 && (p2 < -32768 || p2 > 32767))) st = FELICS_NEED_EXACT;
                else {
                    pl[0] = (int16_t)p1;
                    if (a.npix >= 2) pl[1] = (int16_t)p2;
                }
            }
            uint32_t x = 0, y = 0;
            if (a.npix >= 3) { x = 2 % w; y = 2 / w; }
            for (uint32_t i = 2; i < a.npix && st == FELICS_OK; i++) {
                uint32_t ia, ib;
                if (x > 0 && y > 0) { ia = i - 1; ib = i - w; }
                else if (y == 0) { ia = i - 1; ib = i - 2; }
                else if (y >= 2) { ia = i - w; ib = i - 2 * w; }
                else { ia = i - w; ib = i - w + 1; }
                int v1 = pl[ia], v2 = pl[ib];
                int hi = max(v1, v2), lo = min(v1, v2);
                uint32_t ctx = (uint32_t)(hi - lo);
                if (ctx > 510u) { st = FELICS_ERR_CORRUPT; break; }   // assert!(context <= max_context), parameter_selection.rs:72
                uint32_t *row = tab + ctx * NK;
                int value;
                if (br.read(1)) {                                       // InRange (:208-215)
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    uint32_t nn = ctx + 1;
                    int m = 31 - __clz(nn);
                    uint32_t left_p = nn - (1u << m), right_p = (2u << m) - nn;
                    uint32_t xx = br.read((uint32_t)m);
                    if (xx >= right_p) xx = (xx - right_p) * 2 + right_p + br.read(1);   // phase_in_coding.rs:102-109
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    xx += left_p;                                                        // rotate_left (:55-57)
                    if (xx >= nn) xx -= nn;
                    if (xx >= nn) { st = FELICS_ERR_CORRUPT; break; }
                    value = lo + (int)xx;
                } else {
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    uint32_t above = br.read(1);
                    uint32_t rr[NK];
#pragma unroll
                    for (int k = 0; k < NK; k++) rr[k] = row[k];
                    int k = argmin_last(rr);                              // get_k (:202)
                    uint32_t q = br.read_unary0();
                    uint32_t rem = br.read((uint32_t)k);
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    if (q > 70000u) { st = FELICS_NEED_EXACT; break; }
                    uint32_t e = (q << k) + rem;
                    uint32_t mn = 0xffffffffu;
#pragma unroll
                    for (int kk = 0; kk < NK; kk++) {                     // update (parameter_selection.rs:49-65)
                        rr[kk] += (e >> kk) + 1u + (uint32_t)kk;
                        mn = min(mn, rr[kk]);
                    }
                    if (mn > HALVE_AT) {
#pragma unroll
                        for (int kk = 0; kk < NK; kk++) rr[kk] >>= 1;
                    }
#pragma unroll
                    for (int kk = 0; kk < NK; kk++) row[kk] = rr[kk];
                    value = above ? hi + (int)e + 1 : lo - (int)e - 1;    // (:216-243)
                }
                if (value < -32768 || value > 32767) { st = FELICS_NEED_EXACT; break; }
                pl[i] = (int16_t)value;
                if (++x == w) { x = 0; y++; }
            }
        }
        st = __shfl_sync(0xffffffffu, st, 0);
    }
    if (lane == 0) a.status[img] = st;
}

// ---------------------------------------------------------------------------------------------
// Fast path: rows up to DEC_MAX_W samples.  Lane 0 is the decoder; everything it touches per pixel is
// on chip: the bit stream sits in a 64-bit register window refilled from words that were loaded 4..8
// words ahead, the previous and the current row live in shared memory (the left neighbour in a
// register), the estimator table (511 x 6 counters) in shared memory.  The warp writes each finished
// row to global memory with coalesced stores.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t DEC_MAX_W = 32768;

struct BitWindow {
    const uint32_t *words;
    uint64_t nwords;
    uint64_t wi;          // index of the word held in n0
    uint64_t win;         // unread bits, MSB first
    int navail;           // valid bits in win
    uint64_t used, limit; // bits consumed so far / bits in the file after the header
    uint32_t n0, n1;      // the next two words of the stream, loaded (and byte-swapped) well before they are needed

    __device__ __forceinline__ uint32_t load(uint64_t i) const { return i < nwords ? bswap32(__ldg(words + i)) : 0u; }
    __device__ __forceinline__ void init(const uint32_t *w, uint64_t nw, uint64_t bitpos, uint64_t bitend) {
        words = w; nwords = nw; used = 0; limit = bitend - bitpos;
        wi = bitpos >> 5;
        const uint32_t sh = (uint32_t)(bitpos & 31);
        const uint32_t first = load(wi);
        wi++;
        n0 = load(wi);
        n1 = load(wi + 1);
        win = (uint64_t)(first << sh) << 32;
        navail = 32 - (int)sh;
        refill();
    }
    // keep more than 32 valid bits in the window; the word consumed was requested two refills (>= 32 bits) ago
    __device__ __forceinline__ void refill() {
        if (navail <= 32) {
            win |= (uint64_t)n0 << (32 - navail);
            navail += 32;
            n0 = n1;
            wi++;
            n1 = load(wi + 1);
        }
    }
    __device__ __forceinline__ uint32_t peek32() const { return (uint32_t)(win >> 32); }
    __device__ __forceinline__ void skip(int n) { win <<= n; navail -= n; used += (uint32_t)n; }   // n <= 32, window refilled before
    __device__ __forceinline__ uint32_t read(int n) {   // n in 0..32
        if (n == 0) return 0;
        const uint32_t v = peek32() >> (32 - n);
        skip(n);
        refill();
        return v;
    }
    __device__ __forceinline__ bool eof() const { return used > limit; }
};

// Rows [y0, y1) of one plane, decoded by lane 0 and written out by the warp.  `rows` is the two-row buffer: when y0 > 0 the
// caller has put row y0 - 1 into its ((y0 & 1) ^ 1) half; col0_a / col0_b are the first samples of rows y0 - 1 and y0 - 2.
__device__ __forceinline__ int decode_rows(BitWindow &br, uint32_t *tab, int16_t *rows, int16_t *pl, uint32_t w, uint32_t y0, uint32_t y1,
                                       int32_t p1, int32_t p2, int col0_a, int col0_b, uint32_t lane, int st) {
    for (uint32_t y = y0; y < y1 && st == FELICS_OK; y++) {
        int16_t *cur = rows + (size_t)(y & 1u) * w;
        const int16_t *up = rows + (size_t)((y & 1u) ^ 1u) * w;
        if (lane == 0) {
            // neighbours (misc.rs:6-24) without a per-pixel case split: inside a row v1 is the previous sample and
            // v2 = upp[x]; `upp` is the row above, or this row shifted by two on the first row (i-1, i-2); the
            // first sample of a row takes (v1, v2) = (left, b0): up / up-up, or up / up-right on the second row
            uint32_t x = 0;
            int left = 0, b0 = 0;
            const int16_t *upp = up;
            if (y == 0) {
                cur[0] = (int16_t)p1; left = p1;
                x = 1;
                if (w >= 2) { cur[1] = (int16_t)p2; left = p2; x = 2; }
                upp = cur - 2;
            } else if (y == 1) {
                if (w == 1) { cur[0] = (int16_t)p2; x = 1; }     // 1-wide image: the second raw sample is (0, 1)
                else { left = up[0]; b0 = up[1]; }
            } else {
                left = col0_a; b0 = col0_b;
            }
            for (; x < w; x++) {
                const int v1 = left, v2 = x == 0 ? b0 : upp[x];
                const int hi = max(v1, v2), lo = min(v1, v2);
                const uint32_t ctx = (uint32_t)(hi - lo);
                if (ctx > 510u) { st = FELICS_ERR_CORRUPT; break; }   // assert!(context <= max_context), parameter_selection.rs:72
                const uint32_t top = br.peek32();
                int value;
                int used;   // bits of this pixel's code still to be skipped: one skip and one refill, behind the two classes
                if (top >> 31) {                                        // InRange (:208-215)
                    const uint32_t nn = ctx + 1;
                    const int m = 31 - __clz(nn);
                    const uint32_t left_p = nn - (1u << m), right_p = (2u << m) - nn;
                    // marker + m bits (+ 1): at most 11 bits, all inside the window
                    uint32_t xx = m ? ((top << 1) >> (32 - m)) : 0u;
                    used = 1 + m;
                    if (xx >= right_p) { xx = (xx - right_p) * 2 + right_p + ((top >> (30 - m)) & 1u); used++; }   // phase_in_coding.rs:102-109
                    xx += left_p;                                       // rotate_left (:55-57)
                    if (xx >= nn) xx -= nn;
                    if (xx >= nn) { st = FELICS_ERR_CORRUPT; break; }
                    value = lo + (int)xx;
                } else {
                    const uint32_t above = (top >> 30) & 1u;
                    uint32_t *row = tab + ctx * NK;
                    const uint2 r01 = *reinterpret_cast<const uint2 *>(row), r23 = *reinterpret_cast<const uint2 *>(row + 2), r45 = *reinterpret_cast<const uint2 *>(row + 4);
                    uint32_t rr[NK] = {r01.x, r01.y, r23.x, r23.y, r45.x, r45.y};
                    const uint32_t k = (uint32_t)argmin_last(rr);       // get_k (:202)
                    const uint32_t rest = top << 2;                     // the bits behind the two markers
                    const uint32_t ones = __clz(~rest);                 // the unary run, as far as the window shows it (at most 30)
                    uint32_t q, rem;
                    if (ones + k + 3u <= 32u) {
                        // the whole code lies in the window: two markers, `ones` ones, a zero, k remainder bits
                        q = ones;
                        rem = k ? (rest << (ones + 1u)) >> (32u - k) : 0u;
                        used = (int)(3u + ones + k);
                        if (br.used + (uint32_t)used > br.limit) { st = FELICS_ERR_IO; break; }
                    } else {
                        // a long run: word by word (read_unary0, then the remainder)
                        br.skip(2);
                        br.refill();
                        q = 0;
                        for (;;) {
                            const uint32_t o2 = __clz(~br.peek32());    // 32 when all ones
                            if (o2 < 32) { q += o2; br.skip((int)o2 + 1); br.refill(); break; }
                            q += 32; br.skip(32); br.refill();
                            if (br.eof()) break;
                        }
                        rem = br.read((int)k);
                        used = 0;
                        if (br.eof()) { st = FELICS_ERR_IO; break; }
                        if (q > 70000u) { st = FELICS_NEED_EXACT; break; }
                    }
                    const uint32_t e = (q << k) + rem;
                    uint32_t mn = 0xffffffffu;
#pragma unroll
                    for (int kk = 0; kk < NK; kk++) {                   // update (parameter_selection.rs:49-65)
                        rr[kk] += (e >> kk) + 1u + (uint32_t)kk;
                        mn = min(mn, rr[kk]);
                    }
                    if (mn > HALVE_AT) {
#pragma unroll
                        for (int kk = 0; kk < NK; kk++) rr[kk] >>= 1;
                    }
                    *reinterpret_cast<uint2 *>(row) = make_uint2(rr[0], rr[1]);
                    *reinterpret_cast<uint2 *>(row + 2) = make_uint2(rr[2], rr[3]);
                    *reinterpret_cast<uint2 *>(row + 4) = make_uint2(rr[4], rr[5]);
                    value = above ? hi + (int)e + 1 : lo - (int)e - 1;  // (:216-243)
                }
                br.skip(used);
                br.refill();
                if (value < -32768 || value > 32767) { st = FELICS_NEED_EXACT; break; }
                cur[x] = (int16_t)value;
                left = value;
            }
            if (br.eof()) st = FELICS_ERR_IO;   // the reference fails at the read that runs out of input, before any later check
            col0_b = col0_a;
            col0_a = cur[0];
        }
        st = __shfl_sync(0xffffffffu, st, 0);
        __syncwarp();
        if (st == FELICS_OK) {
            int16_t *dst = pl + (size_t)y * w;
            for (uint32_t x = lane; x < w; x += 32) dst[x] = cur[x];
        }
        __syncwarp();
    }
    return st;
}

__global__ void __launch_bounds__(32) k_decode(DecArgs a, uint32_t n) {
    extern __shared__ __align__(16) unsigned char dec_smem[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(dec_smem);                         // [(NBIN-1) * NK]
    int16_t *rows = reinterpret_cast<int16_t *>(dec_smem + (NBIN - 1) * NK * 4 + 8);  // [2][w]
    const uint32_t img = blockIdx.x;
    if (img >= n) return;
    const uint32_t lane = threadIdx.x;
    const uint64_t off0 = a.offsets[img], off1 = a.offsets[img + 1];
    const uint64_t len = off1 - off0;
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(a.words);
    int st = FELICS_OK;
    // read_header order (format.rs:63-84), then decompress_with_header's checks (compression.rs:289-294)
    if (lane == 0) {
        const uint8_t *hb = bytes + off0;
        if (len < 4) st = FELICS_ERR_IO;
        else if (hb[0] != 'F' || hb[1] != 'L' || hb[2] != 'C' || hb[3] != 'S') st = FELICS_ERR_INVALID_SIGNATURE;
        else if (len < 5) st = FELICS_ERR_IO;
        else if (hb[4] > 1) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (len < 6) st = FELICS_ERR_IO;
        else if (hb[5] > 1) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else if (len < FELICS_HEADER_BYTES) st = FELICS_ERR_IO;
        else if (hb[4] != a.color) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (hb[5] != a.depth) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else {
            uint32_t w = ((uint32_t)hb[6] << 24) | ((uint32_t)hb[7] << 16) | ((uint32_t)hb[8] << 8) | hb[9];
            uint32_t h = ((uint32_t)hb[10] << 24) | ((uint32_t)hb[11] << 16) | ((uint32_t)hb[12] << 8) | hb[13];
            if (w != a.w || h != a.h) st = FELICS_ERR_INVALID_DIMENSIONS;
        }
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    if (st != FELICS_OK) { if (lane == 0) a.status[img] = st; return; }

    const uint32_t w = a.w, h = a.h;
    BitWindow br;
    if (lane == 0) br.init(a.words, a.arena_words, 8 * (off0 + FELICS_HEADER_BYTES), 8 * off1);

    for (uint32_t ch = 0; ch < a.nch && st == FELICS_OK; ch++) {
        for (uint32_t j = lane; j < (NBIN - 1) * NK; j += 32) tab[j] = 0;   // fresh estimator per channel (:186-190)
        __syncwarp();
        int16_t *pl = a.planes + ((size_t)img * a.nch + ch) * a.pstride;
        int32_t p1 = 0, p2 = 0;
        if (lane == 0) {
            p1 = (int32_t)br.read(32);   // read_signed(32) twice (:161-162)
            p2 = (int32_t)br.read(32);
            if (br.eof()) st = FELICS_ERR_IO;
            else if (a.npix >= 1 && (p1 < -32768 || p1 > 32767 || (a.npix >= 2 && (p2 < -32768 || p2 > 32767)))) st = FELICS_NEED_EXACT;
        }
        st = __shfl_sync(0xffffffffu, st, 0);
        if (st != FELICS_OK || a.npix == 0) continue;
        st = decode_rows(br, tab, rows, pl, w, 0, h, p1, p2, 0, 0, lane, st);
    }
    if (lane == 0) a.status[img] = st;
}

// ---------------------------------------------------------------------------------------------
// Batches of 8-bit gray files (BASELINE.json configs[3]): F files per warp, one per lane.  A .fel file is one serial bit
// chain (compression.rs:151-248), so the batch is the only parallelism -- but a warp that decodes one file on one lane
// issues every instruction for a single pixel.  Here lanes 0..F-1 of a warp run the same loop on F different files of the
// same shape: they stay together row by row and pixel by pixel (only the class of the pixel splits them), so an issued
// instruction serves up to F pixels.  Everything a lane touches per pixel is on chip, interleaved by lane so that lanes at
// the same position hit different banks: the estimator rows as u16 pairs plus the k that the NEXT out-of-range pixel of the
// context will use (get_k is computed at update time, off the critical path: one 16-byte load gives the decoder its k), the
// previous and the current row as bytes.  The pixels leave as bytes, row by row, written by the whole warp: no i16 planes
// and no unplane pass.  A sample outside 0..255 (no valid file; the reference goes on in i32, compression.rs:305-310), a raw
// sample outside 0..255 or a residual above 255 hand the file to k_decode_exact (FELICS_NEED_EXACT), which also keeps the
// u16 counters exact: with residuals <= 255 a counter stays below 2 * (256 / 13) * 1030 < 2^16.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t G8_MAX_W = 32768;    // one pixel row per file in shared memory
constexpr int G8_WARPS = 2;
constexpr int G8_CTX = 256;            // contexts of 8-bit gray samples
constexpr int G8_HOT = 16;             // contexts whose estimator rows live in shared memory (measured on the configs[3] tiles, 32 files per warp:
                                       // 8 -> 39.7, 16 -> 47.6, 32 -> 26.1 GPixel/s: fewer rows = more warps per SM, more rows = fewer trips to global memory)

struct G8Args {
    const uint32_t *words;
    uint64_t arena_words;
    const uint64_t *offsets;
    uint8_t *pixels;           // file i at pixels + i * npix
    int *status;
    uint4 *cold;               // estimator rows of the contexts >= HOT: [slot][G8_CTX - HOT], tagged with the file (zeroed before the launch)
    uint32_t *ticket;          // next group of F files (zeroed before the launch)
    uint32_t w, h, npix;       // w a multiple of 4, 4 <= w <= G8_MAX_W
    uint32_t word_out;         // pixels is 4-byte aligned: rows leave as words
};

// read_header order (format.rs:63-84), then decompress_with_header's checks (compression.rs:289-294)
__device__ __forceinline__ int check_file_header(const uint8_t *hb, uint64_t len, uint8_t color, uint8_t depth, uint32_t w, uint32_t h) {
    if (len < 4) return FELICS_ERR_IO;
    if (hb[0] != 'F' || hb[1] != 'L' || hb[2] != 'C' || hb[3] != 'S') return FELICS_ERR_INVALID_SIGNATURE;
    if (len < 5) return FELICS_ERR_IO;
    if (hb[4] > 1) return FELICS_ERR_INVALID_COLOR_TYPE;
    if (len < 6) return FELICS_ERR_IO;
    if (hb[5] > 1) return FELICS_ERR_INVALID_PIXEL_DEPTH;
    if (len < FELICS_HEADER_BYTES) return FELICS_ERR_IO;
    if (hb[4] != color) return FELICS_ERR_INVALID_COLOR_TYPE;
    if (hb[5] != depth) return FELICS_ERR_INVALID_PIXEL_DEPTH;
    const uint32_t fw = ((uint32_t)hb[6] << 24) | ((uint32_t)hb[7] << 16) | ((uint32_t)hb[8] << 8) | hb[9];
    const uint32_t fh = ((uint32_t)hb[10] << 24) | ((uint32_t)hb[11] << 16) | ((uint32_t)hb[12] << 8) | hb[13];
    if (fw != w || fh != h) return FELICS_ERR_INVALID_DIMENSIONS;
    return FELICS_OK;
}

// shared memory of one warp: the hot estimator rows and one pixel row of each of its F files
__host__ __device__ inline size_t g8_warp_bytes(int F, int hot, uint32_t w) { return (size_t)F * hot * 16 + (((size_t)F * (w / 4) * 4 + 15) & ~(size_t)15); }

// F files per warp, one per lane; the estimator rows of the contexts below HOT live in shared memory, the others (rare in
// smooth images, and a file's lanes only meet them one pixel at a time) in a tagged table in global memory.  Warps take
// groups of F files from a ticket until the batch is done.
template <int F, int HOT>
__global__ void __launch_bounds__(32 * G8_WARPS) k_decode_g8(G8Args a, uint32_t n) {
    extern __shared__ __align__(16) unsigned char g8_smem[];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t wq = a.w >> 2;                                          // words per row
    unsigned char *wbase = g8_smem + (size_t)wid * g8_warp_bytes(F, HOT, a.w);
    uint4 *tab = reinterpret_cast<uint4 *>(wbase);                          // [HOT][F]: c0|c1<<16, c2|c3<<16, c4|c5<<16, next k
    uint32_t *rows = reinterpret_cast<uint32_t *>(wbase + (size_t)F * HOT * 16);   // [wq][F]: four samples per word; one row, overwritten in place
                                                                                   // (sample x of the row above is last needed when sample x is decoded)
    uint4 *mytab = tab + lane;
    uint4 *mycold = a.cold + ((size_t)(blockIdx.x * G8_WARPS + wid) * F + (lane < F ? lane : 0)) * (G8_CTX - HOT);

    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(a.ticket, 1u) * F;
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= n) break;
        const uint32_t img = first + lane;
        const bool mine = lane < F && img < n;
        const uint32_t tag = (img + 1u) << 8;                               // marks the cold rows this file has written
        for (uint32_t j = lane; j < (uint32_t)HOT * F; j += 32) tab[j] = make_uint4(0u, 0u, 0u, (uint32_t)(NK - 1));   // fresh estimator: get_k of equal counts is the last k
        __syncwarp();

        int st = FELICS_OK;
        BitWindow br;
        int p1 = 0, p2 = 0;
        if (mine) {
            const uint64_t off0 = a.offsets[img], off1 = a.offsets[img + 1];
            st = check_file_header(reinterpret_cast<const uint8_t *>(a.words) + off0, off1 - off0, 0, 0, a.w, a.h);
            if (st == FELICS_OK) {
                br.init(a.words, a.arena_words, 8 * (off0 + FELICS_HEADER_BYTES), 8 * off1);
                p1 = (int)br.read(32);   // read_signed(32) twice (:161-162)
                p2 = (int)br.read(32);
                if (br.eof()) st = FELICS_ERR_IO;
                else if ((uint32_t)p1 > 255u || (uint32_t)p2 > 255u) st = FELICS_NEED_EXACT;
            }
        }

        // one pixel (compression.rs:193-246) from its two neighbours; leaves st != OK when the file cannot go on here
        auto dec = [&](int v1, int v2) -> int {
            const int hi = max(v1, v2), lo = min(v1, v2);
            const uint32_t ctx = (uint32_t)(hi - lo);                          // <= 255: every sample so far is 0..255
            const uint32_t top = br.peek32();
            // the in-range reading of the window (phase_in_coding.rs:102-109, :55-57) is computed for every lane: it is short, and
            // the lanes whose pixel is out of range then go through their branch alone instead of waiting for this one first
            int value, used;
            {
                const uint32_t nn = ctx + 1;
                const int m = 31 - __clz(nn);
                const uint32_t left_p = nn - (1u << m), right_p = (2u << m) - nn;
                uint32_t xx = m ? ((top << 1) >> (32 - m)) : 0u;               // marker + m bits (+ 1): at most 10 bits, all inside the window
                used = 1 + m;
                if (xx >= right_p) { xx = (xx - right_p) * 2 + right_p + ((top >> (30 - m)) & 1u); used++; }
                xx += left_p;                                                   // rotate_left; xx < nn before
                if (xx >= nn) xx -= nn;
                value = lo + (int)xx;
            }
            if (!(top >> 31)) {                                                 // not InRange (:216-243)
                const uint32_t above = (top >> 30) & 1u;
                uint4 *rowp = ctx < (uint32_t)HOT ? mytab + ctx * F : mycold + (ctx - HOT);
                uint4 r = *rowp;
                if (ctx >= (uint32_t)HOT) {
                    if ((r.w & 0xffffff00u) != tag) r = make_uint4(0u, 0u, 0u, (uint32_t)(NK - 1));   // not yet touched by this file
                    r.w &= 255u;
                }
                const uint32_t k = r.w;                                         // get_k (:202), computed when the row was last updated
                const uint32_t rest = top << 2;                                 // the bits behind the two markers
                const uint32_t ones = __clz(~rest);                             // the unary run, as far as the window shows it (at most 30)
                uint32_t q, rem;
                if (ones + k + 3u <= 32u) {
                    // the whole code lies in the window: two markers, `ones` ones, a zero, k remainder bits
                    q = ones;
                    rem = k ? (rest << (ones + 1u)) >> (32u - k) : 0u;
                    used = (int)(3u + ones + k);
                    if (br.used + (uint32_t)used > br.limit) { st = FELICS_ERR_IO; return 0; }
                } else {
                    // a long run: word by word (read_unary0, then the remainder)
                    br.skip(2);
                    br.refill();
                    q = 0;
                    for (;;) {
                        const uint32_t o2 = __clz(~br.peek32());                // 32 when all ones
                        if (o2 < 32) { q += o2; br.skip((int)o2 + 1); br.refill(); break; }
                        q += 32; br.skip(32); br.refill();
                        if (br.eof() || q > 255u) break;
                    }
                    rem = br.read((int)k);
                    used = 0;
                    if (br.eof()) { st = FELICS_ERR_IO; return 0; }
                }
                const uint32_t e = (q << k) + rem;
                if (q > 255u || e > 255u) { st = FELICS_NEED_EXACT; return 0; }
                // update (parameter_selection.rs:49-65) on u16 pairs: cost of e under k = (e >> k) + 1 + k
                r.x += (e + 1u) | (((e >> 1) + 2u) << 16);
                r.y += ((e >> 2) + 3u) | (((e >> 3) + 4u) << 16);
                r.z += ((e >> 4) + 5u) | (((e >> 5) + 6u) << 16);
                const uint32_t m2 = __vminu2(__vminu2(r.x, r.y), r.z);
                if (min(m2 & 0xffffu, m2 >> 16) > HALVE_AT) {
                    r.x = (r.x >> 1) & 0x7fff7fffu; r.y = (r.y >> 1) & 0x7fff7fffu; r.z = (r.z >> 1) & 0x7fff7fffu;
                }
                const uint32_t cc[NK] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16, r.z & 0xffffu, r.z >> 16};
                r.w = (uint32_t)argmin_last(cc) | (ctx >= (uint32_t)HOT ? tag : 0u);
                *rowp = r;
                value = above ? hi + (int)e + 1 : lo - (int)e - 1;              // (:216-243)
            }
            br.skip(used);
            br.refill();
            if ((uint32_t)value > 255u) st = FELICS_NEED_EXACT;
            return value;
        };

        int col0_b = 0;   // first sample of row y - 2
        for (uint32_t y = 0; y < a.h; y++) {
            uint32_t *cur = rows + lane;                                        // word g of my row at cur[g * F]
            const uint32_t *up = cur;                                           // ... which holds the row above until it is overwritten
            if (mine && st == FELICS_OK) {
                if (y == 0) {
                    // first row: the neighbours are the two samples to the left (misc.rs:8-9)
                    int l2 = p1, l1 = p2;
                    uint32_t wv = (uint32_t)p1 | ((uint32_t)p2 << 8);
                    for (uint32_t x = 2; x < a.w && st == FELICS_OK; x++) {
                        const int v = dec(l1, l2);
                        l2 = l1; l1 = v;
                        if ((x & 3u) == 0) wv = 0;
                        wv |= (uint32_t)(v & 255) << (8 * (x & 3u));
                        if ((x & 3u) == 3u) cur[(x >> 2) * F] = wv;
                    }
                } else {
                    // first column: up and up-right on the second row, up and up-up below it (misc.rs:13-17); then left and up
                    const uint32_t u0 = up[0];
                    int left = dec((int)(u0 & 255u), y == 1 ? (int)((u0 >> 8) & 255u) : col0_b);
                    uint32_t wv = (uint32_t)(left & 255);
                    col0_b = (int)(u0 & 255u);
#pragma unroll
                    for (int j = 1; j < 4; j++) {
                        if (st == FELICS_OK) {
                            left = dec(left, (int)((u0 >> (8 * j)) & 255u));
                            wv |= (uint32_t)(left & 255) << (8 * j);
                        }
                    }
                    cur[0] = wv;
                    for (uint32_t g = 1; g < wq && st == FELICS_OK; g++) {
                        const uint32_t uw = up[g * F];
                        left = dec(left, (int)(uw & 255u));
                        wv = (uint32_t)(left & 255);
#pragma unroll
                        for (int j = 1; j < 4; j++) {
                            if (st == FELICS_OK) {
                                left = dec(left, (int)((uw >> (8 * j)) & 255u));
                                wv |= (uint32_t)(left & 255) << (8 * j);
                            }
                        }
                        cur[g * F] = wv;
                    }
                }
                if (st == FELICS_OK && br.eof()) st = FELICS_ERR_IO;   // the reference fails at the read that runs out of input, before any later check
            }
            __syncwarp();
            // the F finished rows leave, every one written by the whole warp (rows of a file that has failed are never read: the
            // status says so, and a file handed to the exact decoder is written again)
#pragma unroll 1
            for (int f = 0; f < F; f++) {
                if (first + f >= n) break;
                uint8_t *dst = a.pixels + (size_t)(first + f) * a.npix + (size_t)y * a.w;
                const uint32_t *src = rows + f;
                if (a.word_out) {
                    for (uint32_t g = lane; g < wq; g += 32) reinterpret_cast<uint32_t *>(dst)[g] = src[g * F];
                } else {
                    for (uint32_t x = lane; x < a.w; x += 32) dst[x] = (uint8_t)(src[(x >> 2) * F] >> (8 * (x & 3u)));
                }
            }
            __syncwarp();
        }
        if (mine) a.status[img] = st;
    }
}

template <int F, int HOT>
int launch_decode_g8(felics_ctx *ctx, G8Args g, size_t n, cudaStream_t st) {
    const size_t smem = G8_WARPS * g8_warp_bytes(F, HOT, g.w);
    FELICS_CUDA_TRY(cudaFuncSetAttribute(k_decode_g8<F, HOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, sms = 148;
    FELICS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode_g8<F, HOT>, 32 * G8_WARPS, smem));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const size_t groups = (n + F - 1) / F;
    const unsigned blocks = (unsigned)std::min<size_t>((groups + G8_WARPS - 1) / G8_WARPS, (size_t)std::max(per_sm, 1) * sms);
    const size_t cold_bytes = (size_t)blocks * G8_WARPS * F * (G8_CTX - HOT) * sizeof(uint4);
    int rc = ensure_buffer(ctx, &ctx->g8_cold, &ctx->g8_cold_cap, cold_bytes + 256);
    if (rc) return rc;
    g.ticket = (uint32_t *)ctx->g8_cold;
    g.cold = (uint4 *)((uint8_t *)ctx->g8_cold + 256);
    FELICS_CUDA_TRY(cudaMemsetAsync(ctx->g8_cold, 0, cold_bytes + 256, st));
    k_decode_g8<F, HOT><<<blocks, 32 * G8_WARPS, smem, st>>>(g, (uint32_t)n);
    return FELICS_OK;
}

// planes -> pixels with the try_into range checks (compression.rs:305-310, :402-407)
// distance between consecutive planes, in samples.  One warp per file writes its rows in lockstep with the others:
// a power-of-two distance would put every file's stores on the same memory channel.
inline size_t plane_stride8(uint32_t npix) { return (size_t)npix + 4672; }

__global__ void k_unplane_gray8(const int16_t *__restrict__ planes, uint8_t *__restrict__ px, uint32_t npix, size_t pstride, size_t total,
                                int *__restrict__ status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        size_t img = i / npix;
        if (status[img] != FELICS_OK) continue;
        int v = planes[img * pstride + (i - img * npix)];
        if (v < 0 || v > 255) { atomicCAS(&status[img], FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
        px[i] = (uint8_t)v;
    }
}

// color_transform.rs:20-26
__global__ void k_unplane_rgb8(const int16_t *__restrict__ planes, uint8_t *__restrict__ px, uint32_t npix, size_t pstride, size_t total,
                               int *__restrict__ status) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        size_t img = idx / npix;
        if (status[img] != FELICS_OK) continue;
        uint32_t i = (uint32_t)(idx - img * npix);
        const int16_t *base = planes + img * 3 * pstride;
        int y = base[i], co = base[pstride + i], cg = base[2 * pstride + i];
        int t = y - cg / 2;
        int g = cg + t;
        int b = t - co / 2;
        int r = b + co;
        if (r < 0 || r > 255 || g < 0 || g > 255 || b < 0 || b > 255) { atomicCAS(&status[img], FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
        px[3 * idx] = (uint8_t)r; px[3 * idx + 1] = (uint8_t)g; px[3 * idx + 2] = (uint8_t)b;
    }
}

// ---------------------------------------------------------------------------------------------
// 16-bit samples: the reference loop as written, one warp per file (serial16.cuh)
// ---------------------------------------------------------------------------------------------
struct Dec16Args {
    const uint32_t *words;
    uint64_t arena_words;
    const uint64_t *offsets;
    int32_t *planes;           // [n*nch][pstride]
    size_t pstride;
    uint32_t *tables;          // [n][TABLE16_WORDS], tagged rows (serial16.cuh)
    uint32_t tag0;             // first tag of this call (one per channel)
    uint32_t rows_smem;        // 1: two rows of samples fit in shared memory beside the estimator rows
    int *status;
    uint32_t w, h, npix, nch;
    uint8_t color, depth;
};

__global__ void __launch_bounds__(32) k16_decode(Dec16Args a, uint32_t n) {
    const uint32_t img = blockIdx.x;
    if (img >= n) return;
    const uint32_t lane = threadIdx.x;
    const uint64_t off0 = a.offsets[img], off1 = a.offsets[img + 1];
    const uint64_t len = off1 - off0;
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(a.words);
    int st = FELICS_OK;
    if (lane == 0) {   // read_header order (format.rs:63-84), then decompress_with_header's checks
        const uint8_t *hb = bytes + off0;
        if (len < 4) st = FELICS_ERR_IO;
        else if (hb[0] != 'F' || hb[1] != 'L' || hb[2] != 'C' || hb[3] != 'S') st = FELICS_ERR_INVALID_SIGNATURE;
        else if (len < 5) st = FELICS_ERR_IO;
        else if (hb[4] > 1) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (len < 6) st = FELICS_ERR_IO;
        else if (hb[5] > 1) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else if (len < FELICS_HEADER_BYTES) st = FELICS_ERR_IO;
        else if (hb[4] != a.color) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (hb[5] != a.depth) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else {
            uint32_t w = ((uint32_t)hb[6] << 24) | ((uint32_t)hb[7] << 16) | ((uint32_t)hb[8] << 8) | hb[9];
            uint32_t h = ((uint32_t)hb[10] << 24) | ((uint32_t)hb[11] << 16) | ((uint32_t)hb[12] << 8) | hb[13];
            if (w != a.w || h != a.h) st = FELICS_ERR_INVALID_DIMENSIONS;
        }
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    if (st != FELICS_OK) { if (lane == 0) a.status[img] = st; return; }

    BitReader br;
    br.init(a.words, a.arena_words, 8 * (off0 + FELICS_HEADER_BYTES), 8 * off1);
    extern __shared__ __align__(16) unsigned char dec16_smem[];
    Est16 est;
    est.sm = reinterpret_cast<uint32_t *>(dec16_smem);
    est.gl = a.tables + (size_t)img * TABLE16_WORDS;
    est.tag = 0;
    // previous and current row in shared memory when they fit (a.rows_smem), else the neighbours come from the planes
    int32_t *srow = reinterpret_cast<int32_t *>(dec16_smem + SM_TABLE16_BYTES);
    const uint32_t w = a.w, h = a.h;
    for (uint32_t ch = 0; ch < a.nch && st == FELICS_OK; ch++) {
        est.reset(lane, a.tag0 + ch);
        int32_t *pl = a.planes + ((size_t)img * a.nch + ch) * a.pstride;
        int32_t p1 = 0, p2 = 0;
        if (lane == 0) {
            p1 = (int32_t)br.read(32); p2 = (int32_t)br.read(32);   // read_signed(32) twice (:161-162)
            if (br.eof) st = FELICS_ERR_IO;
        }
        st = __shfl_sync(0xffffffffu, st, 0);
        if (st != FELICS_OK || a.npix == 0) continue;
        long long col0_a = 0, col0_b = 0;   // samples at x = 0 of the previous row and of the row before it
        for (uint32_t y = 0; y < h && st == FELICS_OK; y++) {
            int32_t *cur = a.rows_smem ? srow + (size_t)(y & 1u) * w : pl + (size_t)y * w;
            const int32_t *up = a.rows_smem ? srow + (size_t)((y & 1u) ^ 1u) * w : pl + (size_t)(y ? y - 1 : 0) * w;   // unused on the first row
            if (lane == 0) {
                // neighbours (misc.rs:6-24) as in k_decode: v1 = previous sample, v2 = upp[x]; first sample of a row: (left, b0)
                uint32_t x = 0;
                long long left = 0, b0 = 0;
                const int32_t *upp = up;
                if (y == 0) {
                    cur[0] = p1; left = p1;
                    x = 1;
                    if (w >= 2) { cur[1] = p2; left = p2; x = 2; }
                    upp = cur - 2;
                } else if (y == 1) {
                    if (w == 1) { cur[0] = p2; x = 1; }
                    else { left = up[0]; b0 = up[1]; }
                } else {
                    left = col0_a; b0 = col0_b;
                }
                for (; x < w; x++) {
                    const long long v1 = left, v2 = x == 0 ? b0 : (long long)upp[x];
                    const long long hi = max(v1, v2), lo = min(v1, v2);
                    if (hi - lo > (long long)MAXCTX16) { st = FELICS_ERR_CORRUPT; break; }   // assert!(context <= max_context), parameter_selection.rs:72
                    const uint32_t ctx = (uint32_t)(hi - lo);
                    long long value;
                    if (br.read(1)) {                                       // InRange (:208-215)
                        if (br.eof) { st = FELICS_ERR_IO; break; }
                        const uint32_t nn = ctx + 1;
                        const int m = 31 - __clz(nn);
                        const uint32_t left_p = nn - (1u << m), right_p = (2u << m) - nn;
                        uint32_t xx = br.read((uint32_t)m);
                        if (xx >= right_p) xx = (xx - right_p) * 2 + right_p + br.read(1);   // phase_in_coding.rs:102-109
                        if (br.eof) { st = FELICS_ERR_IO; break; }
                        xx += left_p;                                                        // rotate_left (:55-57)
                        if (xx >= nn) xx -= nn;
                        if (xx >= nn) { st = FELICS_ERR_CORRUPT; break; }
                        value = lo + (long long)xx;
                    } else {
                        if (br.eof) { st = FELICS_ERR_IO; break; }
                        const uint32_t above = br.read(1);
                        uint32_t cnt[NK16];
                        est.load(ctx, cnt);
                        const int k = get_k16(cnt);                         // get_k (:202)
                        const uint32_t q = br.read_unary0();
                        const uint32_t rem = br.read((uint32_t)k);
                        if (br.eof) { st = FELICS_ERR_IO; break; }
                        if (((unsigned long long)q << k) > 0xFFFFFFFFull) { st = FELICS_ERR_CORRUPT; break; }   // checked_mul(..).unwrap(), rice_coding.rs:50
                        const uint32_t e = (q << k) + rem;
                        update16(cnt, e);
                        est.store(ctx, cnt);
                        if (e > 0x7FFFFFFFu) { st = FELICS_ERR_INVALID_VALUE; break; }   // u32 -> i32 try_into (compression.rs:222-224, :234-236)
                        value = above ? hi + (long long)e + 1 : lo - (long long)e - 1;   // (:216-243)
                    }
                    if (value < -2147483648ll || value > 2147483647ll) { st = FELICS_ERR_VALUE_OVERFLOW; break; }   // checked_add / checked_sub
                    cur[x] = (int32_t)value;
                    left = value;
                }
                col0_b = col0_a;
                col0_a = cur[0];
            }
            st = __shfl_sync(0xffffffffu, st, 0);
            __syncwarp();
            if (a.rows_smem && st == FELICS_OK) {
                int32_t *dst = pl + (size_t)y * w;
                for (uint32_t x = lane; x < w; x += 32) dst[x] = cur[x];
            }
            __syncwarp();
        }
    }
    if (lane == 0) a.status[img] = st;
}

// i32 planes -> u16 pixels with the try_into range checks (compression.rs:305-310, :402-407)
__global__ void k16_unplane_gray(const int32_t *__restrict__ planes, uint16_t *__restrict__ px, uint32_t npix, size_t pstride, size_t total, int *__restrict__ status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const size_t img = i / npix;
        if (status[img] != FELICS_OK) continue;
        const int v = planes[img * pstride + (i - img * npix)];
        if (v < 0 || v > 65535) { atomicCAS(&status[img], FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
        px[i] = (uint16_t)v;
    }
}
__global__ void k16_unplane_rgb(const int32_t *__restrict__ planes, uint16_t *__restrict__ px, uint32_t npix, size_t pstride, size_t total, int *__restrict__ status) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        const size_t img = idx / npix;
        if (status[img] != FELICS_OK) continue;
        const uint32_t i = (uint32_t)(idx - img * npix);
        const int32_t *base = planes + img * 3 * pstride;
        const long long y = base[i], co = base[pstride + i], cg = base[2 * pstride + i];
        const long long t = y - cg / 2;      // color_transform.rs:20-26
        const long long g = cg + t;
        const long long b = t - co / 2;
        const long long r = b + co;
        if (r < 0 || r > 65535 || g < 0 || g > 65535 || b < 0 || b > 65535) { atomicCAS(&status[img], FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
        px[3 * idx] = (uint16_t)r; px[3 * idx + 1] = (uint16_t)g; px[3 * idx + 2] = (uint16_t)b;
    }
}

static int decode16_batch_device(felics_ctx *ctx, size_t n, const uint8_t *d_arena, const uint64_t *offsets_host, const felics_header &hdr,
                                 void *d_pixels_out, int *status_host) {
    const uint64_t npix64 = (uint64_t)hdr.width * hdr.height;
    cudaStream_t st = ctx->stream;
    const uint32_t npix = (uint32_t)npix64, nch = hdr.color_type ? 3 : 1;
    const size_t sub = std::max<size_t>(1, std::min<size_t>(n, 64));
    size_t off_bytes = align_up((n + 1) * sizeof(uint64_t), 256);
    size_t stat_bytes = align_up(n * sizeof(int), 256);
    const size_t pstride = plane_stride16(npix);
    size_t plane_bytes = align_up((sub * nch * pstride + 8) * sizeof(int32_t), 256);
    int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, off_bytes + stat_bytes + plane_bytes);
    if (rc) return rc;
    uint32_t *d_tables = nullptr;
    if ((rc = tables16(ctx, sub, &d_tables))) return rc;
    const size_t row_bytes = 2 * (size_t)hdr.width * sizeof(int32_t);
    const bool rows_smem = SM_TABLE16_BYTES + row_bytes <= (size_t)200 * 1024;
    const size_t smem = SM_TABLE16_BYTES + (rows_smem ? row_bytes : 0);
    if (smem > ctx->dec16_smem_set) {
        FELICS_CUDA_TRY(cudaFuncSetAttribute(k16_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->dec16_smem_set = smem;
    }
    uint8_t *sb = (uint8_t *)ctx->scratch;
    uint64_t *d_off = (uint64_t *)sb;
    int *d_status = (int *)(sb + off_bytes);
    int32_t *d_planes = (int32_t *)(sb + off_bytes + stat_bytes);
    FELICS_CUDA_TRY(cudaMemcpyAsync(d_off, offsets_host, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const size_t img_bytes = (size_t)npix * nch * 2;
    for (size_t first = 0; first < n; first += sub) {
        const size_t ni = std::min(sub, n - first);
        Dec16Args a;
        a.words = (const uint32_t *)d_arena;
        a.arena_words = (offsets_host[n] + 3) / 4;
        a.offsets = d_off + first; a.planes = d_planes; a.pstride = pstride; a.tables = d_tables; a.status = d_status + first;
        a.w = hdr.width; a.h = hdr.height; a.npix = npix; a.nch = nch; a.color = hdr.color_type; a.depth = hdr.pixel_depth;
        {
            StageScope s(ctx, ST_DECODE);
            a.tag0 = next_tags16(ctx);
            a.rows_smem = rows_smem ? 1u : 0u;
            k16_decode<<<(unsigned)ni, 32, smem, st>>>(a, (uint32_t)ni);
            s.launched();
        }
        if (npix > 0) {
            StageScope s(ctx, ST_UNPLANE);
            const size_t total = ni * (size_t)npix;
            const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
            uint16_t *out = (uint16_t *)((uint8_t *)d_pixels_out + first * img_bytes);
            if (nch == 1) k16_unplane_gray<<<blocks, 256, 0, st>>>(d_planes, out, npix, pstride, total, d_status + first);
            else k16_unplane_rgb<<<blocks, 256, 0, st>>>(d_planes, out, npix, pstride, total, d_status + first);
            s.launched();
        }
    }
    FELICS_CUDA_TRY(cudaMemcpyAsync(status_host, d_status, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    FELICS_CUDA_TRY(cudaGetLastError());
    rc = profile_collect(ctx);
    if (rc) return rc;
    for (size_t i = 0; i < n; i++)
        if (status_host[i] != FELICS_OK) return status_host[i];
    return FELICS_OK;
}

// ---------------------------------------------------------------------------------------------
// Band decode with a sidecar (sidecar.cuh; opt-in, not the reference format): one warp per (plane, band) starts from the
// recorded bit position, estimator table and row above, and must end exactly where the next band starts.
// ---------------------------------------------------------------------------------------------
struct BandArgs {
    const uint32_t *words;
    uint64_t nwords, fel_bytes;
    const uint8_t *entries;
    size_t entry_bytes;
    uint32_t band_rows, nbands;
    int16_t *planes;
    size_t pstride;
    int *status;
    uint32_t w, h, npix, nch;
};

__global__ void __launch_bounds__(32) k_decode_bands(BandArgs a) {
    extern __shared__ __align__(16) unsigned char dec_smem[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(dec_smem);
    int16_t *rows = reinterpret_cast<int16_t *>(dec_smem + (NBIN - 1) * NK * 4 + 8);
    const uint32_t lane = threadIdx.x;
    const uint32_t p = blockIdx.x / a.nbands, j = blockIdx.x - p * a.nbands;
    const uint32_t y0 = j * a.band_rows, y1 = min(a.h, y0 + a.band_rows);
    if (y0 >= a.h) return;
    const uint8_t *ent = a.entries + (size_t)blockIdx.x * a.entry_bytes;
    const uint64_t bit = *reinterpret_cast<const uint64_t *>(ent);
    const bool last = blockIdx.x + 1 == gridDim.x;   // entries are in stream order: the bands of plane 0, then of plane 1, ...
    const uint64_t next_bit = last ? 0 : *reinterpret_cast<const uint64_t *>(ent + a.entry_bytes);
    int st = FELICS_OK;
    if (bit < 8ull * FELICS_HEADER_BYTES || bit > 8ull * a.fel_bytes) st = FELICS_ERR_CORRUPT;
    const uint32_t *stab = reinterpret_cast<const uint32_t *>(ent + 16);
    for (uint32_t t = lane; t < (NBIN - 1) * NK; t += 32) tab[t] = j ? stab[t] : 0u;
    int col0_a = 0, col0_b = 0;
    if (j) {
        const int16_t *srow = reinterpret_cast<const int16_t *>(ent + 16 + (size_t)SIDECAR_TABLE_WORDS * 4);
        int16_t *up = rows + (size_t)((y0 & 1u) ^ 1u) * a.w;
        for (uint32_t x = lane; x < a.w; x += 32) up[x] = srow[x];
        col0_a = srow[0];
        col0_b = reinterpret_cast<const int32_t *>(ent)[2];
    }
    __syncwarp();
    BitWindow br;
    int32_t p1 = 0, p2 = 0;
    if (lane == 0 && st == FELICS_OK) {
        br.init(a.words, a.nwords, bit, 8 * a.fel_bytes);
        if (j == 0) {
            p1 = (int32_t)br.read(32);   // read_signed(32) twice (compression.rs:161-162)
            p2 = (int32_t)br.read(32);
            if (br.eof()) st = FELICS_ERR_IO;
            else if (p1 < -32768 || p1 > 32767 || (a.npix >= 2 && (p2 < -32768 || p2 > 32767))) st = FELICS_NEED_EXACT;
        }
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    int16_t *pl = a.planes + (size_t)p * a.pstride;
    if (st == FELICS_OK) st = decode_rows(br, tab, rows, pl, a.w, y0, y1, p1, p2, col0_a, col0_b, lane, st);
    if (lane == 0) {
        if (st == FELICS_OK && !last && bit + br.used != next_bit) st = FELICS_ERR_CORRUPT;   // the sidecar does not belong to this file
        if (st != FELICS_OK) atomicCAS(a.status, FELICS_OK, st);
    }
}

int decode_sidecar(felics_ctx *ctx, const uint8_t *h_fel, size_t len, const uint8_t *h_side, size_t side_len, void *h_pixels_out, size_t cap,
                   felics_header *hdr_out) {
    ctx->last.valid = false;
    felics_header hdr;
    int rc = felics_read_header(h_fel, len, &hdr);
    if (rc) return rc;
    if (hdr_out) *hdr_out = hdr;
    if (hdr.pixel_depth != 0) { set_error("sidecar decode is built for 8-bit samples"); return FELICS_ERR_UNSUPPORTED; }
    const uint64_t npix64 = (uint64_t)hdr.width * hdr.height;
    if (npix64 > 0x7fff0000ull) return FELICS_ERR_INVALID_DIMENSIONS;
    const uint32_t npix = (uint32_t)npix64, nch = hdr.color_type ? 3 : 1;
    const size_t need = (size_t)npix * nch;
    if (need > cap) { set_error("pixel buffer too small: need %zu", need); return FELICS_ERR_BUFFER_TOO_SMALL; }
    uint32_t sh[8];
    if (side_len < SIDECAR_HEADER_BYTES) { set_error("sidecar too short"); return FELICS_ERR_CORRUPT; }
    memcpy(sh, h_side, sizeof(sh));
    const size_t entry = sidecar_entry_bytes(hdr.width);
    const uint32_t band_rows = sh[5], nbands = sh[6];
    if (sh[0] != SIDECAR_MAGIC || sh[1] != SIDECAR_VERSION || sh[2] != hdr.width || sh[3] != hdr.height || sh[4] != nch || band_rows < 2 ||
        band_rows % sidecar_row_unit(hdr.width ? hdr.width : 1) != 0 || nbands != std::max<uint32_t>(1, (hdr.height + band_rows - 1) / band_rows) ||
        sh[7] != (uint32_t)entry || side_len != SIDECAR_HEADER_BYTES + (size_t)nch * nbands * entry) {
        set_error("sidecar header does not match the file");
        return FELICS_ERR_CORRUPT;
    }
    if (npix < 3 || hdr.width > DEC_MAX_W) {   // nothing to gain: the plain decoder
        uint64_t off[2] = {0, (uint64_t)len};
        int status = 0;
        return felics_decompress_batch(ctx, 1, h_fel, off, &hdr, h_pixels_out, &status);
    }
    cudaStream_t st = ctx->stream;
    const size_t fel_al = align_up(len + 8, 256), ent_bytes = side_len - SIDECAR_HEADER_BYTES;
    if ((rc = ensure_buffer(ctx, &ctx->staging_in, &ctx->staging_in_cap, fel_al + ent_bytes + 16))) return rc;
    if ((rc = ensure_buffer(ctx, &ctx->staging_out, &ctx->staging_out_cap, need + 16))) return rc;
    const size_t pstride = plane_stride8(npix);
    const size_t plane_bytes = align_up(((size_t)nch * pstride + 8) * sizeof(int16_t), 256);
    if ((rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, 256 + plane_bytes))) return rc;
    uint8_t *d_fel = (uint8_t *)ctx->staging_in, *d_ent = d_fel + fel_al;
    int *d_status = (int *)ctx->scratch;
    int16_t *d_planes = (int16_t *)((uint8_t *)ctx->scratch + 256);
    FELICS_CUDA_TRY(cudaMemcpyAsync(d_fel, h_fel, len, cudaMemcpyHostToDevice, st));
    FELICS_CUDA_TRY(cudaMemcpyAsync(d_ent, h_side + SIDECAR_HEADER_BYTES, ent_bytes, cudaMemcpyHostToDevice, st));
    FELICS_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    BandArgs a;
    a.words = (const uint32_t *)d_fel; a.nwords = (len + 3) / 4; a.fel_bytes = len;
    a.entries = d_ent; a.entry_bytes = entry; a.band_rows = band_rows; a.nbands = nbands;
    a.planes = d_planes; a.pstride = pstride; a.status = d_status; a.w = hdr.width; a.h = hdr.height; a.npix = npix; a.nch = nch;
    {
        StageScope s(ctx, ST_DECODE);
        const size_t smem = (size_t)(NBIN - 1) * NK * 4 + 8 + 2 * (size_t)hdr.width * sizeof(int16_t);
        if (smem > 48 * 1024 && smem > ctx->decode_bands_smem_set) {
            FELICS_CUDA_TRY(cudaFuncSetAttribute(k_decode_bands, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->decode_bands_smem_set = smem;
        }
        k_decode_bands<<<nch * nbands, 32, smem, st>>>(a);
        s.launched();
    }
    {
        StageScope s(ctx, ST_UNPLANE);
        const size_t total = npix;
        const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
        if (nch == 1) k_unplane_gray8<<<blocks, 256, 0, st>>>(d_planes, (uint8_t *)ctx->staging_out, npix, pstride, total, d_status);
        else k_unplane_rgb8<<<blocks, 256, 0, st>>>(d_planes, (uint8_t *)ctx->staging_out, npix, pstride, total, d_status);
        s.launched();
    }
    int status = 0;
    FELICS_CUDA_TRY(cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    FELICS_CUDA_TRY(cudaMemcpyAsync(h_pixels_out, ctx->staging_out, need, cudaMemcpyDeviceToHost, st));
    FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    FELICS_CUDA_TRY(cudaGetLastError());
    if ((rc = profile_collect(ctx))) return rc;
    if (status == FELICS_NEED_EXACT) {   // a band left the 16-bit planes: the plain decoder settles what the reference would report
        uint64_t off[2] = {0, (uint64_t)len};
        int st1 = 0;
        return felics_decompress_batch(ctx, 1, h_fel, off, &hdr, h_pixels_out, &st1);
    }
    return status;
}

// offsets[i] <= offsets[i + 1] for every image: the kernels take lengths as differences and read headers at offsets[i]
int check_offsets(size_t n, const uint64_t *offsets) {
    for (size_t i = 0; i < n; i++)
        if (offsets[i] > offsets[i + 1]) {
            set_error("offsets must not decrease (offsets[%zu] = %llu > offsets[%zu] = %llu)", i, (unsigned long long)offsets[i], i + 1, (unsigned long long)offsets[i + 1]);
            return FELICS_ERR_INVALID_ARGUMENT;
        }
    return FELICS_OK;
}

int decode_batch_device(felics_ctx *ctx, size_t n, const uint8_t *d_arena, const uint64_t *offsets_host,
                        const felics_header &hdr, void *d_pixels_out, int *status_host) {
    ctx->last.valid = false;
    if (((uintptr_t)d_arena & 3) != 0) {
        set_error("device arena must be 4-byte aligned");
        return FELICS_ERR_INVALID_ARGUMENT;
    }
    uint64_t npix64 = (uint64_t)hdr.width * hdr.height;
    if (npix64 > 0xffffffffull) return FELICS_ERR_INVALID_DIMENSIONS;   // checked_mul, compression.rs:176-180
    if (npix64 > 0x7fff0000ull) { set_error("image too large for one call"); return FELICS_ERR_INVALID_DIMENSIONS; }
    if (int orc = check_offsets(n, offsets_host)) return orc;
    if (hdr.pixel_depth != 0) return decode16_batch_device(ctx, n, d_arena, offsets_host, hdr, d_pixels_out, status_host);
    cudaStream_t st = ctx->stream;
    const uint32_t npix = (uint32_t)npix64;
    const uint32_t nch = hdr.color_type ? 3 : 1;

    size_t off_bytes = align_up((n + 1) * sizeof(uint64_t), 256);
    size_t stat_bytes = align_up(n * sizeof(int), 256);
    // batches of gray files whose rows are whole words: several files per warp, pixels written directly (k_decode_g8)
    const bool g8 = nch == 1 && hdr.width >= 4 && hdr.width % 4 == 0 && hdr.width <= G8_MAX_W && hdr.height >= 1 && !ctx->no_g8;
    const size_t pstride = plane_stride8(npix);
    size_t plane_bytes = g8 ? 0 : align_up(((size_t)n * nch * pstride + 8) * sizeof(int16_t), 256);
    int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, off_bytes + stat_bytes + plane_bytes);
    if (rc) return rc;
    uint8_t *sb = (uint8_t *)ctx->scratch;
    uint64_t *d_off = (uint64_t *)sb;
    int *d_status = (int *)(sb + off_bytes);
    int16_t *d_planes = (int16_t *)(sb + off_bytes + stat_bytes);
    FELICS_CUDA_TRY(cudaMemcpyAsync(d_off, offsets_host, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));

    DecArgs a;
    a.words = (const uint32_t *)d_arena;
    a.arena_words = (offsets_host[n] + 3) / 4;   // the word holding the last byte must be readable
    a.offsets = d_off; a.planes = d_planes; a.status = d_status;
    a.w = hdr.width; a.h = hdr.height; a.npix = npix; a.nch = nch; a.pstride = pstride;
    a.color = hdr.color_type; a.depth = hdr.pixel_depth;
    if (g8) {
        StageScope s(ctx, ST_DECODE);
        G8Args g;
        g.words = a.words; g.arena_words = a.arena_words; g.offsets = d_off; g.pixels = (uint8_t *)d_pixels_out; g.status = d_status;
        g.w = hdr.width; g.h = hdr.height; g.npix = npix; g.word_out = ((uintptr_t)d_pixels_out & 3) == 0 ? 1u : 0u;
        // files per warp: as many as still leave every SM half a dozen warps of work
        int F = 1;
        const size_t warp_budget = 100 * 1024;   // shared memory of one warp (two warps per block)
        while (F < 32 && n / (size_t)(2 * F) >= (size_t)148 * 6 && g8_warp_bytes(2 * F, G8_HOT, hdr.width) <= warp_budget) F *= 2;
        if (g8_warp_bytes(1, G8_CTX, hdr.width) > warp_budget) F = std::max(F, 2);   // (rows of 32,768 samples: the hot / cold split even for one file)
        if (ctx->g8_files_per_warp) {
            F = ctx->g8_files_per_warp;
            while (F > 1 && g8_warp_bytes(F, G8_HOT, hdr.width) > warp_budget) F /= 2;
        }
#define G8_LAUNCH(FF) launch_decode_g8<FF, G8_HOT>(ctx, g, n, st)
        switch (F) {
            case 1: rc = launch_decode_g8<1, G8_CTX>(ctx, g, n, st); break;   // a warp per file: room for every estimator row
            case 2: rc = G8_LAUNCH(2); break;
            case 4: rc = G8_LAUNCH(4); break;
            case 8: rc = G8_LAUNCH(8); break;
            case 16: rc = G8_LAUNCH(16); break;
            default: rc = G8_LAUNCH(32); break;
        }
#undef G8_LAUNCH
        if (rc) return rc;
        s.launched();
    } else {
        StageScope s(ctx, ST_DECODE);
        const size_t smem = (size_t)(NBIN - 1) * NK * 4 + 8 + 2 * (size_t)hdr.width * sizeof(int16_t);
        if (hdr.width <= DEC_MAX_W && npix >= 1) {
            if (smem > 48 * 1024 && smem > ctx->decode_smem_set) {
                FELICS_CUDA_TRY(cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ctx->decode_smem_set = smem;
            }
            k_decode<<<(unsigned)n, 32, smem, st>>>(a, (uint32_t)n);
        } else {
            k_decode_wide<<<(unsigned)n, 32, 0, st>>>(a, (uint32_t)n);   // very wide rows (or empty images): neighbours from global memory
        }
        s.launched();
    }
    if (npix > 0 && !g8) {
        StageScope s(ctx, ST_UNPLANE);
        size_t total = n * (size_t)npix;
        unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
        if (nch == 1) k_unplane_gray8<<<blocks, 256, 0, st>>>(d_planes, (uint8_t *)d_pixels_out, npix, pstride, total, d_status);
        else k_unplane_rgb8<<<blocks, 256, 0, st>>>(d_planes, (uint8_t *)d_pixels_out, npix, pstride, total, d_status);
        s.launched();
    }
    FELICS_CUDA_TRY(cudaMemcpyAsync(status_host, d_status, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    FELICS_CUDA_TRY(cudaGetLastError());
    rc = profile_collect(ctx);
    if (rc) return rc;
    // files the fast decoder gave up on: again, the reference's way (i32 samples), one at a time
    for (size_t i = 0; i < n; i++) {
        if (status_host[i] != FELICS_NEED_EXACT) continue;
        const size_t pl_bytes = align_up((size_t)nch * npix * sizeof(int32_t) + 16, 256);
        if ((rc = ensure_buffer(ctx, &ctx->exact_buf, &ctx->exact_cap, pl_bytes + 256))) return rc;
        ExactArgs ea;
        ea.words = (const uint32_t *)d_arena; ea.arena_words = a.arena_words; ea.off0 = offsets_host[i]; ea.off1 = offsets_host[i + 1];
        ea.planes = (int32_t *)ctx->exact_buf; ea.status = (int *)((uint8_t *)ctx->exact_buf + pl_bytes);
        ea.w = hdr.width; ea.h = hdr.height; ea.npix = npix; ea.nch = nch;
        k_decode_exact<<<1, 32, 0, st>>>(ea);
        if (npix > 0) k_unplane_exact8<<<(unsigned)std::min<size_t>(((size_t)npix + 255) / 256, 148 * 8), 256, 0, st>>>(
            ea.planes, (uint8_t *)d_pixels_out + i * (size_t)npix * nch, npix, nch, ea.status);
        FELICS_CUDA_TRY(cudaMemcpyAsync(&status_host[i], ea.status, sizeof(int), cudaMemcpyDeviceToHost, st));
        FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    }
    for (size_t i = 0; i < n; i++)
        if (status_host[i] != FELICS_OK) return status_host[i];
    return FELICS_OK;
}

}  // namespace felics
