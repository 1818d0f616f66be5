// Device helpers shared by the encode and decode kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace felics {

constexpr int TILE = 4096;          // pixels per tile (one thread block)
constexpr int TILE_THREADS = 256;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int WARP_PIX = TILE / TILE_WARPS;   // 512 consecutive pixels per warp
constexpr int WARP_ITERS = WARP_PIX / 32;     // 16 rounds of 32 consecutive pixels
constexpr int NBIN = 512;           // contexts 0..510 (traits.rs:29: MAX_CONTEXT = 2*255) padded to 512
constexpr int CHUNK_TILES = 64;     // tiles per histogram chunk
constexpr int GROUP = 1024;         // grouped elements per prefix group (32 blocks of 32)
constexpr uint32_t HALVE_AT = 1024; // traits.rs:31 COUNT_SCALING = Some(1024)
constexpr int NK = 6;               // traits.rs:27 K_VALUES = [0..5]
constexpr uint32_t PAD_E = 0xFFFFu; // marks padding slots of the grouped residual array

// Neighbour geometry + classification of pixel i of a row-major plane
// (misc.rs:6-24, compression.rs:118-145).  cls: 0 in-range, 1 above, 2 below.
struct PixelClass {
    int delta;  // context H - L
    int cls;
    int val;    // P-L (in-range) | P-H-1 (above) | L-P-1 (below)
    int lo;
};

// Raster cursor over one plane: position (i, x, y) and a pointer to the sample, advanced by a fixed step.
// T = uint8_t (gray samples straight from the caller's pixels) or int16_t (Y/Co/Cg planes).
template <typename T>
struct RasterCursor {
    const T *pp;       // &plane[i]
    uint32_t i, x, y, w;
    __device__ __forceinline__ void init(const T *plane, uint32_t i0, uint32_t width) {
        w = width; i = i0; y = i0 / width; x = i0 - y * width; pp = plane + i0;
    }
    __device__ __forceinline__ void step(uint32_t s) {
        i += s; pp += s;
        if (w >= s) {                      // at most one row boundary per step
            x += s;
            if (x >= w) { x -= w; y++; }
        } else {                           // narrow images: one division
            y = i / w; x = i - y * w;
        }
    }
    // the same step with its quotient and remainder by the width worked out once (q = s / w, r = s % w): no division per step
    __device__ __forceinline__ void step_qr(uint32_t s, uint32_t q, uint32_t r) {
        i += s; pp += s;
        y += q; x += r;
        if (x >= w) { x -= w; y++; }
    }
    // neighbours and class of the sample under the cursor (misc.rs:6-24, compression.rs:118-145); needs i >= 2
    __device__ __forceinline__ PixelClass classify() const {
        const int p = pp[0];
        int v1, v2;
        if (x > 0 && y > 0) {              // interior: left, up
            v1 = pp[-1];
            v2 = *(pp - w);
        } else if (y == 0) {               // first row (x >= 2 because i >= 2): i-1, i-2
            v1 = pp[-1]; v2 = pp[-2];
        } else if (y >= 2) {               // first column: up, up-up
            v1 = *(pp - w); v2 = *(pp - 2 * (size_t)w);
        } else {                           // pixel (0,1): up, up-right
            v1 = *(pp - w); v2 = *(pp - w + 1);
        }
        const int h = max(v1, v2), l = min(v1, v2);
        PixelClass r;
        r.delta = h - l;
        r.lo = l;
        const bool below = p < l, above = p > h;
        r.cls = below ? 2 : (above ? 1 : 0);
        r.val = below ? l - p - 1 : (above ? p - h - 1 : p - l);
        return r;
    }
};

// Four consecutive samples at a 4-sample-aligned index: one 32-bit (u8) or 64-bit (i16) load
__device__ __forceinline__ void load4(const uint8_t *p, int (&v)[4]) {
    const uint32_t w = *reinterpret_cast<const uint32_t *>(p);
    v[0] = (int)(w & 255u); v[1] = (int)((w >> 8) & 255u); v[2] = (int)((w >> 16) & 255u); v[3] = (int)(w >> 24);
}
__device__ __forceinline__ void load4(const int16_t *p, int (&v)[4]) {
    const uint2 w = *reinterpret_cast<const uint2 *>(p);
    v[0] = (int)(int16_t)(w.x & 0xffffu); v[1] = (int)(int16_t)(w.x >> 16); v[2] = (int)(int16_t)(w.y & 0xffffu); v[3] = (int)(int16_t)(w.y >> 16);
}
__device__ __forceinline__ void load4(const int32_t *p, int (&v)[4]) {
    const uint4 w = *reinterpret_cast<const uint4 *>(p);
    v[0] = (int)w.x; v[1] = (int)w.y; v[2] = (int)w.z; v[3] = (int)w.w;
}
__device__ __forceinline__ PixelClass classify_from(int p, int v1, int v2) {   // compression.rs:118-145 given the two neighbours
    const int h = max(v1, v2), l = min(v1, v2);
    PixelClass r;
    r.delta = h - l;
    r.lo = l;
    const bool below = p < l, above = p > h;
    r.cls = below ? 2 : (above ? 1 : 0);
    r.val = below ? l - p - 1 : (above ? p - h - 1 : p - l);
    return r;
}
// Classes of the four samples i .. i+3 of a plane whose width is a multiple of four (i, x multiples of four; the group
// never straddles a row).  Interior groups (x >= 4, y >= 1: left neighbour and the row above) take three loads for four
// samples; the first row and the first group of a row go through the per-sample rules of RasterCursor::classify.
// valid[j] is false for samples 0 and 1 of the plane (sent raw, compression.rs:93-108).
template <typename T>
__device__ __forceinline__ void classify4(const T *plane, uint32_t i, uint32_t x, uint32_t y, uint32_t w, PixelClass (&pc)[4], bool (&valid)[4]) {
    if (x >= 4 && y >= 1) {
        int cur[4], up[4];
        load4(plane + i, cur);
        load4(plane + i - w, up);
        const int left = plane[i - 1];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            pc[j] = classify_from(cur[j], j ? cur[j ? j - 1 : 0] : left, up[j]);
            valid[j] = true;
        }
    } else {
        RasterCursor<T> c;
        c.pp = plane + i; c.i = i; c.x = x; c.y = y; c.w = w;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            valid[j] = c.i >= 2;
            if (valid[j]) pc[j] = c.classify();
            else { pc[j].delta = 0; pc[j].cls = 0; pc[j].val = 0; pc[j].lo = 0; }
            c.i++; c.pp++; c.x++;      // stays inside the row: x + 3 < w
        }
    }
}

// Phased-in code of v in [0, n-1] (phase_in_coding.rs:23-84): returns the code value, sets len.
// long codeword = (x - right_p)/2 + right_p in m bits followed by (x - right_p)&1  ==  x + right_p in m+1 bits.
__device__ __forceinline__ uint32_t phase_in_code(uint32_t n, uint32_t v, int &len) {
    int m = 31 - __clz(n);
    uint32_t left_p = n - (1u << m);
    uint32_t right_p = (2u << m) - n;
    uint32_t x = v + n - left_p;
    if (x >= n) x -= n;
    if (x < right_p) { len = m; return x; }
    len = m + 1;
    return x + right_p;
}

// k of a counter row: `<=` scan, ties go to the largest index (parameter_selection.rs:78-83).
__device__ __forceinline__ int argmin_last(const uint32_t v[NK]) {
    uint32_t best = v[0];
    int bi = 0;
#pragma unroll
    for (int k = 1; k < NK; k++)
        if (v[k] <= best) { best = v[k]; bi = k; }
    return bi;
}

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// OR `n` (1..32) bits of val, MSB first, at absolute bit position `bit` of a word
// buffer whose words hold the stream big-endian-in-register (bit 31 = first bit).
__device__ __forceinline__ void put_bits_smem(uint32_t *buf, uint32_t bit, uint32_t val, int n) {
    uint32_t wi = bit >> 5, sh = bit & 31;
    uint32_t left = val << (32 - n);
    atomicOr(&buf[wi], left >> sh);
    if (sh + n > 32) atomicOr(&buf[wi + 1], left << (32 - sh));
}
// Same on the global arena, whose memory is the byte stream itself (so byte-swap).
__device__ __forceinline__ void put_bits_global(uint32_t *arena, uint64_t bit, uint32_t val, int n) {
    uint64_t wi = bit >> 5;
    uint32_t sh = (uint32_t)(bit & 31);
    uint32_t left = val << (32 - n);
    uint32_t hi = left >> sh;
    if (hi) atomicOr(&arena[wi], bswap32(hi));
    if (sh + n > 32) {
        uint32_t lo = left << (32 - sh);
        if (lo) atomicOr(&arena[wi + 1], bswap32(lo));
    }
}

// Per-pixel code record (u32): bits 31..22 = length (0..1023), bits 21..0 = payload.
//   length <= REC_SHORT_MAX: payload is the whole code word, right aligned.
//   otherwise (out-of-range pixel with a long unary run): payload =
//     above(1) << 17 | k(3) << 14 | rem(5) << 9 | q(9)
constexpr int REC_SHORT_MAX = 22;
__device__ __forceinline__ uint32_t rec_len(uint32_t rec) { return rec >> 22; }

}  // namespace felics
