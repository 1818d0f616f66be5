// 16-bit samples (traits.rs:35-43): K = {0..14}, MAX_CONTEXT = 131070, count scaling at 1024.
// The serial 16-bit path: the reference loops run as they are written (compression.rs:76-148 / :151-248), one warp per
// image with lane 0 walking the raster, the 131071 x 15 estimator table (parameter_selection.rs:29-33) split between
// shared and global memory.  This is the DECODER (a file is one serial bit chain) and the cross-check encoder behind
// FELICS_B200_SERIAL16; the product encoder is the parallel pipeline of enc16_par.cuh (DESIGN.md "16-bit").
#pragma once
#include <stdint.h>

namespace felics {

constexpr int NK16 = 15;                  // traits.rs:37
constexpr uint32_t MAXCTX16 = 131070;     // traits.rs:39 (2 * 65535)
constexpr int ROW16 = 16;                 // table row padded to 16 words (64 bytes)
constexpr size_t TABLE16_WORDS = (size_t)(MAXCTX16 + 1) * ROW16;

// Estimator rows: contexts below SM_ROWS16 (nearly all of a natural image's) live in shared memory, the rest in a
// per-image global table whose rows are validated lazily by a tag in their 16th word (tables persist between calls and
// are zeroed once when allocated; every (call, pass, channel) uses a fresh tag), so nothing is cleared per channel but
// the shared part.
constexpr uint32_t SM_ROWS16 = 2048;
constexpr size_t SM_TABLE16_BYTES = (size_t)SM_ROWS16 * ROW16 * sizeof(uint32_t);

struct Est16 {
    uint32_t *sm;      // [SM_ROWS16][16]
    uint32_t *gl;      // [MAXCTX16 + 1][16]
    uint32_t tag;

    __device__ __forceinline__ void load(uint32_t ctx, uint32_t v[NK16]) const {
        const uint32_t *row = ctx < SM_ROWS16 ? sm + (size_t)ctx * ROW16 : gl + (size_t)ctx * ROW16;
        const uint4 a = *reinterpret_cast<const uint4 *>(row), b = *reinterpret_cast<const uint4 *>(row + 4),
                    c = *reinterpret_cast<const uint4 *>(row + 8), d = *reinterpret_cast<const uint4 *>(row + 12);
        const bool fresh = ctx < SM_ROWS16 || d.w == tag;
        const uint32_t t[NK16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z};
#pragma unroll
        for (int k = 0; k < NK16; k++) v[k] = fresh ? t[k] : 0u;
    }
    __device__ __forceinline__ void store(uint32_t ctx, const uint32_t v[NK16]) const {
        uint32_t *row = ctx < SM_ROWS16 ? sm + (size_t)ctx * ROW16 : gl + (size_t)ctx * ROW16;
        reinterpret_cast<uint4 *>(row)[0] = make_uint4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<uint4 *>(row)[1] = make_uint4(v[4], v[5], v[6], v[7]);
        reinterpret_cast<uint4 *>(row)[2] = make_uint4(v[8], v[9], v[10], v[11]);
        reinterpret_cast<uint4 *>(row)[3] = make_uint4(v[12], v[13], v[14], tag);
    }
    // fresh estimator for the next channel (compression.rs:110-114): clear the shared rows, move to the next tag
    __device__ __forceinline__ void reset(uint32_t lane, uint32_t new_tag) {
        uint4 *t4 = reinterpret_cast<uint4 *>(sm);
        for (uint32_t j = lane; j < SM_ROWS16 * ROW16 / 4; j += 32) t4[j] = make_uint4(0u, 0u, 0u, 0u);
        tag = new_tag;
        __syncwarp();
    }
};
// get_k: `<=` scan, ties to the largest k (parameter_selection.rs:78-83)
__device__ __forceinline__ int get_k16(const uint32_t v[NK16]) {
    uint32_t best = v[0];
    int bi = 0;
#pragma unroll
    for (int k = 1; k < NK16; k++)
        if (v[k] <= best) { best = v[k]; bi = k; }
    return bi;
}
// update (parameter_selection.rs:49-64), in registers
__device__ __forceinline__ void update16(uint32_t v[NK16], uint32_t e) {
    uint32_t mn = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < NK16; k++) {
        v[k] += (e >> k) + 1u + (uint32_t)k;
        mn = min(mn, v[k]);
    }
    const int sh = mn > HALVE_AT ? 1 : 0;
#pragma unroll
    for (int k = 0; k < NK16; k++) v[k] >>= sh;
}
// distance between consecutive planes, in samples: never a power of two (2,336 extra samples = 9,344 bytes)
inline size_t plane_stride16(uint32_t npix) { return (size_t)npix + 2336; }

// neighbour indices of raster index i >= 2 (misc.rs:6-24)
__device__ __forceinline__ void neighbours16(uint32_t i, uint32_t x, uint32_t y, uint32_t w, uint32_t &ia, uint32_t &ib) {
    if (x > 0 && y > 0) { ia = i - 1; ib = i - w; }
    else if (y == 0) { ia = i - 1; ib = i - 2; }
    else if (y >= 2) { ia = i - w; ib = i - 2 * w; }
    else { ia = i - w; ib = i - w + 1; }
}

}  // namespace felics
