// 16-bit samples (traits.rs:35-43): K = {0..14}, MAX_CONTEXT = 131070, count scaling at 1024.
// First correct device path for this pixel depth: the reference loops run as they are written
// (compression.rs:76-148 / :151-248), one warp per image with lane 0 walking the raster, the
// 131071 x 15 estimator table (parameter_selection.rs:29-33) in global memory.  Images of a batch
// run in parallel; inside an image nothing is parallel yet (DESIGN.md "16-bit").
#pragma once
#include <stdint.h>

namespace felics {

constexpr int NK16 = 15;                  // traits.rs:37
constexpr uint32_t MAXCTX16 = 131070;     // traits.rs:39 (2 * 65535)
constexpr int ROW16 = 16;                 // table row padded to 16 words (64 bytes)
constexpr size_t TABLE16_WORDS = (size_t)(MAXCTX16 + 1) * ROW16;

// get_k: `<=` scan, ties to the largest k (parameter_selection.rs:78-83)
__device__ __forceinline__ int get_k16(const uint32_t *row) {
    const uint4 a = *reinterpret_cast<const uint4 *>(row), b = *reinterpret_cast<const uint4 *>(row + 4),
                c = *reinterpret_cast<const uint4 *>(row + 8), d = *reinterpret_cast<const uint4 *>(row + 12);
    const uint32_t v[NK16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z};
    uint32_t best = v[0];
    int bi = 0;
#pragma unroll
    for (int k = 1; k < NK16; k++)
        if (v[k] <= best) { best = v[k]; bi = k; }
    return bi;
}
// update (parameter_selection.rs:49-64)
__device__ __forceinline__ void update16(uint32_t *row, uint32_t e) {
    uint32_t v[NK16], mn = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < NK16; k++) {
        v[k] = row[k] + (e >> k) + 1u + (uint32_t)k;
        mn = min(mn, v[k]);
    }
    const int sh = mn > HALVE_AT ? 1 : 0;
#pragma unroll
    for (int k = 0; k < NK16; k++) row[k] = v[k] >> sh;
}
// zero one image's table with the whole warp
__device__ __forceinline__ void clear_table16(uint32_t *tab, uint32_t lane) {
    uint4 *t4 = reinterpret_cast<uint4 *>(tab);
    for (size_t j = lane; j < TABLE16_WORDS / 4; j += 32) t4[j] = make_uint4(0u, 0u, 0u, 0u);
    __threadfence_block();
    __syncwarp();
}
// neighbour indices of raster index i >= 2 (misc.rs:6-24)
__device__ __forceinline__ void neighbours16(uint32_t i, uint32_t x, uint32_t y, uint32_t w, uint32_t &ia, uint32_t &ib) {
    if (x > 0 && y > 0) { ia = i - 1; ib = i - w; }
    else if (y == 0) { ia = i - 1; ib = i - 2; }
    else if (y >= 2) { ia = i - w; ib = i - 2 * w; }
    else { ia = i - w; ib = i - w + 1; }
}

}  // namespace felics
