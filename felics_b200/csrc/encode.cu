// FELICS encode pipeline for sm_100a (8-bit samples; 16-bit samples: encode16.cu).
//
// Replaces the reference's sequential compress_channel loop
// (/root/reference/src/compression.rs:76-148) by data-parallel passes whose
// output is byte-identical:
//
//   planes    RGB only: pixels -> Y/Co/Cg i16 planes (YCoCg-R, color_transform.rs:11-17); gray is read in place
//   hist      per-tile histogram of out-of-range pixels by context
//   chainscan / tilebase   counting-sort offsets: one "chain" per (plane, context)
//   scatter   stable grouping of the residuals e by context (raster order kept)
//   prefix    code cost of e under each k in {0..5} (rice_coding.rs:56-58), prefix-summed per 32-block / group / all
//   spec      KEstimator with count halving (parameter_selection.rs:49-85) for long stationary chains: speculative
//             parallel epoch walk, exact by verification (sp_walk.cuh), plus segment-hop tables (hop_walk.cuh)
//   walk      the same recurrence walked epoch by epoch, one warp per chain (everything the speculation leaves),
//             crossing flip-free segments of rejected chains with one table lookup; runs beside `spec` on a second stream
//   kfill     k = argmin (ties to the largest k) for every out-of-range pixel, in parallel
//   code      marker + phased-in / Rice code word and length per pixel
//   bitscan   exclusive scans: tile -> plane -> image bit offsets (exact output size)
//   pack      MSB-first bit packing, identical to bitstream-io's BigEndian BitWriter
#include "ctx.h"
#include "device_common.cuh"
#include "serial16.cuh"

#include <algorithm>
#include <cstring>

#include "sp_walk.cuh"
#include "hop_walk.cuh"

namespace felics {

// ------------------------------------------------------------------------------------
// planes
// ------------------------------------------------------------------------------------
// color_transform.rs:11-17; C++ int division truncates toward zero like Rust's.
__global__ void k_to_planes_rgb8(const uint8_t *__restrict__ px, int16_t *__restrict__ planes, uint32_t npix, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        size_t img = idx / npix;
        uint32_t i = (uint32_t)(idx - img * npix);
        int r = px[3 * idx], g = px[3 * idx + 1], b = px[3 * idx + 2];
        int co = r - b;
        int t = b + co / 2;
        int cg = g - t;
        int y = t + cg / 2;
        int16_t *base = planes + img * 3 * (size_t)npix;
        base[i] = (int16_t)y;
        base[(size_t)npix + i] = (int16_t)co;
        base[2 * (size_t)npix + i] = (int16_t)cg;
    }
}

// ------------------------------------------------------------------------------------
// hist: one block per tile
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS) k_hist(const T *__restrict__ planes, uint32_t w, uint32_t npix,
                                                       uint32_t tpp, uint32_t nchunks, uint32_t *__restrict__ tile_hist,
                                                       uint32_t *__restrict__ chunk_tot) {
    __shared__ uint32_t h[NBIN];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / tpp, t = bid - p * tpp;
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) h[c] = 0;
    __syncthreads();
    const T *pl = planes + (size_t)p * npix;
    uint32_t start = t * TILE;
    RasterCursor<T> cur;
    cur.init(pl, start + threadIdx.x, w);
#pragma unroll 4
    for (int j = 0; j < TILE / TILE_THREADS; j++) {
        if (cur.i >= 2 && cur.i < npix) {
            PixelClass pc = cur.classify();
            if (pc.cls != 0) atomicAdd(&h[sizeof(T) == 4 ? (pc.delta & (NBIN - 1)) : pc.delta], 1u);   // 16-bit samples: buckets (enc16_par.cuh)
        }
        cur.step(TILE_THREADS);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) {
        uint32_t v = h[c];
        tile_hist[(size_t)bid * NBIN + c] = v;
        if (v) atomicAdd(&chunk_tot[((size_t)p * nchunks + t / CHUNK_TILES) * NBIN + c], v);
    }
}

// The same for planes whose width is a multiple of four: a thread classifies four consecutive samples from three loads.
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS) k_hist4(const T *__restrict__ planes, uint32_t w, uint32_t npix, uint32_t tpp, uint32_t nchunks,
                                                        uint32_t *__restrict__ tile_hist, uint32_t *__restrict__ chunk_tot) {
    __shared__ uint32_t h[NBIN];
    const uint32_t bid = blockIdx.x;
    const uint32_t p = bid / tpp, t = bid - p * tpp;
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) h[c] = 0;
    __syncthreads();
    const T *pl = planes + (size_t)p * npix;
    RasterCursor<T> cur;
    cur.init(pl, t * TILE + 4u * threadIdx.x, w);
    const uint32_t step_q = (4u * TILE_THREADS) / w, step_r = (4u * TILE_THREADS) - step_q * w;
#pragma unroll
    for (int j = 0; j < TILE / (4 * TILE_THREADS); j++) {
        if (cur.i < npix) {
            PixelClass pc[4];
            bool valid[4];
            classify4(pl, cur.i, cur.x, cur.y, w, pc, valid);
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (valid[q] && pc[q].cls != 0) atomicAdd(&h[sizeof(T) == 4 ? (pc[q].delta & (NBIN - 1)) : pc[q].delta], 1u);
        }
        cur.step_qr(4 * TILE_THREADS, step_q, step_r);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) {
        const uint32_t v = h[c];
        tile_hist[(size_t)bid * NBIN + c] = v;
        if (v) atomicAdd(&chunk_tot[((size_t)p * nchunks + t / CHUNK_TILES) * NBIN + c], v);
    }
}

// ------------------------------------------------------------------------------------
// chainscan: one block (512 threads = one per context) per plane
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NBIN) k_chainscan(uint32_t *__restrict__ chunk_tot, uint32_t nchunks,
                                                    uint32_t *__restrict__ chain_count, uint32_t *__restrict__ chain_base,
                                                    uint32_t *__restrict__ plane_used, uint32_t *__restrict__ live,
                                                    uint32_t *__restrict__ counters /* [0]=live count */) {
    __shared__ uint32_t wsum[NBIN / 32];
    uint32_t p = blockIdx.x, c = threadIdx.x;
    uint32_t run = 0;
    for (uint32_t ch = 0; ch < nchunks; ch++) {
        size_t idx = ((size_t)p * nchunks + ch) * NBIN + c;
        uint32_t v = chunk_tot[idx];
        chunk_tot[idx] = run;  // exclusive prefix over chunks
        run += v;
    }
    uint32_t count = run;
    uint32_t aligned = (count + 31u) & ~31u;
    // block-wide exclusive scan of `aligned`
    uint32_t lane = c & 31, wid = c >> 5;
    uint32_t inc = aligned;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += n;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t v = lane < NBIN / 32 ? wsum[lane] : 0;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += n;
        }
        if (lane < NBIN / 32) wsum[lane] = s - v;
    }
    __syncthreads();
    uint32_t base = wsum[wid] + inc - aligned;
    chain_count[(size_t)p * NBIN + c] = count;
    chain_base[(size_t)p * NBIN + c] = base;
    if (c == NBIN - 1) plane_used[p] = base + aligned;
    if (count) {
        uint32_t slot = atomicAdd(&counters[0], 1u);
        live[slot] = p * NBIN + c;
    }
}

// tilebase: one block per (plane, chunk); turns tile_hist into scatter bases (plane-relative)
__global__ void __launch_bounds__(NBIN) k_tilebase(uint32_t *__restrict__ tile_hist, const uint32_t *__restrict__ chunk_tot,
                                                   const uint32_t *__restrict__ chain_base, uint32_t tpp, uint32_t nchunks) {
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / nchunks, ch = bid - p * nchunks, c = threadIdx.x;
    uint32_t run = chunk_tot[(size_t)bid * NBIN + c] + chain_base[(size_t)p * NBIN + c];
    uint32_t t0 = ch * CHUNK_TILES, t1 = min(t0 + CHUNK_TILES, tpp);
    uint32_t *col = tile_hist + ((size_t)p * tpp) * NBIN + c;
    uint32_t t = t0;
    for (; t + 8 <= t1; t += 8) {        // eight independent loads in flight
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = col[(size_t)(t + q) * NBIN];
#pragma unroll
        for (int q = 0; q < 8; q++) { col[(size_t)(t + q) * NBIN] = run; run += v[q]; }
    }
    for (; t < t1; t++) {
        uint32_t v = col[(size_t)t * NBIN];
        col[(size_t)t * NBIN] = run;
        run += v;
    }
}

// ------------------------------------------------------------------------------------
// scatter: stable grouping by context.  Warp w owns pixels [512w, 512w+512) of the tile
// and visits them 32 consecutive pixels at a time, so ranks follow raster order.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS, 4) k_scatter(const T *__restrict__ planes, uint32_t w, uint32_t npix,
                                                          uint32_t tpp, uint32_t cap, const uint32_t *__restrict__ tile_base,
                                                          uint16_t *__restrict__ e_grp, uint32_t *__restrict__ gidx) {
    __shared__ uint16_t wcnt[TILE_WARPS][NBIN];
    __shared__ uint32_t wbase[TILE_WARPS][NBIN];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / tpp, t = bid - p * tpp;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < TILE_WARPS * NBIN; c += TILE_THREADS) (&wcnt[0][0])[c] = 0;
    __syncthreads();
    const T *pl = planes + (size_t)p * npix;
    uint32_t wstart = t * TILE + wid * WARP_PIX;
    uint32_t info[WARP_ITERS];  // rank(13) | delta(9) << 13 | oor << 22
    uint16_t ev[WARP_ITERS];
    const uint32_t lt = (1u << lane) - 1u;
    RasterCursor<T> cur;
    cur.init(pl, wstart + lane, w);
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        const uint32_t i = cur.i;
        bool oor = false;
        int delta = 0, val = 0;
        if (i >= 2 && i < npix) {
            PixelClass pc = cur.classify();
            oor = pc.cls != 0;
            delta = pc.delta;
            val = pc.val;
        }
        uint32_t act = __ballot_sync(0xffffffffu, oor);
        uint32_t rank = 0, grpmask = 0, prev = 0;
        if (oor) {
            grpmask = __match_any_sync(act, delta);
            prev = wcnt[wid][delta];
            rank = prev + __popc(grpmask & lt);
        }
        __syncwarp();
        if (oor && (grpmask & lt) == 0) wcnt[wid][delta] = (uint16_t)(prev + __popc(grpmask));
        __syncwarp();
        info[it] = rank | ((uint32_t)delta << 13) | (oor ? (1u << 22) : 0u);
        ev[it] = (uint16_t)val;
        cur.step(32);
    }
    __syncthreads();
    // exclusive prefix over warps, per context, on top of the tile's base
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) {
        uint32_t run = tile_base[(size_t)bid * NBIN + c];
#pragma unroll
        for (int q = 0; q < TILE_WARPS; q++) {
            wbase[q][c] = run;
            run += wcnt[q][c];
        }
    }
    __syncthreads();
    uint16_t *eg = e_grp + (size_t)p * cap;
    uint32_t *gi = gidx + (size_t)p * npix;
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        if (info[it] >> 22) {
            uint32_t delta = (info[it] >> 13) & 511u;
            uint32_t g = wbase[wid][delta] + (info[it] & 8191u);
            eg[g] = ev[it];
            gi[wstart + it * 32 + lane] = g;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(TILE_THREADS, 4) k_scatter4(const T *__restrict__ planes, uint32_t w, uint32_t npix,
                                                          uint32_t tpp, uint32_t cap, const uint32_t *__restrict__ tile_base,
                                                          uint16_t *__restrict__ e_grp, uint32_t *__restrict__ gidx) {
    __shared__ uint16_t wcnt[TILE_WARPS][NBIN];
    __shared__ uint32_t wbase[TILE_WARPS][NBIN];
    __shared__ __align__(16) uint32_t xch[TILE_WARPS][128];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / tpp, t = bid - p * tpp;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < TILE_WARPS * NBIN; c += TILE_THREADS) (&wcnt[0][0])[c] = 0;
    __syncthreads();
    const T *pl = planes + (size_t)p * npix;
    uint32_t wstart = t * TILE + wid * WARP_PIX;
    uint32_t info[WARP_ITERS];  // rank(13) | delta(9) << 13 | oor << 22
    uint16_t ev[WARP_ITERS];
    const uint32_t lt = (1u << lane) - 1u;
    // classification four consecutive samples per lane (classify4), then a transpose through shared memory so that the
    // ranking below still sees 32 consecutive pixels per round, lane = pixel (ranks follow raster order)
    RasterCursor<T> cur;
    cur.init(pl, wstart + 4u * lane, w);
#pragma unroll
    for (int sup = 0; sup < WARP_ITERS / 4; sup++) {
        uint32_t word[4] = {0u, 0u, 0u, 0u};   // delta | val << 9 | out-of-range << 31
        if (cur.i < npix) {
            PixelClass pc[4];
            bool valid[4];
            classify4(pl, cur.i, cur.x, cur.y, w, pc, valid);
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (valid[q] && pc[q].cls != 0) word[q] = (uint32_t)pc[q].delta | ((uint32_t)pc[q].val << 9) | 0x80000000u;
        }
        *reinterpret_cast<uint4 *>(&xch[wid][4 * lane]) = make_uint4(word[0], word[1], word[2], word[3]);
        __syncwarp();
#pragma unroll
        for (int sl = 0; sl < 4; sl++) {
            const int it = sup * 4 + sl;
            const uint32_t wd = xch[wid][32 * sl + lane];
            const bool oor = wd >> 31;
            const uint32_t delta = wd & 511u, val = (wd >> 9) & 1023u;
            const uint32_t act = __ballot_sync(0xffffffffu, oor);
            uint32_t rank = 0, grpmask = 0, prev = 0;
            if (oor) {
                grpmask = __match_any_sync(act, delta);
                prev = wcnt[wid][delta];
                rank = prev + __popc(grpmask & lt);
            }
            __syncwarp();
            if (oor && (grpmask & lt) == 0) wcnt[wid][delta] = (uint16_t)(prev + __popc(grpmask));
            __syncwarp();
            info[it] = rank | (delta << 13) | (oor ? (1u << 22) : 0u);
            ev[it] = (uint16_t)val;
        }
        cur.step(128);
    }
    __syncthreads();
    // exclusive prefix over warps, per context, on top of the tile's base
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) {
        uint32_t run = tile_base[(size_t)bid * NBIN + c];
#pragma unroll
        for (int q = 0; q < TILE_WARPS; q++) {
            wbase[q][c] = run;
            run += wcnt[q][c];
        }
    }
    __syncthreads();
    uint16_t *eg = e_grp + (size_t)p * cap;
    uint32_t *gi = gidx + (size_t)p * npix;
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        if (info[it] >> 22) {
            uint32_t delta = (info[it] >> 13) & 511u;
            uint32_t g = wbase[wid][delta] + (info[it] & 8191u);
            eg[g] = ev[it];
            gi[wstart + it * 32 + lane] = g;
        }
    }
}

// ------------------------------------------------------------------------------------
// prefix: one block per group of 1024 grouped elements (32 blocks of 32).
// fine[g] = {U01, U23, U45, e}: inclusive prefix, within the 32-block, of the code cost
// (e >> k) + 1 + k for k = 0..5, two 16-bit lanes per word (32 * 516 < 65536).
// blk_rec[blk] (16 x u32): [0..5] cost prefix T at the block start, [8..13] T at the block
// end -- local to the group here, made global by k_blkfinal.
// grp_tot[grp][k]: totals of the group (scanned by k_grpscan1/2).
// All prefix sums are mod 2^32 over the whole grouped array; only differences inside a
// chain are ever used, so chain boundaries need no segmentation.
// ------------------------------------------------------------------------------------
constexpr int BLK_REC = 16;

// code costs (e >> k) + 1 + k of one residual for k = 0..5, two 16-bit lanes per word; padding slots cost nothing
__device__ __forceinline__ void pack_costs(uint32_t e, uint32_t &u01, uint32_t &u23, uint32_t &u45) {
    u01 = u23 = u45 = 0;
    if (e != PAD_E) {
        u01 = (e + 1u) | (((e >> 1) + 2u) << 16);
        u23 = ((e >> 2) + 3u) | (((e >> 3) + 4u) << 16);
        u45 = ((e >> 4) + 5u) | (((e >> 5) + 6u) << 16);
    }
}
// inclusive scan over the 32 lanes of a warp (= one 32-block of grouped elements)
__device__ __forceinline__ void warp_scan3(uint32_t &u01, uint32_t &u23, uint32_t &u45, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(0xffffffffu, u01, o);
        uint32_t b = __shfl_up_sync(0xffffffffu, u23, o);
        uint32_t c = __shfl_up_sync(0xffffffffu, u45, o);
        if (lane >= (uint32_t)o) { u01 += a; u23 += b; u45 += c; }
    }
}

// FINE = true: the per-element records are written (big single images: the speculative walk and the latency-bound
// serial walk read them); FINE = false: only block records, the consumers redo the 32-element scan themselves.
constexpr int PREFIX_THREADS = GROUP / 8;   // one thread per 8 consecutive elements, four threads per 32-block
template <bool FINE>
__global__ void __launch_bounds__(PREFIX_THREADS) k_prefix(const uint16_t *__restrict__ e_grp, uint32_t cap, uint32_t gpp,
                                                           const uint32_t *__restrict__ plane_used, uint4 *__restrict__ fine,
                                                           uint32_t *__restrict__ blk_rec, uint32_t *__restrict__ grp_tot) {
    __shared__ uint32_t tot[32][NK];
    uint32_t grp = blockIdx.x;
    uint32_t p = grp / gpp;
    uint32_t off = (grp - p * gpp) * GROUP;  // plane-relative element offset of this group
    if (off >= plane_used[p]) return;        // grp_tot stays 0 (memset)
    const size_t g = (size_t)p * cap + off + threadIdx.x * 8u;
    const uint32_t lane = threadIdx.x & 31;
    const uint4 ev = *reinterpret_cast<const uint4 *>(e_grp + g);   // eight residuals
    const uint32_t e[8] = {ev.x & 0xffffu, ev.x >> 16, ev.y & 0xffffu, ev.y >> 16, ev.z & 0xffffu, ev.z >> 16, ev.w & 0xffffu, ev.w >> 16};
    uint32_t u01[8], u23[8], u45[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        pack_costs(e[i], u01[i], u23[i], u45[i]);
        if (i) { u01[i] += u01[i - 1]; u23[i] += u23[i - 1]; u45[i] += u45[i - 1]; }   // inclusive inside my eight
    }
    // exclusive prefix over the four threads of my 32-block
    uint32_t x01 = u01[7], x23 = u23[7], x45 = u45[7];
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, x01, o, 4), b = __shfl_up_sync(0xffffffffu, x23, o, 4), c = __shfl_up_sync(0xffffffffu, x45, o, 4);
        if ((lane & 3u) >= (uint32_t)o) { x01 += a; x23 += b; x45 += c; }
    }
    const uint32_t b01 = x01 - u01[7], b23 = x23 - u23[7], b45 = x45 - u45[7];
    if (FINE) {
        // each thread holds 8 consecutive 16-byte records: transpose through shared memory (XOR swizzle against bank
        // conflicts) so that every store instruction writes 512 contiguous bytes per warp
        __shared__ uint4 stage[PREFIX_THREADS * 8];
#pragma unroll
        for (int i = 0; i < 8; i++) stage[threadIdx.x * 8 + (i ^ (threadIdx.x & 7))] = make_uint4(u01[i] + b01, u23[i] + b23, u45[i] + b45, e[i]);
        __syncthreads();
        uint4 *out = fine + ((size_t)p * cap + off);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t f = i * PREFIX_THREADS + threadIdx.x, tt = f >> 3, ii = f & 7u;
            out[f] = stage[tt * 8 + (ii ^ (tt & 7u))];
        }
    }
    if ((lane & 3u) == 3u) {                 // the last quarter holds the block's totals
        const uint32_t bi = threadIdx.x >> 2;
        tot[bi][0] = x01 & 0xffffu; tot[bi][1] = x01 >> 16;
        tot[bi][2] = x23 & 0xffffu; tot[bi][3] = x23 >> 16;
        tot[bi][4] = x45 & 0xffffu; tot[bi][5] = x45 >> 16;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        size_t blk = (size_t)grp * 32 + lane;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            uint32_t v = tot[lane][k];
            uint32_t s = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= (uint32_t)o) s += n;
            }
            blk_rec[blk * BLK_REC + k] = s - v;
            blk_rec[blk * BLK_REC + 8 + k] = s;
            if (lane == 31) grp_tot[(size_t)grp * 8 + k] = s;
        }
    }
}

// Exclusive scan (mod 2^32) of `n` records of 8 x u32 (6 used), in place, by one thread block
// per 1024 records; block totals go to `tot_out` (nullptr: not needed).
__device__ __forceinline__ void block_scan_records(uint32_t *__restrict__ rec, uint32_t n, uint32_t base_idx,
                                                   uint32_t carry_in[NK], uint32_t total_out[NK], uint32_t (*wsum)[NK]) {
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t idx = base_idx + threadIdx.x;
    uint32_t v[NK], s[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        v[k] = idx < n ? rec[(size_t)idx * 8 + k] : 0;
        s[k] = v[k];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < NK; k++) {
            uint32_t t = __shfl_up_sync(0xffffffffu, s[k], o);
            if (lane >= (uint32_t)o) s[k] += t;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < NK; k++) wsum[wid][k] = s[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NK; k++) {
            uint32_t x = wsum[lane][k], y = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= (uint32_t)o) y += t;
            }
            wsum[lane][k] = y - x;  // exclusive over warps
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NK; k++) {
        uint32_t excl = carry_in[k] + wsum[wid][k] + s[k] - v[k];
        if (idx < n) rec[(size_t)idx * 8 + k] = excl;
        total_out[k] = excl + v[k];  // meaningful in the last thread
    }
}

// level 1: every block scans 1024 group totals; its total goes to super_tot
__global__ void __launch_bounds__(1024) k_grpscan1(uint32_t *__restrict__ grp_tot, uint32_t ngroups, uint32_t *__restrict__ super_tot) {
    __shared__ uint32_t wsum[32][NK];
    uint32_t carry[NK] = {0, 0, 0, 0, 0, 0}, total[NK];
    block_scan_records(grp_tot, ngroups, blockIdx.x * 1024u, carry, total, wsum);
    if (threadIdx.x == 1023) {
#pragma unroll
        for (int k = 0; k < NK; k++) super_tot[(size_t)blockIdx.x * 8 + k] = total[k];
    }
}
// level 2: one block scans the super totals
__global__ void __launch_bounds__(1024) k_grpscan2(uint32_t *__restrict__ super_tot, uint32_t nsuper) {
    __shared__ uint32_t wsum[32][NK];
    __shared__ uint32_t carry_s[NK];
    if (threadIdx.x < NK) carry_s[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nsuper; base += 1024) {
        uint32_t carry[NK], total[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) carry[k] = carry_s[k];
        block_scan_records(super_tot, nsuper, base, carry, total, wsum);
        __syncthreads();
        if (threadIdx.x == 1023) {
#pragma unroll
            for (int k = 0; k < NK; k++) carry_s[k] = total[k];
        }
        __syncthreads();
    }
}
// make the block records global: add the group's and the super-group's exclusive prefix.
// One thread per 16-byte quarter of a record.
__global__ void __launch_bounds__(256) k_blkfinal(uint4 *__restrict__ blk_rec4, const uint32_t *__restrict__ grp_pre,
                                                  const uint32_t *__restrict__ super_pre, const uint32_t *__restrict__ plane_used,
                                                  uint32_t gpp, size_t nquarters) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquarters) return;
    size_t blk = q >> 2;
    uint32_t part = (uint32_t)(q & 3);
    uint32_t grp = (uint32_t)(blk >> 5);
    uint32_t p = grp / gpp;
    if ((grp - p * gpp) * GROUP >= plane_used[p]) return;
    const uint32_t *gp = grp_pre + (size_t)grp * 8, *sp = super_pre + (size_t)(grp >> 10) * 8;
    uint4 r = blk_rec4[q];
    if ((part & 1) == 0) { r.x += gp[0] + sp[0]; r.y += gp[1] + sp[1]; r.z += gp[2] + sp[2]; r.w += gp[3] + sp[3]; }
    else { r.x += gp[4] + sp[4]; r.y += gp[5] + sp[5]; }
    blk_rec4[q] = r;
}

// ------------------------------------------------------------------------------------
// walk: one warp per chain.  State S[6] (u32) is replicated across lanes.
// base[k] = S[k] - T(cur)[k] (mod 2^32), T = global exclusive cost prefix, cur = first
// element of the current epoch; the counters before element t are base + T(t).
// Epoch i is recorded as (first element, base) for k_kfill.
// ------------------------------------------------------------------------------------
struct WalkArgs {
    const uint4 *fine;      // per-element prefix records (FINE) ...
    const uint16_t *e_grp;  // ... or the grouped residuals
    const uint4 *blk_rec4;  // 4 x uint4 per 32-block
    const uint32_t *chain_count;
    const uint32_t *chain_base;
    const uint32_t *live;
    uint32_t *counters;     // [0] live count, [1] queue head, [2] error flags
    uint4 *ep_rec;          // 2 x uint4 per epoch: {base0..3}, {base4, base5, first element (global index), 0}
    uint32_t *blk_epoch;    // epoch in effect at the start of each 32-block
    uint32_t cap;           // grouped elements per plane
    uint32_t epcap;         // epoch records per plane
    const uint8_t *resolved; // per (plane, context): 1 = already resolved by the speculative walk (may be null)
    HopTables hop;           // segment hops for rejected chains (hop.ready == nullptr: off)
    const SpDesc *desc;
    const uint32_t *chain_fail;
    uint32_t hop_any = 0;    // hop over any planned chain, without waiting for the speculation's verdict (experiment)
};

// ---- TMA bulk copy (global -> shared) completing on an mbarrier --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

// ---- epoch walk: one warp per chain ---------------------------------------------------------------
// The six counters are replicated in every lane (base[k], re-based: counter before element t =
// base[k] + T_k(t)).  Windows of 32 blocks x 32 elements arrive by TMA, WALK_BUF in flight.  Per epoch:
//   block level    lane l holds the prefixes at the END of block l; the first block (at or after the
//                  current one) where all six counters have passed 1024 holds the halving element
//   element level  lane l looks at element l of that block; the first element where all six counters
//                  have passed 1024 is the halving (parameter_selection.rs:58), every lane has already
//                  computed the re-based halved counters for its own element, the winner's are shuffled out
// No assumption about which counter binds, so the walk is exact for any data.
constexpr int WALK_BUF = 4;
template <bool FINE>
struct WalkSmem {
    uint4 blk[WALK_BUF][32 * 4];
    uint64_t bars[WALK_BUF];
    uint4 fine[FINE ? WALK_BUF : 1][FINE ? GROUP : 1];          // FINE: per-element prefix records
    uint16_t e[FINE ? 1 : WALK_BUF][FINE ? 8 : GROUP];           // otherwise: the residuals, scanned per block on demand
};

template <bool FINE>
__global__ void __launch_bounds__(32) k_walk(WalkArgs a) {
    extern __shared__ __align__(128) unsigned char walk_smem[];
    WalkSmem<FINE> &WS = *reinterpret_cast<WalkSmem<FINE> *>(walk_smem);
    auto &sfine = WS.fine;
    auto &se = WS.e;
    auto &sblk = WS.blk;
    uint64_t *bars = WS.bars;
    const uint32_t lane = threadIdx.x;
    if (lane == 0)
        for (int b = 0; b < WALK_BUF; b++) mbar_init(&bars[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phases = 0;   // bit b = parity of the next completion of buffer b's barrier

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(&a.counters[1], 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= a.counters[0]) break;
        const uint32_t pc = a.live[qi];
        if (a.resolved && a.resolved[pc]) continue;
        const uint32_t p = pc / NBIN, cx = pc % NBIN;
        const uint32_t count = a.chain_count[pc];
        const uint32_t cbase = a.chain_base[pc];
        const uint32_t nblk = (count + 31u) >> 5;
        const size_t gbase = (size_t)p * a.cap + cbase;      // global element index of the chain start
        const size_t blk0 = gbase >> 5;                      // global 32-block index
        const uint32_t ep0 = p * a.epcap + cbase / 8 + 16 * cx;
        uint4 *rec = a.ep_rec + (size_t)ep0 * 2;
        const uint32_t ep_room = nblk * 4 + 16;              // records available to this chain
        const uint32_t nwin = (nblk + 31u) >> 5;

        uint32_t base[NK];
        uint32_t mintot;
        {
            const uint4 r0 = a.blk_rec4[blk0 * 4], r1 = a.blk_rec4[blk0 * 4 + 1];
            const uint4 z0 = a.blk_rec4[(blk0 + nblk - 1) * 4 + 2], z1 = a.blk_rec4[(blk0 + nblk - 1) * 4 + 3];
            base[0] = 0u - r0.x; base[1] = 0u - r0.y; base[2] = 0u - r0.z; base[3] = 0u - r0.w; base[4] = 0u - r1.x; base[5] = 0u - r1.y;
            mintot = min(min(min(z0.x - r0.x, z0.y - r0.y), min(z0.z - r0.z, z0.w - r0.w)), min(z1.x - r1.x, z1.y - r1.y));
        }
        if (lane == 0) {
            rec[0] = make_uint4(base[0], base[1], base[2], base[3]);
            rec[1] = make_uint4(base[4], base[5], (uint32_t)gbase, 0u);
        }
        if (mintot <= HALVE_AT) {
            // the slowest counter never passes 1024: a single epoch, nothing to walk
            for (uint32_t b = lane; b < nblk; b += 32) a.blk_epoch[blk0 + b] = ep0;
            if (lane == 0) rec[3] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
            continue;
        }

        auto prefetch = [&](uint32_t w) {
            if (lane == 0) {
                const uint32_t buf = w % WALK_BUF;
                const uint32_t nb = min(nblk - w * 32u, 32u);
                if (FINE) {
                    mbar_expect_tx(&bars[buf], nb * (32u * 16u + 64u));
                    bulk_g2s(&sfine[buf][0], a.fine + gbase + (size_t)w * GROUP, nb * 32u * 16u, &bars[buf]);
                } else {
                    mbar_expect_tx(&bars[buf], nb * (32u * 2u + 64u));
                    bulk_g2s(&se[buf][0], a.e_grp + gbase + (size_t)w * GROUP, nb * 32u * 2u, &bars[buf]);
                }
                bulk_g2s(&sblk[buf][0], a.blk_rec4 + (blk0 + w * 32u) * 4, nb * 64u, &bars[buf]);
            }
        };
        bool aborted = false;   // the speculative walk (running beside this kernel) resolved the chain
        uint32_t flag = 0;      // resolved[pc], polled ahead of its use so that the load never stalls the walk
        uint32_t nep = 1;       // epochs recorded so far
        uint32_t cur = 0;       // chain-relative first element of the current epoch
        bool overflow = false;
        // segment hops (hop_walk.cuh): once the tables of a chain the speculation rejected are ready, a segment whose
        // entry state passes the threshold tests is crossed with one lookup instead of being walked
        bool hop_on = false;
        uint32_t hop_seg0 = 0, hop_kb = 0, kb_pref = 0, kb_pref_seg = 0xFFFFFFFFu, hop_skip = 0, hop_fails = 0, ready = 0, hops_done = 0, hops_refused = 0;
        bool abs_state = false; // after a hop the counters are held as plain values at the start of window w
        uint32_t S[NK] = {0, 0, 0, 0, 0, 0};
        uint32_t w = 0;         // next window to process
        uint32_t issued = 0;    // next window to copy; windows [w, issued) are in flight
        uint32_t polls = 0;
        while (w < nwin) {
            if (flag == 1) aborted = true;
            if ((polls++ & 31u) == 0) {
                if (a.resolved) flag = *reinterpret_cast<const volatile uint8_t *>(a.resolved + pc);
                if (FINE && a.hop.ready && !hop_on && ready == 0) ready = *reinterpret_cast<const volatile uint32_t *>(a.hop.ready);
            }
            if (FINE && ready == 1 && !hop_on) {
                ready = 2;
                const uint32_t di = a.hop.pc2desc[pc];
                if (di != 0xFFFFFFFFu && (a.hop_any || *reinterpret_cast<const volatile uint32_t *>(a.chain_fail + di) == SP_OK - 1)) {
                    hop_on = true;
                    hop_seg0 = a.desc[di].seg0;
                    hop_skip = (w >> 2) + 1;   // first attempt at the next segment boundary
                }
            }
            if (aborted) {
                // only drain the copies in flight
                if (w < issued) { const uint32_t buf = w % WALK_BUF; mbar_wait(&bars[buf], (phases >> buf) & 1u); phases ^= 1u << buf; w++; continue; }
                break;
            }
            const uint32_t seg = w >> 2;
            if (FINE && hop_on && (w & 3u) == 0 && seg >= hop_skip && issued == w) {
                // ---- try to hop over segment `seg` (windows w .. w+3) ----
                const uint4 t0 = a.blk_rec4[(blk0 + (size_t)w * 32u) * 4], t1 = a.blk_rec4[(blk0 + (size_t)w * 32u) * 4 + 1];
                if (!abs_state) {
                    S[0] = base[0] + t0.x; S[1] = base[1] + t0.y; S[2] = base[2] + t0.z; S[3] = base[3] + t0.w; S[4] = base[4] + t1.x; S[5] = base[5] + t1.y;
                }
                const size_t slot = hop_seg0 + seg;
                hop_kb = kb_pref_seg == seg ? kb_pref : a.hop.segkb[slot];      // the segment's own binding candidate
                if ((seg + 1) * 4u < nwin) { kb_pref = a.hop.segkb[slot + 1]; kb_pref_seg = seg + 1; }   // in flight for the next attempt
                uint32_t x = S[0], mine = S[0];
#pragma unroll
                for (uint32_t k = 1; k < NK; k++) { x = hop_kb == k ? S[k] : x; mine = lane == k ? S[k] : mine; }
                bool good = x >= 1u && x <= HALVE_AT;
                uint32_t xn = 0, th = 0, dt = 0;
                unsigned long long A = 0;
                if (good) {
                    xn = a.hop.xn[slot * 1024 + (x - 1u)];
                    if (lane < NK) {
                        const size_t t = (slot * NK + lane) * 1024 + (x - 1u);
                        th = a.hop.theta[t]; A = a.hop.A[t]; dt = a.hop.dt[t];
                    }
                }
                const uint32_t n = xn >> 16;
                good = good && n != HOP_INVALID && nep + n + 1 < ep_room;
                good = __all_sync(0xffffffffu, good && (lane >= NK || lane == hop_kb || mine >= th));
                if (good) {
                    if (lane < 8) a.hop.log[slot * 8 + lane] = lane < NK ? mine : (lane == 6 ? nep : 1u);
                    const uint32_t mynew = lane == hop_kb ? (xn & 0xffffu) : (uint32_t)(((unsigned long long)mine + A) >> n) + dt;
#pragma unroll
                    for (uint32_t k = 0; k < NK; k++) S[k] = __shfl_sync(0xffffffffu, mynew, k);
                    nep += n;
                    abs_state = true;
                    w = min(w + 4u, nwin);
                    issued = w;
                    hop_fails = 0;
                    hops_done++;
                    continue;
                }
                hop_fails++;
                hops_refused++;
                // back off: 1, 2, 4, .. segments; up to 64 for a chain that has never hopped (it flips all the time), up to 8 otherwise
                hop_skip = seg + 1 + (hop_fails >= 2 ? (1u << min(hop_fails - 2u, hops_done ? 3u : 6u)) : 0u);
            }
            // ---- stream mode: walk window w element-exactly ----
            {
                // copies run ahead, but not across a segment boundary where a hop will be tried
                uint32_t limit = nwin;
                if (FINE && hop_on) {
                    const uint32_t nb = max(seg + 1, hop_skip);      // next segment that will be tried
                    limit = min(nwin, nb * 4u);
                }
                while (issued < limit && issued < w + (uint32_t)WALK_BUF) { prefetch(issued); issued++; }   // window w+3 reuses the buffer of w-1
            }
            const uint32_t buf = w % WALK_BUF;
            mbar_wait(&bars[buf], (phases >> buf) & 1u);
            phases ^= 1u << buf;
            const uint4 *wf = &sfine[FINE ? buf : 0][0];
            const uint16_t *we = &se[FINE ? 0 : buf][0];
            const uint4 *wb = &sblk[buf][0];
            const uint32_t wblk0 = w * 32u;
            const bool valid = wblk0 + lane < nblk;
            if (abs_state) {
                // back from plain values to re-based ones: base = S - T(window start); the open epoch began before this window
                const uint4 t0 = wb[0], t1 = wb[1];
                base[0] = S[0] - t0.x; base[1] = S[1] - t0.y; base[2] = S[2] - t0.z; base[3] = S[3] - t0.w; base[4] = S[4] - t1.x; base[5] = S[5] - t1.y;
                cur = wblk0 * 32u;
                abs_state = false;
            }
            uint32_t incl[NK];
            {
                const uint4 r2 = wb[lane * 4 + 2], r3 = wb[lane * 4 + 3];
                incl[0] = r2.x; incl[1] = r2.y; incl[2] = r2.z; incl[3] = r2.w; incl[4] = r3.x; incl[5] = r3.y;
            }
            uint32_t my_epoch = nep - 1;                              // epoch in effect at the start of my block
            int curblk = (int)(cur >> 5) - (int)wblk0;                // block (window-relative) holding `cur`
            uint32_t blkmask = curblk <= 0 ? 0xffffffffu : (curblk >= 32 ? 0u : (0xffffffffu << curblk));
            uint32_t lanemask = 0xffffffffu << (cur & 31u);
            for (;;) {
                // block level
                const int32_t bm = min(min(min((int32_t)(base[0] + incl[0]), (int32_t)(base[1] + incl[1])), min((int32_t)(base[2] + incl[2]), (int32_t)(base[3] + incl[3]))),
                                       min((int32_t)(base[4] + incl[4]), (int32_t)(base[5] + incl[5])));
                const uint32_t m = __ballot_sync(0xffffffffu, valid && bm > (int32_t)HALVE_AT) & blkmask;
                if (!m) break;
                const int B = __ffs(m) - 1;
                // element level
                uint4 f;
                if (FINE) {
                    f = wf[B * 32 + lane];
                } else {
                    pack_costs(we[B * 32 + lane], f.x, f.y, f.z);
                    warp_scan3(f.x, f.y, f.z, lane);
                }
                const uint4 e0 = wb[B * 4], e1 = wb[B * 4 + 1];
                uint32_t T[NK], v[NK];
                T[0] = e0.x + (f.x & 0xffffu); T[1] = e0.y + (f.x >> 16); T[2] = e0.z + (f.y & 0xffffu);
                T[3] = e0.w + (f.y >> 16); T[4] = e1.x + (f.z & 0xffffu); T[5] = e1.y + (f.z >> 16);
#pragma unroll
                for (int k = 0; k < NK; k++) v[k] = base[k] + T[k];                  // counters after my element's update
                const int32_t em = min(min(min((int32_t)v[0], (int32_t)v[1]), min((int32_t)v[2], (int32_t)v[3])), min((int32_t)v[4], (int32_t)v[5]));
                uint32_t fm = __ballot_sync(0xffffffffu, em > (int32_t)HALVE_AT);
                if (B == curblk) fm &= lanemask;
                const int h = __ffs(fm) - 1;                                          // fm != 0: the block's last element passes
#pragma unroll
                for (int k = 0; k < NK; k++) base[k] = __shfl_sync(0xffffffffu, (v[k] >> 1) - T[k], h);   // parameter_selection.rs:58-63, re-based
                cur = (wblk0 + (uint32_t)B) * 32u + (uint32_t)h + 1u;
                curblk = B + (h == 31 ? 1 : 0);
                blkmask = curblk >= 32 ? 0u : (0xffffffffu << curblk);
                lanemask = 0xffffffffu << ((uint32_t)(h + 1) & 31u);
                if (nep + 1 < ep_room) {
                    if (lane == 0) {
                        rec[2 * nep] = make_uint4(base[0], base[1], base[2], base[3]);
                        rec[2 * nep + 1] = make_uint4(base[4], base[5], (uint32_t)gbase + cur, 0u);
                    }
                } else {
                    overflow = true;
                }
                nep++;
                if ((int)lane > B) my_epoch++;
            }
            if (valid) a.blk_epoch[blk0 + wblk0 + lane] = ep0 + my_epoch;
            __syncwarp();
            w++;
        }
        if (lane == 0 && (hops_done | hops_refused)) { atomicAdd(&a.counters[5], hops_done); atomicAdd(&a.counters[6], hops_refused); }
        if (lane == 0 && !aborted) {
            if (overflow) atomicOr(&a.counters[2], 1u);
            else rec[2 * nep + 1] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);  // sentinel
        }
    }
}

// kfill: k = argmin (ties to the largest k, parameter_selection.rs:78-83) of the counters before every out-of-range
// element.  One thread per eight consecutive elements (four threads per 32-block): the costs are re-derived from the
// residuals and scanned over the block, the epoch bases come from the walk's records.
__global__ void __launch_bounds__(256, 4) k_kfill(const uint16_t *__restrict__ e_grp, const uint4 *__restrict__ blk_rec4,
                                               const uint32_t *__restrict__ blk_epoch, const uint4 *__restrict__ ep_rec,
                                               const uint32_t *__restrict__ plane_used, uint32_t cap, uint32_t np,
                                               uint8_t *__restrict__ k_grp) {
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8u;
    const uint32_t p = (uint32_t)(g / cap);
    if (p >= np) return;
    const uint32_t off = (uint32_t)(g - (size_t)p * cap);
    if (off >= plane_used[p]) return;   // whole warps: plane_used and cap are multiples of 256
    const uint32_t lane = threadIdx.x & 31;
    const uint4 ev = *reinterpret_cast<const uint4 *>(e_grp + g);
    const uint32_t e[8] = {ev.x & 0xffffu, ev.x >> 16, ev.y & 0xffffu, ev.y >> 16, ev.z & 0xffffu, ev.z >> 16, ev.w & 0xffffu, ev.w >> 16};
    uint32_t x01[8], x23[8], x45[8];    // cost of the elements before element i inside my eight
    uint32_t t01 = 0, t23 = 0, t45 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t a, b, c;
        pack_costs(e[i], a, b, c);
        x01[i] = t01; x23[i] = t23; x45[i] = t45;
        t01 += a; t23 += b; t45 += c;
    }
    uint32_t s01 = t01, s23 = t23, s45 = t45;
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, s01, o, 4), b = __shfl_up_sync(0xffffffffu, s23, o, 4), c = __shfl_up_sync(0xffffffffu, s45, o, 4);
        if ((lane & 3u) >= (uint32_t)o) { s01 += a; s23 += b; s45 += c; }
    }
    s01 -= t01; s23 -= t23; s45 -= t45;   // cost of the block's elements before my eight
    const size_t blk = g >> 5;
    const uint4 c0 = blk_rec4[blk * 4], c1 = blk_rec4[blk * 4 + 1];
    // counters before element i = epoch base + prefix at the block start + costs inside the block before i
    const uint32_t q[NK] = {c0.x + (s01 & 0xffffu), c0.y + (s01 >> 16), c0.z + (s23 & 0xffffu), c0.w + (s23 >> 16), c1.x + (s45 & 0xffffu), c1.y + (s45 >> 16)};
    uint32_t ep = blk_epoch[blk];
    uint4 b0 = ep_rec[(size_t)ep * 2], b1 = ep_rec[(size_t)ep * 2 + 1];
    uint32_t next = ep_rec[(size_t)ep * 2 + 3].z;   // first element of the following epoch (0xFFFFFFFF after the last one)
    // A counter x travels as the key x * 8 + (5 - k): the smallest key is the smallest count with ties going to the largest
    // k, i.e. get_k's `<=` scan (parameter_selection.rs:78-83).  Counts stay far below 2^28, so the key fits.
    uint32_t key[NK];
    auto rekey = [&]() {
        key[0] = (b0.x + q[0]) * 8u + 5u; key[1] = (b0.y + q[1]) * 8u + 4u; key[2] = (b0.z + q[2]) * 8u + 3u;
        key[3] = (b0.w + q[3]) * 8u + 2u; key[4] = (b1.x + q[4]) * 8u + 1u; key[5] = (b1.y + q[5]) * 8u;
    };
    rekey();
    uint32_t kk[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t gi = (uint32_t)g + i;
        while (next <= gi) {
            ep++;
            b0 = ep_rec[(size_t)ep * 2]; b1 = ep_rec[(size_t)ep * 2 + 1];
            next = ep_rec[(size_t)ep * 2 + 3].z;
            rekey();
        }
        const uint32_t m = min(min(min(key[0] + ((x01[i] & 0xffffu) << 3), key[1] + ((x01[i] >> 16) << 3)),
                                   min(key[2] + ((x23[i] & 0xffffu) << 3), key[3] + ((x23[i] >> 16) << 3))),
                               min(key[4] + ((x45[i] & 0xffffu) << 3), key[5] + ((x45[i] >> 16) << 3)));
        kk[i] = 5u - (m & 7u);
    }
    uint2 out;
    out.x = kk[0] | (kk[1] << 8) | (kk[2] << 16) | (kk[3] << 24);
    out.y = kk[4] | (kk[5] << 8) | (kk[6] << 16) | (kk[7] << 24);
    *reinterpret_cast<uint2 *>(k_grp + g) = out;
}

// code record of one pixel (device_common.cuh: length << 22 | payload): marker + phased-in code, or marker + Rice code
// with the k of its grouped element (compression.rs:130-145)
__device__ __forceinline__ uint32_t code_record_g(const PixelClass &pc, const uint8_t *__restrict__ kg, uint32_t g);
__device__ __forceinline__ uint32_t code_record(const PixelClass &pc, const uint32_t *__restrict__ gi, const uint8_t *__restrict__ kg, uint32_t i) {
    return code_record_g(pc, kg, pc.cls == 0 ? 0u : gi[i]);
}
// the same with the grouped index already at hand (ignored for in-range pixels)
__device__ __forceinline__ uint32_t code_record_g(const PixelClass &pc, const uint8_t *__restrict__ kg, uint32_t g) {
    if (pc.cls == 0) {
        int len;
        const uint32_t code = phase_in_code((uint32_t)pc.delta + 1u, (uint32_t)pc.val, len);
        return ((uint32_t)(len + 1) << 22) | (1u << len) | code;           // '1' marker then the phased-in code
    }
    const uint32_t k = kg[g];
    const uint32_t e = (uint32_t)pc.val;
    const uint32_t q = e >> k, rem = e & ((1u << k) - 1u);
    const uint32_t above = pc.cls == 1 ? 1u : 0u;
    const uint32_t len = 2u + q + 1u + k;
    if (len <= (uint32_t)REC_SHORT_MAX)                                     // '0', above, q ones, '0', k remainder bits
        return (len << 22) | (above << (q + 1u + k)) | (((1u << q) - 1u) << (k + 1u)) | rem;
    return (len << 22) | (above << 17) | (k << 14) | (rem << 9) | q;
}

// ------------------------------------------------------------------------------------
// code: marker + code word per pixel (compression.rs:130-145), bits per tile
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS) k_code(const T *__restrict__ planes, uint32_t w, uint32_t npix, uint32_t tpp,
                                                       uint32_t cap, const uint32_t *__restrict__ gidx,
                                                       const uint8_t *__restrict__ k_grp, uint32_t *__restrict__ rec,
                                                       uint32_t *__restrict__ tile_bits) {
    __shared__ uint32_t wsum[TILE_WARPS];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / tpp, t = bid - p * tpp;
    const T *pl = planes + (size_t)p * npix;
    uint32_t start = t * TILE;
    uint32_t bits = 0;
    RasterCursor<T> cur;
    cur.init(pl, start + threadIdx.x, w);
#pragma unroll 4
    for (int j = 0; j < TILE / TILE_THREADS; j++) {
        const uint32_t i = cur.i;
        if (i >= npix) break;
        uint32_t r = 0;
        if (i >= 2) r = code_record(cur.classify(), gidx + (size_t)p * npix, k_grp + (size_t)p * cap, i);
        rec[(size_t)p * npix + i] = r;
        bits += rec_len(r);
        cur.step(TILE_THREADS);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < TILE_WARPS; q++) s += wsum[q];
        tile_bits[bid] = s;
    }
}

// The same for planes whose width is a multiple of four: four consecutive pixels per thread, records leave as one 16-byte store.
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS) k_code4(const T *__restrict__ planes, uint32_t w, uint32_t npix, uint32_t tpp, uint32_t cap,
                                                        const uint32_t *__restrict__ gidx, const uint8_t *__restrict__ k_grp, uint32_t *__restrict__ rec,
                                                        uint32_t *__restrict__ tile_bits) {
    __shared__ uint32_t wsum[TILE_WARPS];
    const uint32_t bid = blockIdx.x;
    const uint32_t p = bid / tpp, t = bid - p * tpp;
    const T *pl = planes + (size_t)p * npix;
    const uint32_t *gi = gidx + (size_t)p * npix;
    const uint8_t *kg = k_grp + (size_t)p * cap;
    uint32_t bits = 0;
    RasterCursor<T> cur;
    cur.init(pl, t * TILE + 4u * threadIdx.x, w);
    const uint32_t step_q = (4u * TILE_THREADS) / w, step_r = (4u * TILE_THREADS) - step_q * w;
#pragma unroll
    for (int j = 0; j < TILE / (4 * TILE_THREADS); j++) {
        const uint32_t i = cur.i;
        if (i < npix) {
            PixelClass pc[4];
            bool valid[4];
            classify4(pl, i, cur.x, cur.y, w, pc, valid);
            const uint4 g4 = *reinterpret_cast<const uint4 *>(gi + i);   // grouped indices of the quad (stale for in-range pixels: unused)
            const uint32_t gq[4] = {g4.x, g4.y, g4.z, g4.w};
            uint32_t r[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                r[q] = valid[q] ? code_record_g(pc[q], kg, gq[q]) : 0u;
                bits += rec_len(r[q]);
            }
            *reinterpret_cast<uint4 *>(rec + (size_t)p * npix + i) = make_uint4(r[0], r[1], r[2], r[3]);   // npix and i are multiples of four
        }
        cur.step_qr(4 * TILE_THREADS, step_q, step_r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < TILE_WARPS; q++) sum += wsum[q];
        tile_bits[bid] = sum;
    }
}

// ------------------------------------------------------------------------------------
// bitscan
// ------------------------------------------------------------------------------------
// per plane: exclusive scan of tile_bits -> tile_off (u64, excludes the 64 raw bits); plane_bits = 64 + total
__global__ void __launch_bounds__(1024) k_planebits(const uint32_t *__restrict__ tile_bits, uint32_t tpp,
                                                    uint64_t *__restrict__ tile_off, uint64_t *__restrict__ plane_bits) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    uint32_t p = blockIdx.x;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < tpp; base += 1024) {
        uint32_t t = base + threadIdx.x;
        uint64_t v = t < tpp ? tile_bits[(size_t)p * tpp + t] : 0;
        uint64_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += n;
        }
        if (lane == 31) wsum[wid] = s;
        __syncthreads();
        if (wid == 0) {
            uint64_t x = wsum[lane], y = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t n = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= (uint32_t)o) y += n;
            }
            wsum[lane] = y - x;
        }
        __syncthreads();
        uint64_t excl = carry + wsum[wid] + s - v;
        if (t < tpp) tile_off[(size_t)p * tpp + t] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) plane_bits[p] = 64 + carry;
}

// per image: bytes = 14 + ceil(sum of plane bits / 8); exclusive scan -> img_off[0..ni]
__global__ void __launch_bounds__(1024) k_imgscan(const uint64_t *__restrict__ plane_bits, uint32_t ni, uint32_t nch,
                                                  uint64_t *__restrict__ img_off) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ni; base += 1024) {
        uint32_t im = base + threadIdx.x;
        uint64_t v = 0;
        if (im < ni) {
            uint64_t bits = 0;
            for (uint32_t q = 0; q < nch; q++) bits += plane_bits[(size_t)im * nch + q];
            v = FELICS_HEADER_BYTES + (bits + 7) / 8;
        }
        uint64_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += n;
        }
        if (lane == 31) wsum[wid] = s;
        __syncthreads();
        if (wid == 0) {
            uint64_t x = wsum[lane], y = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t n = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= (uint32_t)o) y += n;
            }
            wsum[lane] = y - x;
        }
        __syncthreads();
        uint64_t excl = carry + wsum[wid] + s - v;
        if (im < ni) img_off[im] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) img_off[ni] = carry;
}

// ------------------------------------------------------------------------------------
// pack
// ------------------------------------------------------------------------------------
constexpr int PACK_WORDS = 8192;  // 32 KB of shared memory = 262144 bits per tile before the slow path

struct PackArgs {
    const uint32_t *rec;
    const uint32_t *tile_bits;
    const uint64_t *tile_off;
    const uint64_t *plane_bits;
    const uint64_t *img_off;
    uint32_t *arena;         // 4-byte aligned
    uint64_t arena_byte0;    // byte offset of image 0 of this sub-batch inside the arena
    uint32_t npix, tpp, nch;
};

__device__ __forceinline__ uint64_t plane_bit_start(const PackArgs &a, uint32_t p) {
    uint32_t img = p / a.nch, ch = p - img * a.nch;
    uint64_t bit = 8ull * (a.arena_byte0 + a.img_off[img] + FELICS_HEADER_BYTES);
    for (uint32_t q = 0; q < ch; q++) bit += a.plane_bits[(size_t)img * a.nch + q];
    return bit;
}

// emit one record at bit offset `off` (relative to buffer word 0) through PUT(bitoff, value, nbits)
template <typename PUT>
__device__ __forceinline__ void emit_record(uint32_t r, uint64_t off, PUT put) {
    uint32_t len = rec_len(r);
    if (len == 0) return;
    if (len <= (uint32_t)REC_SHORT_MAX) {
        put(off, r & 0x3fffffu, (int)len);
        return;
    }
    uint32_t q = r & 511u, rem = (r >> 9) & 31u, k = (r >> 14) & 7u, above = (r >> 17) & 1u;
    put(off, above, 2);  // '0', above
    off += 2;
    while (q >= 32) { put(off, 0xffffffffu, 32); off += 32; q -= 32; }
    if (q) { put(off, (1u << q) - 1u, (int)q); off += q; }
    put(off, rem, (int)k + 1);  // '0' then k remainder bits
}

__global__ void __launch_bounds__(TILE_THREADS) k_pack(PackArgs a) {
    __shared__ uint32_t buf[PACK_WORDS];
    __shared__ uint32_t wtot[TILE_WARPS];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / a.tpp, t = bid - p * a.tpp;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t tbits = a.tile_bits[bid];
    const uint64_t bit0 = plane_bit_start(a, p) + 64 + a.tile_off[bid];
    const uint64_t word0 = bit0 >> 5;
    const uint32_t sh0 = (uint32_t)(bit0 & 31);
    const uint32_t nwords = (sh0 + tbits + 31u) >> 5;
    const bool in_smem = nwords <= (uint32_t)PACK_WORDS;
    if (in_smem)
        for (uint32_t j = threadIdx.x; j < nwords; j += TILE_THREADS) buf[j] = 0;

    // a lane owns 16 consecutive pixels (four 16-byte loads): their codes are concatenated in registers and leave as
    // whole words, so the shared-memory traffic is a few atomicOr per lane instead of one or two per pixel, and the
    // bit offsets need one warp scan per 512 pixels
    const uint32_t *rc = a.rec + (size_t)p * a.npix;
    const uint32_t wstart = t * TILE + wid * WARP_PIX;
    const uint32_t lstart = wstart + lane * WARP_ITERS;
    uint32_t r[WARP_ITERS];
    if (lstart + WARP_ITERS <= a.npix && (((size_t)p * a.npix + lstart) & 3u) == 0) {
        const uint4 *r4 = reinterpret_cast<const uint4 *>(rc + lstart);
#pragma unroll
        for (int q = 0; q < WARP_ITERS / 4; q++) {
            const uint4 v = r4[q];
            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int it = 0; it < WARP_ITERS; it++) r[it] = lstart + it < a.npix ? rc[lstart + it] : 0u;
    }
    uint32_t mylen = 0;
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) mylen += rec_len(r[it]);
    uint32_t inc = mylen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += n;
    }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    uint32_t pos = sh0 + inc - mylen;  // bit offset of my first code relative to word0
    for (uint32_t q = 0; q < wid; q++) pos += wtot[q];

    auto put_s = [&](uint64_t off, uint32_t val, int n) { put_bits_smem(buf, (uint32_t)off, val, n); };
    auto put_g = [&](uint64_t off, uint32_t val, int n) { put_bits_global(a.arena, (word0 << 5) + off, val, n); };
    auto flush = [&](uint32_t w, uint32_t v) {
        if (!v) return;
        if (in_smem) atomicOr(&buf[w], v);
        else atomicOr(&a.arena[word0 + w], bswap32(v));
    };
    uint32_t wi = pos >> 5, sh = pos & 31u, wv = 0;   // the word being filled: index, bits already placed in it, their value
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        const uint32_t len = rec_len(r[it]);
        if (len == 0) continue;
        if (len <= (uint32_t)REC_SHORT_MAX) {
            const uint32_t left = (r[it] & 0x3fffffu) << (32u - len);   // code word, MSB aligned
            wv |= left >> sh;
            if (sh + len >= 32u) {
                flush(wi, wv);
                wi++;
                wv = sh + len > 32u ? left << (32u - sh) : 0u;         // sh > 0 here: len <= 22
                sh = sh + len - 32u;
            } else {
                sh += len;
            }
        } else {
            // long unary run: emitted field by field; my partial word goes out first
            flush(wi, wv);
            const uint32_t at = (wi << 5) + sh;
            if (in_smem) emit_record(r[it], at, put_s);
            else emit_record(r[it], at, put_g);
            const uint32_t np2 = at + len;
            wi = np2 >> 5; sh = np2 & 31u; wv = 0;
        }
    }
    flush(wi, wv);
    if (!in_smem) return;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < nwords; j += TILE_THREADS) {
        uint32_t v = buf[j];
        if (j == 0 || j == nwords - 1) {
            if (v) atomicOr(&a.arena[word0 + j], bswap32(v));
        } else {
            a.arena[word0 + j] = bswap32(v);
        }
    }
}

// header + raw first two samples of every plane (format.rs:51-61, compression.rs:93-108)
template <typename T>
__global__ void k_heads(PackArgs a, const T *__restrict__ planes, uint32_t np, uint32_t width, uint32_t height,
                        uint32_t color, uint32_t depth) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    uint64_t bit = plane_bit_start(a, p);
    int32_t v0 = 0, v1 = 0;
    if (a.npix >= 1) v0 = planes[(size_t)p * a.npix];
    if (a.npix >= 2) v1 = planes[(size_t)p * a.npix + 1];
    put_bits_global(a.arena, bit, (uint32_t)v0, 32);
    put_bits_global(a.arena, bit + 32, (uint32_t)v1, 32);
    if (p % a.nch == 0) {
        uint64_t hb = bit - 8ull * FELICS_HEADER_BYTES;
        put_bits_global(a.arena, hb, 0x464C4353u, 32);  // "FLCS"
        put_bits_global(a.arena, hb + 32, (color << 8) | depth, 16);
        put_bits_global(a.arena, hb + 48, width, 32);
        put_bits_global(a.arena, hb + 80, height, 32);
    }
}

// ------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------
namespace {

struct Carver {
    uint8_t *base;
    size_t off = 0;
    template <typename T>
    T *take(size_t count) {
        off = align_up(off, 256);
        T *ptr = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return ptr;
    }
};

struct Layout {
    int16_t *planes;
    uint32_t *tile_hist, *chunk_tot, *chain_count, *chain_base, *plane_used, *live, *counters;
    uint16_t *e_grp;
    uint32_t *gidx;
    uint4 *fine;
    uint32_t *blk_rec, *grp_tot, *super_tot, *ep_rec, *blk_epoch;
    uint8_t *k_grp;
    uint32_t *rec, *tile_bits;
    uint64_t *tile_off, *plane_bits, *img_off;
    // speculative walk (single big images only; null otherwise)
    bool sp;
    SpSizes spsz;
    SpDesc *sp_desc;
    uint32_t *sp_seg_desc, *sp_grp_desc, *sp_eb_desc, *sp_counts, *sp_map, *sp_pre, *sp_gmap, *sp_grp_xin, *sp_grp_nbefore, *sp_chain_n, *sp_chain_fail;
    uint4 *sp_trow;
    unsigned long long *sp_ablk;
    uint8_t *sp_resolved;
    HopTables hop;
    uint32_t *sp_pc2desc;
    size_t bytes;
};

struct Geom {
    uint32_t w, h, npix, nch, tpp, nchunks, cap, gpp, epcap;
};

Layout carve(uint8_t *base, const Geom &g, size_t ni) {
    Layout L;
    Carver c{base};
    size_t np = ni * g.nch;
    L.planes = c.take<int16_t>(g.nch == 1 ? 8 : np * g.npix + 8);   // gray needs no planes
    L.tile_hist = c.take<uint32_t>(np * g.tpp * NBIN);
    L.chunk_tot = c.take<uint32_t>(np * g.nchunks * NBIN);
    L.chain_count = c.take<uint32_t>(np * NBIN);
    L.chain_base = c.take<uint32_t>(np * NBIN);
    L.plane_used = c.take<uint32_t>(np);
    L.live = c.take<uint32_t>(np * NBIN);
    L.counters = c.take<uint32_t>(8);
    L.e_grp = c.take<uint16_t>(np * g.cap);
    L.gidx = c.take<uint32_t>(np * g.npix + 8);
    L.sp = np <= (size_t)SP_MAX_PLANES && g.npix >= SP_MIN_COUNT;   // big single images: speculative walk, per-element records
    L.fine = c.take<uint4>(L.sp ? np * g.cap : 8);
    L.blk_rec = c.take<uint32_t>(np * (g.cap / 32) * BLK_REC);
    L.grp_tot = c.take<uint32_t>(np * g.gpp * 8 + 8);
    L.super_tot = c.take<uint32_t>((np * g.gpp / 1024 + 2) * 8);
    L.ep_rec = c.take<uint32_t>((np * g.epcap + 8) * 8);
    L.blk_epoch = c.take<uint32_t>(np * (g.cap / 32));
    L.k_grp = c.take<uint8_t>(np * g.cap);
    L.rec = c.take<uint32_t>(np * g.npix + 8);
    L.tile_bits = c.take<uint32_t>(np * g.tpp + 8);
    L.tile_off = c.take<uint64_t>(np * g.tpp + 8);
    L.plane_bits = c.take<uint64_t>(np + 8);
    L.img_off = c.take<uint64_t>(ni + 8);
    if (L.sp) {
        const SpSizes z = sp_sizes((uint32_t)np, g.cap);
        L.spsz = z;
        L.sp_desc = c.take<SpDesc>(z.max_desc);
        L.sp_seg_desc = c.take<uint32_t>(z.max_seg);
        L.sp_grp_desc = c.take<uint32_t>(z.max_grp);
        L.sp_eb_desc = c.take<uint32_t>(z.max_eb);
        L.sp_counts = c.take<uint32_t>(8);
        L.sp_map = c.take<uint32_t>((size_t)z.max_seg * SP_DOM);
        L.sp_pre = c.take<uint32_t>((size_t)z.max_seg * SP_DOM);
        L.sp_gmap = c.take<uint32_t>((size_t)z.max_grp * SP_DOM);
        L.sp_grp_xin = c.take<uint32_t>(z.max_grp);
        L.sp_grp_nbefore = c.take<uint32_t>(z.max_grp);
        L.sp_chain_n = c.take<uint32_t>(z.max_desc);
        L.sp_chain_fail = c.take<uint32_t>(z.max_desc);
        L.sp_trow = c.take<uint4>(((size_t)np * g.epcap + 8) * 2);
        L.sp_pc2desc = c.take<uint32_t>(np * NBIN);
        L.hop.xn = c.take<uint32_t>((size_t)z.max_seg * 1024);
        L.hop.theta = c.take<uint32_t>((size_t)z.max_seg * NK * 1024);
        L.hop.A = c.take<unsigned long long>((size_t)z.max_seg * NK * 1024);
        L.hop.dt = c.take<uint32_t>((size_t)z.max_seg * NK * 1024);
        L.hop.segkb = c.take<uint32_t>(z.max_seg);
        L.hop.log = c.take<uint32_t>((size_t)z.max_seg * 8 + 8);   // + the ready flag
        L.hop.ready = L.hop.log ? L.hop.log + (size_t)z.max_seg * 8 : nullptr;
        L.hop.pc2desc = L.sp_pc2desc;
        L.sp_ablk = c.take<unsigned long long>((size_t)z.max_eb * 8);
    }
    L.sp_resolved = c.take<uint8_t>(np * NBIN);
    L.bytes = align_up(c.off, 256);
    return L;
}

}  // namespace

}  // namespace felics
#include "enc16_par.cuh"
#define FELICS_SIDECAR_BUILDER
#include "sidecar.cuh"
namespace felics {

// Exactly one of d_arena (device memory, 4-byte aligned) / h_arena (host memory) is non-null.
// With h_arena every sub-batch is packed into the context's staging buffer and copied out.
// With h_pixels (host memory in, host arena out) the sub-batches are double buffered: the copy-in of sub-batch i+1 and the
// copy-out of sub-batch i-1 run on their own streams beside the kernels of sub-batch i.
int encode_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr, uint8_t *d_arena,
                        uint8_t *h_arena, size_t arena_cap, uint64_t *offsets_host, const void *h_pixels) {
    ctx->last.valid = false;
    // batches of gray images, device resident: one block per image, everything on chip (stream.cu); any arena alignment
    if (d_arena && !h_pixels && stream_eligible(ctx, n, d_pixels, hdr)) return stream_encode_batch_device(ctx, n, d_pixels, hdr, d_arena, arena_cap, offsets_host);
    if (d_arena && ((uintptr_t)d_arena & 3) != 0) {
        set_error("device arena must be 4-byte aligned");
        return FELICS_ERR_INVALID_ARGUMENT;
    }
    if (hdr.pixel_depth != 0) return encode16_batch_device(ctx, n, d_pixels, hdr, d_arena, h_arena, arena_cap, offsets_host);
    uint64_t npix64 = (uint64_t)hdr.width * hdr.height;
    if (npix64 > 0x7fff0000ull) {
        set_error("image too large for one call: %llu pixels", (unsigned long long)npix64);
        return FELICS_ERR_INVALID_DIMENSIONS;
    }
    cudaStream_t st = ctx->stream;
    Geom g;
    g.w = hdr.width; g.h = hdr.height; g.npix = (uint32_t)npix64;
    g.nch = hdr.color_type ? 3 : 1;
    g.tpp = (g.npix + TILE - 1) / TILE;
    g.nchunks = (g.tpp + CHUNK_TILES - 1) / CHUNK_TILES;
    g.cap = (uint32_t)align_up((size_t)g.npix + NBIN * 32, GROUP);
    g.gpp = g.cap / GROUP;
    g.epcap = g.cap / 8 + NBIN * 16;
    const size_t usable_cap = d_arena ? (arena_cap & ~(size_t)3) : arena_cap;
    const size_t img_pix_bytes = (size_t)g.npix * g.nch;

    // images per sub-batch: bound scratch (and keep every global element index below 2^32)
    size_t per_image = carve(nullptr, g, 1).bytes;
    // the walk of a sub-batch has one warp per chain and is bound by latency: the more images travel together, the better
    // the machine is filled, so the scratch budget is half of what the device has free (the context's own scratch counted as
    // free), between 12 and 48 GB
    size_t budget = (size_t)12 << 30;
    if (n > 1 && per_image * n > budget) {
        if (!ctx->batch_budget) {   // asked once per context: the query itself takes milliseconds
            size_t free_b = 0, total_b = 0;
            ctx->batch_budget = budget;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
                ctx->batch_budget = std::min<size_t>((size_t)48 << 30, std::max<size_t>(budget, (free_b + ctx->scratch_cap) / 2));
        }
        budget = ctx->batch_budget;
    }
    size_t sub = std::max<size_t>(1, std::min<size_t>(n, budget / std::max<size_t>(per_image, 1)));
    while (sub > 1 && (uint64_t)sub * g.nch * g.cap >= 0xffff0000ull) sub--;
    if ((uint64_t)g.nch * g.cap >= 0xffff0000ull) {
        set_error("image too large for one call");
        return FELICS_ERR_INVALID_DIMENSIONS;
    }
    const bool piped = h_pixels != nullptr;
    if (!d_arena && !piped) {
        set_error("a host arena needs host pixels (the pipelined path)");
        return FELICS_ERR_INVALID_ARGUMENT;
    }
    // copies of earlier sub-batches may still be writing the caller's arena: finish them before any early return
    auto drain = [&]() {
        if (piped && ctx->copy_in) { cudaStreamSynchronize(ctx->copy_in); cudaStreamSynchronize(ctx->copy_out); }
    };
    if (piped) {
        if (n >= 64) sub = std::min(sub, std::max<size_t>(32, (n + 7) / 8));   // at least eight sub-batches to overlap
        if (!ctx->copy_in) {
            FELICS_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
            FELICS_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
            for (int i = 0; i < 2; i++) {
                FELICS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
                FELICS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
                FELICS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_pack[i], cudaEventDisableTiming));
                FELICS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
            }
        }
    }
    // copy sub-batch `idx` (images first .. first+ni) into its staging slot, after the kernels that last read the slot
    auto enqueue_in = [&](size_t idx, size_t first, size_t ni) -> int {
        const int slot = (int)(idx & 1);
        const size_t bytes = ni * img_pix_bytes;
        int rc2 = ensure_buffer(ctx, &ctx->stage_in[slot], &ctx->stage_in_cap[slot], bytes + 16);
        if (rc2) return rc2;
        FELICS_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_done[slot], 0));
        if (bytes) FELICS_CUDA_TRY(cudaMemcpyAsync(ctx->stage_in[slot], (const uint8_t *)h_pixels + first * img_pix_bytes, bytes, cudaMemcpyHostToDevice, ctx->copy_in));
        FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_in[slot], ctx->copy_in));
        return FELICS_OK;
    };
    if (piped) {
        // the copy streams start after whatever the caller queued on the context's stream
        FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_done[0], st));
        FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_done[1], st));
        int rc0 = enqueue_in(0, 0, std::min(sub, n));
        if (rc0) return rc0;
    }

    uint64_t arena_off = 0;
    offsets_host[0] = 0;
    size_t idx = 0;
    for (size_t first = 0; first < n; first += sub, idx++) {
        const size_t ni = std::min(sub, n - first);
        const size_t np = ni * g.nch;
        const int slot = (int)(idx & 1);
        Layout L = carve(nullptr, g, ni);
        int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, L.bytes);
        if (rc) return rc;
        L = carve((uint8_t *)ctx->scratch, g, ni);
        const uint8_t *px = (const uint8_t *)d_pixels + first * img_pix_bytes;
        if (piped) {
            if (first + sub < n) {   // the next sub-batch's pixels travel while this one is encoded
                rc = enqueue_in(idx + 1, first + sub, std::min(sub, n - first - sub));
                if (rc) return rc;
            }
            FELICS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_in[slot], 0));
            px = (const uint8_t *)ctx->stage_in[slot];
        }

        // gray samples are classified straight from the caller's pixels; RGB goes through Y/Co/Cg planes
        const bool gray = g.nch == 1;
        // width a multiple of four (and 4-byte aligned samples): four consecutive samples per thread in the order-free kernels
        const bool quads = g.w % 4 == 0 && g.w >= 8 && (!gray || ((uintptr_t)px & 3) == 0) && !ctx->no_quads;
        if (g.npix > 0 && !gray) {
            StageScope s(ctx, ST_PLANES);
            size_t total = ni * (size_t)g.npix;
            unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
            k_to_planes_rgb8<<<blocks, 256, 0, st>>>(px, L.planes, g.npix, total);
            s.launched();
        }
        const unsigned ntiles = (unsigned)(np * g.tpp);
        if (g.npix > 2) {
            {
                StageScope s(ctx, ST_HIST);
                FELICS_CUDA_TRY(cudaMemsetAsync(L.chunk_tot, 0, np * g.nchunks * NBIN * sizeof(uint32_t), st));
                FELICS_CUDA_TRY(cudaMemsetAsync(L.counters, 0, 8 * sizeof(uint32_t), st));
                if (quads && gray) k_hist4<uint8_t><<<ntiles, TILE_THREADS, 0, st>>>(px, g.w, g.npix, g.tpp, g.nchunks, L.tile_hist, L.chunk_tot);
                else if (quads) k_hist4<int16_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.nchunks, L.tile_hist, L.chunk_tot);
                else if (gray) k_hist<uint8_t><<<ntiles, TILE_THREADS, 0, st>>>(px, g.w, g.npix, g.tpp, g.nchunks, L.tile_hist, L.chunk_tot);
                else k_hist<int16_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.nchunks, L.tile_hist, L.chunk_tot);
                s.launched();
            }
            {
                StageScope s(ctx, ST_CHAINSCAN);
                k_chainscan<<<(unsigned)np, NBIN, 0, st>>>(L.chunk_tot, g.nchunks, L.chain_count, L.chain_base, L.plane_used, L.live, L.counters);
                s.launched();
            }
            {
                StageScope s(ctx, ST_TILEBASE);
                k_tilebase<<<(unsigned)(np * g.nchunks), NBIN, 0, st>>>(L.tile_hist, L.chunk_tot, L.chain_base, g.tpp, g.nchunks);
                s.launched();
            }
            {
                StageScope s(ctx, ST_SCATTER);
                FELICS_CUDA_TRY(cudaMemsetAsync(L.e_grp, 0xFF, np * (size_t)g.cap * sizeof(uint16_t), st));
                if (quads && gray) k_scatter4<uint8_t><<<ntiles, TILE_THREADS, 0, st>>>(px, g.w, g.npix, g.tpp, g.cap, L.tile_hist, L.e_grp, L.gidx);
                else if (quads) k_scatter4<int16_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.cap, L.tile_hist, L.e_grp, L.gidx);
                else if (gray) k_scatter<uint8_t><<<ntiles, TILE_THREADS, 0, st>>>(px, g.w, g.npix, g.tpp, g.cap, L.tile_hist, L.e_grp, L.gidx);
                else k_scatter<int16_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.cap, L.tile_hist, L.e_grp, L.gidx);
                s.launched();
            }
            const unsigned ngroups = (unsigned)(np * g.gpp);
            {
                StageScope s(ctx, ST_PREFIX);
                FELICS_CUDA_TRY(cudaMemsetAsync(L.grp_tot, 0, ((size_t)ngroups * 8 + 8) * sizeof(uint32_t), st));
                if (L.sp) k_prefix<true><<<ngroups, PREFIX_THREADS, 0, st>>>(L.e_grp, g.cap, g.gpp, L.plane_used, L.fine, L.blk_rec, L.grp_tot);
                else k_prefix<false><<<ngroups, PREFIX_THREADS, 0, st>>>(L.e_grp, g.cap, g.gpp, L.plane_used, L.fine, L.blk_rec, L.grp_tot);
                s.launched();
            }
            {
                StageScope s(ctx, ST_GRPSCAN);
                const unsigned nsuper = (ngroups + 1023) / 1024;
                k_grpscan1<<<nsuper, 1024, 0, st>>>(L.grp_tot, ngroups, L.super_tot);
                k_grpscan2<<<1, 1024, 0, st>>>(L.super_tot, nsuper);
                const size_t nquarters = (size_t)ngroups * 32 * 4;
                k_blkfinal<<<(unsigned)((nquarters + 255) / 256), 256, 0, st>>>((uint4 *)L.blk_rec, L.grp_tot, L.super_tot, L.plane_used, g.gpp, nquarters);
                s.launched(3);
            }
            FELICS_CUDA_TRY(cudaMemsetAsync(L.sp_resolved, 0, np * NBIN, st));
            // With the speculative walk on, the serial walker starts at the same time on a second stream: it walks
            // every chain and drops a chain as soon as the speculation has resolved it (both write identical
            // records), so the speculation's latency is hidden behind the chains that must be walked serially.
            const bool overlap = L.sp && !ctx->no_spec && !ctx->no_overlap;
            cudaStream_t wst = st;
            const bool hops = L.sp && !ctx->no_spec && !ctx->no_hop;   // segment hops over rejected chains (hop_walk.cuh)
            if (hops) {
                FELICS_CUDA_TRY(cudaMemsetAsync(L.hop.log, 0, ((size_t)L.spsz.max_seg * 8 + 8) * sizeof(uint32_t), st));
                FELICS_CUDA_TRY(cudaMemsetAsync(L.sp_pc2desc, 0xFF, np * NBIN * sizeof(uint32_t), st));
            }
            if (overlap) {
                if (!ctx->side) {
                    FELICS_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
                    FELICS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
                    FELICS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
                }
                wst = ctx->side;
                FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_fork, st));
                FELICS_CUDA_TRY(cudaStreamWaitEvent(wst, ctx->ev_fork, 0));
            }
            auto launch_walk = [&]() -> int {
                StageScope s(ctx, ST_WALK, wst);
                WalkArgs wa;
                wa.resolved = L.sp_resolved;
                wa.hop = L.hop; wa.desc = L.sp_desc; wa.chain_fail = L.sp_chain_fail;
                if (!hops) wa.hop.ready = nullptr;
                wa.hop_any = hops && ctx->early_hops;
                wa.fine = L.fine; wa.e_grp = L.e_grp; wa.blk_rec4 = (const uint4 *)L.blk_rec;
                wa.chain_count = L.chain_count; wa.chain_base = L.chain_base; wa.live = L.live;
                wa.counters = L.counters; wa.ep_rec = (uint4 *)L.ep_rec; wa.blk_epoch = L.blk_epoch;
                wa.cap = g.cap; wa.epcap = g.epcap;
                // beside the speculative kernels every walker warp gets a whole SM (its shared-memory request leaves no room
                // for other blocks): the walk is a latency chain, co-resident blocks would steal its issue slots
                if (!ctx->walk_attr_done) {
                    FELICS_CUDA_TRY(cudaFuncSetAttribute(k_walk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    ctx->walk_attr_done = true;
                }
                if (L.sp) {
                    // ctx->walk_per_sm (FELICS_B200_WALK_PER_SM, experiment): 2 or 3 walkers per SM beside the speculation
                    const unsigned per_sm = overlap ? std::min(3u, std::max(1u, ctx->walk_per_sm)) : 3u;
                    const size_t walk_smem = overlap ? std::max(sizeof(WalkSmem<true>), (size_t)(200 / per_sm) * 1024) : sizeof(WalkSmem<true>);
                    unsigned blocks = (unsigned)std::min<size_t>(np * NBIN, 148 * per_sm);
                    k_walk<true><<<blocks, 32, walk_smem, wst>>>(wa);
                } else {
                    unsigned blocks = (unsigned)std::min<size_t>(np * NBIN, 148 * 12);
                    k_walk<false><<<blocks, 32, sizeof(WalkSmem<false>), wst>>>(wa);
                }
                s.launched();
                return FELICS_OK;
            };
            if (overlap) {
                int wrc = launch_walk();
                if (wrc) return wrc;
                FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_join, wst));
            }
            if (L.sp && !ctx->no_spec) {
                StageScope s(ctx, ST_SPEC);
                SpArgs sa;
                sa.chain_count = L.chain_count; sa.chain_base = L.chain_base; sa.fine = L.fine; sa.blk_rec4 = (const uint4 *)L.blk_rec;
                sa.desc = L.sp_desc; sa.seg_desc = L.sp_seg_desc; sa.grp_desc = L.sp_grp_desc; sa.eb_desc = L.sp_eb_desc; sa.counts = L.sp_counts;
                sa.map = L.sp_map; sa.pre = L.sp_pre; sa.gmap = L.sp_gmap; sa.grp_xin = L.sp_grp_xin; sa.grp_nbefore = L.sp_grp_nbefore;
                sa.chain_n = L.sp_chain_n; sa.chain_fail = L.sp_chain_fail; sa.ablk = L.sp_ablk; sa.trow = L.sp_trow;
                sa.ep_rec = (uint4 *)L.ep_rec; sa.blk_epoch = L.blk_epoch; sa.resolved = L.sp_resolved; sa.dbg = L.counters;
                sa.np = (uint32_t)np; sa.cap = g.cap; sa.epcap = g.epcap; sa.sz = L.spsz; sa.pc2desc = hops ? L.sp_pc2desc : nullptr;
                if (!ctx->sp_attr_done) {
                    FELICS_CUDA_TRY(cudaFuncSetAttribute(k_sp_maps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSegSmem)));
                    ctx->sp_attr_done = true;
                }
                const SpSizes &z = L.spsz;
                k_sp_plan<<<1, NBIN, 0, st>>>(sa);
                const bool early = hops && ctx->early_hops;   // experiment: hop tables of every planned chain first
                if (hops && !ctx->hop_attr_done) {
                    FELICS_CUDA_TRY(cudaFuncSetAttribute(k_hop_maps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HopSmem)));
                    ctx->hop_attr_done = true;
                }
                if (early) {
                    sa.hop_all = 1;
                    k_hop_maps<<<z.max_seg, 1024, sizeof(HopSmem), st>>>(sa, L.hop);
                    k_hop_ready<<<1, 1, 0, st>>>(L.hop);
                    s.launched(2);
                }
                k_sp_maps<<<z.max_seg, 1024, sizeof(SpSegSmem), st>>>(sa);
                k_sp_compose<<<z.max_grp, 1024, 0, st>>>(sa);
                k_sp_scan<<<(z.max_desc + 63) / 64, 64, 0, st>>>(sa);
                k_sp_emit<<<(z.max_seg + 3) / 4, 128, 0, st>>>(sa);
                k_sp_blocksum<<<(z.max_eb + 3) / 4, 128, 0, st>>>(sa);
                k_sp_finish<false><<<(z.max_eb + 3) / 4, 128, 0, st>>>(sa);
                k_sp_finish<true><<<(z.max_eb + 3) / 4, 128, 0, st>>>(sa);
                k_sp_resolve<<<(z.max_desc + 63) / 64, 64, 0, st>>>(sa);
                s.launched(9);
                if (hops && !early) {
                    // tables for segment hops over the chains that failed the verification; the walker (already running
                    // on the other stream) starts using them when the ready flag appears
                    k_hop_maps<<<z.max_seg, 1024, sizeof(HopSmem), st>>>(sa, L.hop);
                    k_hop_ready<<<1, 1, 0, st>>>(L.hop);
                    s.launched(2);
                }
            }
            if (overlap) {
                FELICS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_join, 0));
            } else {
                int wrc = launch_walk();
                if (wrc) return wrc;
            }
            {
                if (hops) {
                    StageScope s(ctx, ST_SPEC);
                    SpArgs sa;
                    sa.fine = L.fine; sa.blk_rec4 = (const uint4 *)L.blk_rec; sa.desc = L.sp_desc; sa.seg_desc = L.sp_seg_desc; sa.counts = L.sp_counts;
                    sa.trow = L.sp_trow; sa.ep_rec = (uint4 *)L.ep_rec; sa.blk_epoch = L.blk_epoch;
                    k_hop_emit<<<(L.spsz.max_seg + 3) / 4, 128, 0, st>>>(sa, L.hop);
                    k_hop_finish<<<(L.spsz.max_seg + 3) / 4, 128, 0, st>>>(sa, L.hop);
                    s.launched(2);
                }
            }
            {
                StageScope s(ctx, ST_KFILL);
                size_t total = np * (size_t)g.cap;
                const unsigned kblocks = (unsigned)((total / 8 + 255) / 256);
                k_kfill<<<kblocks, 256, 0, st>>>(L.e_grp, (const uint4 *)L.blk_rec, L.blk_epoch, (const uint4 *)L.ep_rec, L.plane_used, g.cap, (uint32_t)np, L.k_grp);
                s.launched();
            }
            {
                StageScope s(ctx, ST_CODE);
                if (quads && gray) k_code4<uint8_t><<<ntiles, TILE_THREADS, 0, st>>>(px, g.w, g.npix, g.tpp, g.cap, L.gidx, L.k_grp, L.rec, L.tile_bits);
                else if (quads) k_code4<int16_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.cap, L.gidx, L.k_grp, L.rec, L.tile_bits);
                else if (gray) k_code<uint8_t><<<ntiles, TILE_THREADS, 0, st>>>(px, g.w, g.npix, g.tpp, g.cap, L.gidx, L.k_grp, L.rec, L.tile_bits);
                else k_code<int16_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.cap, L.gidx, L.k_grp, L.rec, L.tile_bits);
                s.launched();
            }
        } else if (ntiles) {
            // 1 or 2 pixels: nothing but the raw words
            FELICS_CUDA_TRY(cudaMemsetAsync(L.rec, 0, np * (size_t)g.npix * sizeof(uint32_t), st));
            FELICS_CUDA_TRY(cudaMemsetAsync(L.tile_bits, 0, ntiles * sizeof(uint32_t), st));
            FELICS_CUDA_TRY(cudaMemsetAsync(L.counters, 0, 8 * sizeof(uint32_t), st));
        } else {
            FELICS_CUDA_TRY(cudaMemsetAsync(L.counters, 0, 8 * sizeof(uint32_t), st));
        }
        {
            StageScope s(ctx, ST_BITSCAN);
            k_planebits<<<(unsigned)np, 1024, 0, st>>>(L.tile_bits, g.tpp, L.tile_off, L.plane_bits);
            k_imgscan<<<1, 1024, 0, st>>>(L.plane_bits, (uint32_t)ni, g.nch, L.img_off);
            s.launched(2);
        }
        // read back the image offsets (exact output size) and the error flags
        rc = ensure_buffer(ctx, &ctx->pinned, &ctx->pinned_cap, (ni + 1 + 8) * sizeof(uint64_t), true);
        if (rc) return rc;
        uint64_t *h_off = (uint64_t *)ctx->pinned;
        uint32_t *h_cnt = (uint32_t *)(h_off + ni + 1);
        FELICS_CUDA_TRY(cudaMemcpyAsync(h_off, L.img_off, (ni + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        FELICS_CUDA_TRY(cudaMemcpyAsync(h_cnt, L.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        FELICS_CUDA_TRY(cudaStreamSynchronize(st));
        for (int i = 0; i < 8; i++) ctx->dbg_counters[i] = h_cnt[i];
        if (h_cnt[2]) {
            drain();
            set_error("internal: epoch record capacity exceeded (flags %u)", h_cnt[2]);
            return FELICS_ERR_CUDA;
        }
        const uint64_t sub_total = h_off[ni];
        for (size_t i = 0; i < ni; i++) offsets_host[first + i + 1] = arena_off + h_off[i + 1];
        if (arena_off + sub_total > usable_cap) {
            // keep sizing: the caller learns the total it needs
            arena_off += sub_total;
            for (size_t f2 = first + ni; f2 < n; f2++) offsets_host[f2 + 1] = arena_off;  // lower bound only
            offsets_host[n] = arena_off;
            drain();
            set_error("output capacity %zu too small (need at least %llu)", arena_cap, (unsigned long long)arena_off);
            return FELICS_ERR_BUFFER_TOO_SMALL;
        }
        uint8_t *target = d_arena;
        uint64_t target_off = arena_off;
        if (!d_arena) {
            rc = ensure_buffer(ctx, &ctx->stage_out[slot], &ctx->stage_out_cap[slot], sub_total + 16);
            if (rc) return rc;
            target = (uint8_t *)ctx->stage_out[slot];
            target_off = 0;
            FELICS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_out[slot], 0));   // the copy-out that last used this slot
        }
        {
            StageScope s(ctx, ST_PACK);
            FELICS_CUDA_TRY(cudaMemsetAsync(target + target_off, 0, sub_total, st));
            PackArgs pa;
            pa.rec = L.rec; pa.tile_bits = L.tile_bits; pa.tile_off = L.tile_off; pa.plane_bits = L.plane_bits; pa.img_off = L.img_off;
            pa.arena = (uint32_t *)target; pa.arena_byte0 = target_off; pa.npix = g.npix; pa.tpp = g.tpp; pa.nch = g.nch;
            if (ntiles && g.npix > 2) {
                k_pack<<<ntiles, TILE_THREADS, 0, st>>>(pa);
                s.launched();
            }
            if (gray) k_heads<uint8_t><<<(unsigned)((np + 127) / 128), 128, 0, st>>>(pa, px, (uint32_t)np, g.w, g.h, hdr.color_type, hdr.pixel_depth);
            else k_heads<int16_t><<<(unsigned)((np + 127) / 128), 128, 0, st>>>(pa, L.planes, (uint32_t)np, g.w, g.h, hdr.color_type, hdr.pixel_depth);
            s.launched();
        }
        if (piped) FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_done[slot], st));   // the staged pixels are free again
        if (!d_arena) {
            FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_pack[slot], st));
            FELICS_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_pack[slot], 0));
            FELICS_CUDA_TRY(cudaMemcpyAsync(h_arena + arena_off, target, sub_total, cudaMemcpyDeviceToHost, ctx->copy_out));
            FELICS_CUDA_TRY(cudaEventRecord(ctx->ev_out[slot], ctx->copy_out));
        }
        if (n == 1 && g.npix > 2) {
            LastEncode &le = ctx->last;
            le.w = g.w; le.h = g.h; le.npix = g.npix; le.nch = g.nch; le.tpp = g.tpp; le.cap = g.cap;
            le.planes = gray ? (const void *)px : (const void *)L.planes;
            le.planes_u8 = gray;
            le.tile_base = L.tile_hist; le.chain_count = L.chain_count; le.chain_base = L.chain_base; le.blk_rec = L.blk_rec;
            le.blk_epoch = L.blk_epoch; le.ep_rec = L.ep_rec; le.e_grp = L.e_grp; le.tile_off = L.tile_off; le.plane_bits = L.plane_bits;
            le.fel_bytes = sub_total;
            le.valid = true;
        }
        ctx->dbg_rec = L.rec;
        ctx->dbg_rec_count = np * (size_t)g.npix;
        arena_off += sub_total;
    }
    FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    if (piped) {
        FELICS_CUDA_TRY(cudaStreamSynchronize(ctx->copy_out));
        // later work on the context's stream is ordered after the copies as well
        FELICS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_out[0], 0));
        FELICS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_out[1], 0));
    }
    FELICS_CUDA_TRY(cudaGetLastError());
    return profile_collect(ctx);
}

}  // namespace felics
