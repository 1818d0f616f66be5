// FELICS encode pipeline for sm_100a.
//
// Replaces the reference's sequential compress_channel loop
// (/root/reference/src/compression.rs:76-148) by data-parallel passes whose
// output is byte-identical:
//
//   planes    pixels -> i16 planes (+ YCoCg-R, color_transform.rs:11-17)
//   hist      per-tile histogram of out-of-range pixels by context
//   chainscan / tilebase   counting-sort offsets: one "chain" per (plane, context)
//   scatter   stable grouping of the residuals e by context (raster order kept)
//   prefix    code cost of e under each k in {0..5} (rice_coding.rs:56-58), prefix-summed
//   walk      KEstimator with count halving (parameter_selection.rs:49-85) recast as an
//             "epoch walk": between two halvings the counters are S0 + prefix differences,
//             so only the halving positions are found sequentially (one warp per chain)
//   kfill     k = argmin (ties to the largest k) for every out-of-range pixel, in parallel
//   code      marker + phased-in / Rice code word and length per pixel
//   bitscan   exclusive scans: tile -> plane -> image bit offsets (exact output size)
//   pack      MSB-first bit packing, identical to bitstream-io's BigEndian BitWriter
#include "ctx.h"
#include "device_common.cuh"

#include <algorithm>
#include <cstring>

namespace felics {

// ------------------------------------------------------------------------------------
// planes
// ------------------------------------------------------------------------------------
__global__ void k_to_planes_gray8(const uint8_t *__restrict__ px, int16_t *__restrict__ planes, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) planes[i] = (int16_t)px[i];
}

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
__global__ void __launch_bounds__(TILE_THREADS) k_hist(const int16_t *__restrict__ planes, uint32_t w, uint32_t npix,
                                                       uint32_t tpp, uint32_t nchunks, uint32_t *__restrict__ tile_hist,
                                                       uint32_t *__restrict__ chunk_tot) {
    __shared__ uint32_t h[NBIN];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / tpp, t = bid - p * tpp;
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) h[c] = 0;
    __syncthreads();
    const int16_t *pl = planes + (size_t)p * npix;
    uint32_t start = t * TILE;
#pragma unroll 4
    for (int j = 0; j < TILE / TILE_THREADS; j++) {
        uint32_t i = start + j * TILE_THREADS + threadIdx.x;
        if (i >= 2 && i < npix) {
            PixelClass pc = classify_pixel(pl, i, w);
            if (pc.cls != 0) atomicAdd(&h[pc.delta], 1u);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) {
        uint32_t v = h[c];
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
    for (uint32_t t = t0; t < t1; t++) {
        size_t idx = ((size_t)p * tpp + t) * NBIN + c;
        uint32_t v = tile_hist[idx];
        tile_hist[idx] = run;
        run += v;
    }
}

// ------------------------------------------------------------------------------------
// scatter: stable grouping by context.  Warp w owns pixels [512w, 512w+512) of the tile
// and visits them 32 consecutive pixels at a time, so ranks follow raster order.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS) k_scatter(const int16_t *__restrict__ planes, uint32_t w, uint32_t npix,
                                                          uint32_t tpp, uint32_t cap, const uint32_t *__restrict__ tile_base,
                                                          uint16_t *__restrict__ e_grp, uint32_t *__restrict__ gidx) {
    __shared__ uint16_t wcnt[TILE_WARPS][NBIN];
    __shared__ uint32_t wbase[TILE_WARPS][NBIN];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / tpp, t = bid - p * tpp;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < TILE_WARPS * NBIN; c += TILE_THREADS) (&wcnt[0][0])[c] = 0;
    __syncthreads();
    const int16_t *pl = planes + (size_t)p * npix;
    uint32_t wstart = t * TILE + wid * WARP_PIX;
    uint32_t info[WARP_ITERS];  // rank(13) | delta(9) << 13 | oor << 22
    uint16_t ev[WARP_ITERS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        uint32_t i = wstart + it * 32 + lane;
        bool oor = false;
        int delta = 0, val = 0;
        if (i >= 2 && i < npix) {
            PixelClass pc = classify_pixel(pl, i, w);
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
// blk_local[blk][k]: exclusive prefix of the block totals inside the group.
// grp_tot[grp][k]: totals of the group (scanned by k_grpscan).
// All prefix sums are mod 2^32 over the whole grouped array; only differences inside a
// chain are ever used, so chain boundaries need no segmentation.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GROUP) k_prefix(const uint16_t *__restrict__ e_grp, uint32_t cap, uint32_t gpp,
                                                  const uint32_t *__restrict__ plane_used, uint4 *__restrict__ fine,
                                                  uint32_t *__restrict__ blk_local, uint32_t *__restrict__ grp_tot) {
    __shared__ uint32_t tot[32][NK];
    uint32_t grp = blockIdx.x;
    uint32_t p = grp / gpp;
    uint32_t off = (grp - p * gpp) * GROUP;  // plane-relative element offset of this group
    if (off >= plane_used[p]) return;        // grp_tot stays 0 (memset)
    size_t g = (size_t)p * cap + off + threadIdx.x;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t e = e_grp[g];
    uint32_t u01 = 0, u23 = 0, u45 = 0;
    if (e != PAD_E) {
        u01 = (e + 1u) | (((e >> 1) + 2u) << 16);
        u23 = ((e >> 2) + 3u) | (((e >> 3) + 4u) << 16);
        u45 = ((e >> 4) + 5u) | (((e >> 5) + 6u) << 16);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(0xffffffffu, u01, o);
        uint32_t b = __shfl_up_sync(0xffffffffu, u23, o);
        uint32_t c = __shfl_up_sync(0xffffffffu, u45, o);
        if (lane >= (uint32_t)o) { u01 += a; u23 += b; u45 += c; }
    }
    fine[g] = make_uint4(u01, u23, u45, e);
    if (lane == 31) {
        tot[wid][0] = u01 & 0xffffu; tot[wid][1] = u01 >> 16;
        tot[wid][2] = u23 & 0xffffu; tot[wid][3] = u23 >> 16;
        tot[wid][4] = u45 & 0xffffu; tot[wid][5] = u45 >> 16;
    }
    __syncthreads();
    if (wid == 0) {
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
            blk_local[blk * 8 + k] = s - v;
            if (lane == 31) grp_tot[(size_t)grp * 8 + k] = s;
        }
    }
}

// grpscan: single block, exclusive scan (mod 2^32) of the group totals, in place.
__global__ void __launch_bounds__(1024) k_grpscan(uint32_t *__restrict__ grp_tot, uint32_t ngroups) {
    __shared__ uint32_t wsum[32][NK];
    __shared__ uint32_t carry_s[NK];
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < NK) carry_s[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ngroups; base += 1024) {
        uint32_t gidx = base + threadIdx.x;
        uint32_t v[NK], s[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) {
            v[k] = gidx < ngroups ? grp_tot[(size_t)gidx * 8 + k] : 0;
            s[k] = v[k];
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < NK; k++) {
                uint32_t n = __shfl_up_sync(0xffffffffu, s[k], o);
                if (lane >= (uint32_t)o) s[k] += n;
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
                    uint32_t n = __shfl_up_sync(0xffffffffu, y, o);
                    if (lane >= (uint32_t)o) y += n;
                }
                wsum[lane][k] = y - x;  // exclusive over warps
            }
        }
        __syncthreads();
        uint32_t chunk_tot_k[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) {
            uint32_t excl = carry_s[k] + wsum[wid][k] + s[k] - v[k];
            if (gidx < ngroups) grp_tot[(size_t)gidx * 8 + k] = excl;
            chunk_tot_k[k] = excl + v[k];  // inclusive at this thread
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
#pragma unroll
            for (int k = 0; k < NK; k++) carry_s[k] = chunk_tot_k[k];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// walk: one warp per chain.  State S[6] (u32) is replicated across lanes.
// base[k] = S[k] - T(cur)[k] (mod 2^32), T = global exclusive cost prefix, cur = first
// element of the current epoch; the counters before element t are base + T(t).
// Epoch i is recorded as (first element, base) for k_kfill.
// ------------------------------------------------------------------------------------
struct WalkArgs {
    const uint4 *fine;
    const uint32_t *blk_local;
    const uint32_t *grp_pre;
    const uint32_t *chain_count;
    const uint32_t *chain_base;
    const uint32_t *live;
    uint32_t *counters;     // [0] live count, [1] queue head, [2] error flags
    uint32_t *ep_start;     // global element index of the first element of each epoch
    uint32_t *ep_base;      // [epoch][8]
    uint32_t *blk_epoch;    // epoch in effect at the start of each 32-block
    uint32_t cap;           // grouped elements per plane
    uint32_t epcap;         // epoch records per plane
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32) k_walk(WalkArgs a) {
    __shared__ uint4 sfine[2][GROUP];
    const uint32_t lane = threadIdx.x;
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(&a.counters[1], 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= a.counters[0]) break;
        const uint32_t pc = a.live[qi];
        const uint32_t p = pc / NBIN, c = pc % NBIN;
        const uint32_t count = a.chain_count[pc];
        const uint32_t cbase = a.chain_base[pc];
        const uint32_t nblk = (count + 31u) >> 5;
        const size_t gbase = (size_t)p * a.cap + cbase;      // global element index of the chain start
        const size_t blk0 = gbase >> 5;                      // global 32-block index
        const uint32_t ep0 = p * a.epcap + cbase / 8 + 16 * c;
        const uint32_t ep_room = nblk * 4 + 16;              // records available to this chain

        uint32_t base[NK];
#pragma unroll
        for (int k = 0; k < NK; k++)
            base[k] = 0u - (a.blk_local[blk0 * 8 + k] + a.grp_pre[(blk0 >> 5) * 8 + k]);
        if (lane < NK) a.ep_base[(size_t)ep0 * 8 + lane] = base[lane];
        if (lane == NK) a.ep_start[ep0] = (uint32_t)gbase;
        uint32_t nep = 1;
        uint32_t cur = 0;  // chain-relative index of the first element of the current epoch
        bool overflow = false;

        // prefetch window 0
        {
            const uint4 *src = a.fine + gbase;
            uint32_t nel = min(nblk * 32u, (uint32_t)GROUP);
            for (uint32_t j = lane; j < nel; j += 32) cp_async16(&sfine[0][j], src + j);
            cp_async_commit();
        }
        uint32_t wi = 0;
        for (uint32_t wb = 0; wb < nblk; wb += 32, wi++) {
            const uint32_t buf = wi & 1;
            if (wb + 32 < nblk) {
                const uint4 *src = a.fine + gbase + (size_t)(wb + 32) * 32;
                uint32_t nel = min((nblk - (wb + 32)) * 32u, (uint32_t)GROUP);
                for (uint32_t j = lane; j < nel; j += 32) cp_async16(&sfine[buf ^ 1][j], src + j);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();

            const uint32_t blk = wb + lane;
            const bool valid = blk < nblk;
            uint32_t cps[NK], cpe[NK];
            {
                size_t gb = blk0 + (valid ? blk : 0);
                uint4 last = sfine[buf][(valid ? lane : 0) * 32 + 31];
                uint32_t bt[NK] = {last.x & 0xffffu, last.x >> 16, last.y & 0xffffu, last.y >> 16, last.z & 0xffffu, last.z >> 16};
#pragma unroll
                for (int k = 0; k < NK; k++) {
                    cps[k] = a.blk_local[gb * 8 + k] + a.grp_pre[(gb >> 5) * 8 + k];
                    cpe[k] = cps[k] + bt[k];
                }
            }
            uint32_t my_epoch = nep - 1;  // epoch in effect at the start of my block
            for (;;) {
                // coarse: first block (at or after the one holding `cur`) at whose END every counter exceeds 1024
                bool ok = valid;
#pragma unroll
                for (int k = 0; k < NK; k++) ok = ok && ((int32_t)(base[k] + cpe[k]) > (int32_t)HALVE_AT);
                uint32_t m = __ballot_sync(0xffffffffu, ok);
                int curblk = (int)(cur >> 5) - (int)wb;  // < 0: the epoch began in an earlier window
                if (curblk >= 32) m = 0;
                else if (curblk > 0) m &= ~((1u << curblk) - 1u);
                if (!m) break;
                const int B = __ffs(m) - 1;
                // fine: first element of block B at which every counter exceeds 1024
                uint4 f = sfine[buf][B * 32 + lane];
                uint32_t U[NK] = {f.x & 0xffffu, f.x >> 16, f.y & 0xffffu, f.y >> 16, f.z & 0xffffu, f.z >> 16};
                uint32_t T[NK], v[NK];
                bool okf = true;
#pragma unroll
                for (int k = 0; k < NK; k++) {
                    T[k] = __shfl_sync(0xffffffffu, cps[k], B) + U[k];  // prefix including this element
                    v[k] = base[k] + T[k];
                    okf = okf && ((int32_t)v[k] > (int32_t)HALVE_AT);
                }
                uint32_t fm = __ballot_sync(0xffffffffu, okf);
                if (B == curblk) fm &= ~((1u << (cur & 31u)) - 1u);
                if (fm == 0) { overflow = true; atomicOr(&a.counters[2], 2u); break; }  // cannot happen: the block's last element satisfies the test
                const int h = __ffs(fm) - 1;
#pragma unroll
                for (int k = 0; k < NK; k++) {
                    uint32_t s = __shfl_sync(0xffffffffu, v[k], h) >> 1;   // parameter_selection.rs:58-63
                    uint32_t th = __shfl_sync(0xffffffffu, T[k], h);
                    base[k] = s - th;
                }
                cur = (wb + (uint32_t)B) * 32u + (uint32_t)h + 1u;
                if (nep + 1 < ep_room) {
                    if (lane < NK) a.ep_base[(size_t)(ep0 + nep) * 8 + lane] = base[lane];
                    if (lane == NK) a.ep_start[ep0 + nep] = (uint32_t)(gbase + cur);
                } else {
                    overflow = true;
                }
                nep++;
                if ((int)lane > B) my_epoch++;
                if (cur >= count) break;
            }
            if (valid) a.blk_epoch[blk0 + blk] = ep0 + my_epoch;
            __syncwarp();
        }
        if (lane == 0) {
            if (overflow) atomicOr(&a.counters[2], 1u);
            else a.ep_start[ep0 + nep] = 0xFFFFFFFFu;  // sentinel
        }
    }
}

// kfill: one thread per grouped element.
__global__ void __launch_bounds__(256) k_kfill(const uint4 *__restrict__ fine, const uint32_t *__restrict__ blk_local,
                                               const uint32_t *__restrict__ grp_pre, const uint32_t *__restrict__ blk_epoch,
                                               const uint32_t *__restrict__ ep_start, const uint32_t *__restrict__ ep_base,
                                               const uint32_t *__restrict__ plane_used, uint32_t cap, uint32_t np,
                                               uint8_t *__restrict__ k_grp) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = (uint32_t)(g / cap);
    if (p >= np) return;
    uint32_t off = (uint32_t)(g - (size_t)p * cap);
    if (off >= plane_used[p]) return;
    uint4 f = fine[g];
    if (f.w == PAD_E) return;
    size_t blk = g >> 5;
    uint32_t ep = blk_epoch[blk];
    while (ep_start[ep + 1] <= (uint32_t)g) ep++;
    uint32_t e = f.w;
    uint32_t U[NK] = {f.x & 0xffffu, f.x >> 16, f.y & 0xffffu, f.y >> 16, f.z & 0xffffu, f.z >> 16};
    uint32_t v[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        uint32_t d = (e >> k) + 1u + (uint32_t)k;
        v[k] = ep_base[(size_t)ep * 8 + k] + blk_local[blk * 8 + k] + grp_pre[(blk >> 5) * 8 + k] + U[k] - d;
    }
    k_grp[g] = (uint8_t)argmin_last(v);
}

// ------------------------------------------------------------------------------------
// code: marker + code word per pixel (compression.rs:130-145), bits per tile
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS) k_code(const int16_t *__restrict__ planes, uint32_t w, uint32_t npix, uint32_t tpp,
                                                       uint32_t cap, const uint32_t *__restrict__ gidx,
                                                       const uint8_t *__restrict__ k_grp, uint32_t *__restrict__ rec,
                                                       uint32_t *__restrict__ tile_bits) {
    __shared__ uint32_t wsum[TILE_WARPS];
    uint32_t bid = blockIdx.x;
    uint32_t p = bid / tpp, t = bid - p * tpp;
    const int16_t *pl = planes + (size_t)p * npix;
    uint32_t start = t * TILE;
    uint32_t bits = 0;
#pragma unroll 4
    for (int j = 0; j < TILE / TILE_THREADS; j++) {
        uint32_t i = start + j * TILE_THREADS + threadIdx.x;
        if (i >= npix) break;
        uint32_t r = 0;
        if (i >= 2) {
            PixelClass pc = classify_pixel(pl, i, w);
            if (pc.cls == 0) {
                int len;
                uint32_t code = phase_in_code((uint32_t)pc.delta + 1u, (uint32_t)pc.val, len);
                // '1' marker then the phased-in code
                r = ((uint32_t)(len + 1) << 22) | (1u << len) | code;
            } else {
                uint32_t g = gidx[(size_t)p * npix + i];
                uint32_t k = k_grp[(size_t)p * cap + g];
                uint32_t e = (uint32_t)pc.val;
                uint32_t q = e >> k, rem = e & ((1u << k) - 1u);
                uint32_t above = pc.cls == 1 ? 1u : 0u;
                uint32_t len = 2u + q + 1u + k;
                if (len <= (uint32_t)REC_SHORT_MAX) {
                    // '0', above, q ones, '0', k remainder bits
                    uint32_t code = (above << (q + 1u + k)) | (((1u << q) - 1u) << (k + 1u)) | rem;
                    r = (len << 22) | code;
                } else {
                    r = (len << 22) | (above << 17) | (k << 14) | (rem << 9) | q;
                }
            }
        }
        rec[(size_t)p * npix + i] = r;
        bits += rec_len(r);
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

    const uint32_t *rc = a.rec + (size_t)p * a.npix;
    uint32_t wstart = t * TILE + wid * WARP_PIX;
    uint32_t r[WARP_ITERS];
    uint32_t mysum = 0;
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        uint32_t i = wstart + it * 32 + lane;
        r[it] = i < a.npix ? rc[i] : 0u;
        mysum += rec_len(r[it]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mysum += __shfl_xor_sync(0xffffffffu, mysum, o);
    if (lane == 0) wtot[wid] = mysum;
    __syncthreads();
    uint32_t run = sh0;  // bit offset relative to word0
    for (uint32_t q = 0; q < wid; q++) run += wtot[q];

    auto put_s = [&](uint64_t off, uint32_t val, int n) { put_bits_smem(buf, (uint32_t)off, val, n); };
    auto put_g = [&](uint64_t off, uint32_t val, int n) { put_bits_global(a.arena, (word0 << 5) + off, val, n); };
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        uint32_t len = rec_len(r[it]);
        uint32_t inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += n;
        }
        uint32_t off = run + inc - len;
        if (in_smem) emit_record(r[it], off, put_s);
        else emit_record(r[it], off, put_g);
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
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
__global__ void k_heads(PackArgs a, const int16_t *__restrict__ planes, uint32_t np, uint32_t width, uint32_t height,
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
    uint32_t *blk_local, *grp_tot, *ep_start, *ep_base, *blk_epoch;
    uint8_t *k_grp;
    uint32_t *rec, *tile_bits;
    uint64_t *tile_off, *plane_bits, *img_off;
    size_t bytes;
};

struct Geom {
    uint32_t w, h, npix, nch, tpp, nchunks, cap, gpp, epcap;
};

Layout carve(uint8_t *base, const Geom &g, size_t ni) {
    Layout L;
    Carver c{base};
    size_t np = ni * g.nch;
    L.planes = c.take<int16_t>(np * g.npix + 8);
    L.tile_hist = c.take<uint32_t>(np * g.tpp * NBIN);
    L.chunk_tot = c.take<uint32_t>(np * g.nchunks * NBIN);
    L.chain_count = c.take<uint32_t>(np * NBIN);
    L.chain_base = c.take<uint32_t>(np * NBIN);
    L.plane_used = c.take<uint32_t>(np);
    L.live = c.take<uint32_t>(np * NBIN);
    L.counters = c.take<uint32_t>(8);
    L.e_grp = c.take<uint16_t>(np * g.cap);
    L.gidx = c.take<uint32_t>(np * g.npix + 8);
    L.fine = c.take<uint4>(np * g.cap);
    L.blk_local = c.take<uint32_t>(np * (g.cap / 32) * 8);
    L.grp_tot = c.take<uint32_t>(np * g.gpp * 8 + 8);
    L.ep_start = c.take<uint32_t>(np * g.epcap + 8);
    L.ep_base = c.take<uint32_t>((np * g.epcap + 8) * 8);
    L.blk_epoch = c.take<uint32_t>(np * (g.cap / 32));
    L.k_grp = c.take<uint8_t>(np * g.cap);
    L.rec = c.take<uint32_t>(np * g.npix + 8);
    L.tile_bits = c.take<uint32_t>(np * g.tpp + 8);
    L.tile_off = c.take<uint64_t>(np * g.tpp + 8);
    L.plane_bits = c.take<uint64_t>(np + 8);
    L.img_off = c.take<uint64_t>(ni + 8);
    L.bytes = align_up(c.off, 256);
    return L;
}

}  // namespace

// Exactly one of d_arena (device memory, 4-byte aligned) / h_arena (host memory) is non-null.
// With h_arena every sub-batch is packed into the context's staging buffer and copied out.
int encode_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr, uint8_t *d_arena,
                        uint8_t *h_arena, size_t arena_cap, uint64_t *offsets_host) {
    if (hdr.pixel_depth != 0) {
        set_error("16-bit samples are not built yet (traits.rs:35-43 is a 'next' row)");
        return FELICS_ERR_UNSUPPORTED;
    }
    if (d_arena && ((uintptr_t)d_arena & 3) != 0) {
        set_error("device arena must be 4-byte aligned");
        return FELICS_ERR_INVALID_ARGUMENT;
    }
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
    size_t budget = (size_t)12 << 30;
    size_t sub = std::max<size_t>(1, std::min<size_t>(n, budget / std::max<size_t>(per_image, 1)));
    while (sub > 1 && (uint64_t)sub * g.nch * g.cap >= 0xffff0000ull) sub--;
    if ((uint64_t)g.nch * g.cap >= 0xffff0000ull) {
        set_error("image too large for one call");
        return FELICS_ERR_INVALID_DIMENSIONS;
    }

    uint64_t arena_off = 0;
    offsets_host[0] = 0;
    for (size_t first = 0; first < n; first += sub) {
        const size_t ni = std::min(sub, n - first);
        const size_t np = ni * g.nch;
        Layout L = carve(nullptr, g, ni);
        int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, L.bytes);
        if (rc) return rc;
        L = carve((uint8_t *)ctx->scratch, g, ni);
        const uint8_t *px = (const uint8_t *)d_pixels + first * img_pix_bytes;

        if (g.npix > 0) {
            StageScope s(ctx, ST_PLANES);
            size_t total = ni * (size_t)g.npix;
            unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
            if (g.nch == 1) k_to_planes_gray8<<<blocks, 256, 0, st>>>(px, L.planes, total);
            else k_to_planes_rgb8<<<blocks, 256, 0, st>>>(px, L.planes, g.npix, total);
            s.launched();
        }
        const unsigned ntiles = (unsigned)(np * g.tpp);
        if (g.npix > 2) {
            {
                StageScope s(ctx, ST_HIST);
                FELICS_CUDA_TRY(cudaMemsetAsync(L.chunk_tot, 0, np * g.nchunks * NBIN * sizeof(uint32_t), st));
                FELICS_CUDA_TRY(cudaMemsetAsync(L.counters, 0, 8 * sizeof(uint32_t), st));
                k_hist<<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.nchunks, L.tile_hist, L.chunk_tot);
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
                k_scatter<<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.cap, L.tile_hist, L.e_grp, L.gidx);
                s.launched();
            }
            const unsigned ngroups = (unsigned)(np * g.gpp);
            {
                StageScope s(ctx, ST_PREFIX);
                FELICS_CUDA_TRY(cudaMemsetAsync(L.grp_tot, 0, ((size_t)ngroups * 8 + 8) * sizeof(uint32_t), st));
                k_prefix<<<ngroups, GROUP, 0, st>>>(L.e_grp, g.cap, g.gpp, L.plane_used, L.fine, L.blk_local, L.grp_tot);
                s.launched();
            }
            {
                StageScope s(ctx, ST_GRPSCAN);
                k_grpscan<<<1, 1024, 0, st>>>(L.grp_tot, ngroups);
                s.launched();
            }
            {
                StageScope s(ctx, ST_WALK);
                WalkArgs wa;
                wa.fine = L.fine; wa.blk_local = L.blk_local; wa.grp_pre = L.grp_tot;
                wa.chain_count = L.chain_count; wa.chain_base = L.chain_base; wa.live = L.live;
                wa.counters = L.counters; wa.ep_start = L.ep_start; wa.ep_base = L.ep_base; wa.blk_epoch = L.blk_epoch;
                wa.cap = g.cap; wa.epcap = g.epcap;
                unsigned blocks = (unsigned)std::min<size_t>(np * NBIN, 148 * 7);
                k_walk<<<blocks, 32, 0, st>>>(wa);
                s.launched();
            }
            {
                StageScope s(ctx, ST_KFILL);
                size_t total = np * (size_t)g.cap;
                k_kfill<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L.fine, L.blk_local, L.grp_tot, L.blk_epoch, L.ep_start, L.ep_base,
                                                                         L.plane_used, g.cap, (uint32_t)np, L.k_grp);
                s.launched();
            }
            {
                StageScope s(ctx, ST_CODE);
                k_code<<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.cap, L.gidx, L.k_grp, L.rec, L.tile_bits);
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
        if (h_cnt[2]) {
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
            set_error("output capacity %zu too small (need at least %llu)", arena_cap, (unsigned long long)arena_off);
            return FELICS_ERR_BUFFER_TOO_SMALL;
        }
        uint8_t *target = d_arena;
        uint64_t target_off = arena_off;
        if (!d_arena) {
            rc = ensure_buffer(ctx, &ctx->staging_out, &ctx->staging_out_cap, sub_total + 16);
            if (rc) return rc;
            target = (uint8_t *)ctx->staging_out;
            target_off = 0;
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
            k_heads<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(pa, L.planes, (uint32_t)np, g.w, g.h, hdr.color_type, hdr.pixel_depth);
            s.launched();
        }
        if (!d_arena) {
            FELICS_CUDA_TRY(cudaMemcpyAsync(h_arena + arena_off, target, sub_total, cudaMemcpyDeviceToHost, st));
            FELICS_CUDA_TRY(cudaStreamSynchronize(st));   // staging_out is reused by the next sub-batch
        }
        ctx->dbg_rec = L.rec;
        ctx->dbg_rec_count = np * (size_t)g.npix;
        arena_off += sub_total;
    }
    FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    FELICS_CUDA_TRY(cudaGetLastError());
    return profile_collect(ctx);
}

}  // namespace felics
