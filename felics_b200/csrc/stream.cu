// Streaming band encoder for batches of 8-bit gray images (sm_100a): ONE thread block encodes ONE whole image.
//
// The multi-kernel pipeline of encode.cu is built for a single big image: every stage is a kernel over the whole
// plane and every intermediate (grouped residuals, grouped indices, code records ...) travels through HBM.  A batch
// of small images does not need that: the estimator rows of an image (parameter_selection.rs:24-85) are 511 x 6
// small counters, so a block can walk the image in raster order, band by band (4096 pixels), and keep everything it
// needs between the pixels and the bit stream in shared memory:
//
//   load      band + the row above it, cp.async into shared memory (double buffered)
//   classify  neighbours, context, class, residual of every pixel (misc.rs:6-24, compression.rs:118-145)
//   rank      stable grouping of the band's out-of-range pixels by context: match_any ranks inside 32 consecutive
//             pixels, per-warp counters, one prefix over the warps -> the band's "chains" (one per context)
//   walk      KEstimator for every chain of the band, starting from the counters the previous band left: a warp
//             walks a long chain 128 elements at a time (prefix sums of the six code costs, halvings found by ballot,
//             exact for any data), short chains are walked one per lane
//   code      marker + phased-in / Rice code of every pixel (compression.rs:130-145) in registers, lengths scanned
//   pack      MSB-first packing into a shared-memory window, whole words leave for the image's slot in HBM
//
// The pixels are read once and the code bits are written once; the only other HBM traffic is the copy of every
// finished stream from its slot to its final place in the arena (k_stream_compact), because an image's offset is
// known only when all images before it are.  Output is byte-identical to the reference's compress_image
// (compression.rs:255-282) for every image; images whose stream does not fit the slot are re-encoded by the
// general pipeline (encode.cu).
#include "ctx.h"
#include "device_common.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace felics {

namespace {

constexpr int SE_THREADS = 256;
constexpr int SE_WARPS = SE_THREADS / 32;
constexpr int SE_PPT = 16;                                  // pixels per thread in the code / pack phase
constexpr int SE_BAND = SE_THREADS * SE_PPT;                // 4096 pixels per band
constexpr int SE_WPIX = SE_BAND / SE_WARPS;                 // 512 consecutive pixels per warp in the rank phase
constexpr int SE_WSTEPS = SE_WPIX / 32;
constexpr int SE_INFO_WORDS = SE_BAND + SE_BAND / 16 * 4;   // one word per pixel, 4 words of padding per 16 (bank spread)
constexpr int SE_EC_BYTES = SE_BAND + NBIN * 4 + 256;       // chain starts are 4-aligned; slack for the walker's last load
constexpr int SE_OUT_WORDS = 2048;                          // bit window: 65536 bits = 16 bits per pixel of a band
constexpr uint32_t SE_LONG = 48;                            // chains this long (per band) are walked by a whole warp
constexpr uint32_t SE_HALVE_KEY = (HALVE_AT + 1u) << 3;     // key of a count that has passed 1024
constexpr uint32_t SE_NONE = 0xFFFFFFFFu;
constexpr uint32_t SE_MAX_W = 8192;                         // widest row whose predecessor fits in front of a band

// info word of a pixel:  [31:30] class (0 in range, 1 above, 2 below, 3 not coded)  [29:21] P-L | P-H-1 | L-P-1
//                        [20:12] context  [11:0] rank inside the warp's pixels of that context (rank phase)
// after the scatter an out-of-range pixel keeps [31:21] and carries its chain position in [12:0]
__device__ __forceinline__ uint32_t info_index(uint32_t p) { return p + ((p >> 4) << 2); }

struct SeSmem {
    uint32_t info[SE_INFO_WORDS];
    uint32_t state[NBIN][4];            // estimator row of a context: counts as u16 pairs (k0 | k1 << 16, k2 | k3 << 16, k4 | k5 << 16)
    uint16_t wcnt[SE_WARPS][NBIN];      // per warp and context: count, then exclusive prefix over the warps
    uint32_t cb[NBIN];                  // chain of a context in this band: base | count << 16
    uint16_t longlist[NBIN], shortlist[NBIN];
    uint32_t out[SE_OUT_WORDS + 4];
    uint8_t ec[SE_EC_BYTES];            // residuals in chain order; the walk overwrites them with k
    uint32_t wsum[SE_WARPS];
    uint32_t nlong, nshort, task, plane;
};

struct StreamArgs {
    const uint8_t *pixels;      // image p at pixels + p * npix
    uint8_t *temp;              // slot of image p at temp + p * slot_bytes
    size_t slot_bytes;          // multiple of 16
    uint32_t *sizes;            // out: .fel bytes of image p
    uint32_t *flags;            // out: 1 = the stream did not fit the slot (size is still exact)
    uint32_t *ticket;
    uint32_t w, h, npix, nplanes;
    uint32_t halo_cap;          // bytes in front of a band in the pixel buffer (>= w, multiple of 16)
    uint32_t vec16;             // pixels, w multiples of 16: 16-byte copies
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// pixels [start - hl, start + cnt) of the plane -> buf[halo_cap - hl ...): the band starts at buf + halo_cap
__device__ __forceinline__ void load_band(uint8_t *buf, const StreamArgs &a, const uint8_t *plane, uint32_t start, uint32_t cnt) {
    const uint32_t hl = min(a.w, start);
    const uint32_t total = hl + cnt;
    const uint8_t *src = plane + start - hl;
    uint8_t *dst = buf + a.halo_cap - hl;
    if (a.vec16) {
        for (uint32_t o = 16u * threadIdx.x; o < total; o += 16u * SE_THREADS) cp_async16(dst + o, src + o);
    } else {
        for (uint32_t o = 4u * threadIdx.x; o < total; o += 4u * SE_THREADS) cp_async4(dst + o, src + o);
    }
    cp_async_commit();
}

// compression.rs:118-145 for one pixel given its two neighbours
__device__ __forceinline__ uint32_t make_info(int p, int v1, int v2) {
    const int h = max(v1, v2), l = min(v1, v2);
    const bool below = p < l, above = p > h;
    const uint32_t cls = below ? 2u : (above ? 1u : 0u);
    const int val = below ? l - p - 1 : (above ? p - h - 1 : p - l);
    return (cls << 30) | ((uint32_t)val << 21) | ((uint32_t)(h - l) << 12);
}
// any pixel (misc.rs:6-24); pb = first pixel of the band, j = offset in the band, i = index in the plane
__device__ __forceinline__ uint32_t classify_slow(const uint8_t *pb, int j, uint32_t i, uint32_t x, uint32_t y, int w, const uint8_t *plane) {
    if (i < 2) return 3u << 30;
    const int p = pb[j];
    int v1, v2;
    if (x > 0 && y > 0) { v1 = pb[j - 1]; v2 = pb[j - w]; }
    else if (y == 0) { v1 = pb[j - 1]; v2 = pb[j - 2]; }
    else if (y >= 2) { v1 = pb[j - w]; v2 = plane[i - 2u * (uint32_t)w]; }
    else { v1 = pb[j - w]; v2 = pb[j - w + 1]; }
    return make_info(p, v1, v2);
}

// ---- estimator rows as keys: count * 8 + (5 - k).  The smallest key is the smallest count with ties going to the
// largest k (get_k's `<=` scan, parameter_selection.rs:78-83); costs are added as cost * 8.
__device__ __forceinline__ void load_state(const uint32_t *row, uint32_t (&v)[NK]) {
    const uint4 s = *reinterpret_cast<const uint4 *>(row);
    v[0] = ((s.x & 0xffffu) << 3) | 5u; v[1] = ((s.x >> 16) << 3) | 4u;
    v[2] = ((s.y & 0xffffu) << 3) | 3u; v[3] = ((s.y >> 16) << 3) | 2u;
    v[4] = ((s.z & 0xffffu) << 3) | 1u; v[5] = ((s.z >> 16) << 3);
}
__device__ __forceinline__ void store_state(uint32_t *row, const uint32_t (&v)[NK]) {
    // counts stay below 2^16 for 8-bit data: c0 <= 25 * (1024 + 510)
    *reinterpret_cast<uint4 *>(row) = make_uint4((v[0] >> 3) | ((v[1] >> 3) << 16), (v[2] >> 3) | ((v[3] >> 3) << 16), (v[4] >> 3) | ((v[5] >> 3) << 16), 0u);
}
__device__ __forceinline__ void cost_keys(uint32_t e, uint32_t (&c)[NK]) {   // rice_coding.rs:56-58 times 8
#pragma unroll
    for (int k = 0; k < NK; k++) c[k] = ((e >> k) + 1u + (uint32_t)k) << 3;
}
__device__ __forceinline__ uint32_t min6(const uint32_t (&v)[NK]) { return min(min(min(v[0], v[1]), min(v[2], v[3])), min(v[4], v[5])); }
__device__ __forceinline__ void halve_keys(uint32_t (&v)[NK]) {              // parameter_selection.rs:58-63
#pragma unroll
    for (int k = 0; k < NK; k++) v[k] = ((v[k] >> 4) << 3) | (uint32_t)(5 - k);
}

// one chain per lane (short chains): the reference's loop as written
__device__ __forceinline__ void serial_walk(SeSmem &S, uint32_t c) {
    if (c == SE_NONE) return;
    const uint32_t cbv = S.cb[c];
    const uint32_t base = cbv & 0xffffu, n = cbv >> 16;
    uint32_t st[NK];
    load_state(S.state[c], st);
    uint32_t m = min6(st);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t e = S.ec[base + i];
        S.ec[base + i] = (uint8_t)(5u - (m & 7u));
        uint32_t c6[NK];
        cost_keys(e, c6);
#pragma unroll
        for (int k = 0; k < NK; k++) st[k] += c6[k];
        m = min6(st);
        if (m >= SE_HALVE_KEY) { halve_keys(st); m = min6(st); }
    }
    store_state(S.state[c], st);
}

// one chain per warp, 128 elements per step (four consecutive elements per lane).  With T the prefix sum of the costs
// inside the step, the counters before element g are B + T(g) as long as no halving lies between; the first element
// after which all six counters have passed 1024 is found by one ballot (counters only grow inside an epoch), the
// halved counters become the new base for the elements behind it, and the search repeats.
__device__ __forceinline__ void coop_walk(SeSmem &S, uint32_t c, uint32_t lane) {
    const uint32_t cbv = S.cb[c];
    const uint32_t base = cbv & 0xffffu, n = cbv >> 16;
    uint32_t st[NK];
    load_state(S.state[c], st);
    uint32_t *ec4 = reinterpret_cast<uint32_t *>(S.ec + base);
    for (uint32_t s0 = 0; s0 < n; s0 += 128) {
        const uint32_t word = ec4[(s0 >> 2) + lane];
        const int nvalid = min(4, max(0, (int)n - (int)s0 - 4 * (int)lane));
        uint32_t P[4][NK];   // inclusive cost prefix inside my four elements
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t c6[NK];
            cost_keys((word >> (8 * j)) & 255u, c6);
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const uint32_t v = j < nvalid ? c6[k] : 0u;
                P[j][k] = j ? P[j - 1][k] + v : v;
            }
        }
        // exclusive prefix of the lane totals over the warp, two 16-bit sums per register (128 * 510 < 65536)
        uint32_t t01 = (P[3][0] >> 3) | ((P[3][1] >> 3) << 16), t23 = (P[3][2] >> 3) | ((P[3][3] >> 3) << 16), t45 = (P[3][4] >> 3) | ((P[3][5] >> 3) << 16);
        uint32_t x01 = t01, x23 = t23, x45 = t45;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, x01, o), b = __shfl_up_sync(0xffffffffu, x23, o), d = __shfl_up_sync(0xffffffffu, x45, o);
            if (lane >= (uint32_t)o) { x01 += a; x23 += b; x45 += d; }
        }
        x01 -= t01; x23 -= t23; x45 -= t45;
        const uint32_t X[NK] = {(x01 & 0xffffu) << 3, (x01 >> 16) << 3, (x23 & 0xffffu) << 3, (x23 >> 16) << 3, (x45 & 0xffffu) << 3, (x45 >> 16) << 3};
        uint32_t B[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) B[k] = st[k] + X[k];
        uint32_t kw = 0;
        int done = 0;   // elements of the step already behind a halving: their k is final
        for (;;) {
            uint32_t m[5];
            m[0] = min6(B);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t t[NK];
#pragma unroll
                for (int k = 0; k < NK; k++) t[k] = B[k] + P[j][k];
                m[j + 1] = min6(t);
            }
            const uint32_t knew = (5u - (m[0] & 7u)) | ((5u - (m[1] & 7u)) << 8) | ((5u - (m[2] & 7u)) << 16) | ((5u - (m[3] & 7u)) << 24);
            const int sh = done - 4 * (int)lane;   // how many of my elements are final
            const uint32_t mask = sh <= 0 ? 0xffffffffu : (sh >= 4 ? 0u : 0xffffffffu << (8 * sh));
            kw = (kw & ~mask) | (knew & mask);
            int first = 4;
#pragma unroll
            for (int j = 3; j >= 0; j--)
                if (m[j + 1] >= SE_HALVE_KEY && j < nvalid && j >= sh) first = j;
            const uint32_t bal = __ballot_sync(0xffffffffu, first < 4);
            if (!bal) break;
            const int L = __ffs(bal) - 1;
            const int jh = __shfl_sync(0xffffffffu, first, L);
            uint32_t after[NK], tin[NK];
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const uint32_t pj = jh == 0 ? P[0][k] : (jh == 1 ? P[1][k] : (jh == 2 ? P[2][k] : P[3][k]));
                after[k] = B[k] + pj;
                tin[k] = X[k] + pj;
            }
            halve_keys(after);
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const uint32_t av = __shfl_sync(0xffffffffu, after[k], L), tv = __shfl_sync(0xffffffffu, tin[k], L);
                B[k] = av + X[k] - tv;   // counters before my first element, for lanes behind the halving
            }
            done = 4 * L + jh + 1;
        }
        if (nvalid > 0) ec4[(s0 >> 2) + lane] = kw;
#pragma unroll
        for (int k = 0; k < NK; k++) st[k] = __shfl_sync(0xffffffffu, B[k] + P[3][k], 31);
    }
    if (lane == 0) store_state(S.state[c], st);
}

// code record of a pixel (device_common.cuh: length << 22 | payload)
__device__ __forceinline__ uint32_t make_record(uint32_t wd, const uint8_t *ec) {
    const uint32_t cls = wd >> 30;
    if (cls == 3u) return 0u;
    const uint32_t val = (wd >> 21) & 511u;
    if (cls == 0u) {
        int len;
        const uint32_t code = phase_in_code(((wd >> 12) & 511u) + 1u, val, len);
        return ((uint32_t)(len + 1) << 22) | (1u << len) | code;            // '1' marker then the phased-in code
    }
    const uint32_t k = ec[wd & 8191u];
    const uint32_t q = val >> k, rem = val & ((1u << k) - 1u);
    const uint32_t above = cls == 1u ? 1u : 0u;
    const uint32_t len = 2u + q + 1u + k;
    if (len <= (uint32_t)REC_SHORT_MAX) return (len << 22) | (above << (q + 1u + k)) | (((1u << q) - 1u) << (k + 1u)) | rem;
    return (len << 22) | (above << 17) | (k << 14) | (rem << 9) | q;
}

// bits [off, off + n) of the band's stream, clipped to the window [w0, w1)
__device__ __forceinline__ void put_clipped(uint32_t *out, uint32_t off, uint32_t val, uint32_t n, uint32_t w0, uint32_t w1) {
    const uint32_t lo = max(off, w0), hi = min(off + n, w1);
    if (lo >= hi) return;
    const uint32_t nb = hi - lo;
    uint32_t v = val >> (off + n - hi);
    if (nb < 32u) v &= (1u << nb) - 1u;
    put_bits_smem(out, lo - w0, v, (int)nb);
}
template <typename PUT>
__device__ __forceinline__ void emit_fields(uint32_t r, uint32_t off, PUT put) {
    const uint32_t len = rec_len(r);
    if (len == 0) return;
    if (len <= (uint32_t)REC_SHORT_MAX) { put(off, r & 0x3fffffu, len); return; }
    uint32_t q = r & 511u;
    const uint32_t rem = (r >> 9) & 31u, k = (r >> 14) & 7u, above = (r >> 17) & 1u;
    put(off, above, 2u);   // '0', above
    off += 2;
    while (q >= 32) { put(off, 0xffffffffu, 32u); off += 32; q -= 32; }
    if (q) { put(off, (1u << q) - 1u, q); off += q; }
    put(off, rem, k + 1u);   // '0' then k remainder bits
}

__global__ void __launch_bounds__(SE_THREADS, 3) k_stream_encode(StreamArgs a) {
    extern __shared__ __align__(16) unsigned char se_smem[];
    SeSmem &S = *reinterpret_cast<SeSmem *>(se_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t buf_bytes = a.halo_cap + SE_BAND + 16u;
    uint8_t *const pixbuf0 = se_smem + ((sizeof(SeSmem) + 15) & ~(size_t)15);   // two buffers of buf_bytes
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t step_q = (4u * SE_THREADS) / a.w, step_r = (4u * SE_THREADS) - step_q * a.w;
    const uint32_t slot_words = (uint32_t)(a.slot_bytes >> 2);
    const int w = (int)a.w;

    for (;;) {
        __syncthreads();
        if (tid == 0) S.plane = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t p = S.plane;
        if (p >= a.nplanes) break;
        const uint8_t *plane = a.pixels + (size_t)p * a.npix;
        uint32_t *slot = reinterpret_cast<uint32_t *>(a.temp + (size_t)p * a.slot_bytes);

        for (uint32_t i = tid; i < NBIN * 4; i += SE_THREADS) (&S.state[0][0])[i] = 0u;
        // header (format.rs:51-61) and the two raw samples (compression.rs:93-108): 22 bytes = five words and a half
        const uint32_t v0 = a.npix >= 1 ? plane[0] : 0u, v1 = a.npix >= 2 ? plane[1] : 0u;
        if (tid == 0) {
            slot[0] = bswap32(0x464C4353u);                                  // "FLCS"
            slot[1] = bswap32((a.w >> 16) & 0xffffu);                        // colour 0, depth 0, width bytes 3, 2
            slot[2] = bswap32((a.w << 16) | (a.h >> 16));
            slot[3] = bswap32(a.h << 16);                                    // height bytes 1, 0, raw sample 0 bytes 3, 2
            slot[4] = bswap32(v0 << 16);                                     // raw sample 0 bytes 1, 0, raw sample 1 bytes 3, 2
        }
        uint32_t carry = v1 << 16, carrybits = 16, wpos = 5, ovf = 0;
        uint64_t total_bits = 176;

        const uint32_t nbands = (a.npix + SE_BAND - 1) / SE_BAND;
        load_band(pixbuf0, a, plane, 0, min((uint32_t)SE_BAND, a.npix));
        for (uint32_t b = 0; b < nbands; b++) {
            const uint32_t start = b * SE_BAND, cnt = min((uint32_t)SE_BAND, a.npix - start);
            const uint8_t *pb = pixbuf0 + (b & 1u) * buf_bytes + a.halo_cap;
            cp_async_wait_all();
            __syncthreads();
            if (b + 1 < nbands) load_band(pixbuf0 + ((b + 1) & 1u) * buf_bytes, a, plane, start + SE_BAND, min((uint32_t)SE_BAND, a.npix - start - SE_BAND));

            // ---- classify: four consecutive pixels per thread ------------------------------------------------
            for (uint32_t i = tid; i < SE_WARPS * NBIN / 2; i += SE_THREADS) reinterpret_cast<uint32_t *>(&S.wcnt[0][0])[i] = 0u;
            if (tid == 0) { S.nlong = 0; S.nshort = 0; S.task = 0; }
            {
                const uint32_t i0 = start + 4u * tid;
                uint32_t y = i0 / a.w, x = i0 - y * a.w;
#pragma unroll
                for (int it = 0; it < SE_PPT / 4; it++) {
                    const uint32_t j = (uint32_t)it * 4u * SE_THREADS + 4u * tid;
                    uint4 o = make_uint4(3u << 30, 3u << 30, 3u << 30, 3u << 30);
                    if (j < cnt) {
                        if (x >= 4 && y >= 1) {
                            const uint32_t cur = *reinterpret_cast<const uint32_t *>(pb + j), up = *reinterpret_cast<const uint32_t *>(pb + (int)j - w);
                            const int left = pb[(int)j - 1];
                            const int c0 = cur & 255u, c1 = (cur >> 8) & 255u, c2 = (cur >> 16) & 255u, c3 = cur >> 24;
                            o.x = make_info(c0, left, up & 255u);
                            o.y = make_info(c1, c0, (up >> 8) & 255u);
                            o.z = make_info(c2, c1, (up >> 16) & 255u);
                            o.w = make_info(c3, c2, up >> 24);
                        } else {
                            o.x = classify_slow(pb, (int)j, start + j, x, y, w, plane);
                            o.y = classify_slow(pb, (int)j + 1, start + j + 1, x + 1, y, w, plane);
                            o.z = classify_slow(pb, (int)j + 2, start + j + 2, x + 2, y, w, plane);
                            o.w = classify_slow(pb, (int)j + 3, start + j + 3, x + 3, y, w, plane);
                        }
                    }
                    *reinterpret_cast<uint4 *>(&S.info[info_index(j)]) = o;
                    x += step_r; y += step_q;
                    if (x >= a.w) { x -= a.w; y++; }
                }
            }
            __syncthreads();

            // ---- rank: stable position of every out-of-range pixel among the warp's pixels of its context ------
            for (int s = 0; s < SE_WSTEPS; s++) {
                const uint32_t ix = info_index(wid * SE_WPIX + s * 32 + lane);
                const uint32_t wd = S.info[ix];
                const bool oor = ((wd >> 30) - 1u) < 2u;
                const uint32_t act = __ballot_sync(0xffffffffu, oor);
                if (act == 0) continue;
                const uint32_t delta = (wd >> 12) & 511u;
                uint32_t grp = 0, prev = 0;
                if (oor) {
                    grp = __match_any_sync(act, delta);
                    prev = S.wcnt[wid][delta];
                }
                __syncwarp();
                if (oor) {
                    if ((grp & lt) == 0) S.wcnt[wid][delta] = (uint16_t)(prev + __popc(grp));
                    S.info[ix] = wd | (prev + __popc(grp & lt));
                }
                __syncwarp();
            }
            __syncthreads();

            // ---- chains: exclusive prefix over the warps, chain bases, work lists ------------------------------
            {
                uint32_t run0 = 0, run1 = 0;
#pragma unroll
                for (int q = 0; q < SE_WARPS; q++) {
                    uint32_t *pw = reinterpret_cast<uint32_t *>(&S.wcnt[q][2 * tid]);
                    const uint32_t v = *pw;
                    *pw = run0 | (run1 << 16);
                    run0 += v & 0xffffu; run1 += v >> 16;
                }
                const uint32_t a0 = (run0 + 3u) & ~3u, a1 = (run1 + 3u) & ~3u;
                uint32_t inc = a0 + a1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += t;
                }
                if (lane == 31) S.wsum[wid] = inc;
                __syncthreads();
                uint32_t excl = inc - (a0 + a1);
                for (uint32_t q = 0; q < wid; q++) excl += S.wsum[q];
                S.cb[2 * tid] = excl | (run0 << 16);
                S.cb[2 * tid + 1] = (excl + a0) | (run1 << 16);
                if (run0 >= SE_LONG) S.longlist[atomicAdd(&S.nlong, 1u)] = (uint16_t)(2 * tid);
                else if (run0) S.shortlist[atomicAdd(&S.nshort, 1u)] = (uint16_t)(2 * tid);
                if (run1 >= SE_LONG) S.longlist[atomicAdd(&S.nlong, 1u)] = (uint16_t)(2 * tid + 1);
                else if (run1) S.shortlist[atomicAdd(&S.nshort, 1u)] = (uint16_t)(2 * tid + 1);
            }
            __syncthreads();

            // ---- scatter: residuals into chain order, chain position into the pixel's word ---------------------
            for (int s = 0; s < SE_WSTEPS; s++) {
                const uint32_t ix = info_index(wid * SE_WPIX + s * 32 + lane);
                const uint32_t wd = S.info[ix];
                if (((wd >> 30) - 1u) < 2u) {
                    const uint32_t delta = (wd >> 12) & 511u;
                    const uint32_t pos = (S.cb[delta] & 0xffffu) + S.wcnt[wid][delta] + (wd & 0xfffu);
                    S.ec[pos] = (uint8_t)((wd >> 21) & 511u);
                    S.info[ix] = (wd & 0xffe00000u) | pos;
                }
            }
            __syncthreads();

            // ---- walk: long chains one per warp, short chains one per lane ------------------------------------
            {
                const uint32_t nl = S.nlong, ns = S.nshort;
                for (;;) {
                    uint32_t t = 0;
                    if (lane == 0) t = atomicAdd(&S.task, 1u);
                    t = __shfl_sync(0xffffffffu, t, 0);
                    if (t < nl) {
                        coop_walk(S, S.longlist[t], lane);
                    } else {
                        const uint32_t si = (t - nl) * 32u;
                        if (si >= ns) break;
                        serial_walk(S, si + lane < ns ? (uint32_t)S.shortlist[si + lane] : SE_NONE);
                        __syncwarp();
                    }
                }
            }
            // zero the bit window while the walkers finish (its last reader was the previous band's flush)
            for (uint32_t i = tid; i < SE_OUT_WORDS + 4; i += SE_THREADS) S.out[i] = 0u;
            __syncthreads();

            // ---- code: 16 consecutive pixels per thread, records stay in registers -------------------------------
            uint32_t r[SE_PPT];
            uint32_t mylen = 0;
            {
                const uint4 *iw = reinterpret_cast<const uint4 *>(&S.info[info_index(tid * SE_PPT)]);
#pragma unroll
                for (int q = 0; q < SE_PPT / 4; q++) {
                    const uint4 v = iw[q];
                    r[4 * q] = make_record(v.x, S.ec); r[4 * q + 1] = make_record(v.y, S.ec);
                    r[4 * q + 2] = make_record(v.z, S.ec); r[4 * q + 3] = make_record(v.w, S.ec);
                }
#pragma unroll
                for (int q = 0; q < SE_PPT; q++) mylen += rec_len(r[q]);
            }
            uint32_t inc = mylen;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += t;
            }
            if (lane == 31) S.wsum[wid] = inc;
            if (tid == 0) S.out[0] = carry;
            __syncthreads();
            uint32_t pos = carrybits + inc - mylen, band_bits = 0;
#pragma unroll
            for (int q = 0; q < SE_WARPS; q++) {
                const uint32_t v = S.wsum[q];
                if ((uint32_t)q < wid) pos += v;
                band_bits += v;
            }
            const uint32_t win_bits = carrybits + band_bits;   // bits in the window buffers of this band, carry included

            // ---- pack ---------------------------------------------------------------------------------------
            if (win_bits <= (uint32_t)SE_OUT_WORDS * 32u) {
                // one window: my codes are concatenated in registers and leave as whole words; the first and the last
                // word of my range are shared with my neighbours (atomicOr), the words between are mine alone
                uint32_t wi = pos >> 5, sh = pos & 31u, wv = 0;
                bool shared = true;
                auto flush = [&](uint32_t wgt, uint32_t v) {
                    if (shared) { if (v) atomicOr(&S.out[wgt], v); shared = false; }
                    else S.out[wgt] = v;
                };
#pragma unroll
                for (int q = 0; q < SE_PPT; q++) {
                    const uint32_t len = rec_len(r[q]);
                    if (len == 0) continue;
                    if (len <= (uint32_t)REC_SHORT_MAX) {
                        const uint32_t left = (r[q] & 0x3fffffu) << (32u - len);
                        wv |= left >> sh;
                        if (sh + len >= 32u) {
                            flush(wi, wv);
                            wi++;
                            wv = sh + len > 32u ? left << (32u - sh) : 0u;
                            sh = sh + len - 32u;
                        } else {
                            sh += len;
                        }
                    } else {
                        // long unary run: my partial word first, then field by field
                        if (wv) atomicOr(&S.out[wi], wv);
                        const uint32_t at = (wi << 5) + sh;
                        emit_fields(r[q], at, [&](uint32_t off, uint32_t val, uint32_t nb) { put_bits_smem(S.out, off, val, (int)nb); });
                        const uint32_t np2 = at + len;
                        wi = np2 >> 5; sh = np2 & 31u; wv = 0;
                        shared = true;
                    }
                }
                if (wv) atomicOr(&S.out[wi], wv);
                __syncthreads();
                const uint32_t nfull = win_bits >> 5;
                for (uint32_t j = tid; j < nfull; j += SE_THREADS) {
                    if (wpos + j < slot_words) slot[wpos + j] = bswap32(S.out[j]);
                }
                if (wpos + nfull >= slot_words) ovf = 1;
                carry = S.out[nfull];
                wpos += nfull;
            } else {
                // more than 16 bits per pixel: the band leaves through several windows, every code clipped to the window
                const uint32_t wbits = (uint32_t)SE_OUT_WORDS * 32u;
                for (uint32_t w0 = 0; w0 < win_bits; w0 += wbits) {
                    const uint32_t w1 = w0 + wbits;
                    if (w0) {
                        __syncthreads();
                        for (uint32_t i = tid; i < SE_OUT_WORDS + 4; i += SE_THREADS) S.out[i] = 0u;
                        __syncthreads();
                    }
                    uint32_t off = pos;
#pragma unroll
                    for (int q = 0; q < SE_PPT; q++) {
                        const uint32_t len = rec_len(r[q]);
                        if (len && off < w1 && off + len > w0)
                            emit_fields(r[q], off, [&](uint32_t o2, uint32_t val, uint32_t nb) { put_clipped(S.out, o2, val, nb, w0, w1); });
                        off += len;
                    }
                    __syncthreads();
                    const uint32_t nfull = w1 <= win_bits ? (uint32_t)SE_OUT_WORDS : (win_bits - w0) >> 5;
                    for (uint32_t j = tid; j < nfull; j += SE_THREADS) {
                        if (wpos + j < slot_words) slot[wpos + j] = bswap32(S.out[j]);
                    }
                    if (wpos + nfull >= slot_words) ovf = 1;
                    carry = w1 <= win_bits ? 0u : S.out[nfull];
                    wpos += nfull;
                }
            }
            carrybits = win_bits & 31u;
            total_bits += band_bits;
        }
        // byte_align + flush (compression.rs:279-280): the last partial word leaves zero padded
        if (tid == 0) {
            if (carrybits) {
                if (wpos < slot_words) slot[wpos] = bswap32(carry);
                else ovf = 1;
            }
            a.sizes[p] = (uint32_t)((total_bits + 7) >> 3);
            a.flags[p] = ovf;
        }
    }
}

// offsets of a sub-batch: exclusive scan of the image sizes on top of a running total kept on the device
__global__ void __launch_bounds__(1024) k_stream_scan(const uint32_t *__restrict__ sizes, uint32_t n, uint64_t *__restrict__ img_off, uint64_t *__restrict__ running) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = *running;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n ? sizes[i] : 0;
        uint64_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += t;
        }
        if (lane == 31) wsum[wid] = s;
        __syncthreads();
        if (wid == 0) {
            const uint64_t x = wsum[lane];
            uint64_t y = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= (uint32_t)o) y += t;
            }
            wsum[lane] = y - x;
        }
        __syncthreads();
        const uint64_t excl = carry + wsum[wid] + s - v;
        if (i < n) img_off[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { img_off[n] = carry; *running = carry; }
}

// slot -> final place in the arena (any byte alignment); images whose stream overflowed its slot or whose end lies
// beyond the capacity are skipped (the host deals with both)
__global__ void __launch_bounds__(256) k_stream_compact(const uint8_t *__restrict__ temp, size_t slot_bytes, const uint32_t *__restrict__ sizes,
                                                         const uint32_t *__restrict__ flags, const uint64_t *__restrict__ img_off, uint32_t n,
                                                         uint8_t *__restrict__ arena, const uint64_t *__restrict__ base_ptr, uint64_t cap) {
    const uint64_t base_off = base_ptr ? *base_ptr : 0;   // arena holds the stream from this absolute offset on (host path: one sub-batch)
    for (uint32_t im = blockIdx.x; im < n; im += gridDim.x) {
        const uint32_t bytes = sizes[im];
        const uint64_t at = img_off[im] - base_off;
        if (flags[im] || img_off[im] + bytes - base_off > cap) continue;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(temp + (size_t)im * slot_bytes);
        uint8_t *dst = arena + at;
        const uint32_t head = min(bytes, (uint32_t)((0u - (uint32_t)(uintptr_t)dst) & 15u));   // bytes up to the first 16-byte boundary of dst
        const uint8_t *srcb = reinterpret_cast<const uint8_t *>(src);
        if (threadIdx.x < head) dst[threadIdx.x] = srcb[threadIdx.x];
        const uint32_t body = (bytes - head) >> 4;   // 16-byte chunks
        const uint32_t sw = head >> 2, sb = (head & 3u) * 8u;
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
        for (uint32_t c = threadIdx.x; c < body; c += blockDim.x) {
            const uint32_t *s = src + sw + 4u * c;
            const uint32_t w0 = s[0], w1 = s[1], w2 = s[2], w3 = s[3];
            uint4 o;
            if (sb) {
                const uint32_t w4 = s[4];   // inside the slot: the encoder keeps at least one spare word behind a stream it reports as fitting
                o.x = __funnelshift_r(w0, w1, sb); o.y = __funnelshift_r(w1, w2, sb); o.z = __funnelshift_r(w2, w3, sb); o.w = __funnelshift_r(w3, w4, sb);
            } else {
                o = make_uint4(w0, w1, w2, w3);
            }
            d4[c] = o;
        }
        const uint32_t tail0 = head + (body << 4);
        if (tail0 + threadIdx.x < bytes) dst[tail0 + threadIdx.x] = srcb[tail0 + threadIdx.x];
    }
}

}  // namespace

bool stream_eligible(const felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr) {
    if (ctx->no_stream) return false;
    if (hdr.color_type != 0 || hdr.pixel_depth != 0) return false;
    const uint64_t npix = (uint64_t)hdr.width * hdr.height;
    if (hdr.width % 4 != 0 || hdr.width < 8 || hdr.width > SE_MAX_W || npix <= 2 || npix > (1ull << 26)) return false;   // sizes are u32: 2^26 pixels * 257 bits < 2^32 bytes
    if (d_pixels && ((uintptr_t)d_pixels & 3) != 0) return false;
    return n >= ctx->stream_min;
}

namespace {

struct StreamPlan {
    uint32_t w, h, npix;
    size_t slot_bytes, smem;
    int per_sm;
    StreamArgs a;
};

int stream_plan(felics_ctx *ctx, const felics_header &hdr, bool vec16, StreamPlan &pl) {
    pl.w = hdr.width; pl.h = hdr.height; pl.npix = hdr.width * hdr.height;
    pl.slot_bytes = align_up((size_t)pl.npix + pl.npix / 4 + 64, 16);   // 10 bits per pixel; longer streams go through the general pipeline
    StreamArgs &a = pl.a;
    a.w = pl.w; a.h = pl.h; a.npix = pl.npix;
    a.halo_cap = (uint32_t)align_up(pl.w, 16);
    a.vec16 = vec16 && pl.w % 16 == 0 ? 1u : 0u;
    a.slot_bytes = pl.slot_bytes;
    pl.smem = ((sizeof(SeSmem) + 15) & ~(size_t)15) + 2 * (size_t)(a.halo_cap + SE_BAND + 16);
    if (!ctx->stream_attr_done) {
        FELICS_CUDA_TRY(cudaFuncSetAttribute(k_stream_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        ctx->sm_count = sms;
        ctx->stream_attr_done = true;
    }
    pl.per_sm = 1;
    FELICS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pl.per_sm, k_stream_encode, SE_THREADS, pl.smem));
    pl.per_sm = std::max(pl.per_sm, 1);
    return FELICS_OK;
}

// encode `ni` images, scan their sizes on top of *d_running, copy the streams to `target` (which starts at absolute offset
// *base_ptr of the whole batch's stream, nullptr = 0)
int stream_launch(felics_ctx *ctx, StreamPlan &pl, size_t ni, const uint8_t *d_pixels, uint32_t *d_sizes, uint32_t *d_flags, uint64_t *d_off,
                  uint32_t *d_ticket, uint64_t *d_running, uint8_t *d_temp, uint8_t *target, const uint64_t *base_ptr, uint64_t target_cap) {
    cudaStream_t st = ctx->stream;
    {
        StageScope s(ctx, ST_STREAM);
        FELICS_CUDA_TRY(cudaMemsetAsync(d_ticket, 0, sizeof(uint32_t), st));
        StreamArgs a = pl.a;
        a.pixels = d_pixels; a.temp = d_temp; a.sizes = d_sizes; a.flags = d_flags; a.ticket = d_ticket; a.nplanes = (uint32_t)ni;
        const unsigned blocks = (unsigned)std::min<size_t>(ni, (size_t)ctx->sm_count * pl.per_sm);
        k_stream_encode<<<blocks, SE_THREADS, pl.smem, st>>>(a);
        s.launched();
    }
    {
        StageScope s(ctx, ST_COMPACT);
        k_stream_scan<<<1, 1024, 0, st>>>(d_sizes, (uint32_t)ni, d_off, d_running);
        const unsigned blocks = (unsigned)std::min<size_t>(ni, (size_t)ctx->sm_count * 8);
        k_stream_compact<<<blocks, 256, 0, st>>>(d_temp, pl.slot_bytes, d_sizes, d_flags, d_off, (uint32_t)ni, target, base_ptr, target_cap);
        s.launched(2);
    }
    return FELICS_OK;
}

struct StreamScratch {
    uint32_t *sizes, *flags, *ticket;
    uint64_t *off, *running;
    uint8_t *temp;
};

int stream_scratch(felics_ctx *ctx, size_t n, size_t temp_bytes, StreamScratch &sc) {
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 256); const size_t at = off; off += bytes; return at; };
    const size_t o_sizes = take(n * sizeof(uint32_t)), o_flags = take(n * sizeof(uint32_t)), o_off = take((n + 1) * sizeof(uint64_t));
    const size_t o_ctr = take(64), o_temp = take(temp_bytes);
    int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, off);
    if (rc) return rc;
    uint8_t *b = (uint8_t *)ctx->scratch;
    sc.sizes = (uint32_t *)(b + o_sizes); sc.flags = (uint32_t *)(b + o_flags); sc.off = (uint64_t *)(b + o_off);
    sc.ticket = (uint32_t *)(b + o_ctr); sc.running = (uint64_t *)(b + o_ctr + 16); sc.temp = b + o_temp;
    return FELICS_OK;
}

}  // namespace

// Device-resident batch: pixels and arena in device memory.  offsets_host[0..n] filled on return; one host
// synchronisation, at the end.
int stream_encode_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr, uint8_t *d_arena, size_t arena_cap,
                               uint64_t *offsets_host) {
    cudaStream_t st = ctx->stream;
    StreamPlan pl;
    int rc = stream_plan(ctx, hdr, ((uintptr_t)d_pixels & 15) == 0, pl);
    if (rc) return rc;
    const size_t sub_max = std::max<size_t>(1, std::min<size_t>(n, ((size_t)4 << 30) / pl.slot_bytes));
    StreamScratch sc;
    if ((rc = stream_scratch(ctx, n, sub_max * pl.slot_bytes, sc))) return rc;
    FELICS_CUDA_TRY(cudaMemsetAsync(sc.ticket, 0, 64, st));
    for (size_t first = 0; first < n; first += sub_max) {
        const size_t ni = std::min(sub_max, n - first);
        rc = stream_launch(ctx, pl, ni, (const uint8_t *)d_pixels + first * (size_t)pl.npix, sc.sizes + first, sc.flags + first, sc.off + first, sc.ticket,
                           sc.running, sc.temp, d_arena, nullptr, arena_cap);
        if (rc) return rc;
    }
    // one read-back at the end: offsets and overflow flags
    rc = ensure_buffer(ctx, &ctx->pinned, &ctx->pinned_cap, (n + 1) * sizeof(uint64_t) + n * sizeof(uint32_t) + 64, true);
    if (rc) return rc;
    uint64_t *h_off = (uint64_t *)ctx->pinned;
    uint32_t *h_flags = (uint32_t *)(h_off + n + 1);
    FELICS_CUDA_TRY(cudaMemcpyAsync(h_off, sc.off, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    FELICS_CUDA_TRY(cudaMemcpyAsync(h_flags, sc.flags, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    FELICS_CUDA_TRY(cudaGetLastError());
    std::memcpy(offsets_host, h_off, (n + 1) * sizeof(uint64_t));
    if (offsets_host[n] > arena_cap) {
        set_error("output capacity %zu too small (need %llu)", arena_cap, (unsigned long long)offsets_host[n]);
        profile_collect(ctx);
        return FELICS_ERR_BUFFER_TOO_SMALL;
    }
    // streams that did not fit their slot (more than 10 bits per pixel): the general pipeline encodes them into the hole
    std::vector<size_t> redo;
    for (size_t i = 0; i < n; i++)
        if (h_flags[i]) redo.push_back(i);
    if (!redo.empty()) {
        const bool keep = ctx->no_stream;
        ctx->no_stream = true;
        for (size_t i : redo) {
            const size_t bytes = (size_t)(offsets_host[i + 1] - offsets_host[i]);
            rc = ensure_buffer(ctx, &ctx->staging_out, &ctx->staging_out_cap, bytes + 64);
            uint64_t o2[2] = {0, 0};
            if (!rc) rc = encode_batch_device(ctx, 1, (const uint8_t *)d_pixels + i * (size_t)pl.npix, hdr, (uint8_t *)ctx->staging_out, nullptr, bytes + 16, o2);
            if (!rc && o2[1] != bytes) { set_error("internal: stream size mismatch for image %zu (%llu vs %zu)", i, (unsigned long long)o2[1], bytes); rc = FELICS_ERR_CUDA; }
            if (!rc && cudaMemcpyAsync(d_arena + offsets_host[i], ctx->staging_out, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = FELICS_ERR_CUDA;
            if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = FELICS_ERR_CUDA;
            if (rc) break;
        }
        ctx->no_stream = keep;
        if (rc) return rc;
        ctx->stream_redone += redo.size();
    }
    return profile_collect(ctx);
}

// Host-resident batch (felics_compress_batch): sub-batches travel in, are encoded and travel out on three streams,
// double buffered; the host waits for the sizes of sub-batch i-1 (it needs them to place the copy-out) only after
// the kernels of sub-batch i are queued, so the device never waits for the host.
int stream_encode_batch_host(felics_ctx *ctx, size_t n, const void *h_pixels, const felics_header &hdr, uint8_t *h_arena, size_t arena_cap,
                             uint64_t *offsets_host) {
    cudaStream_t st = ctx->stream;
    StreamPlan pl;
    int rc = stream_plan(ctx, hdr, true, pl);
    if (rc) return rc;
    size_t sub = std::max<size_t>(32, (n + 7) / 8);                                       // at least eight sub-batches to overlap
    sub = std::min(sub, std::max<size_t>(1, ((size_t)1 << 30) / pl.slot_bytes));           // at most 1 GB of slots
    sub = std::min(sub, n);
    const size_t nsub = (n + sub - 1) / sub;
    StreamScratch sc;
    if ((rc = stream_scratch(ctx, n, sub * pl.slot_bytes, sc))) return rc;
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
    for (int i = 0; i < 2; i++)
        if (!ctx->ev_sizes[i]) FELICS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_sizes[i], cudaEventDisableTiming));
    rc = ensure_buffer(ctx, &ctx->pinned, &ctx->pinned_cap, (n + 1) * sizeof(uint64_t) + n * sizeof(uint32_t) + 64, true);
    if (rc) return rc;
    uint64_t *h_off = (uint64_t *)ctx->pinned;
    uint32_t *h_flags = (uint32_t *)(h_off + n + 1);
    const size_t img_bytes = pl.npix;
    for (int i = 0; i < 2; i++) {
        if ((rc = ensure_buffer(ctx, &ctx->stage_in[i], &ctx->stage_in_cap[i], sub * img_bytes + 16))) return rc;
        if ((rc = ensure_buffer(ctx, &ctx->stage_out[i], &ctx->stage_out_cap[i], sub * pl.slot_bytes + 16))) return rc;
    }
    // every exit below this line first waits for the copies that read the caller's pixels / write the caller's arena
    auto drain = [&]() { cudaStreamSynchronize(ctx->copy_in); cudaStreamSynchronize(st); cudaStreamSynchronize(ctx->copy_out); };
    auto fail = [&](int code) { drain(); return code; };
#define SE_TRY(expr)                                                                                                   \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) {                                                                                       \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                     \
            return fail(FELICS_ERR_CUDA);                                                                              \
        }                                                                                                              \
    } while (0)
    SE_TRY(cudaMemsetAsync(sc.ticket, 0, 64, st));
    SE_TRY(cudaEventRecord(ctx->ev_done[0], st));   // the copy streams start after whatever the caller queued on the context's stream
    SE_TRY(cudaEventRecord(ctx->ev_done[1], st));
    SE_TRY(cudaEventRecord(ctx->ev_out[0], st));
    SE_TRY(cudaEventRecord(ctx->ev_out[1], st));
    auto enqueue_in = [&](size_t idx) -> cudaError_t {
        const int slot = (int)(idx & 1);
        const size_t first = idx * sub, ni = std::min(sub, n - first);
        cudaError_t e = cudaStreamWaitEvent(ctx->copy_in, ctx->ev_done[slot], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->stage_in[slot], (const uint8_t *)h_pixels + first * img_bytes, ni * img_bytes, cudaMemcpyHostToDevice, ctx->copy_in);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_in[slot], ctx->copy_in);
        return e;
    };
    bool too_small = false;
    // sizes of sub-batch idx are on the host: queue the copy of its streams into the caller's arena
    auto complete = [&](size_t idx) -> cudaError_t {
        const int slot = (int)(idx & 1);
        const size_t first = idx * sub, ni = std::min(sub, n - first);
        cudaError_t e = cudaEventSynchronize(ctx->ev_sizes[slot]);
        if (e != cudaSuccess) return e;
        const uint64_t lo = h_off[first], hi = h_off[first + ni];
        if (hi > arena_cap) { too_small = true; return cudaEventRecord(ctx->ev_out[slot], ctx->copy_out); }
        e = cudaStreamWaitEvent(ctx->copy_out, ctx->ev_pack[slot], 0);
        if (e == cudaSuccess && hi > lo) e = cudaMemcpyAsync(h_arena + lo, ctx->stage_out[slot], hi - lo, cudaMemcpyDeviceToHost, ctx->copy_out);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_out[slot], ctx->copy_out);
        return e;
    };
    SE_TRY(enqueue_in(0));
    for (size_t idx = 0; idx < nsub; idx++) {
        const int slot = (int)(idx & 1);
        const size_t first = idx * sub, ni = std::min(sub, n - first);
        if (idx + 1 < nsub) SE_TRY(enqueue_in(idx + 1));
        SE_TRY(cudaStreamWaitEvent(st, ctx->ev_in[slot], 0));
        SE_TRY(cudaStreamWaitEvent(st, ctx->ev_out[slot], 0));   // the copy-out that last read this slot's streams
        rc = stream_launch(ctx, pl, ni, (const uint8_t *)ctx->stage_in[slot], sc.sizes + first, sc.flags + first, sc.off + first, sc.ticket, sc.running,
                           sc.temp, (uint8_t *)ctx->stage_out[slot], sc.off + first, sub * pl.slot_bytes);
        if (rc) return fail(rc);
        SE_TRY(cudaEventRecord(ctx->ev_done[slot], st));
        SE_TRY(cudaEventRecord(ctx->ev_pack[slot], st));
        SE_TRY(cudaMemcpyAsync(h_off + first, sc.off + first, (ni + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        SE_TRY(cudaMemcpyAsync(h_flags + first, sc.flags + first, ni * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SE_TRY(cudaEventRecord(ctx->ev_sizes[slot], st));
        if (idx > 0) SE_TRY(complete(idx - 1));
    }
    SE_TRY(complete(nsub - 1));
    SE_TRY(cudaStreamSynchronize(st));
    SE_TRY(cudaStreamSynchronize(ctx->copy_out));
    SE_TRY(cudaStreamSynchronize(ctx->copy_in));
    SE_TRY(cudaGetLastError());
#undef SE_TRY
    std::memcpy(offsets_host, h_off, (n + 1) * sizeof(uint64_t));
    if (too_small || offsets_host[n] > arena_cap) {
        set_error("output capacity %zu too small (need %llu)", arena_cap, (unsigned long long)offsets_host[n]);
        profile_collect(ctx);
        return FELICS_ERR_BUFFER_TOO_SMALL;
    }
    std::vector<size_t> redo;
    for (size_t i = 0; i < n; i++)
        if (h_flags[i]) redo.push_back(i);
    if (!redo.empty()) {
        const bool keep = ctx->no_stream;
        ctx->no_stream = true;
        for (size_t i : redo) {
            const size_t bytes = (size_t)(offsets_host[i + 1] - offsets_host[i]);
            uint64_t o2[2] = {0, 0};
            rc = felics_compress_batch(ctx, 1, (const uint8_t *)h_pixels + i * img_bytes, &hdr, h_arena + offsets_host[i], bytes, o2);
            if (!rc && o2[1] != bytes) { set_error("internal: stream size mismatch for image %zu", i); rc = FELICS_ERR_CUDA; }
            if (rc) break;
        }
        ctx->no_stream = keep;
        if (rc) return rc;
        ctx->stream_redone += redo.size();
    }
    return profile_collect(ctx);
}

}  // namespace felics
