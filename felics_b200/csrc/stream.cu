// Streaming band encoder for batches of 8-bit gray images (sm_100a): ONE thread block encodes ONE whole image.
//
// The multi-kernel pipeline of encode.cu is built for a single big image: every stage is a kernel over the whole
// plane and every intermediate (grouped residuals, grouped indices, code records ...) travels through HBM.  A batch
// of small images does not need that: the estimator rows of an image (parameter_selection.rs:24-85) are 256 x 6
// small counters, so a block can walk the image in raster order, band by band (4096 pixels), and keep everything it
// needs between the pixels and the bit stream in shared memory:
//
//   load      band + the row above it, cp.async into shared memory (double buffered)
//   group     every warp takes 512 consecutive pixels, 32 at a time with lane = pixel: neighbours, context, class and
//             residual (misc.rs:6-24, compression.rs:118-145), then the stable rank of every out-of-range pixel among
//             the warp's pixels of its context (match_any + one counter per context), one scan over the warp's
//             counters, and the residuals land in the warp's region of the chain array, grouped by context, raster
//             order kept.  A context's chain in this band = its segments in the eight regions, in warp order
//   walk      KEstimator for every chain, starting from the counters the previous band left.  A long chain is cut
//             into steps of 128 elements; a step is one task: any warp loads its elements, looks up their six code
//             costs and prefix-sums them (all of that is independent of the counters), then waits until the step
//             before it has published the counters, finds the halvings by ballot (counters only grow inside an epoch)
//             and the k of every element, and publishes the counters for the next step -- so the steps of one chain
//             are prepared side by side and only their short second halves run one after the other.  Short chains
//             are walked one per lane, the reference's loop as written.  Exact for any data
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
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <vector>

namespace felics {

namespace {

#ifndef SE_THREADS_N
#define SE_THREADS_N 256
#endif
constexpr int SE_THREADS = SE_THREADS_N;
#ifndef SE_CTAS_PER_SM
#define SE_CTAS_PER_SM 4
#endif
constexpr int SE_WARPS = SE_THREADS / 32;
constexpr int SE_PPT = 16;                                  // pixels per thread in the code / pack phase
constexpr int SE_BAND = SE_THREADS * SE_PPT;                // 4096 pixels per band
constexpr int SE_WPIX = SE_BAND / SE_WARPS;                 // 512 consecutive pixels per warp in the group phase
constexpr int SE_WSTEPS = SE_WPIX / 32;
constexpr int SE_INFO_WORDS = SE_BAND + SE_BAND / 16 * 4;   // one word per pixel, 4 words of padding per 16 (bank spread)
constexpr int SE_NCTX = 256;                                // contexts of 8-bit gray samples: H - L <= 255
constexpr int SE_WREG = SE_WPIX + 3 * SE_NCTX;              // bytes of a warp's region of the chain array (segments are padded to 4)
constexpr int SE_OUT_WORDS = SE_BAND / 2;                    // bit window: 16 bits per pixel of a band
constexpr uint32_t SE_LONG = 32;                            // chains this long (padded, per band) are walked in 128-element steps
constexpr uint32_t SE_HALVE_KEY = (HALVE_AT + 1u) << 3;     // key of a count that has passed 1024
constexpr uint32_t SE_NONE = 0xFFFFFFFFu;
constexpr uint32_t SE_MAX_W = 8192;                         // widest row whose predecessor fits in front of a band
constexpr uint32_t SE_NULL_E = 255;                         // padding element of a chain segment (a gray residual is at most 254)

// info word of a pixel:  [31] out of range  [30] above (with [31]); 01 = not coded  [29:21] P-L | P-H-1 | L-P-1
//                        [20:12] context  [11:0] rank among the warp's pixels of that context, then position in the warp's region
constexpr uint32_t SE_INFO_NONE = 0x40000000u;
__device__ __forceinline__ uint32_t info_index(uint32_t p) { return p + ((p >> 4) << 2); }

struct SeSmem {
    uint32_t info[SE_INFO_WORDS];
    uint32_t wseg[SE_WARPS][SE_NCTX];   // per warp and context: count while ranking, then base << 16 | count inside the warp's region
    uint32_t stepdone[SE_NCTX];         // steps of a chain walked so far in this band
    uint32_t state[SE_NCTX][3];         // estimator row of a context: counts as u16 pairs (k0 | k1 << 16, k2 | k3 << 16, k4 | k5 << 16)
    uint2 lut03[SE_NCTX];               // code costs of a residual under k = 0..3, times 8, as u16 pairs (nothing for SE_NULL_E)
    uint32_t lut45[SE_NCTX];            // ... under k = 4, 5
    uint16_t longc[SE_NCTX];            // long chains of the band: context | steps << 8
    uint16_t shortlist[SE_NCTX];
    uint32_t roundcnt[SE_BAND / 128 + 8];   // chains that have a step s (a chain of a band has at most SE_BAND / 128 + 2 steps)
    uint8_t ec[SE_WARPS][SE_WREG];      // residuals grouped by context, one region per warp; the walk overwrites them with k
    uint32_t wsum[SE_WARPS];
    uint32_t ntask, nlong, nshort, task, plane;
    uint64_t bar[2];                    // one mbarrier per pixel buffer (TMA bulk loads)
};
// the bit window of the pack phase (SE_OUT_WORDS + 4 words) lies over wseg and stepdone: both are dead once the walk is over
static_assert(sizeof(SeSmem::wseg) + sizeof(SeSmem::stepdone) >= (SE_OUT_WORDS + 4) * sizeof(uint32_t), "bit window does not fit");
static_assert(offsetof(SeSmem, stepdone) == offsetof(SeSmem, wseg) + sizeof(SeSmem::wseg), "bit window must be contiguous");

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
    uint32_t dbg;               // timing experiments only (FELICS_B200_STREAM_DBG): 1 no walk, 2 no code/pack, 4 no group/walk, 8 phase clocks
    unsigned long long *clocks; // dbg & 8: cycles per phase, summed over the blocks (thread 0 of every block)
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// TMA bulk copy (cp.async.bulk, 1-D) completing on an mbarrier: one thread moves a whole band
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes),
                 "r"(smem_addr(bar))
                 : "memory");
}

// pixels [start - hl, start + cnt) of the plane -> buf[halo_cap - hl ...): the band starts at buf + halo_cap.  Aligned images
// (vec16: base, width and pixel count multiples of 16) travel as one TMA bulk copy issued by thread 0 that completes on `bar`;
// the others as 4-byte cp.async copies by all threads
__device__ __forceinline__ void load_band(uint8_t *buf, const StreamArgs &a, const uint8_t *plane, uint32_t start, uint32_t cnt, uint64_t *bar) {
    const uint32_t hl = min(a.w, start);
    const uint32_t total = hl + cnt;
    const uint8_t *src = plane + start - hl;
    uint8_t *dst = buf + a.halo_cap - hl;
    if (a.vec16) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, total);
            bulk_load(dst, src, total, bar);
        }
    } else {
        for (uint32_t o = 4u * threadIdx.x; o < total; o += 4u * SE_THREADS) cp_async4(dst + o, src + o);
        cp_async_commit();
    }
}

// compression.rs:118-145 for one pixel given its two neighbours.  At most one of L-P-1 and P-H-1 is non-negative:
// it is the residual of an out-of-range pixel; both negative = in range, the code is P-L.
__device__ __forceinline__ uint32_t make_info(int p, int v1, int v2) {
    const int h = max(v1, v2), l = min(v1, v2);
    const int below = l - p - 1, above = p - h - 1;
    const int m = max(below, above);
    const uint32_t val = (uint32_t)(m >= 0 ? m : p - l);
    const uint32_t top = (~(uint32_t)(below & above) & 0x80000000u) | ((~(uint32_t)above >> 1) & 0x40000000u);   // [31] out of range, [30] above
    return top + val * 0x200000u + (uint32_t)(h - l) * 0x1000u;   // disjoint fields: sums the compiler can turn into multiply-adds (the ALU pipe is the busy one)
}
// any pixel (misc.rs:6-24): steps that touch a row start, the first row or the end of the image; pb = first pixel of the band,
// j = offset in the band, i = index in the plane
__device__ __noinline__ uint32_t classify_any(const uint8_t *pb, int j, uint32_t i, int w, uint32_t npix, const uint8_t *plane) {
    if (i < 2 || i >= npix) return SE_INFO_NONE;
    const uint32_t y = i / (uint32_t)w, x = i - y * (uint32_t)w;
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
    const uint32_t s0 = row[0], s1 = row[1], s2 = row[2];
    v[0] = ((s0 & 0xffffu) << 3) | 5u; v[1] = ((s0 >> 16) << 3) | 4u;
    v[2] = ((s1 & 0xffffu) << 3) | 3u; v[3] = ((s1 >> 16) << 3) | 2u;
    v[4] = ((s2 & 0xffffu) << 3) | 1u; v[5] = ((s2 >> 16) << 3);
}
__device__ __forceinline__ void store_state(uint32_t *row, const uint32_t (&v)[NK]) {
    // counts stay below 2^16 for 8-bit data: c0 <= 25 * (1024 + 510)
    row[0] = (v[0] >> 3) | ((v[1] >> 3) << 16); row[1] = (v[2] >> 3) | ((v[3] >> 3) << 16); row[2] = (v[4] >> 3) | ((v[5] >> 3) << 16);
}
// the six code costs (e >> k) + 1 + k of a residual (rice_coding.rs:56-58), times 8, from the table
__device__ __forceinline__ void cost_keys(const SeSmem &S, uint32_t e, uint32_t (&c)[NK]) {
    const uint2 t = S.lut03[e];
    const uint32_t u = S.lut45[e];
    c[0] = t.x & 0xffffu; c[1] = t.x >> 16; c[2] = t.y & 0xffffu; c[3] = t.y >> 16; c[4] = u & 0xffffu; c[5] = u >> 16;
}
__device__ __forceinline__ uint32_t min6(const uint32_t (&v)[NK]) { return min(min(min(v[0], v[1]), min(v[2], v[3])), min(v[4], v[5])); }
__device__ __forceinline__ void halve_keys(uint32_t (&v)[NK]) {              // parameter_selection.rs:58-63
#pragma unroll
    for (int k = 0; k < NK; k++) v[k] = ((v[k] >> 4) << 3) | (uint32_t)(5 - k);
}

// one chain per lane (short chains): the reference's loop as written, segment after segment
__device__ __forceinline__ void serial_walk(SeSmem &S, uint32_t c) {
    if (c == SE_NONE) return;
    uint32_t st[NK];
    load_state(S.state[c], st);
    uint32_t m = min6(st);
    for (int q = 0; q < SE_WARPS; q++) {
        const uint32_t sw = S.wseg[q][c];
        uint8_t *seg = S.ec[q] + (sw >> 16);
        const uint32_t n = sw & 0xffffu;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t e = seg[i];
            seg[i] = (uint8_t)(5u - (m & 7u));
            uint32_t c6[NK];
            cost_keys(S, e, c6);
#pragma unroll
            for (int k = 0; k < NK; k++) st[k] += c6[k];
            m = min6(st);
            if (m >= SE_HALVE_KEY) { halve_keys(st); m = min6(st); }
        }
    }
    store_state(S.state[c], st);
}

// One step of a long chain: 128 consecutive elements of the chain (four per lane; the chain is the concatenation of its
// segments in the eight warp regions, every segment padded to a multiple of four with null elements that cost nothing).
// With T the prefix sum of the costs inside the step, the counters before element g are B + T(g) as long as no halving
// lies between; the first element after which all six counters have passed 1024 is found by one ballot (counters only
// grow inside an epoch), the halved counters become the new base for the elements behind it, and the search repeats.
__device__ __forceinline__ void coop_step(SeSmem &S, uint32_t c, uint32_t step, uint32_t lane) {
    // where my four elements live
    const uint32_t sw = lane < SE_WARPS ? S.wseg[lane][c] : 0u;
    const uint32_t pc = ((sw & 0xffffu) + 3u) & ~3u;          // padded length of segment `lane`
    uint32_t incl = pc;
#pragma unroll
    for (int o = 1; o < SE_WARPS; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t v = step * 128u + 4u * lane;               // my first element's index in the chain
    uint32_t seg = 0;
#pragma unroll
    for (int q = 0; q < SE_WARPS - 1; q++) seg += __shfl_sync(0xffffffffu, incl, q) <= v ? 1u : 0u;
    const uint32_t vtot = __shfl_sync(0xffffffffu, incl, SE_WARPS - 1);
    const uint32_t seg_first = __shfl_sync(0xffffffffu, incl - pc, seg), seg_base = __shfl_sync(0xffffffffu, sw >> 16, seg);
    const bool valid = v < vtot;
    uint32_t *slot = reinterpret_cast<uint32_t *>(S.ec[seg] + seg_base + (v - seg_first));
    const uint32_t word = valid ? *slot : 0xffffffffu;

    uint32_t P[4][NK];   // inclusive cost prefix inside my four elements
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t c6[NK];
        cost_keys(S, (word >> (8 * j)) & 255u, c6);
#pragma unroll
        for (int k = 0; k < NK; k++) P[j][k] = j ? P[j - 1][k] + c6[k] : c6[k];
    }
    // exclusive prefix of the lane totals over the warp, two 16-bit sums per register (128 * 255 < 65536)
    const uint32_t t01 = (P[3][0] >> 3) | ((P[3][1] >> 3) << 16), t23 = (P[3][2] >> 3) | ((P[3][3] >> 3) << 16), t45 = (P[3][4] >> 3) | ((P[3][5] >> 3) << 16);
    uint32_t x01 = t01, x23 = t23, x45 = t45;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, x01, o), b = __shfl_up_sync(0xffffffffu, x23, o), d = __shfl_up_sync(0xffffffffu, x45, o);
        if (lane >= (uint32_t)o) { x01 += a; x23 += b; x45 += d; }
    }
    x01 -= t01; x23 -= t23; x45 -= t45;
    const uint32_t X[NK] = {(x01 & 0xffffu) << 3, (x01 >> 16) << 3, (x23 & 0xffffu) << 3, (x23 >> 16) << 3, (x45 & 0xffffu) << 3, (x45 >> 16) << 3};

    // everything above is independent of the counters: now wait for the step before this one
    if (step) {
        while (*reinterpret_cast<volatile uint32_t *>(&S.stepdone[c]) < step) __nanosleep(100);
        __threadfence_block();
    }
    uint32_t B[NK];
    load_state(S.state[c], B);
#pragma unroll
    for (int k = 0; k < NK; k++) B[k] += X[k];
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
            if (m[j + 1] >= SE_HALVE_KEY && j >= sh) first = j;
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
    if (valid) *slot = kw;
    uint32_t st[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) st[k] = __shfl_sync(0xffffffffu, B[k] + P[3][k], 31);
    if (lane == 0) {
        store_state(S.state[c], st);
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t *>(&S.stepdone[c]) = step + 1u;
    }
}

// code record of a pixel: length << 22 | payload.  length <= 22: the payload is the whole code word, right aligned
// (in range: '1' marker + phased-in code, phase_in_coding.rs:64-84; out of range: '0', above, q ones, '0', k remainder bits,
// rice_coding.rs:26-39).  Longer codes (long unary runs) carry above << 12 | k << 9 | e and are expanded by the packer.
// Straight-line code: both classes are evaluated for every pixel and selected, a divergent branch costs more.
__device__ __forceinline__ uint32_t make_record(uint32_t wd, const uint8_t *ecw) {
    const uint32_t top = wd >> 30;                       // 0 in range, 1 not coded, 2 below, 3 above
    const uint32_t val = (wd >> 21) & 511u;
    const bool oor = top >= 2u;
    // in range: x = (v + n - left_p) mod n = (v + 2^m) mod n; short codes (x < right_p) have m bits, the others m + 1 bits holding x + right_p
    const uint32_t n = ((wd >> 12) & 511u) + 1u;
    const uint32_t m = 31u - (uint32_t)__clz(n);
    const uint32_t p2 = 1u << m;
    uint32_t x = val + p2;
    x = min(x, x - n);
    const uint32_t rp = 2u * p2 - n;
    const bool lng = x >= rp;
    const uint32_t rin = ((m + 1u + (lng ? 1u : 0u)) << 22) + x + (lng ? rp + 2u * p2 : p2);
    // out of range
    const uint32_t k = ecw[oor ? (wd & 0xfffu) : 0u] & 7u;
    const uint32_t q = val >> k;
    const uint32_t rem = val - (q << k);
    const uint32_t len = q + k + 3u;
    const uint32_t ones = ((top - 1u) << min(q, 24u)) - 1u;   // above: q + 1 ones, the top one is the `above` bit; below: q ones
    const uint32_t rshort = (len << 22) | (ones << (k + 1u)) | rem;
    const uint32_t rlong = len * 0x400000u + (top & 1u) * 0x1000u + k * 0x200u + val;
    const uint32_t roor = len <= (uint32_t)REC_SHORT_MAX ? rshort : rlong;
    const uint32_t r = oor ? roor : rin;
    return top == 1u ? 0u : r;
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
__device__ __forceinline__ void emit_fields(uint32_t r, uint32_t off, PUT put) {   // any record, field by field
    const uint32_t len = rec_len(r);
    if (len == 0) return;
    if (len <= (uint32_t)REC_SHORT_MAX) { put(off, r & 0x3fffffu, len); return; }
    const uint32_t e = r & 511u, k = (r >> 9) & 7u, above = (r >> 12) & 1u;
    uint32_t q = e >> k;
    put(off, above, 2u);   // '0', above
    off += 2;
    while (q >= 32) { put(off, 0xffffffffu, 32u); off += 32; q -= 32; }
    if (q) { put(off, (1u << q) - 1u, q); off += q; }
    put(off, e & ((1u << k) - 1u), k + 1u);   // '0' then k remainder bits
}

// a long unary run (more than 22 bits) straight into the window, field by field; rare, kept out of the packer's loop
__device__ __noinline__ void pack_long(uint32_t *out, uint32_t r, uint32_t at) {
    emit_fields(r, at, [&](uint32_t off, uint32_t val, uint32_t nb) { put_bits_smem(out, off, val, (int)nb); });
}

// A band of more than SE_OUT_WORDS * 32 bits (over 16 bits per pixel): the band leaves through several windows, every code
// clipped to the window.  Called by the whole block; `recs` = the caller's SE_PPT records, `pos` = bit offset of its first one.
__device__ __noinline__ void pack_windows(uint32_t *out, const uint32_t *recs, uint32_t pos, uint32_t win_bits, uint32_t *slot, uint32_t slot_words,
                                          uint32_t &wpos, uint32_t &carry, uint32_t &ovf) {
    const uint32_t tid = threadIdx.x;
    const uint32_t wbits = (uint32_t)SE_OUT_WORDS * 32u;
    for (uint32_t w0 = 0; w0 < win_bits; w0 += wbits) {
        const uint32_t w1 = w0 + wbits;
        if (w0) {
            __syncthreads();
            for (uint32_t i = tid; i < SE_OUT_WORDS + 4; i += SE_THREADS) out[i] = 0u;
            __syncthreads();
        }
        uint32_t off = pos;
#pragma unroll 1
        for (int q = 0; q < SE_PPT; q++) {
            const uint32_t rq = recs[q], len = rec_len(rq);
            if (len && off < w1 && off + len > w0)
                emit_fields(rq, off, [&](uint32_t o2, uint32_t val, uint32_t nb) { put_clipped(out, o2, val, nb, w0, w1); });
            off += len;
        }
        __syncthreads();
        const uint32_t nfull = w1 <= win_bits ? (uint32_t)SE_OUT_WORDS : (win_bits - w0) >> 5;
        for (uint32_t j = tid; j < nfull; j += SE_THREADS) {
            if (wpos + j < slot_words) slot[wpos + j] = bswap32(out[j]);
        }
        if (wpos + nfull >= slot_words) ovf = 1;
        carry = w1 <= win_bits ? 0u : out[nfull];
        wpos += nfull;
    }
}

__global__ void __launch_bounds__(SE_THREADS, SE_CTAS_PER_SM) k_stream_encode(StreamArgs a) {
    extern __shared__ __align__(16) unsigned char se_smem[];
    SeSmem &S = *reinterpret_cast<SeSmem *>(se_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t buf_bytes = a.halo_cap + SE_BAND + 16u;
    uint8_t *const pixbuf0 = se_smem + ((sizeof(SeSmem) + 15) & ~(size_t)15);   // two buffers of buf_bytes
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t *const out = &S.wseg[0][0];   // the bit window (see SeSmem)
    const uint32_t slot_words = (uint32_t)(a.slot_bytes >> 2);
    const int w = (int)a.w;
    long long tclk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
    const bool clk = (a.dbg & 8u) && tid == 0;
#define SE_CLK(i) do { if (clk) { const long long t_ = clock64(); tclk[i] += t_ - tlast; tlast = t_; } } while (0)

    // code costs of every residual (rice_coding.rs:56-58) under k = 0..5, times 8 (the estimator keys of coop_step)
    for (uint32_t e = tid; e < SE_NCTX; e += SE_THREADS) {
        uint32_t c[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) c[k] = e == SE_NULL_E ? 0u : ((e >> k) + 1u + (uint32_t)k) << 3;
        S.lut03[e] = make_uint2(c[0] | (c[1] << 16), c[2] | (c[3] << 16));
        S.lut45[e] = c[4] | (c[5] << 16);
    }

    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t parity = 0;   // bit i: phase of S.bar[i] the next wait is for

    for (;;) {
        __syncthreads();
        if (tid == 0) S.plane = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t p = S.plane;
        if (p >= a.nplanes) break;
        const uint8_t *plane = a.pixels + (size_t)p * a.npix;
        uint32_t *slot = reinterpret_cast<uint32_t *>(a.temp + (size_t)p * a.slot_bytes);

        for (uint32_t i = tid; i < SE_NCTX * 3; i += SE_THREADS) (&S.state[0][0])[i] = 0u;
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
        load_band(pixbuf0, a, plane, 0, min((uint32_t)SE_BAND, a.npix), &S.bar[0]);
        for (uint32_t b = 0; b < nbands; b++) {
            const uint32_t start = b * SE_BAND, cnt = min((uint32_t)SE_BAND, a.npix - start);
            const uint8_t *pb = pixbuf0 + (b & 1u) * buf_bytes + a.halo_cap;
            if (clk) tlast = clock64();
            if (a.vec16) {
                mbar_wait(&S.bar[b & 1u], (parity >> (b & 1u)) & 1u);
                parity ^= 1u << (b & 1u);
            } else {
                cp_async_wait_all();
            }
            __syncthreads();
            SE_CLK(0);
            if (b + 1 < nbands)
                load_band(pixbuf0 + ((b + 1) & 1u) * buf_bytes, a, plane, start + SE_BAND, min((uint32_t)SE_BAND, a.npix - start - SE_BAND), &S.bar[(b + 1) & 1u]);
            if (tid == 0) { S.ntask = 0; S.nlong = 0; S.nshort = 0; S.task = 0; }
            if (tid < (uint32_t)(SE_BAND / 128 + 8)) S.roundcnt[tid] = 0u;

            // ---- group: classify, rank and scatter, every warp on its own 512 pixels --------------------------------
            if (!(a.dbg & 4u)) {
                uint32_t *cntw = S.wseg[wid];
                uint8_t *ecw = S.ec[wid];
#pragma unroll
                for (int q = 0; q < SE_NCTX / 32; q++) cntw[q * 32 + lane] = 0u;
#pragma unroll
                for (int q = 0; q < SE_WREG / 128; q++) reinterpret_cast<uint32_t *>(ecw)[q * 32 + lane] = 0xffffffffu;   // null elements
                __syncwarp();
                const uint32_t wp0 = wid * SE_WPIX;
                // classify: every thread takes 16 consecutive pixels (the 512 of a warp's 32 threads are the warp's own).  A run
                // inside one row below the second one reads its 16 samples and the 16 above them as eight words and carries the
                // left neighbour in a register; when it starts a row its first pixel takes up and up-up (misc.rs:15-17).
                // Everything else (the first two rows, runs across a row end, the ragged end of the image) goes pixel by pixel
                // through classify_any
                {
                    const uint32_t wr = a.w;
                    const uint32_t j0 = tid * SE_PPT, i0 = start + j0;
                    const uint32_t y0 = i0 / wr, x0 = i0 - y0 * wr;
                    uint4 *iw = reinterpret_cast<uint4 *>(&S.info[info_index(j0)]);
                    if (y0 >= 2 && x0 + (uint32_t)SE_PPT <= wr && j0 + (uint32_t)SE_PPT <= cnt) {
                        const uint32_t *cw = reinterpret_cast<const uint32_t *>(pb + j0);       // the band starts on a 16-byte boundary
                        const uint32_t *uw = reinterpret_cast<const uint32_t *>(pb + j0 - w);   // the width is a multiple of four
                        int left, upup = 0;
                        if (x0 == 0) { left = -1; upup = j0 >= wr ? (int)pb[(int)j0 - 2 * w] : (int)plane[i0 - 2u * wr]; }
                        else left = pb[(int)j0 - 1];
#pragma unroll
                        for (int q = 0; q < SE_PPT / 4; q++) {
                            const uint32_t c = cw[q], u = uw[q];
                            uint32_t wd[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const int p = (int)((c >> (8 * k)) & 255u), up = (int)((u >> (8 * k)) & 255u);
                                if (q == 0 && k == 0) wd[k] = left < 0 ? make_info(p, up, upup) : make_info(p, left, up);
                                else wd[k] = make_info(p, left, up);
                                left = p;
                            }
                            iw[q] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                        }
                    } else {
#pragma unroll 1
                        for (int k = 0; k < SE_PPT; k++)
                            S.info[info_index(j0) + k] = classify_any(pb, (int)(j0 + k), i0 + k, w, a.npix, plane);
                    }
                }
                __syncwarp();
                // rank: lane = pixel, 32 consecutive pixels per step: the stable rank of an out-of-range pixel among the warp's
                // pixels of its context is one match over the warp (pixels that are not coded out of range get a key of their
                // own: a group of one, nothing counted) and one counter per context
                {
                    uint32_t *ij = &S.info[info_index(wp0 + lane)];          // info_index advances by 40 words per 32 pixels
#pragma unroll 1
                    for (int s0 = 0; s0 < SE_WSTEPS; s0 += 4, ij += 160) {
                        // four steps at a time: the matches are issued back to back (their latency is long), the counters follow in order
                        uint32_t wd[4], grp[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) wd[u] = ij[40 * u];
#pragma unroll
                        for (int u = 0; u < 4; u++) grp[u] = __match_any_sync(0xffffffffu, (wd[u] >> 31) ? (wd[u] >> 12) & 255u : 256u);   // one group for all the others: the match costs by distinct keys
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const bool oor = wd[u] >> 31;
                            const uint32_t delta = (wd[u] >> 12) & 255u;
                            const int leader = __ffs(grp[u]) - 1;
                            uint32_t prev = 0;
                            if (oor && (int)lane == leader) { prev = cntw[delta]; cntw[delta] = prev + __popc(grp[u]); }
                            prev = __shfl_sync(0xffffffffu, prev, leader);
                            if (oor) ij[40 * u] = wd[u] | (prev + __popc(grp[u] & lt));
                            __syncwarp();
                        }
                    }
                }
                // bases of the warp's segments: exclusive prefix of the counts, every segment padded to a multiple of four
                {
                    uint4 *c4 = reinterpret_cast<uint4 *>(cntw + 8u * lane);
                    const uint4 ca = c4[0], cb4 = c4[1];
                    const uint32_t n[8] = {ca.x, ca.y, ca.z, ca.w, cb4.x, cb4.y, cb4.z, cb4.w};
                    uint32_t sum = 0, ex[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) { ex[q] = sum; sum += (n[q] + 3u) & ~3u; }
                    uint32_t inc = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= (uint32_t)o) inc += t;
                    }
                    const uint32_t base = inc - sum;
                    c4[0] = make_uint4(((base + ex[0]) << 16) | n[0], ((base + ex[1]) << 16) | n[1], ((base + ex[2]) << 16) | n[2], ((base + ex[3]) << 16) | n[3]);
                    c4[1] = make_uint4(((base + ex[4]) << 16) | n[4], ((base + ex[5]) << 16) | n[5], ((base + ex[6]) << 16) | n[6], ((base + ex[7]) << 16) | n[7]);
                }
                __syncwarp();
                {
                    uint32_t *iw = &S.info[info_index(wp0 + lane)];
#pragma unroll 4
                    for (int s = 0; s < SE_WSTEPS; s++, iw += 40) {
                        const uint32_t wd = *iw;
                        if (wd >> 31) {
                            const uint32_t pos = (cntw[(wd >> 12) & 255u] >> 16) + (wd & 0xfffu);
                            ecw[pos] = (uint8_t)((wd >> 21) & 511u);
                            *iw = (wd & 0xffe00000u) | pos;
                        }
                    }
                }
            } else {
                for (int s = 0; s < SE_WSTEPS; s++) S.info[info_index(wid * SE_WPIX + 32u * s + lane)] = SE_INFO_NONE;
            }
            __syncthreads();
            SE_CLK(1);

            // ---- chains: one thread per context: length of its chain in this band, work lists -------------------------
            if (!(a.dbg & 4u) && tid < (uint32_t)SE_NCTX) {
                uint32_t vlen = 0;
#pragma unroll
                for (int q = 0; q < SE_WARPS; q++) vlen += ((S.wseg[q][tid] & 0xffffu) + 3u) & ~3u;
                if (vlen >= SE_LONG) {
                    const uint32_t nst = (vlen + 127u) >> 7;
                    S.longc[atomicAdd(&S.nlong, 1u)] = (uint16_t)(tid | (nst << 8));
                    atomicAdd(&S.ntask, nst);
                    for (uint32_t s = 0; s < nst; s++) atomicAdd(&S.roundcnt[s], 1u);
                    S.stepdone[tid] = 0u;
                } else if (vlen) {
                    S.shortlist[atomicAdd(&S.nshort, 1u)] = (uint16_t)tid;
                }
            }
            __syncthreads();
            SE_CLK(2);

            // ---- walk: steps of long chains as tasks, short chains one per lane -----------------------------------------
            if (!(a.dbg & 1u)) {
                const uint32_t nt = S.ntask, nl = S.nlong, ns = S.nshort;
                for (;;) {
                    uint32_t t = 0;
                    if (lane == 0) t = atomicAdd(&S.task, 1u);
                    t = __shfl_sync(0xffffffffu, t, 0);
                    if (t < nt) {
                        // tasks in step-major order (step 0 of every long chain, then step 1 of those that have one, ...): by the
                        // time a warp has prepared step s of a chain, step s - 1 was claimed a round earlier and is done or nearly
                        uint32_t s = 0, i = t;
                        for (uint32_t n = S.roundcnt[0]; i >= n; n = S.roundcnt[++s]) i -= n;
                        uint32_t c = 0;
                        for (uint32_t base = 0;; base += 32) {   // the i-th long chain that has a step s
                            const uint32_t e = base + lane < nl ? (uint32_t)S.longc[base + lane] : 0u;
                            const uint32_t bal = __ballot_sync(0xffffffffu, (e >> 8) > s);
                            const uint32_t nb = __popc(bal);
                            if (i < nb) { c = __shfl_sync(0xffffffffu, e, __fns(bal, 0, i + 1)) & 255u; break; }
                            i -= nb;
                        }
                        coop_step(S, c, s, lane);
                    } else {
                        const uint32_t si = (t - nt) * 32u;
                        if (si >= ns) break;
                        serial_walk(S, si + lane < ns ? (uint32_t)S.shortlist[si + lane] : SE_NONE);
                        __syncwarp();
                    }
                }
            }
            SE_CLK(3);
            __syncthreads();
            SE_CLK(4);
            // the walk is over: its segment tables become the bit window (zeroed here, filled after the next barrier)
            for (uint32_t i = tid; i < SE_OUT_WORDS + 4; i += SE_THREADS) out[i] = 0u;

            // ---- code: 16 consecutive pixels per thread, four at a time; the records take the place of the info words -------
            uint32_t mylen = 0;
            uint4 *const iw = reinterpret_cast<uint4 *>(&S.info[info_index(tid * SE_PPT)]);   // my 16 words are contiguous
            {
                const uint8_t *ecw = S.ec[tid >> 5];   // the warp region my 16 pixels were grouped into: (16 * tid) / 512
#pragma unroll 1
                for (int q = 0; q < SE_PPT / 4; q++) {
                    const uint4 v = iw[q];
                    uint4 r;
                    if (a.dbg & 2u) { r.x = r.y = r.z = r.w = (v.x & 1u) << 22; }
                    else { r.x = make_record(v.x, ecw); r.y = make_record(v.y, ecw); r.z = make_record(v.z, ecw); r.w = make_record(v.w, ecw); }
                    mylen += rec_len(r.x) + rec_len(r.y) + rec_len(r.z) + rec_len(r.w);
                    iw[q] = r;
                }
            }
            uint32_t inc = mylen;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += t;
            }
            if (lane == 31) S.wsum[wid] = inc;
            if (tid == 0) out[0] = carry;
            __syncthreads();
            SE_CLK(5);
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
                // one window: my codes are concatenated in registers and leave as whole words (atomicOr: the first and the last
                // word of my range are shared with my neighbours; about two words per thread and band)
                uint32_t wi = pos >> 5, sh = pos & 31u, wv = 0;
#pragma unroll 1
                for (int q = 0; q < SE_PPT / 4; q++) {
                    const uint4 r4 = iw[q];
                    const uint32_t r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t len = rec_len(r[k]);
                        if (len > (uint32_t)REC_SHORT_MAX) {
                            // long unary run: my partial word first, then field by field
                            if (wv) atomicOr(&out[wi], wv);
                            const uint32_t at = (wi << 5) + sh;
                            pack_long(out, r[k], at);
                            const uint32_t np2 = at + len;
                            wi = np2 >> 5; sh = np2 & 31u; wv = 0;
                            continue;
                        }
                        const uint32_t left = ((r[k] & 0x3fffffu) << 1) << (31u - len);   // code word, MSB aligned (nothing for len 0)
                        wv |= left >> sh;
                        const uint32_t nsh = sh + len;
                        if (nsh >= 32u) {                                                  // the word is full: sh >= 10 here
                            atomicOr(&out[wi], wv);
                            wi++;
                            wv = left << (32u - sh);
                        }
                        sh = nsh & 31u;
                    }
                }
                if (wv) atomicOr(&out[wi], wv);
                __syncthreads();
                const uint32_t nfull = win_bits >> 5;
                for (uint32_t j = tid; j < nfull; j += SE_THREADS) {
                    if (wpos + j < slot_words) slot[wpos + j] = bswap32(out[j]);
                }
                if (wpos + nfull >= slot_words) ovf = 1;
                carry = out[nfull];
                wpos += nfull;
            } else {
                // more than 16 bits per pixel (rare): the band leaves through several windows, out of line
                pack_windows(out, reinterpret_cast<const uint32_t *>(iw), pos, win_bits, slot, slot_words, wpos, carry, ovf);
            }
            carrybits = win_bits & 31u;
            total_bits += band_bits;
            SE_CLK(6);
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
    if (clk)
        for (int i = 0; i < 8; i++) atomicAdd(a.clocks + i, (unsigned long long)tclk[i]);
#undef SE_CLK
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
    a.dbg = ctx->stream_dbg;
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
                  uint32_t *d_ticket, uint64_t *d_running, uint8_t *d_temp, uint8_t *target, const uint64_t *base_ptr, uint64_t target_cap,
                  unsigned long long *d_clocks) {
    cudaStream_t st = ctx->stream;
    {
        StageScope s(ctx, ST_STREAM);
        FELICS_CUDA_TRY(cudaMemsetAsync(d_ticket, 0, sizeof(uint32_t), st));
        StreamArgs a = pl.a;
        a.pixels = d_pixels; a.temp = d_temp; a.sizes = d_sizes; a.flags = d_flags; a.ticket = d_ticket; a.nplanes = (uint32_t)ni; a.clocks = d_clocks;
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
    unsigned long long *clocks;
    uint8_t *temp;
};

int stream_scratch(felics_ctx *ctx, size_t n, size_t temp_bytes, StreamScratch &sc) {
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 256); const size_t at = off; off += bytes; return at; };
    const size_t o_sizes = take(n * sizeof(uint32_t)), o_flags = take(n * sizeof(uint32_t)), o_off = take((n + 1) * sizeof(uint64_t));
    const size_t o_ctr = take(256), o_temp = take(temp_bytes);
    int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, off);
    if (rc) return rc;
    uint8_t *b = (uint8_t *)ctx->scratch;
    sc.sizes = (uint32_t *)(b + o_sizes); sc.flags = (uint32_t *)(b + o_flags); sc.off = (uint64_t *)(b + o_off);
    sc.ticket = (uint32_t *)(b + o_ctr); sc.running = (uint64_t *)(b + o_ctr + 16); sc.clocks = (unsigned long long *)(b + o_ctr + 64); sc.temp = b + o_temp;
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
    FELICS_CUDA_TRY(cudaMemsetAsync(sc.ticket, 0, 256, st));
    for (size_t first = 0; first < n; first += sub_max) {
        const size_t ni = std::min(sub_max, n - first);
        rc = stream_launch(ctx, pl, ni, (const uint8_t *)d_pixels + first * (size_t)pl.npix, sc.sizes + first, sc.flags + first, sc.off + first, sc.ticket,
                           sc.running, sc.temp, d_arena, nullptr, arena_cap, sc.clocks);
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
    if (ctx->stream_dbg & 8u) {   // timing experiment: share of every phase in the blocks' cycles
        unsigned long long hc[8];
        if (cudaMemcpy(hc, sc.clocks, sizeof(hc), cudaMemcpyDeviceToHost) == cudaSuccess) {
            double tot = 0;
            for (int i = 0; i < 7; i++) tot += (double)hc[i];
            fprintf(stderr, "stream phases (%% of block cycles): load-wait %.1f group %.1f chains %.1f walk %.1f walk-barrier %.1f code+scan %.1f pack+flush %.1f\n", 100 * hc[0] / tot,
                    100 * hc[1] / tot, 100 * hc[2] / tot, 100 * hc[3] / tot, 100 * hc[4] / tot, 100 * hc[5] / tot, 100 * hc[6] / tot);
        }
    }
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
    // eight to thirty-two sub-batches to overlap (what does not overlap is the first copy in, the last encode and the last
    // copy out), none smaller than two waves of blocks
    const size_t nsub_want = std::min<size_t>(32, std::max<size_t>(8, n / ((size_t)2 * 148 * SE_CTAS_PER_SM)));
    size_t sub = std::max<size_t>(32, (n + nsub_want - 1) / nsub_want);
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
    SE_TRY(cudaMemsetAsync(sc.ticket, 0, 256, st));
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
                           sc.temp, (uint8_t *)ctx->stage_out[slot], sc.off + first, sub * pl.slot_bytes, sc.clocks);
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
