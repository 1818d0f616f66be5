// Segment hops for chains the speculative walk rejected (sp_walk.cuh): exact, one table lookup per flip-free segment.
//
// A rejected chain still follows its presumed binding counter kb through most segments; only where another counter
// binds ("flip") is the one-dimensional table wrong.  For a segment and an entry value x of kb the solo trajectory
// (halving elements h_1 < ... < h_n) is known.  It is also the TRUE trajectory iff at every h_j every other counter k
// has passed 1024.  With D_i the cost of counter k between the halvings and S_in its value at the segment start,
//     value at h_j = ((S_in + sum_{i<j} D_i 2^(i-1)) >> (j-1)) + D_j          (nested floors compose exactly)
// is monotone in S_in, so "no flip in this segment" is S_in >= theta_k with
//     theta_k = max_j ((1025 - D_j) 2^(j-1) - sum_{i<j} D_i 2^(i-1)),
// and the value at the segment end is ((S_in + A_k) >> n) + Dtail_k with A_k = sum_{i<=n} D_i 2^(i-1).
// k_hop_maps tabulates (n, x_out, theta_k, A_k, Dtail_k) for every x of every segment of the rejected chains; the
// serial walker (k_walk) then crosses a segment whose entry state passes the six threshold tests with one lookup,
// and walks the others element-exactly as before.  k_hop_emit / k_hop_finish afterwards write the epoch records
// of the hopped segments in parallel (positions from the solo walk, counters from the closed form).
#pragma once

namespace felics {

constexpr int HOP_MAX_N = 44;                 // halvings per segment representable in 64-bit sums
constexpr uint32_t HOP_INVALID = 0xFFFFu;

struct HopTables {
    uint32_t *xn;                 // [seg slot][1024]: x_out | n << 16 (n = HOP_INVALID: not usable)
    uint32_t *theta;              // [seg slot][6][1024]
    unsigned long long *A;        // [seg slot][6][1024]
    uint32_t *dt;                 // [seg slot][6][1024]
    uint32_t *log;                // [seg slot][8]: entry values of the six counters, epochs recorded at entry, 1 = hopped
    uint32_t *segkb;              // [seg slot]: the segment's own presumed binding counter (smallest cost inside the segment)
    uint32_t *ready;              // [0]: 1 once the tables are complete
    const uint32_t *pc2desc;      // [plane*512 + context] -> descriptor index (0xFFFFFFFF: none)
};

struct HopSmem {
    uint32_t T[(SP_SEG + 1) * 8];             // per boundary j: prefixes of the six counters relative to the segment start
    uint16_t inv[SP_INV_CAP];                 // inverse prefix table of kb
};

// 1. tables: one block per segment of a rejected chain, one thread per entry value x = 1..1024
__global__ void __launch_bounds__(1024, 1) k_hop_maps(SpArgs a, HopTables h) {
    extern __shared__ __align__(16) unsigned char hop_smem[];
    HopSmem &S = *reinterpret_cast<HopSmem *>(hop_smem);
    const uint32_t slot = blockIdx.x;
    if (slot >= a.counts[1]) return;
    const uint32_t di = a.seg_desc[slot];
    if (!a.hop_all && a.chain_fail[di] != SP_OK - 1) return;     // only chains that failed the verification
    const SpDesc d = a.desc[di];
    const uint32_t seg = slot - d.seg0;
    const uint32_t e0 = seg * SP_SEG;
    const uint32_t nel = min(d.count - e0, (uint32_t)SP_SEG);
    const uint32_t g0 = d.gbase + e0;
    const uint4 s0 = a.blk_rec4[(size_t)(g0 >> 5) * 4], s1 = a.blk_rec4[(size_t)(g0 >> 5) * 4 + 1];
    uint4 *rows = reinterpret_cast<uint4 *>(S.T);
    if (threadIdx.x == 0) { rows[0] = make_uint4(0, 0, 0, 0); rows[1] = make_uint4(0, 0, 0, 0); }
#pragma unroll
    for (int i = 0; i < SP_SEG / 1024; i++) {
        const uint32_t j = threadIdx.x + i * 1024u;
        if (j < nel) {
            const uint32_t g = g0 + j;
            const uint4 f = a.fine[g];
            const uint4 x0 = a.blk_rec4[(size_t)(g >> 5) * 4], x1 = a.blk_rec4[(size_t)(g >> 5) * 4 + 1];
            rows[(j + 1) * 2] = make_uint4(x0.x + (f.x & 0xffffu) - s0.x, x0.y + (f.x >> 16) - s0.y, x0.z + (f.y & 0xffffu) - s0.z, x0.w + (f.y >> 16) - s0.w);
            rows[(j + 1) * 2 + 1] = make_uint4(x1.x + (f.z & 0xffffu) - s1.x, x1.y + (f.z >> 16) - s1.y, 0u, 0u);
        }
    }
    __syncthreads();
    // the segment's own binding candidate: the counter with the smallest cost inside the segment (ties: larger k);
    // natural images are not stationary, the chain-wide choice of the speculative walk is often wrong locally
    uint32_t kb = 0;
    {
        uint32_t best = 0xffffffffu;
        for (uint32_t k = 0; k < NK; k++) {
            const uint32_t tot = S.T[nel * 8 + k];
            if (tot <= best) { best = tot; kb = k; }
        }
    }
    if (threadIdx.x == 0) h.segkb[slot] = kb;
#pragma unroll
    for (int i = 0; i < SP_SEG / 1024; i++) {
        const uint32_t j = threadIdx.x + i * 1024u;
        if (j < nel) {
            const uint32_t lo = S.T[j * 8 + kb], hi = min(S.T[(j + 1) * 8 + kb], (uint32_t)SP_INV_CAP);
            for (uint32_t q = lo; q < hi; q++) S.inv[q] = (uint16_t)j;
        }
    }
    __syncthreads();
    const uint32_t tend = S.T[nel * 8 + kb];
    const uint32_t x0 = threadIdx.x + 1u;
    uint32_t x = x0, p = 0, n = 0;
    uint32_t tp[NK] = {0, 0, 0, 0, 0, 0};          // prefixes at the start of the current epoch
    unsigned long long A[NK] = {0, 0, 0, 0, 0, 0};
    long long th[NK] = {0, 0, 0, 0, 0, 0};
    bool ok = true;
    for (;;) {
        uint32_t tkb = tp[0];
        tkb = kb == 1 ? tp[1] : tkb; tkb = kb == 2 ? tp[2] : tkb; tkb = kb == 3 ? tp[3] : tkb; tkb = kb == 4 ? tp[4] : tkb; tkb = kb == 5 ? tp[5] : tkb;
        const uint32_t q = tkb + (HALVE_AT - x);               // next halving of kb: first j with T[j + 1][kb] > q
        if (q >= tend) break;
        uint32_t j;
        if (q < (uint32_t)SP_INV_CAP) {
            j = S.inv[q];
        } else {
            uint32_t lo = p, hi = nel - 1;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (S.T[(mid + 1) * 8 + kb] > q) hi = mid; else lo = mid + 1; }
            j = lo;
        }
        if (n >= (uint32_t)HOP_MAX_N) { ok = false; break; }
        const uint4 r0 = rows[(j + 1) * 2], r1 = rows[(j + 1) * 2 + 1];
        const uint32_t row[NK] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
        uint32_t dkb = 0;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            const uint32_t D = row[k] - tp[k];
            const long long cand = ((1025ll - (long long)D) << n) - (long long)A[k];   // smallest entry value that has passed 1024 here
            th[k] = max(th[k], cand);
            A[k] += (unsigned long long)D << n;
            dkb = (uint32_t)k == kb ? D : dkb;
            tp[k] = row[k];
        }
        x = (x + dkb) >> 1;
        p = j + 1;
        n++;
    }
    const uint4 z0 = rows[nel * 2], z1 = rows[nel * 2 + 1];
    const uint32_t zend[NK] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y};
    const size_t e = (size_t)slot * 1024 + threadIdx.x;
    uint32_t xo = x;
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const uint32_t dtail = zend[k] - tp[k];
        if ((uint32_t)k == kb) xo = x + dtail;
        const size_t t = ((size_t)slot * NK + k) * 1024 + threadIdx.x;
        h.theta[t] = th[k] > 0x7fffffffll ? 0xffffffffu : (uint32_t)th[k];
        h.A[t] = A[k];
        h.dt[t] = dtail;
    }
    h.xn[e] = (xo & 0xffffu) | ((ok ? n : HOP_INVALID) << 16);
}

__global__ void k_hop_ready(HopTables h) {
    __threadfence();
    *h.ready = 1u;
}

// 2. positions of the hopped segments: the solo walk from the logged entry value (one warp per segment, as k_sp_emit)
__global__ void __launch_bounds__(128) k_hop_emit(SpArgs a, HopTables h) {
    const uint32_t slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (slot >= a.counts[1]) return;
    const uint32_t *lg = h.log + (size_t)slot * 8;
    if (lg[7] != 1u) return;
    const uint32_t di = a.seg_desc[slot];
    const SpDesc d = a.desc[di];
    const uint32_t seg = slot - d.seg0, kb = h.segkb[slot];
    const uint32_t e0 = seg * SP_SEG;
    const uint32_t nel = min(d.count - e0, (uint32_t)SP_SEG);
    const uint32_t g0 = d.gbase + e0, b0 = g0 >> 5, nblk = (nel + 31u) >> 5;
    const uint32_t tstart = reinterpret_cast<const uint32_t *>(a.blk_rec4 + (size_t)b0 * 4)[kb];
    uint32_t tend[SP_SEG / 1024];
#pragma unroll
    for (int i = 0; i < SP_SEG / 1024; i++) {
        const uint32_t b = lane + 32u * i;
        tend[i] = b < nblk ? reinterpret_cast<const uint32_t *>(a.blk_rec4 + (size_t)(b0 + b) * 4)[8 + kb] - tstart : 0u;
    }
    uint4 *rows = a.trow + (size_t)d.ep0 * 2;
    uint32_t x = lg[kb], n = lg[6] - 1u, tp = 0;      // the first new epoch gets index lg[6]
    for (;;) {
        const uint32_t q = tp + (HALVE_AT - x);
        int B = -1;
#pragma unroll
        for (int i = 0; i < SP_SEG / 1024; i++) {
            if (B < 0) {
                const uint32_t m = __ballot_sync(0xffffffffu, lane + 32u * i < nblk && tend[i] > q);
                if (m) B = 32 * i + __ffs(m) - 1;
            }
        }
        if (B < 0) break;
        const uint32_t g = g0 + (uint32_t)B * 32u + lane;
        const uint4 f = a.fine[g];
        const uint4 x0 = a.blk_rec4[(size_t)(b0 + B) * 4], x1 = a.blk_rec4[(size_t)(b0 + B) * 4 + 1];
        const uint32_t T[NK] = {x0.x + (f.x & 0xffffu), x0.y + (f.x >> 16), x0.z + (f.y & 0xffffu), x0.w + (f.y >> 16), x1.x + (f.z & 0xffffu), x1.y + (f.z >> 16)};
        uint32_t tkb = T[0];
        tkb = kb == 1 ? T[1] : tkb; tkb = kb == 2 ? T[2] : tkb; tkb = kb == 3 ? T[3] : tkb; tkb = kb == 4 ? T[4] : tkb; tkb = kb == 5 ? T[5] : tkb;
        tkb -= tstart;
        const uint32_t m = __ballot_sync(0xffffffffu, (uint32_t)B * 32u + lane < nel && tkb > q);
        if (!m) break;                                   // cannot happen: the block's last element passes
        const int hh = __ffs(m) - 1;
        const uint32_t t1 = __shfl_sync(0xffffffffu, tkb, hh);
        x = (x + t1 - tp) >> 1;
        tp = t1;
        n++;
        if ((int)lane == hh) {
            rows[(size_t)n * 2] = make_uint4(T[0], T[1], T[2], T[3]);
            rows[(size_t)n * 2 + 1] = make_uint4(T[4], T[5], g + 1u, 0u);
        }
    }
}

// 3. epoch records and block epochs of the hopped segments: counters at every new epoch start in closed form
__global__ void __launch_bounds__(128) k_hop_finish(SpArgs a, HopTables h) {
    const uint32_t slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (slot >= a.counts[1]) return;
    const uint32_t *lg = h.log + (size_t)slot * 8;
    if (lg[7] != 1u) return;
    const uint32_t di = a.seg_desc[slot];
    const SpDesc d = a.desc[di];
    const uint32_t seg = slot - d.seg0;
    const uint32_t e0 = seg * SP_SEG;
    const uint32_t nel = min(d.count - e0, (uint32_t)SP_SEG);
    const uint32_t g0 = d.gbase + e0;
    const uint32_t x_in = lg[h.segkb[slot]];
    const uint32_t n = h.xn[(size_t)slot * 1024 + (x_in - 1u)] >> 16;   // halvings inside the segment
    const uint32_t nep0 = lg[6];                                         // index of the first new epoch
    const uint4 s0 = a.blk_rec4[(size_t)(g0 >> 5) * 4], s1 = a.blk_rec4[(size_t)(g0 >> 5) * 4 + 1];
    const uint32_t Tb[NK] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y};        // prefixes at the segment start
    const uint4 *rows = a.trow + (size_t)d.ep0 * 2;
    unsigned long long carry[NK] = {0, 0, 0, 0, 0, 0};                   // sum of D_i 2^(i-1) over the chunks done
    uint32_t prev_last[NK];                                               // prefixes at the start of the last epoch of the previous chunk
#pragma unroll
    for (int k = 0; k < NK; k++) prev_last[k] = Tb[k];
    uint32_t first_start = g0 + nel;                                      // start of the first new epoch (segment end if none)
    for (uint32_t c0 = 0; c0 < n; c0 += 32) {
        const uint32_t j = c0 + lane;                                     // 0-based index of my new epoch inside the segment
        const bool have = j < n;
        uint32_t T[NK] = {0, 0, 0, 0, 0, 0}, start = 0, nxt = 0;
        if (have) {
            const uint4 r0 = rows[(size_t)(nep0 + j) * 2], r1 = rows[(size_t)(nep0 + j) * 2 + 1];
            T[0] = r0.x; T[1] = r0.y; T[2] = r0.z; T[3] = r0.w; T[4] = r1.x; T[5] = r1.y;
            start = r1.z;
            nxt = j + 1 < n ? rows[(size_t)(nep0 + j + 1) * 2 + 1].z : g0 + nel;
        }
        if (c0 == 0) first_start = __shfl_sync(0xffffffffu, start, 0);
        uint32_t S[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) {
            // D of the epoch that ends with my halving = T(my start) - T(previous epoch start)
            uint32_t tprev = __shfl_up_sync(0xffffffffu, T[k], 1);
            if (lane == 0) tprev = prev_last[k];
            unsigned long long v = have ? ((unsigned long long)(T[k] - tprev) << j) : 0ull;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= (uint32_t)o) v += t;
            }
            S[k] = (uint32_t)(((unsigned long long)lg[k] + carry[k] + v) >> (j + 1));   // counter k at the start of my epoch
            carry[k] += __shfl_sync(0xffffffffu, v, 31);
            prev_last[k] = __shfl_sync(0xffffffffu, T[k], 31);
        }
        if (have) {
            uint4 *rec = a.ep_rec + (size_t)(d.ep0 + nep0 + j) * 2;
            rec[0] = make_uint4(S[0] - T[0], S[1] - T[1], S[2] - T[2], S[3] - T[3]);
            rec[1] = make_uint4(S[4] - T[4], S[5] - T[5], start, 0u);
            for (uint32_t B = (start + 31) >> 5; B < ((nxt + 31) >> 5); B++) a.blk_epoch[B] = d.ep0 + nep0 + j;
        }
    }
    // blocks from the segment start up to the first new epoch still belong to the epoch that was open at entry
    for (uint32_t B = (g0 >> 5) + lane; B < ((first_start + 31) >> 5); B += 32) a.blk_epoch[B] = d.ep0 + nep0 - 1u;
}

}  // namespace felics
