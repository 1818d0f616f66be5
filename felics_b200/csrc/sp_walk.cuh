// Speculative parallel epoch walk ("SP") for long chains, exact by verification.
//
// The halving recurrence of KEstimator (parameter_selection.rs:49-64) is serial per context
// chain, but on stationary data the counter that passes 1024 LAST (the "binding" counter) is the
// same for thousands of epochs.  If only that counter kb mattered, the recurrence would be one
// dimensional: x += cost_kb(e); if x > 1024: x >>= 1.  Its value at an element boundary lies in
// [0, 1024], so for a segment of 4096 elements the whole behaviour is a table
//     x_in (0..1024)  ->  (x_out, number of halvings)
// (1025 independent trajectories per segment, all segments in parallel).  Composing the tables
// gives the value of kb at every segment start, a second pass re-walks each segment from its true
// x_in and emits the halving positions, and then the states of ALL SIX counters at every epoch
// start follow from the positions in closed form: with D_e the cost sums of epoch e,
//     S_{e+1} = (S_e + D_e) >> 1     =>     S_{e+32} = (S_e + sum_j D_{e+j} << j) >> 32   (exact)
// Every halving is then checked exactly (all six counters > 1024 after the element, kb <= 1024
// just before it; no halving left in the open last epoch).  By induction over the epochs a chain
// whose every epoch passes is identical to the sequential reference; a chain with any failing
// epoch is left to the serial walker (k_walk).  Results are therefore always exact.
#pragma once

namespace felics {

constexpr uint32_t SP_MIN_COUNT = 32768;   // chains at least this long are tried
constexpr int SP_SEG = 4096;               // elements per segment
constexpr int SP_GRP = 32;                 // segments per composed group
constexpr int SP_DOM = 1025;               // x in 0..1024
constexpr int SP_INV_CAP = 16384;
constexpr int SP_MAX_PLANES = 8;           // only single big images / small batches take this path
constexpr uint32_t SP_OK = 0xFFFFFFFFu;    // chain_fail value of a chain that is (so far) accepted

struct SpDesc {
    uint32_t pc;        // plane * 512 + context
    uint32_t count;     // elements in the chain
    uint32_t gbase;     // global element index of the chain start (multiple of 32)
    uint32_t ep0;       // first epoch record of the chain
    uint32_t ep_room;   // epoch records available
    uint32_t seg0, nseg;
    uint32_t grp0, ngrp;
    uint32_t eb0, neb;  // 32-epoch blocks (upper bound)
    uint32_t kb;        // presumed binding counter
};

struct SpSizes {        // upper bounds used for allocation and grid sizes
    uint32_t max_desc, max_seg, max_grp, max_eb;
};

struct SpArgs {
    const uint32_t *chain_count;
    const uint32_t *chain_base;
    const uint4 *fine;
    const uint4 *blk_rec4;
    SpDesc *desc;
    uint32_t *seg_desc;         // segment slot -> descriptor index
    uint32_t *grp_desc;
    uint32_t *eb_desc;
    uint32_t *counts;           // [0] descriptors, [1] segment slots, [2] group slots, [3] epoch-block slots
    uint32_t *map;              // [seg slot][1025]: x_out | halvings << 16
    uint32_t *pre;              // [seg slot][1025]: state at the segment start, as a function of the group-entry x
    uint32_t *gmap;             // [group slot][1025]
    uint32_t *grp_xin;          // [group slot]: x at the group start
    uint32_t *grp_nbefore;      // [group slot]: halvings before the group
    uint32_t *chain_n;          // [desc]: halvings in the whole chain
    uint32_t *chain_fail;       // [desc]: SP_OK or a rejection marker
    unsigned long long *ablk;   // [epoch block][8]: sum_j D(epoch 32b + j) << j
    uint4 *trow;                // [epoch record][2]: cost prefixes T0..T5 at the epoch start, start element
    uint4 *ep_rec;
    uint32_t *blk_epoch;
    uint8_t *resolved;          // [plane*512 + context]: 1 when the chain needs no serial walk
    uint32_t *dbg;              // encode counters: [3] chains tried, [4] chains resolved
    uint32_t *pc2desc;          // [plane*512 + context] -> descriptor index (may be null)
    uint32_t hop_all = 0;       // hop tables for every planned chain, before the verification has a verdict (experiment)
    uint32_t np, cap, epcap;
    SpSizes sz;
};

__host__ __device__ inline SpSizes sp_sizes(uint32_t np, uint32_t cap) {
    SpSizes s;
    s.max_desc = np * (cap / SP_MIN_COUNT + 1 < (uint32_t)NBIN ? cap / SP_MIN_COUNT + 1 : (uint32_t)NBIN);
    s.max_seg = np * (cap / SP_SEG) + s.max_desc;
    s.max_grp = s.max_seg / SP_GRP + s.max_desc;
    s.max_eb = (np * (cap / 8) + 16 * s.max_desc) / 32 + s.max_desc;
    return s;
}

// cost prefix T_k after element g - 1 (g > chain start), mod 2^32
__device__ __forceinline__ uint32_t sp_prefix_end(const uint4 *__restrict__ fine, const uint4 *__restrict__ blk_rec4, uint32_t g, uint32_t k) {
    const uint32_t last = g - 1;
    return reinterpret_cast<const uint32_t *>(blk_rec4 + (size_t)(last >> 5) * 4)[k] + reinterpret_cast<const uint16_t *>(fine + last)[k];
}
// cost prefix T_k before the first element of the chain (block aligned)
__device__ __forceinline__ uint32_t sp_prefix_chain_start(const uint4 *__restrict__ blk_rec4, uint32_t gbase, uint32_t k) {
    return reinterpret_cast<const uint32_t *>(blk_rec4 + (size_t)(gbase >> 5) * 4)[k];
}

// 0. plan: one descriptor per long chain, slot tables for the per-segment / per-group / per-epoch-block kernels
__global__ void __launch_bounds__(NBIN) k_sp_plan(SpArgs a) {
    __shared__ uint32_t s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t p = 0; p < a.np; p++) {
        const uint32_t pc = p * NBIN + threadIdx.x;
        const uint32_t count = a.chain_count[pc];
        if (count >= SP_MIN_COUNT) {
            SpDesc d;
            d.pc = pc;
            d.count = count;
            const uint32_t cbase = a.chain_base[pc];
            d.gbase = p * a.cap + cbase;
            d.ep0 = p * a.epcap + cbase / 8 + 16 * threadIdx.x;
            d.ep_room = ((count + 31u) >> 5) * 4 + 16;
            d.nseg = (count + SP_SEG - 1) / SP_SEG;
            d.ngrp = (d.nseg + SP_GRP - 1) / SP_GRP;
            d.neb = d.ep_room / 32 + 1;
            const uint32_t di = atomicAdd(&s_cnt[0], 1u);
            d.seg0 = atomicAdd(&s_cnt[1], d.nseg);
            d.grp0 = atomicAdd(&s_cnt[2], d.ngrp);
            d.eb0 = atomicAdd(&s_cnt[3], d.neb);
            // presumed binding counter: the smallest total cost over the chain (ties: larger k)
            uint32_t best = 0xffffffffu, kb = 0;
            for (uint32_t k = 0; k < NK; k++) {
                const uint32_t tot = sp_prefix_end(a.fine, a.blk_rec4, d.gbase + count, k) - sp_prefix_chain_start(a.blk_rec4, d.gbase, k);
                if (tot <= best) { best = tot; kb = k; }
            }
            d.kb = kb;
            a.desc[di] = d;
            a.chain_fail[di] = SP_OK;
            if (a.pc2desc) a.pc2desc[pc] = di;
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) a.counts[threadIdx.x] = s_cnt[threadIdx.x];
    const uint32_t nd = s_cnt[0];
    __threadfence_block();
    for (uint32_t di = 0; di < nd; di++) {
        const SpDesc d = a.desc[di];   // written above by this block
        for (uint32_t j = threadIdx.x; j < d.nseg; j += NBIN) a.seg_desc[d.seg0 + j] = di;
        for (uint32_t j = threadIdx.x; j < d.ngrp; j += NBIN) a.grp_desc[d.grp0 + j] = di;
        for (uint32_t j = threadIdx.x; j < d.neb; j += NBIN) a.eb_desc[d.eb0 + j] = di;
    }
}

// shared-memory image of one segment for counter kb: tk[j] = cost of the first j elements,
// inv[q] = first element j with tk[j + 1] > q (q < min(total, SP_INV_CAP))
struct SpSegSmem {
    uint32_t tk[SP_SEG + 4];
    uint16_t inv[SP_INV_CAP];
};

template <int THREADS>
__device__ __forceinline__ uint32_t sp_build_segment(SpSegSmem &S, const SpDesc &d, uint32_t seg, const uint4 *__restrict__ fine,
                                                     const uint4 *__restrict__ blk_rec4) {
    constexpr int PER = SP_SEG / THREADS;
    const uint32_t kb = d.kb;
    const uint32_t e0 = seg * SP_SEG;
    const uint32_t nel = min(d.count - e0, (uint32_t)SP_SEG);
    const uint32_t g0 = d.gbase + e0;   // block aligned
    const uint32_t tstart = reinterpret_cast<const uint32_t *>(blk_rec4 + (size_t)(g0 >> 5) * 4)[kb];
    if (threadIdx.x == 0) S.tk[0] = 0;
    uint32_t incl[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {      // all loads in flight together
        const uint32_t j = threadIdx.x + i * THREADS;
        const uint32_t g = g0 + min(j, nel - 1);
        incl[i] = reinterpret_cast<const uint32_t *>(blk_rec4 + (size_t)(g >> 5) * 4)[kb] + reinterpret_cast<const uint16_t *>(fine + g)[kb] - tstart;
    }
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const uint32_t j = threadIdx.x + i * THREADS;
        if (j < nel) S.tk[j + 1] = incl[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const uint32_t j = threadIdx.x + i * THREADS;
        if (j < nel) {
            const uint32_t lo = S.tk[j], hi = min(incl[i], (uint32_t)SP_INV_CAP);
            for (uint32_t q = lo; q < hi; q++) S.inv[q] = (uint16_t)j;
        }
    }
    __syncthreads();
    return nel;
}

// 1. segment maps: one block per segment, one thread per phase x
__global__ void __launch_bounds__(1024) k_sp_maps(SpArgs a) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    SpSegSmem &S = *reinterpret_cast<SpSegSmem *>(sp_smem);
    const uint32_t slot = blockIdx.x;
    if (slot >= a.counts[1]) return;
    const SpDesc d = a.desc[a.seg_desc[slot]];
    const uint32_t seg = slot - d.seg0;
    const uint32_t nel = sp_build_segment<1024>(S, d, seg, a.fine, a.blk_rec4);
    const uint32_t tend = S.tk[nel];
    auto run = [&](uint32_t x0) {
        uint32_t x = x0, p = 0, tp = 0, n = 0;   // tp = tk[p]
        for (;;) {
            const uint32_t q = tp + (HALVE_AT - x);          // next halving: first j with tk[j + 1] > q
            if (q >= tend) { x += tend - tp; break; }
            uint32_t j;
            if (q < (uint32_t)SP_INV_CAP) {
                j = S.inv[q];
            } else {                                          // beyond the table: bisection
                uint32_t lo = p, hi = nel - 1;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (S.tk[mid + 1] > q) hi = mid; else lo = mid + 1; }
                j = lo;
            }
            const uint32_t t1 = S.tk[j + 1];
            x = (x + t1 - tp) >> 1;
            tp = t1;
            p = j + 1;
            n++;
        }
        a.map[(size_t)slot * SP_DOM + x0] = x | (n << 16);
    };
    // a counter is at least 1 at every segment boundary but the chain start, so x = 0 only matters for segment 0
    run(threadIdx.x + 1);
    if (seg == 0 && threadIdx.x == 0) run(0);
}

// 2. compose the maps of 32 consecutive segments; pre[] keeps the state at every segment start
__global__ void __launch_bounds__(1024) k_sp_compose(SpArgs a) {
    const uint32_t gslot = blockIdx.x;
    if (gslot >= a.counts[2]) return;
    const SpDesc d = a.desc[a.grp_desc[gslot]];
    const uint32_t g = gslot - d.grp0;
    const uint32_t s0 = g * SP_GRP, s1 = min(s0 + SP_GRP, d.nseg);
    for (uint32_t x0 = threadIdx.x; x0 < SP_DOM; x0 += blockDim.x) {
        if (x0 == 0 && g != 0) continue;   // x = 0 exists only at the chain start
        uint32_t x = x0, n = 0;
        for (uint32_t s = s0; s < s1; s++) {
            a.pre[(size_t)(d.seg0 + s) * SP_DOM + x0] = x | (n << 16);   // n <= 32 * 4096 / 2 fits 16 bits only if epochs >= 2 elements: checked in k_sp_scan
            const uint32_t m = a.map[(size_t)(d.seg0 + s) * SP_DOM + x];
            x = m & 0xffffu;
            n += m >> 16;
        }
        a.gmap[(size_t)gslot * SP_DOM + x0] = x | (min(n, 0xffffu) << 16);
    }
}

// 3. serial over the groups of one chain (a few dozen dependent loads)
__global__ void k_sp_scan(SpArgs a) {
    const uint32_t di = blockIdx.x * blockDim.x + threadIdx.x;
    if (di >= a.counts[0]) return;
    const SpDesc d = a.desc[di];
    uint32_t x = 0, n = 0;
    bool sat = false;
    for (uint32_t g = 0; g < d.ngrp; g++) {
        a.grp_xin[d.grp0 + g] = x;
        a.grp_nbefore[d.grp0 + g] = n;
        const uint32_t m = a.gmap[(size_t)(d.grp0 + g) * SP_DOM + x];
        x = m & 0xffffu;
        sat = sat || (m >> 16) >= 0xffffu;
        n += m >> 16;
    }
    a.chain_n[di] = n;
    // epoch records needed: n + 1 epochs + sentinel
    a.chain_fail[di] = (!sat && n + 2 < d.ep_room) ? SP_OK : 0u;
}

// 4. re-walk every segment from its true x_in (one warp per segment, two-level search straight from global
//    memory: block-end prefixes, then the 32 elements of the block found); for every epoch that starts inside
//    the segment write the start element and the six cost prefixes at the start (trow)
__global__ void __launch_bounds__(128) k_sp_emit(SpArgs a) {
    const uint32_t slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (slot >= a.counts[1]) return;
    const uint32_t di = a.seg_desc[slot];
    if (a.chain_fail[di] != SP_OK) return;
    const SpDesc d = a.desc[di];
    const uint32_t seg = slot - d.seg0, kb = d.kb;
    const uint32_t e0 = seg * SP_SEG;
    const uint32_t nel = min(d.count - e0, (uint32_t)SP_SEG);
    const uint32_t g0 = d.gbase + e0, b0 = g0 >> 5, nblk = (nel + 31u) >> 5;
    const uint32_t tstart = reinterpret_cast<const uint32_t *>(a.blk_rec4 + (size_t)b0 * 4)[kb];
    uint32_t tend[SP_SEG / 1024];          // relative prefix at the end of blocks lane, lane + 32, ...
#pragma unroll
    for (int i = 0; i < SP_SEG / 1024; i++) {
        const uint32_t b = lane + 32u * i;
        tend[i] = b < nblk ? reinterpret_cast<const uint32_t *>(a.blk_rec4 + (size_t)(b0 + b) * 4)[8 + kb] - tstart : 0u;
    }
    uint4 *rows = a.trow + (size_t)d.ep0 * 2;
    if (seg == 0 && lane == 0) {
        const uint4 r0 = a.blk_rec4[(size_t)b0 * 4], r1 = a.blk_rec4[(size_t)b0 * 4 + 1];
        rows[0] = r0;
        rows[1] = make_uint4(r1.x, r1.y, d.gbase, 0u);
    }
    const uint32_t grp = seg / SP_GRP;
    const uint32_t pr = a.pre[(size_t)slot * SP_DOM + a.grp_xin[d.grp0 + grp]];
    uint32_t x = pr & 0xffffu, n = a.grp_nbefore[d.grp0 + grp] + (pr >> 16), tp = 0;
    for (;;) {
        const uint32_t q = tp + (HALVE_AT - x);              // next halving: first element whose inclusive prefix passes q
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
        if (!m) { atomicMin(&a.chain_fail[di], 0u); break; }   // cannot happen: the block's last element passes
        const int h = __ffs(m) - 1;
        const uint32_t t1 = __shfl_sync(0xffffffffu, tkb, h);
        x = (x + t1 - tp) >> 1;
        tp = t1;
        n++;
        if ((int)lane == h) {                                   // epoch n starts right after my element
            rows[(size_t)n * 2] = make_uint4(T[0], T[1], T[2], T[3]);
            rows[(size_t)n * 2 + 1] = make_uint4(T[4], T[5], g + 1u, 0u);
        }
    }
}

// 5. per block of 32 epochs: A[k] = sum_j D_k(epoch 32b + j) << j   (D = cost of the whole epoch)
__global__ void __launch_bounds__(128) k_sp_blocksum(SpArgs a) {
    const uint32_t slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (slot >= a.counts[3]) return;
    const uint32_t di = a.eb_desc[slot];
    if (a.chain_fail[di] != SP_OK) return;
    const SpDesc d = a.desc[di];
    const uint32_t n = a.chain_n[di];            // halvings; epochs 0..n, epochs 0..n-1 end with a halving
    const uint32_t b = slot - d.eb0;
    if (b * 32 >= n) return;
    const uint32_t e = b * 32 + lane;
    unsigned long long acc[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) acc[k] = 0;
    if (e < n) {
        const uint4 *r = a.trow + (size_t)(d.ep0 + e) * 2;
        const uint4 a0 = r[0], a1 = r[1], b0 = r[2], b1 = r[3];
        const uint32_t D[NK] = {b0.x - a0.x, b0.y - a0.y, b0.z - a0.z, b0.w - a0.w, b1.x - a1.x, b1.y - a1.y};
#pragma unroll
        for (int k = 0; k < NK; k++) acc[k] = (unsigned long long)D[k] << lane;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < NK; k++) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (lane < NK) {
        unsigned long long v = acc[0];
        v = lane == 1 ? acc[1] : v; v = lane == 2 ? acc[2] : v; v = lane == 3 ? acc[3] : v; v = lane == 4 ? acc[4] : v; v = lane == 5 ? acc[5] : v;
        a.ablk[(size_t)slot * 8 + lane] = v;
    }
}

// Counter k at the start of epoch 32b of a chain: S_{b} = (S_{b-1} + A_{b-1}) >> 32 = hi(A_{b-1}) + carry, where the
// carry out of the low word depends on S_{b-1} only when lo(A_{b-1}) + hi(A_{b-2}) is exactly 2^32 - 1; the look-back
// stops as soon as the carry is certain (block 0 starts from zero).  Exact, and parallel over the blocks.
__device__ __forceinline__ uint32_t sp_block_start(const unsigned long long *__restrict__ A /* ablk of the chain, slot k */, uint32_t b) {
    if (b == 0) return 0;
    // find the newest block c <= b - 1 whose incoming state does not matter, then run forward
    uint32_t c = b - 1;
    for (;;) {
        if (c == 0) break;                                   // S_0 = 0 is known
        const unsigned long long lo = A[(size_t)c * 8] & 0xffffffffull, hi_prev = A[(size_t)(c - 1) * 8] >> 32;
        // S_c is hi_prev or hi_prev + 1: the carry of block c is certain unless lo + hi_prev == 2^32 - 1
        if (lo + hi_prev != 0xffffffffull) break;
        c--;
    }
    unsigned long long S = c == 0 ? 0ull : (A[(size_t)(c - 1) * 8] >> 32);   // exact when c == 0, else a value whose carry effect on block c is the same as the true one
    for (uint32_t i = c; i < b; i++) S = (S + A[(size_t)i * 8]) >> 32;
    return (uint32_t)S;
}

// 6. per block of 32 epochs: exact counters at every epoch start.  WRITE = false: verification only (any failing epoch
//    rejects the chain); WRITE = true, after every block has been verified: epoch records and block epochs of the accepted
//    chains.  Rejected chains are never written, so the serial walker may run beside this kernel.
template <bool WRITE>
__global__ void __launch_bounds__(128) k_sp_finish(SpArgs a) {
    const uint32_t slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (slot >= a.counts[3]) return;
    const uint32_t di = a.eb_desc[slot];
    const uint32_t fail = a.chain_fail[di];
    if (WRITE ? fail != SP_OK : (fail != SP_OK && fail != SP_OK - 1)) return;   // rejected
    const SpDesc d = a.desc[di];
    const uint32_t n = a.chain_n[di];
    const uint32_t b = slot - d.eb0;
    if (b * 32 > n) return;
    const uint32_t e = b * 32 + lane;            // my epoch (valid if e <= n)
    const uint32_t kb = d.kb;
    const bool has_epoch = e <= n, has_halving = e < n;
    uint32_t s0 = 0, s1 = 0, T0[NK], D[NK];
    unsigned long long pre[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) { T0[k] = 0; D[k] = 0; pre[k] = 0; }
    if (has_epoch) {
        const uint4 *r = a.trow + (size_t)(d.ep0 + e) * 2;
        const uint4 a0 = r[0], a1 = r[1];
        T0[0] = a0.x; T0[1] = a0.y; T0[2] = a0.z; T0[3] = a0.w; T0[4] = a1.x; T0[5] = a1.y;
        s0 = a1.z;
        if (has_halving) {
            const uint4 b0 = r[2], b1 = r[3];
            D[0] = b0.x - a0.x; D[1] = b0.y - a0.y; D[2] = b0.z - a0.z; D[3] = b0.w - a0.w; D[4] = b1.x - a1.x; D[5] = b1.y - a1.y;
            s1 = b1.z;
        } else {
            s1 = d.gbase + d.count;
            if (s1 > s0) {
#pragma unroll
                for (int k = 0; k < NK; k++) D[k] = sp_prefix_end(a.fine, a.blk_rec4, s1, k) - T0[k];
            }
        }
#pragma unroll
        for (int k = 0; k < NK; k++)
            if (has_halving) pre[k] = (unsigned long long)D[k] << lane;
    }
    // inclusive prefix over the lanes: sum_{i <= lane} D_i << i
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < NK; k++) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, pre[k], o);
            if (lane >= (uint32_t)o) pre[k] += t;
        }
    }
    // counters at the start of the block's first epoch: lane k computes counter k, then broadcast
    uint32_t sb = 0;
    if (lane < NK) sb = sp_block_start(a.ablk + (size_t)d.eb0 * 8 + lane, b);
    bool bad = false;
    uint32_t S[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const unsigned long long sblk = __shfl_sync(0xffffffffu, sb, k);
        const unsigned long long excl = pre[k] - (has_halving ? ((unsigned long long)D[k] << lane) : 0ull);
        S[k] = (uint32_t)((sblk + excl) >> lane);                                                // counters at the start of my epoch
        if (has_halving && !((int32_t)(S[k] + D[k]) > (int32_t)HALVE_AT)) bad = true;           // every counter must be past 1024 after the halving element
    }
    uint32_t vkb = S[0] + D[0];
    vkb = kb == 1 ? S[1] + D[1] : vkb; vkb = kb == 2 ? S[2] + D[2] : vkb; vkb = kb == 3 ? S[3] + D[3] : vkb;
    vkb = kb == 4 ? S[4] + D[4] : vkb; vkb = kb == 5 ? S[5] + D[5] : vkb;
    if (has_halving) {
        // the presumed binding counter must still be <= 1024 just before the halving element
        const uint32_t eh = a.fine[s1 - 1].w;
        const uint32_t dk = (eh >> kb) + 1u + kb;
        if ((int32_t)(vkb - dk) > (int32_t)HALVE_AT) bad = true;
    } else if (has_epoch) {
        // open last epoch: no halving may be left (kb <= 1024 after the last element)
        if ((int32_t)vkb > (int32_t)HALVE_AT) bad = true;
    }
    if (!WRITE) {
        if (bad) atomicMin(&a.chain_fail[di], SP_OK - 1);   // any failure rejects the chain
        return;
    }
    if (has_epoch) {
        uint4 *rec = a.ep_rec + (size_t)(d.ep0 + e) * 2;
        rec[0] = make_uint4(S[0] - T0[0], S[1] - T0[1], S[2] - T0[2], S[3] - T0[3]);
        rec[1] = make_uint4(S[4] - T0[4], S[5] - T0[5], s0, 0u);
        if (e == n) rec[3] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);   // sentinel after the last epoch
        // 32-element blocks whose first element lies in my epoch
        for (uint32_t B = (s0 + 31) >> 5; B < ((s1 + 31) >> 5); B++) a.blk_epoch[B] = d.ep0 + e;
    }
}

// 8. chains whose every epoch verified need no serial walk
__global__ void k_sp_resolve(SpArgs a) {
    const uint32_t di = blockIdx.x * blockDim.x + threadIdx.x;
    if (di >= a.counts[0]) return;
    const bool ok = a.chain_fail[di] == SP_OK;
    a.resolved[a.desc[di].pc] = ok ? 1 : 0;
    atomicAdd(&a.dbg[3], 1u);
    if (ok) atomicAdd(&a.dbg[4], 1u);
}

}  // namespace felics
