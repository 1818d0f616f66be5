// Parallel FELICS encode of 16-bit samples (traits.rs:35-43: K = {0..14}, MAX_CONTEXT = 131070).  Included by encode.cu.
//
// The 131071 x 15 estimator table (parameter_selection.rs:29-33) is too big for the 8-bit pipeline's
// one-chain-per-context layout, and real 16-bit images spread their out-of-range pixels over thousands of contexts.
// Rows of different contexts never interact (parameter_selection.rs:49-64 touches one row), so the plane is cut
// into 512 BUCKETS by the low nine bits of the context: a bucket holds, in raster order, every out-of-range pixel whose
// context is congruent to it, i.e. at most 256 estimator rows (context >> 9), which fit in 16 KB of shared memory.
// One warp walks one bucket exactly as the reference walks the plane (lane k owns counter k; get_k and the halving test
// are two warp reductions), 512 buckets per plane run side by side.  Everything around the walk is data parallel and
// shared with the 8-bit pipeline: histogram -> bases -> stable scatter in front, bit lengths -> offsets -> packing behind.
#pragma once

namespace felics {

constexpr int ROWS16B = 256;              // estimator rows of one bucket: contexts b, b + 512, ... (131071 / 512 rounded up)
constexpr uint32_t E16_BITS = 17;         // residuals of Co/Cg planes reach 131069

// Stable grouping by bucket: k_scatter's scheme (warp w owns 512 consecutive pixels, 32 at a time, ranks by match_any),
// with {residual | row << 17, pixel index} records.
__global__ void __launch_bounds__(TILE_THREADS, 3) k16_scatter(const int32_t *__restrict__ planes, uint32_t w, uint32_t npix, uint32_t tpp,
                                                               uint32_t cap, const uint32_t *__restrict__ tile_base, uint2 *__restrict__ grp) {
    __shared__ uint16_t wcnt[TILE_WARPS][NBIN];
    __shared__ uint32_t wbase[TILE_WARPS][NBIN];
    const uint32_t bid = blockIdx.x;
    const uint32_t p = bid / tpp, t = bid - p * tpp;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < TILE_WARPS * NBIN; c += TILE_THREADS) (&wcnt[0][0])[c] = 0;
    __syncthreads();
    const int32_t *pl = planes + (size_t)p * npix;
    const uint32_t wstart = t * TILE + wid * WARP_PIX;
    uint32_t info[WARP_ITERS];  // rank(13) | bucket(9) << 13 | oor << 22 | row(8) << 23
    uint32_t ev[WARP_ITERS];
    const uint32_t lt = (1u << lane) - 1u;
    RasterCursor<int32_t> cur;
    cur.init(pl, wstart + lane, w);
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        const uint32_t i = cur.i;
        bool oor = false;
        uint32_t delta = 0, val = 0;
        if (i >= 2 && i < npix) {
            const PixelClass pc = cur.classify();
            oor = pc.cls != 0;
            delta = (uint32_t)pc.delta;
            val = (uint32_t)pc.val;
        }
        const uint32_t bucket = delta & (NBIN - 1u);
        const uint32_t act = __ballot_sync(0xffffffffu, oor);
        uint32_t rank = 0, grpmask = 0, prev = 0;
        if (oor) {
            grpmask = __match_any_sync(act, bucket);
            prev = wcnt[wid][bucket];
            rank = prev + __popc(grpmask & lt);
        }
        __syncwarp();
        if (oor && (grpmask & lt) == 0) wcnt[wid][bucket] = (uint16_t)(prev + __popc(grpmask));
        __syncwarp();
        info[it] = rank | (bucket << 13) | (oor ? (1u << 22) : 0u) | ((delta >> 9) << 23);
        ev[it] = val;
        cur.step(32);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NBIN; c += TILE_THREADS) {
        uint32_t run = tile_base[(size_t)bid * NBIN + c];
#pragma unroll
        for (int q = 0; q < TILE_WARPS; q++) {
            wbase[q][c] = run;
            run += wcnt[q][c];
        }
    }
    __syncthreads();
    uint2 *out = grp + (size_t)p * cap;
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        if ((info[it] >> 22) & 1u) {
            const uint32_t bucket = (info[it] >> 13) & 511u;
            const uint32_t g = wbase[wid][bucket] + (info[it] & 8191u);
            out[g] = make_uint2(ev[it] | ((info[it] >> 23) << E16_BITS), wstart + it * 32 + lane);
        }
    }
}

// Bucket walk: the reference's estimator loop over one bucket (parameter_selection.rs:49-85), 32 elements at a time.
// A batch is split by estimator row (context >> 9; rows never interact, so their order inside a batch is free):
//   scan pass (few distinct rows in the batch, at least eight elements per row: small contexts all sit in row 0)
//       lane j holds element j: the 15 code costs (rice_coding.rs:56) are prefix-summed over the lanes of the row, so the
//       counters after element j are V + P(j) with V the (re-based) counters at the batch start.  A halving
//       (parameter_selection.rs:58-63) is the first lane where all 15 have passed 1024: one ballot per halving, V is
//       re-based with that lane's prefix, later lanes pick their k again.  get_k is the `<=` scan of
//       parameter_selection.rs:78-83 on the counters BEFORE the element.
//   serial pass (many distinct rows: wide contexts, e.g. noise)
//       lane k < 15 owns counter k of the row in use; get_k is a warp minimum over (count << 4 | 14 - k), which resolves
//       ties to the largest k; the halving test is a second warp minimum.
// k goes straight to its pixel.
// inclusive scan over the lanes of the first N counters' costs; the others cost 1 + k per element of the row (prefix = a count)
template <int N>
__device__ __forceinline__ void scan_costs16(uint32_t (&P)[NK16], uint32_t lane, uint32_t cnt) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < N; k++) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, P[k], o);
            if (lane >= (uint32_t)o) P[k] += t;
        }
    }
#pragma unroll
    for (int k = N; k < NK16; k++) P[k] = (1u + (uint32_t)k) * cnt;
}

// Wide pass of the bucket walk: 128 consecutive elements of the bucket, four consecutive elements per lane, of which the
// ones flagged `in` belong to the row being walked (the others cost nothing and are never a halving or a get_k).  The costs
// of k < N are prefix-summed inside the lane and then over the lanes (N: every residual of the row is below 2^N, or 15);
// for k >= N the cost is 1 + k per element of the row, so the prefix is a multiple of the running element count.
// All sums travel times 16 with 14 - k in the low bits: a counter x is the key x * 16 + (14 - k), whose minimum over k is
// get_k's answer (smallest count, ties to the largest k: parameter_selection.rs:78-83) and, shifted back, the smallest
// count itself (the halving test of parameter_selection.rs:58).  Within an epoch the counters only grow: the halving
// element is in the first lane whose LAST element has all counters past 1024.
template <int N>
__device__ __forceinline__ void bw16_pass128(const uint32_t (&e)[4], const bool (&in)[4], const uint32_t (&pix)[4], uint32_t lane,
                                             uint32_t *__restrict__ trow /* estimator row, shared */, uint32_t *__restrict__ sx /* 16 words, shared */,
                                             uint8_t *__restrict__ kp) {
    constexpr int NS = N < NK16 ? N + 1 : N;    // scanned quantities: the costs and, when a tail exists, the element count
    uint32_t q[4][NS], o[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t c = k < N ? (((e[i] >> k) + 1u + (uint32_t)k) << 4) : 16u;
            q[i][k] = (i ? q[i ? i - 1 : 0][k] : 0u) + (in[i] ? c : 0u);
        }
        o[k] = q[3][k];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int k = 0; k < NS; k++) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, o[k], d);
            if (lane >= (uint32_t)d) o[k] += t;
        }
    }
#pragma unroll
    for (int k = 0; k < NS; k++) o[k] -= q[3][k];   // exclusive over lanes
    uint32_t V[NK16], Vo[NK16];
    {
        const uint4 *t4 = reinterpret_cast<const uint4 *>(trow);
        const uint4 a = t4[0], b = t4[1], c = t4[2], d = t4[3];
        V[0] = a.x; V[1] = a.y; V[2] = a.z; V[3] = a.w; V[4] = b.x; V[5] = b.y; V[6] = b.z; V[7] = b.w;
        V[8] = c.x; V[9] = c.y; V[10] = c.z; V[11] = c.w; V[12] = d.x; V[13] = d.y; V[14] = d.z;
    }
    // key of counter k after my element i: Vo[k] + (k < N ? q[i][k] : (1 + k) * q[i][N]); before it: the same with i - 1 (nothing for i = 0)
    auto rebase = [&]() {
#pragma unroll
        for (int k = 0; k < NK16; k++) Vo[k] = V[k] * 16u + (14u - (uint32_t)k) + (k < N ? o[k < N ? k : 0] : (1u + (uint32_t)k) * o[NS - 1]);
    };
    auto key_min = [&](int i) {   // smallest key after element i (i = -1: before my first element)
        uint32_t m = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < NK16; k++) m = min(m, Vo[k] + (i < 0 ? 0u : (k < N ? q[i < 0 ? 0 : i][k < N ? k : 0] : (1u + (uint32_t)k) * q[i < 0 ? 0 : i][NS - 1])));
        return m;
    };
    rebase();
    uint32_t myk[4];
#pragma unroll
    for (int i = 0; i < 4; i++) myk[i] = 14u - (key_min(i - 1) & 15u);
    int done = -1;   // my elements up to `done` lie at or before the last halving
    for (;;) {
        const uint32_t hit = __ballot_sync(0xffffffffu, done < 3 && (key_min(3) >> 4) > HALVE_AT);
        if (!hit) break;
        const uint32_t L = (uint32_t)__ffs(hit) - 1u;
        int ih = 3;
#pragma unroll
        for (int i = 3; i >= 0; i--)
            if (i > done && in[i] && (key_min(i) >> 4) > HALVE_AT) ih = i;
        if (lane == L) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i == ih) {
#pragma unroll
                    for (int k = 0; k < NK16; k++)
                        sx[k] = k < N ? (o[k < N ? k : 0] + q[i][k < N ? k : 0]) >> 4 : (1u + (uint32_t)k) * ((o[NS - 1] + q[i][NS - 1]) >> 4);
                }
            sx[15] = (uint32_t)ih;
        }
        __syncwarp();
        {
            const uint4 *x4 = reinterpret_cast<const uint4 *>(sx);
            const uint4 a = x4[0], b = x4[1], c = x4[2], d = x4[3];
            const uint32_t Ph[NK16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z};
#pragma unroll
            for (int k = 0; k < NK16; k++) V[k] = ((V[k] + Ph[k]) >> 1) - Ph[k];   // parameter_selection.rs:58-63, re-based
            ih = (int)d.w;
        }
        __syncwarp();
        rebase();
        done = lane < L ? 3 : (lane == L ? ih : done);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i > done) myk[i] = 14u - (key_min(i - 1) & 15u);
    }
    if (lane == 31) {
        uint32_t A[NK16];
#pragma unroll
        for (int k = 0; k < NK16; k++) A[k] = (Vo[k] + (k < N ? q[3][k < N ? k : 0] : (1u + (uint32_t)k) * q[3][NS - 1])) >> 4;
        uint4 *t4 = reinterpret_cast<uint4 *>(trow);
        t4[0] = make_uint4(A[0], A[1], A[2], A[3]); t4[1] = make_uint4(A[4], A[5], A[6], A[7]);
        t4[2] = make_uint4(A[8], A[9], A[10], A[11]); t4[3] = make_uint4(A[12], A[13], A[14], 0u);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (in[i]) kp[pix[i]] = (uint8_t)myk[i];
}

// Rows are independent, so two warps share a bucket: warp `way` takes the rows with (row & 1) == way.  (Four warps per
// bucket halve the resident blocks at the wide pass's 235 registers: 1.36 vs 1.14 ms on the gray16 bench image.)
constexpr uint32_t BW_SCAN_MIN_PER_ROW = 8;
constexpr int BW_WAYS = 2;
__global__ void __launch_bounds__(32 * BW_WAYS) k16_bwalk(const uint2 *__restrict__ grp, const uint32_t *__restrict__ chain_count,
                                                          const uint32_t *__restrict__ chain_base, const uint32_t *__restrict__ live,
                                                          uint32_t *__restrict__ counters, uint32_t cap, uint32_t npix,
                                                          uint8_t *__restrict__ kpix, uint32_t opts) {
    __shared__ __align__(16) uint32_t tab[ROWS16B][16];
    __shared__ __align__(16) uint32_t pref[BW_WAYS][32][16];
    __shared__ uint32_t s_qi;
    const uint32_t lane = threadIdx.x & 31, way = threadIdx.x >> 5;
    uint32_t(*sP)[16] = pref[way];
    const bool mine = lane < (uint32_t)NK16;
    for (;;) {
        if (threadIdx.x == 0) s_qi = atomicAdd(&counters[1], 1u);
        __syncthreads();
        const uint32_t qi = s_qi;
        if (qi >= counters[0]) break;
        const uint32_t pc = live[qi];
        const uint32_t p = pc / NBIN;
        const uint32_t count = chain_count[pc];
        const uint2 *src = grp + (size_t)p * cap + chain_base[pc];
        uint8_t *kp = kpix + (size_t)p * npix;
        for (uint32_t j = threadIdx.x; j < ROWS16B * 16; j += 32 * BW_WAYS) (&tab[0][0])[j] = 0u;   // KEstimator::new: all counts zero
        __syncthreads();
        const uint32_t waymask = (opts & 3u) == 0 ? (uint32_t)BW_WAYS - 1u : 0u;   // experiment switch: one warp takes every row
        if (way > waymask) { __syncthreads(); continue; }
        uint4 pa = make_uint4(0u, 0u, 0u, 0u), pb = pa;   // the next 128 elements, in flight while these are walked
        uint32_t pf = 0xffffffffu;
        for (uint32_t first = 0; first < count; first += 32) {
            if ((first & 127u) == 0 && first + 128 <= count && (opts & 8u) == 0) {
                // 128 elements whose rows are all below 8 (contexts below 4096): at most four rows are mine, one wide pass for each that occurs
                const uint4 *s4 = reinterpret_cast<const uint4 *>(src + first) + 2 * lane;
                uint4 ra = pa, rb = pb;
                if (pf != first) { ra = s4[0]; rb = s4[1]; }
                if (first + 256 <= count) { pa = s4[64]; pb = s4[65]; pf = first + 128; }
                const uint32_t emask = (1u << E16_BITS) - 1u;
                const uint32_t rows[4] = {ra.x >> E16_BITS, ra.z >> E16_BITS, rb.x >> E16_BITS, rb.z >> E16_BITS};
                if (__reduce_or_sync(0xffffffffu, rows[0] | rows[1] | rows[2] | rows[3]) < 8u && waymask == (uint32_t)BW_WAYS - 1u) {
                    const uint32_t e[4] = {ra.x & emask, ra.z & emask, rb.x & emask, rb.z & emask};
                    const uint32_t pix[4] = {ra.y, ra.w, rb.y, rb.w};
#pragma unroll 1
                    for (uint32_t prow = way; prow < 8u; prow += (uint32_t)BW_WAYS) {
                        const bool in[4] = {rows[0] == prow, rows[1] == prow, rows[2] == prow, rows[3] == prow};
                        const uint32_t mybits = (in[0] ? e[0] : 0u) | (in[1] ? e[1] : 0u) | (in[2] ? e[2] : 0u) | (in[3] ? e[3] : 0u);
                        if (!__any_sync(0xffffffffu, in[0] || in[1] || in[2] || in[3])) continue;
                        const uint32_t bits = __reduce_or_sync(0xffffffffu, mybits);
                        if (bits < 16u) bw16_pass128<4>(e, in, pix, lane, tab[prow], sP[0], kp);
                        else if (bits < 256u) bw16_pass128<8>(e, in, pix, lane, tab[prow], sP[0], kp);
                        else bw16_pass128<NK16>(e, in, pix, lane, tab[prow], sP[0], kp);
                    }
                    first += 96;
                    continue;
                }
            }
            const uint32_t nv = min(32u, count - first);
            uint2 r = make_uint2(0u, 0u);
            if (lane < nv) r = src[first + lane];
            const uint32_t e = r.x & ((1u << E16_BITS) - 1u), row = r.x >> E16_BITS;
            const bool valid = lane < nv && (row & waymask) == way;
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            if (!vmask) continue;
            const uint32_t same = __match_any_sync(0xffffffffu, valid ? row : 0xffffffffu);
            uint32_t todo = __ballot_sync(0xffffffffu, valid && (uint32_t)(__ffs(same) - 1) == lane);   // one leader per row
            uint32_t myk = 0;
            const uint32_t rule = (opts >> 4) & 15u ? (opts >> 4) & 15u : BW_SCAN_MIN_PER_ROW;
            if ((uint32_t)__popc(todo) * rule <= (uint32_t)__popc(vmask)) {   // a scan pass per row pays off from a few elements per row on
                while (todo) {
                    const uint32_t leader = (uint32_t)__ffs(todo) - 1u;
                    todo &= todo - 1u;
                    const uint32_t M = __shfl_sync(0xffffffffu, same, leader);       // lanes of this row
                    const uint32_t prow = __shfl_sync(0xffffffffu, row, leader);
                    const bool in = (M >> lane) & 1u;
                    // for k >= K0 every residual of the row has e >> k == 0: the cost is 1 + k and the prefix a count
                    const int K0 = 32 - __clz(__reduce_max_sync(0xffffffffu, in ? e : 0u));
                    const uint32_t cnt = (uint32_t)__popc(M & (0xffffffffu >> (31u - lane)));
                    uint32_t P[NK16], V[NK16];
#pragma unroll
                    for (int k = 0; k < NK16; k++) P[k] = in ? (e >> k) + 1u + (uint32_t)k : 0u;
                    if (K0 <= 4 && (opts & 4u) == 0) scan_costs16<4>(P, lane, cnt);
                    else if (K0 <= 8 && (opts & 4u) == 0) scan_costs16<8>(P, lane, cnt);
                    else scan_costs16<NK16>(P, lane, cnt);
                    {
                        uint4 *d = reinterpret_cast<uint4 *>(sP[lane]);
                        d[0] = make_uint4(P[0], P[1], P[2], P[3]); d[1] = make_uint4(P[4], P[5], P[6], P[7]);
                        d[2] = make_uint4(P[8], P[9], P[10], P[11]); d[3] = make_uint4(P[12], P[13], P[14], 0u);
                        const uint4 *t4 = reinterpret_cast<const uint4 *>(tab[prow]);
                        const uint4 a = t4[0], b = t4[1], c = t4[2], dd = t4[3];
                        V[0] = a.x; V[1] = a.y; V[2] = a.z; V[3] = a.w; V[4] = b.x; V[5] = b.y; V[6] = b.z; V[7] = b.w;
                        V[8] = c.x; V[9] = c.y; V[10] = c.z; V[11] = c.w; V[12] = dd.x; V[13] = dd.y; V[14] = dd.z;
                    }
                    __syncwarp();
                    auto pick = [&]() {
                        uint32_t best = V[0] + P[0] - (e + 1u), bi = 0;   // counters before my element
#pragma unroll
                        for (int k = 1; k < NK16; k++) {
                            const uint32_t x = V[k] + P[k] - ((e >> k) + 1u + (uint32_t)k);
                            if (x <= best) { best = x; bi = (uint32_t)k; }
                        }
                        return bi;
                    };
                    if (in) myk = pick();
                    uint32_t after = 0xffffffffu;   // lanes behind the last halving
                    for (;;) {
                        uint32_t m = V[0] + P[0];
#pragma unroll
                        for (int k = 1; k < NK16; k++) m = min(m, V[k] + P[k]);
                        const uint32_t hit = __ballot_sync(0xffffffffu, in && m > HALVE_AT) & after;
                        if (!hit) break;
                        const uint32_t h = (uint32_t)__ffs(hit) - 1u;
                        const uint4 *q4 = reinterpret_cast<const uint4 *>(sP[h]);
                        const uint4 a = q4[0], b = q4[1], c = q4[2], dd = q4[3];
                        const uint32_t Ph[NK16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, dd.x, dd.y, dd.z};
#pragma unroll
                        for (int k = 0; k < NK16; k++) V[k] = ((V[k] + Ph[k]) >> 1) - Ph[k];
                        after = h == 31u ? 0u : (0xffffffffu << (h + 1u));
                        if (in && lane > h) myk = pick();
                    }
                    if (lane == 0) {
                        const uint4 *q4 = reinterpret_cast<const uint4 *>(sP[31]);
                        const uint4 a = q4[0], b = q4[1], c = q4[2], dd = q4[3];
                        uint4 *t4 = reinterpret_cast<uint4 *>(tab[prow]);
                        t4[0] = make_uint4(V[0] + a.x, V[1] + a.y, V[2] + a.z, V[3] + a.w);
                        t4[1] = make_uint4(V[4] + b.x, V[5] + b.y, V[6] + b.z, V[7] + b.w);
                        t4[2] = make_uint4(V[8] + c.x, V[9] + c.y, V[10] + c.z, V[11] + c.w);
                        t4[3] = make_uint4(V[12] + dd.x, V[13] + dd.y, V[14] + dd.z, 0u);
                    }
                    __syncwarp();
                }
            } else {
                uint32_t prow = way, v = mine ? tab[way][lane] : 0u;   // v: my counter of row prow
                for (uint32_t left = vmask; left; left &= left - 1u) {
                    const uint32_t j = (uint32_t)__ffs(left) - 1u;
                    const uint32_t ej = __shfl_sync(0xffffffffu, e, j), rj = __shfl_sync(0xffffffffu, row, j);
                    if (rj != prow) {              // warp uniform
                        if (mine) { tab[prow][lane] = v; v = tab[rj][lane]; }
                        prow = rj;
                    }
                    const uint32_t best = __reduce_min_sync(0xffffffffu, mine ? ((v << 4) | (14u - lane)) : 0xffffffffu);
                    if (lane == j) myk = 14u - (best & 15u);
                    v += (ej >> lane) + 1u + lane;   // rice_coding.rs:56 code_length with k = lane
                    const uint32_t mn = __reduce_min_sync(0xffffffffu, mine ? v : 0xffffffffu);
                    if (mn > HALVE_AT) v >>= 1;
                }
                if (mine) tab[prow][lane] = v;
                __syncwarp();
            }
            if (valid) kp[r.y] = (uint8_t)myk;
        }
        __syncthreads();   // the table and s_qi are reused
    }
}

// Code of one pixel as {length, payload}.  In range: payload = '1' marker and the phased-in code, right aligned
// (compression.rs:130-134).  Out of range: payload = e | k << 17 | above << 21 | 1 << 31, length = 2 + q + 1 + k
// (compression.rs:136-145, rice_coding.rs:26-39).
struct Code16 {
    uint32_t len, pay;
};
__device__ __forceinline__ Code16 code16_of(const PixelClass &pc, const uint8_t *__restrict__ kp, uint32_t i) {
    Code16 c;
    if (pc.cls == 0) {
        int len;
        const uint32_t code = phase_in_code((uint32_t)pc.delta + 1u, (uint32_t)pc.val, len);
        c.len = (uint32_t)len + 1u;
        c.pay = (1u << len) | code;
    } else {
        const uint32_t k = kp[i], e = (uint32_t)pc.val;
        c.len = 2u + (e >> k) + 1u + k;
        c.pay = e | (k << 17) | ((pc.cls == 1 ? 1u : 0u) << 21) | 0x80000000u;
    }
    return c;
}

// bits per tile; QUADS (width a multiple of four): four consecutive samples per thread (classify4)
template <bool QUADS>
__global__ void __launch_bounds__(TILE_THREADS) k16_code(const int32_t *__restrict__ planes, uint32_t w, uint32_t npix, uint32_t tpp,
                                                         const uint8_t *__restrict__ kpix, uint32_t *__restrict__ tile_bits) {
    __shared__ uint32_t wsum[TILE_WARPS];
    const uint32_t bid = blockIdx.x;
    const uint32_t p = bid / tpp, t = bid - p * tpp;
    const uint8_t *kp = kpix + (size_t)p * npix;
    const int32_t *pl = planes + (size_t)p * npix;
    uint32_t bits = 0;
    RasterCursor<int32_t> cur;
    if (QUADS) {
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
                    if (valid[q]) bits += code16_of(pc[q], kp, cur.i + q).len;
            }
            cur.step_qr(4 * TILE_THREADS, step_q, step_r);
        }
    } else {
        cur.init(pl, t * TILE + threadIdx.x, w);
#pragma unroll 4
        for (int j = 0; j < TILE / TILE_THREADS; j++) {
            const uint32_t i = cur.i;
            if (i >= npix) break;
            if (i >= 2) bits += code16_of(cur.classify(), kp, i).len;
            cur.step(TILE_THREADS);
        }
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

// Packing: a thread owns 16 consecutive pixels; lengths are scanned over the tile, codes are OR-ed into a shared-memory
// image of the tile's bits (or straight into the arena when a tile is longer than the buffer: unary runs of 16-bit
// residuals reach 131069 bits).
template <bool QUADS>
__global__ void __launch_bounds__(TILE_THREADS) k16_pack(PackArgs a, const int32_t *__restrict__ planes, uint32_t w,
                                                         const uint8_t *__restrict__ kpix) {
    __shared__ uint32_t buf[PACK_WORDS];
    __shared__ uint32_t wtot[TILE_WARPS];
    const uint32_t bid = blockIdx.x;
    const uint32_t p = bid / a.tpp, t = bid - p * a.tpp;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t tbits = a.tile_bits[bid];
    const uint64_t bit0 = plane_bit_start(a, p) + 64 + a.tile_off[bid];
    const uint64_t word0 = bit0 >> 5;
    const uint32_t sh0 = (uint32_t)(bit0 & 31);
    const uint64_t nwords64 = ((uint64_t)sh0 + tbits + 31u) >> 5;
    const bool in_smem = nwords64 <= (uint64_t)PACK_WORDS;
    const uint32_t nwords = (uint32_t)nwords64;
    if (in_smem)
        for (uint32_t j = threadIdx.x; j < nwords; j += TILE_THREADS) buf[j] = 0;

    const uint8_t *kp = kpix + (size_t)p * a.npix;
    const uint32_t lstart = t * TILE + wid * WARP_PIX + lane * WARP_ITERS;
    Code16 c[WARP_ITERS];
    uint32_t mylen = 0;
    {
        const int32_t *pl = planes + (size_t)p * a.npix;
        RasterCursor<int32_t> cur;
        cur.init(pl, lstart, w);
        if (QUADS) {
#pragma unroll
            for (int qd = 0; qd < WARP_ITERS / 4; qd++) {
                PixelClass pc[4];
                bool valid[4] = {false, false, false, false};
                if (cur.i < a.npix) classify4(pl, cur.i, cur.x, cur.y, w, pc, valid);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int it = qd * 4 + q;
                    c[it].len = 0; c[it].pay = 0;
                    if (valid[q]) c[it] = code16_of(pc[q], kp, cur.i + q);
                    mylen += c[it].len;
                }
                cur.step(4);
            }
        } else {
#pragma unroll
            for (int it = 0; it < WARP_ITERS; it++) {
                c[it].len = 0; c[it].pay = 0;
                if (cur.i >= 2 && cur.i < a.npix) c[it] = code16_of(cur.classify(), kp, cur.i);
                mylen += c[it].len;
                cur.step(1);
            }
        }
    }
    uint32_t inc = mylen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += n;
    }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    uint64_t pos = (uint64_t)sh0 + inc - mylen;   // bit offset of my first code relative to word0
    for (uint32_t q = 0; q < wid; q++) pos += wtot[q];

    auto put = [&](uint64_t off, uint32_t val, int n) {
        if (in_smem) put_bits_smem(buf, (uint32_t)off, val, n);
        else put_bits_global(a.arena, (word0 << 5) + off, val, n);
    };
#pragma unroll
    for (int it = 0; it < WARP_ITERS; it++) {
        const uint32_t len = c[it].len;
        if (len == 0) continue;
        if (!(c[it].pay >> 31)) {
            put(pos, c[it].pay, (int)len);
        } else {
            const uint32_t e = c[it].pay & 0x1ffffu, k = (c[it].pay >> 17) & 15u, above = (c[it].pay >> 21) & 1u;
            uint32_t q = e >> k;
            const uint32_t rem = e & ((1u << k) - 1u);
            if (len <= 32u) {
                // '0', above, q ones, '0', k remainder bits
                put(pos, (above << (q + 1u + k)) | (((1u << q) - 1u) << (k + 1u)) | rem, (int)len);
            } else {
                uint64_t at = pos;
                put(at, above, 2);
                at += 2;
                while (q >= 32u) { put(at, 0xffffffffu, 32); at += 32; q -= 32u; }
                if (q) { put(at, (1u << q) - 1u, (int)q); at += q; }
                put(at, rem, (int)k + 1);
            }
        }
        pos += len;
    }
    if (!in_smem) return;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < nwords; j += TILE_THREADS) {
        const uint32_t v = buf[j];
        if (j == 0 || j == nwords - 1) {
            if (v) atomicOr(&a.arena[word0 + j], bswap32(v));
        } else {
            a.arena[word0 + j] = bswap32(v);
        }
    }
}

__global__ void k16_planes_gray(const uint16_t *__restrict__ px, int32_t *__restrict__ planes, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) planes[idx] = (int32_t)px[idx];
}
// color_transform.rs:11-17; `/` truncates toward zero in C++ as in Rust
__global__ void k16_planes_rgb(const uint16_t *__restrict__ px, int32_t *__restrict__ planes, uint32_t npix, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        const size_t img = idx / npix;
        const uint32_t i = (uint32_t)(idx - img * npix);
        const int r = px[3 * idx], g = px[3 * idx + 1], b = px[3 * idx + 2];
        const int co = r - b;
        const int t = b + co / 2;
        const int cg = g - t;
        const int y = t + cg / 2;
        int32_t *base = planes + img * 3 * (size_t)npix;
        base[i] = y;
        base[(size_t)npix + i] = co;
        base[2 * (size_t)npix + i] = cg;
    }
}

namespace {
struct Layout16 {
    int32_t *planes;
    uint32_t *tile_hist, *chunk_tot, *chain_count, *chain_base, *plane_used, *live, *counters;
    uint2 *grp;
    uint8_t *kpix;
    uint32_t *tile_bits;
    uint64_t *tile_off, *plane_bits, *img_off;
    size_t bytes;
};
Layout16 carve16(uint8_t *base, const Geom &g, size_t ni) {
    Layout16 L;
    Carver c{base};
    const size_t np = ni * g.nch;
    L.planes = c.take<int32_t>(np * g.npix + 8);
    L.tile_hist = c.take<uint32_t>(np * g.tpp * NBIN);
    L.chunk_tot = c.take<uint32_t>(np * g.nchunks * NBIN);
    L.chain_count = c.take<uint32_t>(np * NBIN);
    L.chain_base = c.take<uint32_t>(np * NBIN);
    L.plane_used = c.take<uint32_t>(np);
    L.live = c.take<uint32_t>(np * NBIN);
    L.counters = c.take<uint32_t>(8);
    L.grp = c.take<uint2>(np * g.cap);
    L.kpix = c.take<uint8_t>(np * g.npix + 8);
    L.tile_bits = c.take<uint32_t>(np * g.tpp + 8);
    L.tile_off = c.take<uint64_t>(np * g.tpp + 8);
    L.plane_bits = c.take<uint64_t>(np + 8);
    L.img_off = c.take<uint64_t>(ni + 8);
    L.bytes = align_up(c.off, 256);
    return L;
}
}  // namespace

// Same contract as the 8-bit encode_batch_device (exactly one of d_arena / h_arena).
int encode16_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr, uint8_t *d_arena, uint8_t *h_arena,
                          size_t arena_cap, uint64_t *offsets_host) {
    if (ctx->serial16) return encode16_serial_batch_device(ctx, n, d_pixels, hdr, d_arena, h_arena, arena_cap, offsets_host);
    const uint64_t npix64 = (uint64_t)hdr.width * hdr.height;
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
    g.epcap = 0;
    if ((uint64_t)g.nch * g.cap >= 0xffff0000ull) {
        set_error("image too large for one call");
        return FELICS_ERR_INVALID_DIMENSIONS;
    }
    const size_t usable_cap = d_arena ? (arena_cap & ~(size_t)3) : arena_cap;
    const size_t img_bytes = (size_t)g.npix * g.nch * 2;
    const size_t per_image = carve16(nullptr, g, 1).bytes;
    size_t sub = std::max<size_t>(1, std::min<size_t>(n, ((size_t)12 << 30) / std::max<size_t>(per_image, 1)));
    while (sub > 1 && (uint64_t)sub * g.nch * g.cap >= 0xffff0000ull) sub--;

    uint64_t arena_off = 0;
    offsets_host[0] = 0;
    for (size_t first = 0; first < n; first += sub) {
        const size_t ni = std::min(sub, n - first);
        const size_t np = ni * g.nch;
        Layout16 L = carve16(nullptr, g, ni);
        int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, L.bytes);
        if (rc) return rc;
        L = carve16((uint8_t *)ctx->scratch, g, ni);
        const uint16_t *px = (const uint16_t *)((const uint8_t *)d_pixels + first * img_bytes);
        const unsigned ntiles = (unsigned)(np * g.tpp);
        const bool quads = g.w % 4 == 0 && g.w >= 8 && !ctx->no_quads;   // the i32 planes are ours: always aligned
        FELICS_CUDA_TRY(cudaMemsetAsync(L.counters, 0, 8 * sizeof(uint32_t), st));
        if (g.npix > 0) {
            StageScope s(ctx, ST_PLANES);
            const size_t total = ni * (size_t)g.npix;
            const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
            if (g.nch == 1) k16_planes_gray<<<blocks, 256, 0, st>>>(px, L.planes, total);
            else k16_planes_rgb<<<blocks, 256, 0, st>>>(px, L.planes, g.npix, total);
            s.launched();
        }
        if (g.npix > 2) {
            {
                StageScope s(ctx, ST_HIST);
                FELICS_CUDA_TRY(cudaMemsetAsync(L.chunk_tot, 0, np * g.nchunks * NBIN * sizeof(uint32_t), st));
                if (quads) k_hist4<int32_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.nchunks, L.tile_hist, L.chunk_tot);
                else k_hist<int32_t><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.nchunks, L.tile_hist, L.chunk_tot);
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
                k16_scatter<<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, g.cap, L.tile_hist, L.grp);
                s.launched();
            }
            {
                StageScope s(ctx, ST_WALK);
                const unsigned blocks = (unsigned)std::min<size_t>(np * NBIN, 148 * 8);
                k16_bwalk<<<blocks, 32 * BW_WAYS, 0, st>>>(L.grp, L.chain_count, L.chain_base, L.live, L.counters, g.cap, g.npix, L.kpix, ctx->bw16_opts);
                s.launched();
            }
            {
                StageScope s(ctx, ST_CODE);
                if (quads) k16_code<true><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, L.kpix, L.tile_bits);
                else k16_code<false><<<ntiles, TILE_THREADS, 0, st>>>(L.planes, g.w, g.npix, g.tpp, L.kpix, L.tile_bits);
                s.launched();
            }
        } else if (ntiles) {
            FELICS_CUDA_TRY(cudaMemsetAsync(L.tile_bits, 0, ntiles * sizeof(uint32_t), st));
        }
        {
            StageScope s(ctx, ST_BITSCAN);
            k_planebits<<<(unsigned)np, 1024, 0, st>>>(L.tile_bits, g.tpp, L.tile_off, L.plane_bits);
            k_imgscan<<<1, 1024, 0, st>>>(L.plane_bits, (uint32_t)ni, g.nch, L.img_off);
            s.launched(2);
        }
        rc = ensure_buffer(ctx, &ctx->pinned, &ctx->pinned_cap, (ni + 1 + 8) * sizeof(uint64_t), true);
        if (rc) return rc;
        uint64_t *h_off = (uint64_t *)ctx->pinned;
        FELICS_CUDA_TRY(cudaMemcpyAsync(h_off, L.img_off, (ni + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        FELICS_CUDA_TRY(cudaStreamSynchronize(st));
        const uint64_t sub_total = h_off[ni];
        for (size_t i = 0; i < ni; i++) offsets_host[first + i + 1] = arena_off + h_off[i + 1];
        if (arena_off + sub_total > usable_cap) {
            // keep sizing: the caller learns the total it needs (lower bound when more sub-batches follow)
            arena_off += sub_total;
            for (size_t f2 = first + ni; f2 < n; f2++) offsets_host[f2 + 1] = arena_off;
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
            pa.rec = nullptr; pa.tile_bits = L.tile_bits; pa.tile_off = L.tile_off; pa.plane_bits = L.plane_bits; pa.img_off = L.img_off;
            pa.arena = (uint32_t *)target; pa.arena_byte0 = target_off; pa.npix = g.npix; pa.tpp = g.tpp; pa.nch = g.nch;
            if (ntiles && g.npix > 2) {
                if (quads) k16_pack<true><<<ntiles, TILE_THREADS, 0, st>>>(pa, L.planes, g.w, L.kpix);
                else k16_pack<false><<<ntiles, TILE_THREADS, 0, st>>>(pa, L.planes, g.w, L.kpix);
                s.launched();
            }
            k_heads<int32_t><<<(unsigned)((np + 127) / 128), 128, 0, st>>>(pa, L.planes, (uint32_t)np, g.w, g.h, hdr.color_type, hdr.pixel_depth);
            s.launched();
        }
        if (!d_arena) FELICS_CUDA_TRY(cudaMemcpyAsync(h_arena + arena_off, target, sub_total, cudaMemcpyDeviceToHost, st));
        FELICS_CUDA_TRY(cudaStreamSynchronize(st));   // pinned offsets / staging are reused by the next sub-batch
        arena_off += sub_total;
    }
    FELICS_CUDA_TRY(cudaGetLastError());
    return profile_collect(ctx);
}

}  // namespace felics
