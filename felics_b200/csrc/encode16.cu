// Serial FELICS encode of 16-bit samples (see serial16.cuh): size pass, then a write pass at exact offsets.
// Kept behind FELICS_B200_SERIAL16=1 as a cross-check of the parallel path (enc16_par.cuh); also home of the decoder's tables.
#include "ctx.h"
#include "device_common.cuh"
#include "serial16.cuh"

#include <algorithm>

namespace felics {

__global__ void k16_to_planes_gray(const uint16_t *__restrict__ px, int32_t *__restrict__ planes, uint32_t npix, size_t pstride, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        const size_t img = idx / npix;
        planes[img * pstride + (idx - img * npix)] = (int32_t)px[idx];
    }
}
// color_transform.rs:11-17; `/` truncates toward zero in C++ as in Rust
__global__ void k16_to_planes_rgb(const uint16_t *__restrict__ px, int32_t *__restrict__ planes, uint32_t npix, size_t pstride, size_t total) {
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
        int32_t *base = planes + img * 3 * pstride;
        base[i] = y;
        base[pstride + i] = co;
        base[2 * pstride + i] = cg;
    }
}

// MSB-first byte sink: bitstream-io's BigEndian BitWriter semantics (write, write_bit, write_unary0, byte_align)
template <bool WRITE>
struct Sink16 {
    uint8_t *out;
    uint64_t pos, cap;   // bytes produced so far (also counted when nothing is stored)
    uint64_t acc;
    int nacc;
    __device__ __forceinline__ void put(int n, uint32_t v) {   // n in 0..32
        if (n == 0) return;
        acc = (acc << n) | (uint64_t)v;
        nacc += n;
        while (nacc >= 8) {
            nacc -= 8;
            if (WRITE && pos < cap) out[pos] = (uint8_t)(acc >> nacc);
            pos++;
        }
        acc &= (1ull << nacc) - 1ull;
    }
    __device__ __forceinline__ void unary0(uint32_t q) {
        while (q >= 32) { put(32, 0xffffffffu); q -= 32; }
        put((int)q + 1, ((1u << q) - 1u) << 1);   // q ones, then the zero
    }
    __device__ __forceinline__ void align() { if (nacc) put(8 - nacc, 0u); }
};

struct Enc16Args {
    const int32_t *planes;      // [n * nch][pstride]: planes padded so that images do not sit a power of two apart
    size_t pstride;
    uint32_t *tables;           // [n][TABLE16_WORDS], tagged rows (serial16.cuh)
    uint32_t tag0;              // first tag of this pass (one per channel)
    uint8_t *arena;             // write pass: image i goes to arena + off[i]
    const uint64_t *off;        // write pass: n + 1 byte offsets (device)
    uint64_t *sizes;            // size pass: bytes of every image
    uint64_t arena_cap;
    uint32_t w, h, npix, nch, color;
};

template <bool WRITE>
__global__ void __launch_bounds__(32) k16_encode(Enc16Args a, uint32_t n) {
    extern __shared__ __align__(16) unsigned char enc16_smem[];
    const uint32_t img = blockIdx.x, lane = threadIdx.x;
    if (img >= n) return;
    Est16 est;
    est.sm = reinterpret_cast<uint32_t *>(enc16_smem);
    est.gl = a.tables + (size_t)img * TABLE16_WORDS;
    est.tag = 0;
    Sink16<WRITE> s;
    s.out = WRITE ? a.arena + a.off[img] : nullptr;
    s.cap = WRITE ? a.off[img + 1] - a.off[img] : 0;
    s.pos = 0; s.acc = 0; s.nacc = 0;
    if (lane == 0) {
        // format.rs:51-61
        s.put(32, 0x464C4353u);
        s.put(8, a.color);
        s.put(8, 1u);
        s.put(32, a.w);
        s.put(32, a.h);
    }
    const uint32_t w = a.w;
    for (uint32_t ch = 0; ch < a.nch; ch++) {
        est.reset(lane, a.tag0 + ch);                              // fresh estimator per channel (compression.rs:110-114)
        if (lane != 0) continue;
        const int32_t *pl = a.planes + ((size_t)img * a.nch + ch) * a.pstride;
        // compression.rs:93-108: two raw samples, 32 bits each
        s.put(32, a.npix >= 1 ? (uint32_t)pl[0] : 0u);
        s.put(32, a.npix >= 2 ? (uint32_t)pl[1] : 0u);
        uint32_t x = 0, y = 0;
        if (a.npix >= 3) { y = 2 / w; x = 2 - y * w; }
        for (uint32_t i = 2; i < a.npix; i++) {
            uint32_t ia, ib;
            neighbours16(i, x, y, w, ia, ib);
            const int v1 = pl[ia], v2 = pl[ib], p = pl[i];
            const int hi = max(v1, v2), lo = min(v1, v2);
            const uint32_t ctx = (uint32_t)(hi - lo);
            if (p >= lo && p <= hi) {                              // InRange: '1' + phased-in code (compression.rs:130-134)
                int len;
                const uint32_t code = phase_in_code(ctx + 1u, (uint32_t)(p - lo), len);
                s.put(len + 1, (1u << len) | code);
            } else {
                const uint32_t above = p > hi ? 1u : 0u;
                const uint32_t e = above ? (uint32_t)(p - hi - 1) : (uint32_t)(lo - p - 1);
                uint32_t cnt[NK16];
                est.load(ctx, cnt);
                const int k = get_k16(cnt);
                s.put(2, above);                                   // '01' above, '00' below (compression.rs:34-41)
                s.unary0(e >> k);                                  // rice_coding.rs:26-39
                s.put(k, e & ((1u << k) - 1u));
                update16(cnt, e);
                est.store(ctx, cnt);
            }
            if (++x == w) { x = 0; y++; }
        }
    }
    if (lane == 0) {
        s.align();                                                 // compression.rs:279 / :368
        if (!WRITE) a.sizes[img] = s.pos;
    }
}

// The per-image estimator tables of the 16-bit path: their own allocation, zeroed once, rows validated by tags afterwards.
int tables16(felics_ctx *ctx, size_t images, uint32_t **out) {
    const size_t need = images * TABLE16_WORDS * sizeof(uint32_t);
    if (need > ctx->tables16_cap) {
        if (ctx->tables16) cudaFree(ctx->tables16);
        ctx->tables16 = nullptr; ctx->tables16_cap = 0;
        FELICS_CUDA_TRY(cudaMalloc(&ctx->tables16, need));
        FELICS_CUDA_TRY(cudaMemsetAsync(ctx->tables16, 0, need, ctx->stream));
        ctx->tables16_cap = need;
    }
    *out = (uint32_t *)ctx->tables16;
    return FELICS_OK;
}
// four fresh tags (one per channel of a pass); tag 0 is the "never written" value of a zeroed table
uint32_t next_tags16(felics_ctx *ctx) {
    if (ctx->tag16 > 0xfffffff0u) {   // wrapped: start over from a clean table
        cudaMemsetAsync(ctx->tables16, 0, ctx->tables16_cap, ctx->stream);
        ctx->tag16 = 1;
    }
    const uint32_t t = ctx->tag16;
    ctx->tag16 += 4;
    return t;
}

int encode16_serial_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr, uint8_t *d_arena, uint8_t *h_arena,
                          size_t arena_cap, uint64_t *offsets_host) {
    const uint64_t npix64 = (uint64_t)hdr.width * hdr.height;
    if (npix64 > 0x7fff0000ull) {
        set_error("image too large for one call: %llu pixels", (unsigned long long)npix64);
        return FELICS_ERR_INVALID_DIMENSIONS;
    }
    cudaStream_t st = ctx->stream;
    const uint32_t npix = (uint32_t)npix64, nch = hdr.color_type ? 3 : 1;
    const size_t img_bytes = (size_t)npix * nch * 2;
    const size_t sub = std::max<size_t>(1, std::min<size_t>(n, 64));   // 8.4 MB of table per image in flight
    uint64_t arena_off = 0;
    offsets_host[0] = 0;
    for (size_t first = 0; first < n; first += sub) {
        const size_t ni = std::min(sub, n - first);
        size_t o_planes = 0;
        // one warp per image walks its planes in lockstep with the others: a power-of-two distance between planes
        // would put all of them on the same memory channel
        const size_t pstride = plane_stride16(npix);
        size_t o_sizes = align_up(o_planes + (ni * nch * pstride + 8) * sizeof(int32_t), 256);
        size_t o_off = align_up(o_sizes + ni * sizeof(uint64_t), 256);
        size_t total = align_up(o_off + (ni + 1) * sizeof(uint64_t), 256);
        int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, total);
        if (rc) return rc;
        uint8_t *sb = (uint8_t *)ctx->scratch;
        Enc16Args a;
        a.planes = (const int32_t *)(sb + o_planes);
        rc = tables16(ctx, ni, &a.tables);
        if (rc) return rc;
        a.sizes = (uint64_t *)(sb + o_sizes);
        a.off = (const uint64_t *)(sb + o_off);
        a.arena = nullptr; a.arena_cap = 0;
        a.w = hdr.width; a.h = hdr.height; a.npix = npix; a.nch = nch; a.color = hdr.color_type; a.pstride = pstride;
        const uint16_t *px = (const uint16_t *)((const uint8_t *)d_pixels + first * img_bytes);
        if (npix) {
            StageScope s(ctx, ST_PLANES);
            const size_t tot = ni * (size_t)npix;
            const unsigned blocks = (unsigned)std::min<size_t>((tot * (nch == 1 ? 1 : 1) + 255) / 256, 148 * 32);
            if (nch == 1) k16_to_planes_gray<<<blocks, 256, 0, st>>>(px, (int32_t *)(sb + o_planes), npix, pstride, tot);
            else k16_to_planes_rgb<<<blocks, 256, 0, st>>>(px, (int32_t *)(sb + o_planes), npix, pstride, tot);
            s.launched();
        }
        if (!ctx->enc16_attr_done) {
            FELICS_CUDA_TRY(cudaFuncSetAttribute(k16_encode<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TABLE16_BYTES));
            FELICS_CUDA_TRY(cudaFuncSetAttribute(k16_encode<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TABLE16_BYTES));
            ctx->enc16_attr_done = true;
        }
        {
            StageScope s(ctx, ST_CODE);   // size pass
            a.tag0 = next_tags16(ctx);
            k16_encode<false><<<(unsigned)ni, 32, SM_TABLE16_BYTES, st>>>(a, (uint32_t)ni);
            s.launched();
        }
        rc = ensure_buffer(ctx, &ctx->pinned, &ctx->pinned_cap, (2 * ni + 2) * sizeof(uint64_t), true);
        if (rc) return rc;
        uint64_t *h_sizes = (uint64_t *)ctx->pinned, *h_off = h_sizes + ni;
        FELICS_CUDA_TRY(cudaMemcpyAsync(h_sizes, a.sizes, ni * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        FELICS_CUDA_TRY(cudaStreamSynchronize(st));
        h_off[0] = 0;
        for (size_t i = 0; i < ni; i++) h_off[i + 1] = h_off[i] + h_sizes[i];
        const uint64_t sub_total = h_off[ni];
        for (size_t i = 0; i < ni; i++) offsets_host[first + i + 1] = arena_off + h_off[i + 1];
        if (arena_off + sub_total > arena_cap) {
            // keep sizing: the caller learns the total it needs (lower bound when more sub-batches follow)
            arena_off += sub_total;
            for (size_t f2 = first + ni; f2 < n; f2++) offsets_host[f2 + 1] = arena_off;
            offsets_host[n] = arena_off;
            set_error("output capacity %zu too small (need at least %llu)", arena_cap, (unsigned long long)arena_off);
            return FELICS_ERR_BUFFER_TOO_SMALL;
        }
        uint8_t *target = d_arena ? d_arena + arena_off : nullptr;
        if (!d_arena) {
            rc = ensure_buffer(ctx, &ctx->staging_out, &ctx->staging_out_cap, sub_total + 16);
            if (rc) return rc;
            target = (uint8_t *)ctx->staging_out;
        }
        FELICS_CUDA_TRY(cudaMemcpyAsync((void *)a.off, h_off, (ni + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        a.arena = target; a.arena_cap = sub_total;
        {
            StageScope s(ctx, ST_PACK);   // write pass
            a.tag0 = next_tags16(ctx);
            k16_encode<true><<<(unsigned)ni, 32, SM_TABLE16_BYTES, st>>>(a, (uint32_t)ni);
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
