// FELICS decode for sm_100a.
//
// Replaces decompress_channel (/root/reference/src/compression.rs:151-248).  A plane's
// bit stream is strictly serial (every code length depends on the previously decoded
// samples and on the adaptive k), and the three planes of an RGB file share one
// unaligned stream with no recorded lengths (compression.rs:392-394), so one decoder
// walks one file.  Parallelism comes from the batch: one warp per file, the estimator
// table (511 contexts x 6 counters, parameter_selection.rs:29-33) in shared memory.
#include "ctx.h"
#include "device_common.cuh"

#include <algorithm>

namespace felics {

struct DecArgs {
    const uint32_t *words;     // the arena viewed as 32-bit words (4-byte aligned)
    uint64_t arena_words;      // number of whole words that may be read
    const uint64_t *offsets;   // device copy, n+1 entries (bytes)
    int16_t *planes;           // [n*nch][npix]
    int *status;               // [n]
    uint32_t w, h, npix, nch;
    uint8_t color, depth;
};

struct BitReader {
    const uint32_t *words;
    uint64_t nwords;
    uint64_t pos, end;   // absolute bit positions in the arena
    uint64_t cwi;        // cached word index
    uint32_t c0, c1;
    bool eof;

    __device__ __forceinline__ void init(const uint32_t *w, uint64_t nw, uint64_t p, uint64_t e) {
        words = w; nwords = nw; pos = p; end = e; cwi = ~0ull; c0 = c1 = 0; eof = false;
    }
    __device__ __forceinline__ uint32_t load(uint64_t wi) const { return wi < nwords ? bswap32(__ldg(words + wi)) : 0u; }
    // next 32 bits of the stream, MSB first (bits past `end` are whatever follows; callers bound by `end`)
    __device__ __forceinline__ uint32_t peek32() {
        uint64_t wi = pos >> 5;
        if (wi != cwi) {
            c0 = (wi == cwi + 1) ? c1 : load(wi);
            c1 = load(wi + 1);
            cwi = wi;
        }
        uint32_t sh = (uint32_t)(pos & 31);
        return sh ? ((c0 << sh) | (c1 >> (32 - sh))) : c0;
    }
    // BitRead::read(n), n in 0..32
    __device__ __forceinline__ uint32_t read(uint32_t n) {
        if (n == 0) return 0;
        if (pos + n > end) { eof = true; pos = end; return 0; }
        uint32_t v = peek32() >> (32 - n);
        pos += n;
        return v;
    }
    // BitRead::read_unary0: ones up to the terminating zero
    __device__ __forceinline__ uint32_t read_unary0() {
        uint32_t q = 0;
        for (;;) {
            if (pos >= end) { eof = true; return q; }
            uint64_t avail64 = end - pos;
            uint32_t avail = avail64 > 32 ? 32u : (uint32_t)avail64;
            uint32_t ones = __clz(~peek32());   // 32 when all ones
            if (ones >= avail) { q += avail; pos += avail; if (avail < 32) { eof = true; return q; } continue; }
            q += ones;
            pos += ones + 1;
            return q;
        }
    }
};

__global__ void __launch_bounds__(32) k_decode(DecArgs a, uint32_t n) {
    __shared__ uint32_t tab[(NBIN - 1) * NK];
    const uint32_t img = blockIdx.x;
    if (img >= n) return;
    const uint32_t lane = threadIdx.x;
    const uint64_t off0 = a.offsets[img], off1 = a.offsets[img + 1];
    const uint64_t len = off1 - off0;
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(a.words);
    int st = FELICS_OK;
    // read_header order (format.rs:63-84), then decompress_with_header's checks (compression.rs:289-294)
    if (lane == 0) {
        const uint8_t *hb = bytes + off0;
        if (len < 4) st = FELICS_ERR_IO;
        else if (hb[0] != 'F' || hb[1] != 'L' || hb[2] != 'C' || hb[3] != 'S') st = FELICS_ERR_INVALID_SIGNATURE;
        else if (len < 5) st = FELICS_ERR_IO;
        else if (hb[4] > 1) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (len < 6) st = FELICS_ERR_IO;
        else if (hb[5] > 1) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else if (len < FELICS_HEADER_BYTES) st = FELICS_ERR_IO;
        else if (hb[4] != a.color) st = FELICS_ERR_INVALID_COLOR_TYPE;
        else if (hb[5] != a.depth) st = FELICS_ERR_INVALID_PIXEL_DEPTH;
        else {
            uint32_t w = ((uint32_t)hb[6] << 24) | ((uint32_t)hb[7] << 16) | ((uint32_t)hb[8] << 8) | hb[9];
            uint32_t h = ((uint32_t)hb[10] << 24) | ((uint32_t)hb[11] << 16) | ((uint32_t)hb[12] << 8) | hb[13];
            if (w != a.w || h != a.h) st = FELICS_ERR_INVALID_DIMENSIONS;
        }
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    if (st != FELICS_OK) { if (lane == 0) a.status[img] = st; return; }

    BitReader br;
    br.init(a.words, a.arena_words, 8 * (off0 + FELICS_HEADER_BYTES), 8 * off1);
    const uint32_t w = a.w;

    for (uint32_t ch = 0; ch < a.nch && st == FELICS_OK; ch++) {
        for (uint32_t j = lane; j < (NBIN - 1) * NK; j += 32) tab[j] = 0;   // fresh estimator per channel (:186-190)
        __syncwarp();
        if (lane == 0) {
            int16_t *pl = a.planes + ((size_t)img * a.nch + ch) * a.npix;
            int32_t p1 = (int32_t)br.read(32);   // read_signed(32) twice (:161-162)
            int32_t p2 = (int32_t)br.read(32);
            if (br.eof) st = FELICS_ERR_IO;
            else if (a.npix >= 1) {
                if (p1 < -32768 || p1 > 32767 || (a.npix >= 2 && (p2 < -32768 || p2 > 32767))) st = FELICS_ERR_INVALID_VALUE;
                else {
                    pl[0] = (int16_t)p1;
                    if (a.npix >= 2) pl[1] = (int16_t)p2;
                }
            }
            uint32_t x = 0, y = 0;
            if (a.npix >= 3) { x = 2 % w; y = 2 / w; }
            for (uint32_t i = 2; i < a.npix && st == FELICS_OK; i++) {
                uint32_t ia, ib;
                if (x > 0 && y > 0) { ia = i - 1; ib = i - w; }
                else if (y == 0) { ia = i - 1; ib = i - 2; }
                else if (y >= 2) { ia = i - w; ib = i - 2 * w; }
                else { ia = i - w; ib = i - w + 1; }
                int v1 = pl[ia], v2 = pl[ib];
                int hi = max(v1, v2), lo = min(v1, v2);
                uint32_t ctx = (uint32_t)(hi - lo);
                if (ctx > 510u) { st = FELICS_ERR_CORRUPT; break; }   // assert!(context <= max_context), parameter_selection.rs:72
                uint32_t *row = tab + ctx * NK;
                int value;
                if (br.read(1)) {                                       // InRange (:208-215)
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    uint32_t nn = ctx + 1;
                    int m = 31 - __clz(nn);
                    uint32_t left_p = nn - (1u << m), right_p = (2u << m) - nn;
                    uint32_t xx = br.read((uint32_t)m);
                    if (xx >= right_p) xx = (xx - right_p) * 2 + right_p + br.read(1);   // phase_in_coding.rs:102-109
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    xx += left_p;                                                        // rotate_left (:55-57)
                    if (xx >= nn) xx -= nn;
                    if (xx >= nn) { st = FELICS_ERR_CORRUPT; break; }
                    value = lo + (int)xx;
                } else {
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    uint32_t above = br.read(1);
                    uint32_t rr[NK];
#pragma unroll
                    for (int k = 0; k < NK; k++) rr[k] = row[k];
                    int k = argmin_last(rr);                              // get_k (:202)
                    uint32_t q = br.read_unary0();
                    uint32_t rem = br.read((uint32_t)k);
                    if (br.eof) { st = FELICS_ERR_IO; break; }
                    if (q > 70000u) { st = FELICS_ERR_INVALID_VALUE; break; }
                    uint32_t e = (q << k) + rem;
                    uint32_t mn = 0xffffffffu;
#pragma unroll
                    for (int kk = 0; kk < NK; kk++) {                     // update (parameter_selection.rs:49-65)
                        rr[kk] += (e >> kk) + 1u + (uint32_t)kk;
                        mn = min(mn, rr[kk]);
                    }
                    if (mn > HALVE_AT) {
#pragma unroll
                        for (int kk = 0; kk < NK; kk++) rr[kk] >>= 1;
                    }
#pragma unroll
                    for (int kk = 0; kk < NK; kk++) row[kk] = rr[kk];
                    value = above ? hi + (int)e + 1 : lo - (int)e - 1;    // (:216-243)
                }
                if (value < -32768 || value > 32767) { st = FELICS_ERR_INVALID_VALUE; break; }
                pl[i] = (int16_t)value;
                if (++x == w) { x = 0; y++; }
            }
        }
        st = __shfl_sync(0xffffffffu, st, 0);
    }
    if (lane == 0) a.status[img] = st;
}

// planes -> pixels with the try_into range checks (compression.rs:305-310, :402-407)
__global__ void k_unplane_gray8(const int16_t *__restrict__ planes, uint8_t *__restrict__ px, uint32_t npix, size_t total,
                                int *__restrict__ status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        size_t img = i / npix;
        if (status[img] != FELICS_OK) continue;
        int v = planes[i];
        if (v < 0 || v > 255) { atomicCAS(&status[img], FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
        px[i] = (uint8_t)v;
    }
}

// color_transform.rs:20-26
__global__ void k_unplane_rgb8(const int16_t *__restrict__ planes, uint8_t *__restrict__ px, uint32_t npix, size_t total,
                               int *__restrict__ status) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        size_t img = idx / npix;
        if (status[img] != FELICS_OK) continue;
        uint32_t i = (uint32_t)(idx - img * npix);
        const int16_t *base = planes + img * 3 * (size_t)npix;
        int y = base[i], co = base[(size_t)npix + i], cg = base[2 * (size_t)npix + i];
        int t = y - cg / 2;
        int g = cg + t;
        int b = t - co / 2;
        int r = b + co;
        if (r < 0 || r > 255 || g < 0 || g > 255 || b < 0 || b > 255) { atomicCAS(&status[img], FELICS_OK, FELICS_ERR_INVALID_VALUE); continue; }
        px[3 * idx] = (uint8_t)r; px[3 * idx + 1] = (uint8_t)g; px[3 * idx + 2] = (uint8_t)b;
    }
}

int decode_batch_device(felics_ctx *ctx, size_t n, const uint8_t *d_arena, const uint64_t *offsets_host,
                        const felics_header &hdr, void *d_pixels_out, int *status_host) {
    if (hdr.pixel_depth != 0) {
        set_error("16-bit samples are not built yet (traits.rs:35-43 is a 'next' row)");
        return FELICS_ERR_UNSUPPORTED;
    }
    if (((uintptr_t)d_arena & 3) != 0) {
        set_error("device arena must be 4-byte aligned");
        return FELICS_ERR_INVALID_ARGUMENT;
    }
    uint64_t npix64 = (uint64_t)hdr.width * hdr.height;
    if (npix64 > 0xffffffffull) return FELICS_ERR_INVALID_DIMENSIONS;   // checked_mul, compression.rs:176-180
    if (npix64 > 0x7fff0000ull) { set_error("image too large for one call"); return FELICS_ERR_INVALID_DIMENSIONS; }
    cudaStream_t st = ctx->stream;
    const uint32_t npix = (uint32_t)npix64;
    const uint32_t nch = hdr.color_type ? 3 : 1;

    size_t off_bytes = align_up((n + 1) * sizeof(uint64_t), 256);
    size_t stat_bytes = align_up(n * sizeof(int), 256);
    size_t plane_bytes = align_up(((size_t)n * nch * npix + 8) * sizeof(int16_t), 256);
    int rc = ensure_buffer(ctx, &ctx->scratch, &ctx->scratch_cap, off_bytes + stat_bytes + plane_bytes);
    if (rc) return rc;
    uint8_t *sb = (uint8_t *)ctx->scratch;
    uint64_t *d_off = (uint64_t *)sb;
    int *d_status = (int *)(sb + off_bytes);
    int16_t *d_planes = (int16_t *)(sb + off_bytes + stat_bytes);
    FELICS_CUDA_TRY(cudaMemcpyAsync(d_off, offsets_host, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));

    DecArgs a;
    a.words = (const uint32_t *)d_arena;
    a.arena_words = (offsets_host[n] + 3) / 4;   // the word holding the last byte must be readable
    a.offsets = d_off; a.planes = d_planes; a.status = d_status;
    a.w = hdr.width; a.h = hdr.height; a.npix = npix; a.nch = nch;
    a.color = hdr.color_type; a.depth = hdr.pixel_depth;
    {
        StageScope s(ctx, ST_DECODE);
        k_decode<<<(unsigned)n, 32, 0, st>>>(a, (uint32_t)n);
        s.launched();
    }
    if (npix > 0) {
        StageScope s(ctx, ST_UNPLANE);
        size_t total = n * (size_t)npix;
        unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
        if (nch == 1) k_unplane_gray8<<<blocks, 256, 0, st>>>(d_planes, (uint8_t *)d_pixels_out, npix, total, d_status);
        else k_unplane_rgb8<<<blocks, 256, 0, st>>>(d_planes, (uint8_t *)d_pixels_out, npix, total, d_status);
        s.launched();
    }
    FELICS_CUDA_TRY(cudaMemcpyAsync(status_host, d_status, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    FELICS_CUDA_TRY(cudaStreamSynchronize(st));
    FELICS_CUDA_TRY(cudaGetLastError());
    rc = profile_collect(ctx);
    if (rc) return rc;
    for (size_t i = 0; i < n; i++)
        if (status_host[i] != FELICS_OK) return status_host[i];
    return FELICS_OK;
}

}  // namespace felics
