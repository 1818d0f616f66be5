// C ABI of the B200 FELICS engine (include/felics_b200.h): contexts, container header,
// host-memory entry points (staging copies around the device pipeline), profiling.
#include "ctx.h"
#include "../../include/felics_b200_debug.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <new>
#include <utility>
#include <vector>

namespace felics {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int bind_device(felics_ctx *ctx) {
    FELICS_CUDA_TRY(cudaSetDevice(ctx->device));
    return FELICS_OK;
}

int ensure_buffer(felics_ctx *ctx, void **buf, size_t *cap, size_t need, bool pinned_host) {
    (void)ctx;
    if (need <= *cap && *buf) return FELICS_OK;
    size_t ncap = need + need / 8 + 4096;
    if (*buf) {
        if (pinned_host) cudaFreeHost(*buf); else cudaFree(*buf);
        *buf = nullptr; *cap = 0;
    }
    void *p = nullptr;
    cudaError_t e = pinned_host ? cudaMallocHost(&p, ncap) : cudaMalloc(&p, ncap);
    if (e != cudaSuccess) {
        set_error("allocation of %zu bytes failed: %s", ncap, cudaGetErrorString(e));
        return FELICS_ERR_CUDA;
    }
    *buf = p; *cap = ncap;
    return FELICS_OK;
}

StageScope::StageScope(felics_ctx *c, int st, cudaStream_t stream) : ctx(c), stage(st), on(stream ? stream : c->stream) {
    if (!ctx->prof) return;
    auto take = [&]() {
        cudaEvent_t e = nullptr;
        if (!ctx->event_pool.empty()) { e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    };
    a = take(); b = take();
    cudaEventRecord(a, on);
}
StageScope::~StageScope() {
    static const bool sync_each = getenv("FELICS_B200_SYNC") != nullptr;   // debug: localise a faulting stage
    if (sync_each) {
        cudaError_t e = cudaStreamSynchronize(on);
        if (e != cudaSuccess) fprintf(stderr, "felics_b200: stage %d failed: %s\n", stage, cudaGetErrorString(e));
    }
    if (!a) return;
    cudaEventRecord(b, on);
    ctx->prof_pending.push_back({stage, a, b});
}

int profile_collect(felics_ctx *ctx) {
    if (ctx->prof_pending.empty()) return FELICS_OK;
    for (auto &pe : ctx->prof_pending) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(pe.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, pe.a, pe.b);
        if (e == cudaSuccess) ctx->stage_ms[pe.stage] += ms;
        ctx->event_pool.push_back(pe.a);
        ctx->event_pool.push_back(pe.b);
    }
    ctx->prof_pending.clear();
    return FELICS_OK;
}

static const char *kStageNames[ST_COUNT] = {"planes", "hist", "chainscan", "tilebase", "scatter", "prefix", "grpscan",
                                            "spec", "walk", "kfill", "code", "bitscan", "pack", "decode", "unplane", "stream", "compact"};

static int check_header(const felics_header *hdr) {
    if (!hdr) { set_error("null header"); return FELICS_ERR_INVALID_ARGUMENT; }
    if (hdr->color_type > 1) return FELICS_ERR_INVALID_COLOR_TYPE;
    if (hdr->pixel_depth > 1) return FELICS_ERR_INVALID_PIXEL_DEPTH;
    return FELICS_OK;
}

}  // namespace felics

using namespace felics;

extern "C" {
#pragma GCC visibility push(default)

const char *felics_version(void) { return "felics_b200 0.1 (sm_100a)"; }
const char *felics_last_error(void) { return g_err; }

int felics_ctx_create(int device, felics_ctx **out) {
    if (!out) return FELICS_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device: %s (this engine has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return FELICS_ERR_CUDA;
    }
    if (device < 0) FELICS_CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) { set_error("device %d out of range (%d devices)", device, count); return FELICS_ERR_INVALID_ARGUMENT; }
    FELICS_CUDA_TRY(cudaSetDevice(device));
    felics_ctx *ctx = new (std::nothrow) felics_ctx();
    if (!ctx) return FELICS_ERR_CUDA;
    ctx->device = device;
    {
        const char *ns = getenv("FELICS_B200_NO_SPEC");   // debug switch: force the serial epoch walk
        ctx->no_spec = ns && ns[0] == '1';
        const char *s16 = getenv("FELICS_B200_SERIAL16");   // debug switch: serial 16-bit encoder
        ctx->serial16 = s16 && s16[0] == '1';
        const char *eh = getenv("FELICS_B200_EARLY_HOPS");
        ctx->early_hops = eh && eh[0] == '1';
        const char *wp = getenv("FELICS_B200_WALK_PER_SM");
        if (wp) ctx->walk_per_sm = (unsigned)strtoul(wp, nullptr, 0);
        const char *nq = getenv("FELICS_B200_NO_QUADS");
        ctx->no_quads = nq && nq[0] == '1';
        const char *bw = getenv("FELICS_B200_BW16");
        ctx->bw16_opts = bw ? (uint32_t)strtoul(bw, nullptr, 0) : 0u;
        const char *no = getenv("FELICS_B200_NO_OVERLAP");   // profiling switch: one stream, kernels back to back
        ctx->no_overlap = no && no[0] == '1';
        const char *nh = getenv("FELICS_B200_NO_HOP");       // debug switch: no segment hops
        ctx->no_hop = nh && nh[0] == '1';
        const char *ng8 = getenv("FELICS_B200_NO_G8");       // debug switch: gray batches decode one file per warp
        if (ng8 && ng8[0] == '1') ctx->no_g8 = true;
        const char *g8f = getenv("FELICS_B200_G8_FILES");    // experiment switch: files per warp of the gray batch decoder (1, 2, 4, 8)
        if (g8f) { const int f = atoi(g8f); if (f == 1 || f == 2 || f == 4 || f == 8 || f == 16 || f == 32) ctx->g8_files_per_warp = f; }
        const char *vch = getenv("FELICS_B200_V_CHUNK_KB");  // test switch: transfer chunk of the mixed-shape batch calls (default 1 GiB)
        if (vch && atol(vch) > 0) ctx->v_chunk_bytes = (size_t)atol(vch) << 10;
        const char *nst = getenv("FELICS_B200_NO_STREAM");   // debug/bench switch: gray batches through the multi-kernel pipeline
        ctx->no_stream = nst && nst[0] == '1';
        const char *sdbg = getenv("FELICS_B200_STREAM_DBG");
        if (sdbg) ctx->stream_dbg = (uint32_t)strtoul(sdbg, nullptr, 0);
        const char *smin = getenv("FELICS_B200_STREAM_MIN");
        if (smin) ctx->stream_min = (size_t)strtoull(smin, nullptr, 0);
    }
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); delete ctx; return FELICS_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return FELICS_OK;
}

void felics_ctx_destroy(felics_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &pe : ctx->prof_pending) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->staging_in) cudaFree(ctx->staging_in);
    if (ctx->staging_out) cudaFree(ctx->staging_out);
    if (ctx->tables16) cudaFree(ctx->tables16);
    if (ctx->exact_buf) cudaFree(ctx->exact_buf);
    if (ctx->g8_cold) cudaFree(ctx->g8_cold);
    if (ctx->v_in) cudaFree(ctx->v_in);
    if (ctx->v_out) cudaFree(ctx->v_out);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
        if (ctx->ev_pack[i]) cudaEventDestroy(ctx->ev_pack[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
        if (ctx->ev_sizes[i]) cudaEventDestroy(ctx->ev_sizes[i]);
        if (ctx->stage_in[i]) cudaFree(ctx->stage_in[i]);
        if (ctx->stage_out[i]) cudaFree(ctx->stage_out[i]);
    }
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int felics_ctx_set_stream(felics_ctx *ctx, void *cuda_stream) {
    if (!ctx) return FELICS_ERR_INVALID_ARGUMENT;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return FELICS_OK;
}

// format.rs:63-84 (check order: magic, colour, depth, then the two u32)
int felics_read_header(const uint8_t *buf, size_t len, felics_header *out) {
    if (!out || (!buf && len)) return FELICS_ERR_INVALID_ARGUMENT;
    if (len < 4) return FELICS_ERR_IO;
    if (memcmp(buf, "FLCS", 4) != 0) return FELICS_ERR_INVALID_SIGNATURE;
    if (len < 5) return FELICS_ERR_IO;
    if (buf[4] > 1) return FELICS_ERR_INVALID_COLOR_TYPE;
    if (len < 6) return FELICS_ERR_IO;
    if (buf[5] > 1) return FELICS_ERR_INVALID_PIXEL_DEPTH;
    if (len < FELICS_HEADER_BYTES) return FELICS_ERR_IO;
    out->color_type = buf[4];
    out->pixel_depth = buf[5];
    out->width = ((uint32_t)buf[6] << 24) | ((uint32_t)buf[7] << 16) | ((uint32_t)buf[8] << 8) | buf[9];
    out->height = ((uint32_t)buf[10] << 24) | ((uint32_t)buf[11] << 16) | ((uint32_t)buf[12] << 8) | buf[13];
    return FELICS_OK;
}

// format.rs:51-61
int felics_write_header(const felics_header *hdr, uint8_t *o) {
    int rc = check_header(hdr);
    if (rc) return rc;
    if (!o) return FELICS_ERR_INVALID_ARGUMENT;
    memcpy(o, "FLCS", 4);
    o[4] = hdr->color_type; o[5] = hdr->pixel_depth;
    o[6] = (uint8_t)(hdr->width >> 24); o[7] = (uint8_t)(hdr->width >> 16); o[8] = (uint8_t)(hdr->width >> 8); o[9] = (uint8_t)hdr->width;
    o[10] = (uint8_t)(hdr->height >> 24); o[11] = (uint8_t)(hdr->height >> 16); o[12] = (uint8_t)(hdr->height >> 8); o[13] = (uint8_t)hdr->height;
    return FELICS_OK;
}

size_t felics_pixel_bytes(const felics_header *hdr) {
    if (!hdr) return 0;
    return (size_t)hdr->width * hdr->height * (hdr->color_type ? 3 : 1) * (hdr->pixel_depth ? 2 : 1);
}

size_t felics_compress_bound(const felics_header *hdr) {
    if (!hdr) return 0;
    size_t npix = (size_t)hdr->width * hdr->height, nch = hdr->color_type ? 3 : 1;
    // marker 2 + unary (max_context >> 0) + 1 + k bits per sample, 64 raw bits per channel
    size_t per = hdr->pixel_depth ? 131100 : 520;
    return FELICS_HEADER_BYTES + nch * 8 + (npix * nch * per + 7) / 8 + 8;
}

int felics_compress_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header *hdr, uint8_t *d_arena,
                                 size_t arena_cap, uint64_t *offsets) {
    if (!ctx || !offsets || (n && (!d_arena))) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = check_header(hdr);
    if (rc) return rc;
    if ((rc = bind_device(ctx))) return rc;
    offsets[0] = 0;
    if (n == 0) return FELICS_OK;
    if (!d_pixels && felics_pixel_bytes(hdr)) { set_error("null pixels"); return FELICS_ERR_INVALID_ARGUMENT; }
    return encode_batch_device(ctx, n, d_pixels, *hdr, d_arena, nullptr, arena_cap, offsets);
}

int felics_compress_device(felics_ctx *ctx, const void *d_pixels, const felics_header *hdr, uint8_t *d_out, size_t cap, size_t *out_len) {
    uint64_t off[2] = {0, 0};
    int rc = felics_compress_batch_device(ctx, 1, d_pixels, hdr, d_out, cap, off);
    if (out_len) *out_len = (size_t)off[1];
    return rc;
}

int felics_compress_batch(felics_ctx *ctx, size_t n, const void *pixels, const felics_header *hdr, uint8_t *arena, size_t arena_cap,
                          uint64_t *offsets) {
    if (!ctx || !offsets || (n && !arena)) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = check_header(hdr);
    if (rc) return rc;
    if ((rc = bind_device(ctx))) return rc;
    offsets[0] = 0;
    if (n == 0) return FELICS_OK;
    size_t in_bytes = felics_pixel_bytes(hdr) * n;
    if (in_bytes && !pixels) { set_error("null pixels"); return FELICS_ERR_INVALID_ARGUMENT; }
    // batches of gray images: one block per image (stream.cu), sub-batches double buffered over three streams
    if (stream_eligible(ctx, n, nullptr, *hdr)) return stream_encode_batch_host(ctx, n, pixels, *hdr, arena, arena_cap, offsets);
    // 8-bit samples: sub-batches stream through the device (copy in / encode / copy out overlap)
    // (images without pixels: any non-null host pointer selects the streaming path, nothing is read through it)
    if (hdr->pixel_depth == 0) return encode_batch_device(ctx, n, nullptr, *hdr, nullptr, arena, arena_cap, offsets, pixels ? pixels : (const void *)arena);
    if ((rc = ensure_buffer(ctx, &ctx->staging_in, &ctx->staging_in_cap, in_bytes + 16))) return rc;
    if (in_bytes) FELICS_CUDA_TRY(cudaMemcpyAsync(ctx->staging_in, pixels, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    return encode_batch_device(ctx, n, ctx->staging_in, *hdr, nullptr, arena, arena_cap, offsets);
}

int felics_compress(felics_ctx *ctx, const void *pixels, const felics_header *hdr, uint8_t *out, size_t cap, size_t *out_len) {
    uint64_t off[2] = {0, 0};
    int rc = felics_compress_batch(ctx, 1, pixels, hdr, out, cap, off);
    if (out_len) *out_len = (size_t)off[1];
    return rc;
}

int felics_decompress_batch_device(felics_ctx *ctx, size_t n, const uint8_t *d_arena, const uint64_t *offsets, const felics_header *hdr,
                                   void *d_pixels_out, int *status) {
    if (!ctx || !offsets || !status || (n && !d_arena)) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = check_header(hdr);
    if (rc) return rc;
    if ((rc = bind_device(ctx))) return rc;
    if (n == 0) return FELICS_OK;
    return decode_batch_device(ctx, n, d_arena, offsets, *hdr, d_pixels_out, status);
}

int felics_decompress_batch(felics_ctx *ctx, size_t n, const uint8_t *arena, const uint64_t *offsets, const felics_header *hdr,
                            void *pixels_out, int *status) {
    if (!ctx || !offsets || !status || (n && !arena)) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = check_header(hdr);
    if (rc) return rc;
    if ((rc = bind_device(ctx))) return rc;
    if (n == 0) return FELICS_OK;
    if ((rc = check_offsets(n, offsets))) return rc;   // before offsets[n] bytes are taken to be the whole arena
    size_t in_bytes = (size_t)offsets[n];
    size_t out_bytes = felics_pixel_bytes(hdr) * n;
    if ((rc = ensure_buffer(ctx, &ctx->staging_in, &ctx->staging_in_cap, in_bytes + 16))) return rc;
    if ((rc = ensure_buffer(ctx, &ctx->staging_out, &ctx->staging_out_cap, out_bytes + 16))) return rc;
    FELICS_CUDA_TRY(cudaMemcpyAsync(ctx->staging_in, arena, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = decode_batch_device(ctx, n, (const uint8_t *)ctx->staging_in, offsets, *hdr, ctx->staging_out, status);
    // images that decoded are returned even when others failed
    if (out_bytes && pixels_out && (rc == FELICS_OK || rc > FELICS_ERR_CUDA || rc == FELICS_ERR_CORRUPT)) {
        FELICS_CUDA_TRY(cudaMemcpyAsync(pixels_out, ctx->staging_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        FELICS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return rc;
}

int felics_decompress(felics_ctx *ctx, const uint8_t *fel, size_t len, void *pixels_out, size_t cap, felics_header *hdr_out) {
    felics_header hdr;
    int rc = felics_read_header(fel, len, &hdr);
    if (rc) return rc;
    if (hdr_out) *hdr_out = hdr;
    uint64_t npix = (uint64_t)hdr.width * hdr.height;
    if (npix > 0xffffffffull) return FELICS_ERR_INVALID_DIMENSIONS;   // checked_mul (compression.rs:176-180)
    size_t need = felics_pixel_bytes(&hdr);
    if (need > cap) { set_error("pixel buffer too small: need %zu", need); return FELICS_ERR_BUFFER_TOO_SMALL; }
    uint64_t off[2] = {0, (uint64_t)len};
    int status = 0;
    return felics_decompress_batch(ctx, 1, fel, off, &hdr, pixels_out, &status);
}

// ---- batches of mixed shapes: images that share a header travel through the device together -------------------------
namespace {
struct HeaderKey {
    uint8_t color, depth;
    uint32_t w, h;
    bool operator<(const HeaderKey &o) const {
        if (color != o.color) return color < o.color;
        if (depth != o.depth) return depth < o.depth;
        if (w != o.w) return w < o.w;
        return h < o.h;
    }
};
}  // namespace

// Images travel in chunks of consecutive images (about 1 GiB of pixels): every image is copied from the caller's memory
// straight into a device buffer, next to the other images of its shape; each shape is encoded as one device batch; the
// streams go from the device straight to their places in the caller's arena.  No host-side staging copy.
int felics_compress_batch_v(felics_ctx *ctx, size_t n, const void *const *pixels, const felics_header *hdrs, uint8_t *arena, size_t arena_cap,
                            uint64_t *offsets) {
    if (!ctx || !offsets || (n && (!pixels || !hdrs || !arena))) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = bind_device(ctx);
    if (rc) return rc;
    offsets[0] = 0;
    if (n == 0) return FELICS_OK;
    for (size_t i = 0; i < n; i++) {
        if ((rc = check_header(&hdrs[i]))) return rc;
        if (!pixels[i] && felics_pixel_bytes(&hdrs[i])) { set_error("null pixels for image %zu", i); return FELICS_ERR_INVALID_ARGUMENT; }
    }
    cudaStream_t st = ctx->stream;
    const size_t chunk_bytes = ctx->v_chunk_bytes;
    bool too_small = false;
    struct Group { felics_header hdr; std::vector<size_t> idx; size_t in_base, out_base, out_cap; std::vector<uint64_t> off; std::vector<uint8_t> bounce; };
    for (size_t c0 = 0; c0 < n;) {
        size_t c1 = c0, bytes = 0;
        while (c1 < n && (c1 == c0 || bytes + felics_pixel_bytes(&hdrs[c1]) <= chunk_bytes)) bytes += felics_pixel_bytes(&hdrs[c1++]);
        std::map<HeaderKey, size_t> index;
        std::vector<Group> groups;
        for (size_t i = c0; i < c1; i++) {
            const HeaderKey key{hdrs[i].color_type, hdrs[i].pixel_depth, hdrs[i].width, hdrs[i].height};
            auto it = index.find(key);
            if (it == index.end()) { it = index.emplace(key, groups.size()).first; groups.push_back(Group{hdrs[i], {}, 0, 0, 0, {}, {}}); }
            groups[it->second].idx.push_back(i);
        }
        size_t in_total = 0, out_total = 0;
        for (Group &g : groups) {
            const size_t per = felics_pixel_bytes(&g.hdr), m = g.idx.size();
            g.in_base = in_total; in_total = align_up(in_total + per * m + 16, 256);
            g.out_cap = per * m + per * m / 4 + 256 * m + 4096;      // uniform noise needs about 9 / 8 of the input
            g.out_base = out_total; out_total = align_up(out_total + g.out_cap, 256);
            g.off.assign(m + 1, 0);
        }
        if ((rc = ensure_buffer(ctx, &ctx->v_in, &ctx->v_in_cap, in_total + 256))) return rc;
        if ((rc = ensure_buffer(ctx, &ctx->v_out, &ctx->v_out_cap, out_total + 256))) return rc;
        for (Group &g : groups) {
            const size_t per = felics_pixel_bytes(&g.hdr);
            if (!per) continue;
            for (size_t j = 0; j < g.idx.size(); j++)
                FELICS_CUDA_TRY(cudaMemcpyAsync((uint8_t *)ctx->v_in + g.in_base + j * per, pixels[g.idx[j]], per, cudaMemcpyHostToDevice, st));
        }
        for (Group &g : groups) {
            rc = encode_batch_device(ctx, g.idx.size(), (const uint8_t *)ctx->v_in + g.in_base, g.hdr, (uint8_t *)ctx->v_out + g.out_base, nullptr, g.out_cap,
                                     g.off.data());
            if (rc == FELICS_ERR_BUFFER_TOO_SMALL) {
                // the sizes are exact: a buffer of its own for this group (only data that expands by more than a quarter gets here)
                const size_t need = (size_t)g.off[g.idx.size()] + 256;
                void *big = nullptr;
                FELICS_CUDA_TRY(cudaMalloc(&big, need));
                rc = encode_batch_device(ctx, g.idx.size(), (const uint8_t *)ctx->v_in + g.in_base, g.hdr, (uint8_t *)big, nullptr, need, g.off.data());
                if (!rc && !too_small) {
                    // copied out below needs the stream in v_out: it does not fit there, so place this group's images now, at offsets known only
                    // once the earlier images of the chunk are sized -- fall back to a host bounce for this rare case
                    std::vector<uint8_t> bounce((size_t)g.off[g.idx.size()]);
                    if (cudaMemcpyAsync(bounce.data(), big, bounce.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) rc = FELICS_ERR_CUDA;
                    g.bounce.swap(bounce);
                }
                cudaFree(big);
            }
            if (rc) return rc;
        }
        // sizes of the chunk's images, in image order
        std::vector<std::pair<uint32_t, uint32_t>> where(c1 - c0);   // image -> (group, position in the group)
        for (size_t gi = 0; gi < groups.size(); gi++)
            for (size_t j = 0; j < groups[gi].idx.size(); j++) where[groups[gi].idx[j] - c0] = {(uint32_t)gi, (uint32_t)j};
        for (size_t i = c0; i < c1; i++) {
            const Group &g = groups[where[i - c0].first];
            const size_t j = where[i - c0].second;
            offsets[i + 1] = offsets[i] + (g.off[j + 1] - g.off[j]);
        }
        if (offsets[c1] > arena_cap) too_small = true;
        if (!too_small) {
            for (size_t i = c0; i < c1; i++) {
                const Group &g = groups[where[i - c0].first];
                const size_t j = where[i - c0].second;
                const size_t len = (size_t)(g.off[j + 1] - g.off[j]);
                if (!len) continue;
                if (!g.bounce.empty()) memcpy(arena + offsets[i], g.bounce.data() + g.off[j], len);
                else FELICS_CUDA_TRY(cudaMemcpyAsync(arena + offsets[i], (const uint8_t *)ctx->v_out + g.out_base + g.off[j], len, cudaMemcpyDeviceToHost, st));
            }
        }
        FELICS_CUDA_TRY(cudaStreamSynchronize(st));   // the caller's pixels have been read, its arena written; v_in / v_out are free again
        c0 = c1;
    }
    if (too_small) {
        set_error("output capacity %zu too small (need %llu)", arena_cap, (unsigned long long)offsets[n]);
        return FELICS_ERR_BUFFER_TOO_SMALL;
    }
    return FELICS_OK;
}

int felics_decompress_batch_v(felics_ctx *ctx, size_t n, const uint8_t *arena, const uint64_t *offsets, void *const *pixels_out, const size_t *caps,
                              felics_header *hdrs_out, int *status) {
    if (!ctx || !offsets || !status || (n && (!arena || !pixels_out || !caps))) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = bind_device(ctx);
    if (rc) return rc;
    if (n == 0) return FELICS_OK;
    if ((rc = check_offsets(n, offsets))) return rc;
    cudaStream_t st = ctx->stream;
    std::vector<felics_header> hs(n);
    std::vector<uint8_t> ok(n, 0);
    for (size_t i = 0; i < n; i++) {
        status[i] = felics_read_header(arena + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), &hs[i]);   // read_header's own errors (format.rs:63-84)
        if (status[i]) continue;
        if (hdrs_out) hdrs_out[i] = hs[i];
        const uint64_t npix = (uint64_t)hs[i].width * hs[i].height;
        if (npix > 0xffffffffull) { status[i] = FELICS_ERR_INVALID_DIMENSIONS; continue; }   // checked_mul (compression.rs:176-180)
        if (felics_pixel_bytes(&hs[i]) > caps[i]) { status[i] = FELICS_ERR_BUFFER_TOO_SMALL; continue; }
        ok[i] = 1;
    }
    const size_t chunk_bytes = ctx->v_chunk_bytes;
    struct Group { felics_header hdr; std::vector<size_t> idx; std::vector<uint64_t> off; size_t in_base, out_base; };
    for (size_t c0 = 0; c0 < n;) {
        size_t c1 = c0, bytes = 0;
        while (c1 < n && (c1 == c0 || bytes + (ok[c1] ? felics_pixel_bytes(&hs[c1]) : 0) <= chunk_bytes)) { bytes += ok[c1] ? felics_pixel_bytes(&hs[c1]) : 0; c1++; }
        std::map<HeaderKey, size_t> index;
        std::vector<Group> groups;
        for (size_t i = c0; i < c1; i++) {
            if (!ok[i]) continue;
            const HeaderKey key{hs[i].color_type, hs[i].pixel_depth, hs[i].width, hs[i].height};
            auto it = index.find(key);
            if (it == index.end()) { it = index.emplace(key, groups.size()).first; groups.push_back(Group{hs[i], {}, {}, 0, 0}); }
            groups[it->second].idx.push_back(i);
        }
        size_t in_total = 0, out_total = 0;
        for (Group &g : groups) {
            const size_t m = g.idx.size();
            g.off.assign(m + 1, 0);
            for (size_t j = 0; j < m; j++) g.off[j + 1] = g.off[j] + (offsets[g.idx[j] + 1] - offsets[g.idx[j]]);
            g.in_base = in_total; in_total = align_up(in_total + (size_t)g.off[m] + 16, 256);
            g.out_base = out_total; out_total = align_up(out_total + felics_pixel_bytes(&g.hdr) * m + 16, 256);
        }
        if ((rc = ensure_buffer(ctx, &ctx->v_in, &ctx->v_in_cap, in_total + 256))) return rc;
        if ((rc = ensure_buffer(ctx, &ctx->v_out, &ctx->v_out_cap, out_total + 256))) return rc;
        for (Group &g : groups)
            for (size_t j = 0; j < g.idx.size(); j++)
                if (g.off[j + 1] > g.off[j])
                    FELICS_CUDA_TRY(cudaMemcpyAsync((uint8_t *)ctx->v_in + g.in_base + g.off[j], arena + offsets[g.idx[j]], (size_t)(g.off[j + 1] - g.off[j]), cudaMemcpyHostToDevice, st));
        for (Group &g : groups) {
            const size_t per = felics_pixel_bytes(&g.hdr), m = g.idx.size();
            std::vector<int> sg(m, 0);
            rc = decode_batch_device(ctx, m, (const uint8_t *)ctx->v_in + g.in_base, g.off.data(), g.hdr, (uint8_t *)ctx->v_out + g.out_base, sg.data());
            bool per_image = false;
            for (size_t j = 0; j < m; j++) per_image |= sg[j] == rc;
            if (rc && !per_image) return rc;   // a failure of the call, not of one file
            for (size_t j = 0; j < m; j++) {
                status[g.idx[j]] = sg[j];
                if (sg[j] == FELICS_OK && per)
                    FELICS_CUDA_TRY(cudaMemcpyAsync(pixels_out[g.idx[j]], (const uint8_t *)ctx->v_out + g.out_base + j * per, per, cudaMemcpyDeviceToHost, st));
            }
        }
        FELICS_CUDA_TRY(cudaStreamSynchronize(st));
        c0 = c1;
    }
    for (size_t i = 0; i < n; i++)
        if (status[i]) return status[i];
    return FELICS_OK;
}

int felics_sidecar_build(felics_ctx *ctx, uint32_t band_rows, uint8_t *sidecar_out, size_t cap, size_t *out_len) {
    if (!ctx || (!sidecar_out && cap)) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = bind_device(ctx);
    if (rc) return rc;
    return sidecar_build(ctx, band_rows, sidecar_out, cap, out_len);
}

int felics_decompress_sidecar(felics_ctx *ctx, const uint8_t *fel, size_t len, const uint8_t *sidecar, size_t sidecar_len, void *pixels_out,
                              size_t cap, felics_header *hdr_out) {
    if (!ctx || !fel || !sidecar || (!pixels_out && cap)) { set_error("null argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = bind_device(ctx);
    if (rc) return rc;
    return decode_sidecar(ctx, fel, len, sidecar, sidecar_len, pixels_out, cap, hdr_out);
}

int felics_decompress_device(felics_ctx *ctx, const uint8_t *d_fel, size_t len, void *d_pixels_out, size_t cap, felics_header *hdr_out) {
    if (!ctx || !d_fel) return FELICS_ERR_INVALID_ARGUMENT;
    int rc = bind_device(ctx);
    if (rc) return rc;
    uint8_t hb[FELICS_HEADER_BYTES];
    size_t hl = len < FELICS_HEADER_BYTES ? len : FELICS_HEADER_BYTES;
    if (hl) FELICS_CUDA_TRY(cudaMemcpy(hb, d_fel, hl, cudaMemcpyDeviceToHost));
    felics_header hdr;
    if ((rc = felics_read_header(hb, hl, &hdr))) return rc;
    if (hdr_out) *hdr_out = hdr;
    uint64_t npix = (uint64_t)hdr.width * hdr.height;
    if (npix > 0xffffffffull) return FELICS_ERR_INVALID_DIMENSIONS;
    size_t need = felics_pixel_bytes(&hdr);
    if (need > cap) { set_error("pixel buffer too small: need %zu", need); return FELICS_ERR_BUFFER_TOO_SMALL; }
    uint64_t off[2] = {0, (uint64_t)len};
    int status = 0;
    return felics_decompress_batch_device(ctx, 1, d_fel, off, &hdr, d_pixels_out, &status);
}

int felics_profile_enable(felics_ctx *ctx, int on) {
    if (!ctx) return FELICS_ERR_INVALID_ARGUMENT;
    ctx->prof = on != 0;
    return FELICS_OK;
}
int felics_profile_reset(felics_ctx *ctx) {
    if (!ctx) return FELICS_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < ST_COUNT; i++) { ctx->stage_ms[i] = 0; ctx->stage_launches[i] = 0; }
    ctx->total_launches = 0;
    return FELICS_OK;
}
int felics_profile_stage_count(void) { return ST_COUNT; }
const char *felics_profile_stage_name(int stage) { return stage >= 0 && stage < ST_COUNT ? kStageNames[stage] : ""; }
double felics_profile_stage_ms(felics_ctx *ctx, int stage) { return ctx && stage >= 0 && stage < ST_COUNT ? ctx->stage_ms[stage] : 0.0; }
uint64_t felics_profile_stage_launches(felics_ctx *ctx, int stage) { return ctx && stage >= 0 && stage < ST_COUNT ? ctx->stage_launches[stage] : 0; }
uint64_t felics_profile_total_launches(felics_ctx *ctx) { return ctx ? ctx->total_launches : 0; }

// Debug aid for the parity tests (not in the public header): copy the per-pixel code
// records of the last encode call (length in bits 31..22) to the host.
int felics_debug_last_records(felics_ctx *ctx, uint32_t *out, size_t count) {
    if (!ctx || !ctx->dbg_rec || count > ctx->dbg_rec_count) return FELICS_ERR_INVALID_ARGUMENT;
    FELICS_CUDA_TRY(cudaMemcpy(out, ctx->dbg_rec, count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return FELICS_OK;
}

// Debug aid: the 8 device counters of the last encode sub-batch (queue sizes, flags, walker cycle counts).
int felics_debug_counters(felics_ctx *ctx, uint32_t *out8) {
    if (!ctx || !out8) return FELICS_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 8; i++) out8[i] = ctx->dbg_counters[i];
    return FELICS_OK;
}

#pragma GCC visibility pop
}  // extern "C"
