// Internal context of the B200 FELICS engine (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <string>

#include "../../include/felics_b200.h"

namespace felics {

// Pipeline stages that get their own CUDA-event bracket when profiling is on.
enum Stage : int {
    ST_PLANES = 0,   // pixels -> i16 planes (+ YCoCg-R for RGB)
    ST_HIST,         // per-tile context histograms of out-of-range pixels
    ST_CHAINSCAN,    // per-plane chain sizes / bases
    ST_TILEBASE,     // per-tile, per-context scatter bases
    ST_SCATTER,      // stable grouping of residuals by context
    ST_PREFIX,       // per-32-block code-cost prefix sums for the 6 k candidates
    ST_GRPSCAN,      // scan of the per-1024 group totals
    ST_SPEC,         // speculative parallel epoch walk of long chains (verified; falls back to ST_WALK)
    ST_WALK,         // halving-epoch walk (the sequential part of the estimator)
    ST_KFILL,        // k per out-of-range pixel
    ST_CODE,         // code word + length per pixel, bits per tile
    ST_BITSCAN,      // bit offsets: tiles -> planes -> images
    ST_PACK,         // MSB-first bit packing
    ST_DECODE,       // bit-serial decode, one stream per warp
    ST_UNPLANE,      // i16 planes -> pixels (+ inverse YCoCg-R), range checks
    ST_STREAM,       // streaming band encoder: one block per image, everything between pixels and bits on chip (stream.cu)
    ST_COMPACT,      // image offsets + copy of the finished streams to their place in the arena
    ST_COUNT
};

struct ProfEntry { int stage; cudaEvent_t a, b; };

// What the sidecar builder (sidecar.cuh) needs from the last single-image 8-bit encode: pointers into the context's scratch
// (valid until the next call on the context) and the geometry they belong to.
struct LastEncode {
    bool valid = false;
    uint32_t w = 0, h = 0, npix = 0, nch = 0, tpp = 0, cap = 0;
    const void *planes = nullptr;    // u8 samples (gray) or i16 planes (RGB)
    bool planes_u8 = false;
    const uint32_t *tile_base = nullptr, *chain_count = nullptr, *chain_base = nullptr, *blk_rec = nullptr, *blk_epoch = nullptr, *ep_rec = nullptr;
    const uint16_t *e_grp = nullptr;
    const uint64_t *tile_off = nullptr, *plane_bits = nullptr;
    uint64_t fel_bytes = 0;
};

void set_error(const char *fmt, ...);

}  // namespace felics

struct felics_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;      // second stream: the serial epoch walk runs beside the speculative one
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    void *scratch = nullptr;      // device scratch, grown on demand
    size_t scratch_cap = 0;
    size_t batch_budget = 0;      // scratch budget of batches through the multi-kernel pipeline (0 = not yet asked)
    size_t v_chunk_bytes = (size_t)1 << 30;   // pixels per transfer chunk of the mixed-shape batch calls
    void *v_in = nullptr;         // device buffers of the mixed-shape batch calls (felics_*_batch_v)
    size_t v_in_cap = 0;
    void *v_out = nullptr;
    size_t v_out_cap = 0;
    void *staging_in = nullptr;   // device staging for host-memory entry points
    size_t staging_in_cap = 0;
    void *staging_out = nullptr;
    size_t staging_out_cap = 0;
    // host-memory batches: sub-batches are copied in, encoded and copied out on three streams (double buffered)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_sizes[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_pack[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    void *stage_in[2] = {nullptr, nullptr};
    size_t stage_in_cap[2] = {0, 0};
    void *stage_out[2] = {nullptr, nullptr};
    size_t stage_out_cap[2] = {0, 0};
    void *exact_buf = nullptr;    // i32 planes of one image: files the fast decoders hand to k_decode_exact
    size_t exact_cap = 0;
    void *pinned = nullptr;       // small pinned host buffer for read-backs
    size_t pinned_cap = 0;

    // per-context (hence per-device) one-time kernel attribute settings
    bool walk_attr_done = false, sp_attr_done = false, hop_attr_done = false, enc16_attr_done = false, dec16_attr_done = false;
    // 16-bit path: per-image estimator tables (tagged rows, serial16.cuh)
    void *tables16 = nullptr;
    size_t tables16_cap = 0;
    uint32_t tag16 = 1;
    size_t decode_smem_set = 0, dec16_smem_set = 0, decode_bands_smem_set = 0;

    bool prof = false;
    bool no_overlap = false;      // debug/profiling switch: run the serial walk after the speculative one, on the same stream
    bool no_hop = false;          // debug switch: no segment hops in the serial walker
    bool no_spec = false;         // debug/bench switch: skip the speculative walk
    uint32_t bw16_opts = 0;       // experiment switches of the 16-bit bucket walk (FELICS_B200_BW16)
    bool early_hops = false;      // experiment: hop tables for every planned chain, built before the speculative walk
    unsigned walk_per_sm = 1;     // walker blocks per SM while the speculative kernels run beside them
    bool no_quads = false;        // debug switch: one sample per thread in the histogram / code kernels
    bool serial16 = false;        // debug switch: the one-warp-per-image 16-bit encoder instead of the parallel one
    bool no_g8 = false;           // debug switch: gray batches through the one-file-per-warp decoder (k_decode) instead of k_decode_g8
    void *g8_cold = nullptr;      // k_decode_g8: ticket + estimator rows of the other contexts
    size_t g8_cold_cap = 0;
    int g8_files_per_warp = 0;    // experiment switch: files per warp of k_decode_g8 (0 = chosen from the batch size)
    bool no_stream = false;       // debug/bench switch: batches of gray images through the multi-kernel pipeline instead of stream.cu
    size_t stream_min = 96;       // smallest batch the streaming band encoder takes (one block per image: fewer leave SMs idle)
    bool stream_attr_done = false;
    uint32_t stream_dbg = 0;      // timing experiments (wrong output): see StreamArgs::dbg
    int sm_count = 148;
    uint64_t stream_redone = 0;   // images re-encoded by the general pipeline because their stream overflowed its slot
    std::vector<felics::ProfEntry> prof_pending;
    std::vector<cudaEvent_t> event_pool;
    double stage_ms[felics::ST_COUNT] = {0};
    uint64_t stage_launches[felics::ST_COUNT] = {0};
    uint64_t total_launches = 0;

    // debug: device pointer / count of the per-pixel code records of the last encode
    const uint32_t *dbg_rec = nullptr;
    size_t dbg_rec_count = 0;
    uint32_t dbg_counters[8] = {0};
    felics::LastEncode last;          // for felics_sidecar_build
};

namespace felics {

#define FELICS_CUDA_TRY(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            felics::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FELICS_ERR_CUDA;                                                        \
        }                                                                                  \
    } while (0)

int ensure_buffer(felics_ctx *ctx, void **buf, size_t *cap, size_t need, bool pinned_host = false);
int bind_device(felics_ctx *ctx);

// RAII-less stage bracket: begin() records an event if profiling, end() likewise.
struct StageScope {
    felics_ctx *ctx;
    int stage;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t on;
    StageScope(felics_ctx *c, int st, cudaStream_t stream = nullptr);
    ~StageScope();
    void launched(int n = 1) { ctx->stage_launches[stage] += n; ctx->total_launches += n; }
};
int profile_collect(felics_ctx *ctx);  // after a stream sync: fold pending events into stage_ms

// encode.cu
int encode_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr,
                        uint8_t *d_arena, uint8_t *h_arena, size_t arena_cap, uint64_t *offsets_host, const void *h_pixels = nullptr);
// stream.cu: batches of 8-bit gray images, one block per image
bool stream_eligible(const felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr);
int stream_encode_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr, uint8_t *d_arena, size_t arena_cap,
                               uint64_t *offsets_host);
int stream_encode_batch_host(felics_ctx *ctx, size_t n, const void *h_pixels, const felics_header &hdr, uint8_t *h_arena, size_t arena_cap,
                             uint64_t *offsets_host);
// encode16.cu: 16-bit samples
int encode16_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr,
                          uint8_t *d_arena, uint8_t *h_arena, size_t arena_cap, uint64_t *offsets_host);
int encode16_serial_batch_device(felics_ctx *ctx, size_t n, const void *d_pixels, const felics_header &hdr,
                                 uint8_t *d_arena, uint8_t *h_arena, size_t arena_cap, uint64_t *offsets_host);
int tables16(felics_ctx *ctx, size_t images, uint32_t **out);
uint32_t next_tags16(felics_ctx *ctx);
// sidecar (opt-in, NOT the reference format): band snapshots that let one big image decode in parallel
int sidecar_build(felics_ctx *ctx, uint32_t band_rows, uint8_t *h_out, size_t cap, size_t *out_len);
int decode_sidecar(felics_ctx *ctx, const uint8_t *h_fel, size_t len, const uint8_t *h_side, size_t side_len, void *h_pixels_out, size_t cap,
                   felics_header *hdr_out);
// decode.cu
int check_offsets(size_t n, const uint64_t *offsets);
int decode_batch_device(felics_ctx *ctx, size_t n, const uint8_t *d_arena, const uint64_t *offsets_host,
                        const felics_header &hdr, void *d_pixels_out, int *status_host);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace felics
