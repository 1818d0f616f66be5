// Band sidecar (opt-in, NOT part of the reference format): a .fel file is one serial bit chain per
// plane, so a single big image decodes on one warp.  The sidecar cuts every plane into bands of `band_rows` rows and records,
// for the start of each band, what a decoder needs to start there: the bit position, the estimator table
// (parameter_selection.rs:24-33, all 511 x 6 counts as they stand before the band's first pixel), the row above and the
// first sample of the row above that (misc.rs:6-24).  Bands then decode side by side.  The .fel bytes are untouched.
//
// Layout (little endian):  header 32 B {"FLSC", version 2, width, height, channels, band_rows, bands per plane, entry bytes}
//                          entries, plane-major: {u64 bit position in the file, i32 col0_b, u32 0, u32 table[511 * 6], i16 row[width], padded to 8 bytes}
// (version 1 padded the row to 4 bytes only: odd entries of widths with width % 4 in {1, 2} were 4-byte aligned under 8-byte accesses)
// Included by encode.cu (the builder reads the scratch of the last encode) and by decode.cu (constants only).
#pragma once
#include <stdint.h>

namespace felics {

constexpr uint32_t SIDECAR_MAGIC = 0x43534C46u;   // "FLSC"
constexpr uint32_t SIDECAR_HEADER_BYTES = 32;
constexpr uint32_t SIDECAR_VERSION = 2;
constexpr uint32_t SIDECAR_TABLE_WORDS = (NBIN - 1) * NK;
inline size_t sidecar_entry_bytes(uint32_t w) { return 16 + (size_t)SIDECAR_TABLE_WORDS * 4 + (((size_t)w * 2 + 7) & ~(size_t)7); }   // a multiple of 8: every entry starts 8-byte aligned
// rows per band must put every band start on a tile boundary (the encoder knows bit offsets and chain positions per tile)
inline uint32_t sidecar_row_unit(uint32_t w) {
    uint32_t a = w, b = TILE;
    while (b) { const uint32_t t = a % b; a = b; b = t; }
    return TILE / a;   // TILE / gcd(w, TILE)
}

#ifdef FELICS_SIDECAR_BUILDER
struct SidecarArgs {
    LastEncode le;
    const uint64_t *img_off;
    uint64_t arena_byte0;
    uint32_t band_rows, nbands;
    size_t entry_bytes;
    uint8_t *out;          // device, entries only
};

// one block per (plane, band), thread c = context c
__global__ void __launch_bounds__(NBIN) k_sidecar(SidecarArgs a) {
    const uint32_t p = blockIdx.x / a.nbands, j = blockIdx.x - p * a.nbands, c = threadIdx.x;
    const LastEncode &le = a.le;
    uint8_t *ent = a.out + (size_t)blockIdx.x * a.entry_bytes;
    const uint32_t y0 = j * a.band_rows;
    const uint32_t tile = (uint32_t)(((uint64_t)y0 * le.w) / TILE);   // exact: band starts sit on tile boundaries
    if (c == 0) {
        uint64_t bit = 8ull * FELICS_HEADER_BYTES;
        for (uint32_t q = 0; q < p; q++) bit += le.plane_bits[q];
        if (j) bit += 64 + le.tile_off[(size_t)p * le.tpp + tile];
        *reinterpret_cast<uint64_t *>(ent) = bit;
        int col0_b = 0;
        if (y0 >= 2) {
            const size_t i = (size_t)p * le.npix + (size_t)(y0 - 2) * le.w;
            col0_b = le.planes_u8 ? (int)reinterpret_cast<const uint8_t *>(le.planes)[i] : (int)reinterpret_cast<const int16_t *>(le.planes)[i];
        }
        reinterpret_cast<int32_t *>(ent)[2] = col0_b;
        reinterpret_cast<uint32_t *>(ent)[3] = 0u;
    }
    uint32_t *tab = reinterpret_cast<uint32_t *>(ent + 16);
    int16_t *row = reinterpret_cast<int16_t *>(ent + 16 + (size_t)SIDECAR_TABLE_WORDS * 4);
    if (j == 0) {
        // the plane start: fresh estimator, no rows above (the decoder reads the two raw samples itself)
        for (uint32_t t = c; t < SIDECAR_TABLE_WORDS; t += NBIN) tab[t] = 0u;
        for (uint32_t x = c; x < le.w; x += NBIN) row[x] = 0;
        return;
    }
    for (uint32_t x = c; x < le.w; x += NBIN) {
        const size_t i = (size_t)p * le.npix + (size_t)(y0 - 1) * le.w + x;
        row[x] = le.planes_u8 ? (int16_t)reinterpret_cast<const uint8_t *>(le.planes)[i] : reinterpret_cast<const int16_t *>(le.planes)[i];
    }
    if (c >= NBIN - 1) return;
    // counts of context c after every out-of-range pixel before the band = after the element preceding the tile's scatter base
    uint32_t v[NK] = {0, 0, 0, 0, 0, 0};
    const uint32_t cb = le.chain_base[(size_t)p * NBIN + c], cnt = le.chain_count[(size_t)p * NBIN + c];
    const uint32_t g = le.tile_base[((size_t)p * le.tpp + tile) * NBIN + c];   // plane-relative index of the tile's first element of c
    if (cnt != 0 && g > cb) {
        const size_t gi = (size_t)p * le.cap + g - 1;          // the last element before the band
        const size_t blk = gi >> 5;
        uint32_t s[NK] = {0, 0, 0, 0, 0, 0};                   // costs of the block's elements before gi
        for (size_t t = blk << 5; t < gi; t++) {
            const uint32_t e = le.e_grp[t];
            if (e != PAD_E)
#pragma unroll
                for (int k = 0; k < NK; k++) s[k] += (e >> k) + 1u + (uint32_t)k;
        }
        const uint32_t *br = le.blk_rec + blk * 16;
        const uint4 *er = reinterpret_cast<const uint4 *>(le.ep_rec);
        uint32_t ep = le.blk_epoch[blk];
        while (er[(size_t)ep * 2 + 3].z <= (uint32_t)gi) ep++;   // first element of the following epoch (0xFFFFFFFF after the last)
        const uint4 b0 = er[(size_t)ep * 2], b1 = er[(size_t)ep * 2 + 1];
        const uint32_t base[NK] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y};
        const uint32_t e = le.e_grp[gi];
        uint32_t mn = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            v[k] = base[k] + br[k] + s[k] + (e >> k) + 1u + (uint32_t)k;   // parameter_selection.rs:52-55
            mn = min(mn, v[k]);
        }
        if (mn > HALVE_AT)
#pragma unroll
            for (int k = 0; k < NK; k++) v[k] >>= 1;                       // parameter_selection.rs:58-63
    }
#pragma unroll
    for (int k = 0; k < NK; k++) tab[c * NK + k] = v[k];
}

int sidecar_build(felics_ctx *ctx, uint32_t band_rows, uint8_t *h_out, size_t cap, size_t *out_len) {
    const LastEncode &le = ctx->last;
    if (!le.valid) {
        set_error("no single-image 8-bit encode on this context to build a sidecar from");
        return FELICS_ERR_INVALID_ARGUMENT;
    }
    const uint32_t unit = sidecar_row_unit(le.w);
    if (band_rows == 0) {
        const uint32_t want = std::max<uint32_t>(2, (le.h + 63) / 64);   // about 64 bands per plane
        band_rows = (want + unit - 1) / unit * unit;
    }
    if (band_rows % unit != 0 || band_rows < 2) {
        set_error("band_rows must be a multiple of %u for width %u (band starts sit on 4096-pixel tile boundaries)", unit, le.w);
        return FELICS_ERR_INVALID_ARGUMENT;
    }
    const uint32_t nbands = std::max<uint32_t>(1, (le.h + band_rows - 1) / band_rows);
    const size_t entry = sidecar_entry_bytes(le.w);
    const size_t total = SIDECAR_HEADER_BYTES + (size_t)le.nch * nbands * entry;
    if (out_len) *out_len = total;
    if (total > cap) {
        set_error("sidecar capacity %zu too small (need %zu)", cap, total);
        return FELICS_ERR_BUFFER_TOO_SMALL;
    }
    int rc = ensure_buffer(ctx, &ctx->staging_out, &ctx->staging_out_cap, total + 16);
    if (rc) return rc;
    SidecarArgs a;
    a.le = le; a.img_off = nullptr; a.arena_byte0 = 0; a.band_rows = band_rows; a.nbands = nbands; a.entry_bytes = entry;
    a.out = (uint8_t *)ctx->staging_out;
    k_sidecar<<<le.nch * nbands, NBIN, 0, ctx->stream>>>(a);
    uint32_t *hh = reinterpret_cast<uint32_t *>(h_out);
    const uint32_t hdr[8] = {SIDECAR_MAGIC, SIDECAR_VERSION, le.w, le.h, le.nch, band_rows, nbands, (uint32_t)entry};
    memcpy(hh, hdr, sizeof(hdr));
    FELICS_CUDA_TRY(cudaMemcpyAsync(h_out + SIDECAR_HEADER_BYTES, a.out, total - SIDECAR_HEADER_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    FELICS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    FELICS_CUDA_TRY(cudaGetLastError());
    return FELICS_OK;
}
#endif  // FELICS_SIDECAR_BUILDER

}  // namespace felics
