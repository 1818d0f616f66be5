/*
 * felics_b200_debug.h -- test and bench aids of libfelics_b200.so.  NOT part of the drop-in boundary
 * (include/felics_b200.h): nothing here has a counterpart in visanalexandru/felics.
 */
#ifndef FELICS_B200_DEBUG_H
#define FELICS_B200_DEBUG_H

#include "felics_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-pixel code records (length in bits 31..22) of the last multi-kernel encode on this context, copied to the host. */
int felics_debug_last_records(felics_ctx *ctx, uint32_t *out, size_t count);
/* The 8 device counters of the last multi-kernel encode sub-batch (queue sizes, flags, speculation / hop statistics). */
int felics_debug_counters(felics_ctx *ctx, uint32_t *out8);
/* Images the streaming band encoder handed to the general pipeline since the context was created
 * (their stream was longer than 10 bits per pixel). */
uint64_t felics_debug_stream_redone(felics_ctx *ctx);

/* Synthetic input of BASELINE.json configs[3], generated on the device: tiles
 * first_tile .. first_tile + n_tiles - 1 of 512 x 512 gray8 samples, tile t at d_out + (t - first_tile) * 262144.
 *   base  = 96 + 64 * tri(x + 37 t, 256) / 256 + 64 * tri(y + 53 t, 384) / 384,   tri(u, p) = p - |u mod 2p - p|
 *   noise = popcount(splitmix64(seed ^ t << 40 ^ y << 20 ^ x) & 0xFFFF) - 8
 *   sample = clip(base + noise, 0, 255)
 * Integer arithmetic only; felics_b200/synth.py holds the bit-identical numpy twin. */
int felics_debug_generate_tiles(felics_ctx *ctx, uint8_t *d_out, uint64_t first_tile, uint64_t n_tiles, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* FELICS_B200_DEBUG_H */
