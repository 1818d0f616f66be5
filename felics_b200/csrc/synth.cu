// Synthetic tile generator of BASELINE.json configs[3] -- a bench / test aid, declared in
// include/felics_b200_debug.h.  Integer arithmetic only, so the numpy twin (felics_b200/synth.py) is bit-identical.
#include "ctx.h"
#include "../../include/felics_b200_debug.h"

namespace felics {
namespace {

constexpr uint32_t GT = 512;   // tile edge

__device__ __forceinline__ uint32_t tri(uint64_t u, uint32_t p) {
    const int32_t m = (int32_t)(u % (2u * p)) - (int32_t)p;
    return p - (uint32_t)(m < 0 ? -m : m);
}

// one thread per four consecutive samples
__global__ void __launch_bounds__(256) k_generate_tiles(uint8_t *__restrict__ out, uint64_t first_tile, uint64_t n_tiles, uint64_t seed) {
    const uint64_t quads = n_tiles * (GT * GT / 4);
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t tl = q / (GT * GT / 4);
        const uint32_t r = (uint32_t)(q - tl * (GT * GT / 4));
        const uint32_t y = r / (GT / 4), x0 = (r - y * (GT / 4)) * 4u;
        const uint64_t t = first_tile + tl;
        const uint32_t by = 64u * tri((uint64_t)y + 53u * t, 384) / 384u;
        uint32_t word = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t x = x0 + j;
            const uint32_t base = 96u + 64u * tri((uint64_t)x + 37u * t, 256) / 256u + by;
            uint64_t z = (seed ^ (t << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)x) + 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z = z ^ (z >> 31);
            const int v = (int)base + __popc((uint32_t)(z & 0xFFFFu)) - 8;
            word |= (uint32_t)min(max(v, 0), 255) << (8u * j);
        }
        reinterpret_cast<uint32_t *>(out)[q] = word;
    }
}

}  // namespace
}  // namespace felics

using namespace felics;

extern "C" {
#pragma GCC visibility push(default)

int felics_debug_generate_tiles(felics_ctx *ctx, uint8_t *d_out, uint64_t first_tile, uint64_t n_tiles, uint64_t seed) {
    if (!ctx || (!d_out && n_tiles) || ((uintptr_t)d_out & 3)) { set_error("bad argument"); return FELICS_ERR_INVALID_ARGUMENT; }
    int rc = bind_device(ctx);
    if (rc) return rc;
    if (n_tiles == 0) return FELICS_OK;
    k_generate_tiles<<<148 * 16, 256, 0, ctx->stream>>>(d_out, first_tile, n_tiles, seed);
    FELICS_CUDA_TRY(cudaGetLastError());
    return FELICS_OK;
}

uint64_t felics_debug_stream_redone(felics_ctx *ctx) { return ctx ? ctx->stream_redone : 0; }

#pragma GCC visibility pop
}
