"""Synthetic inputs of BASELINE.json's configs: numpy twins of the device generators.

Bench and test infrastructure -- the product path never imports this module."""
from __future__ import annotations

import numpy as np

TILE_W = TILE_H = 512


def tile_batch(n: int, first: int = 0, seed: int = 1) -> np.ndarray:
    """configs[3] integer generator: tiles first .. first+n-1, shape (n, 512, 512) uint8.
    Bit-identical to felics_debug_generate_tiles (felics_b200/csrc/synth.cu)."""
    t = (np.arange(first, first + n, dtype=np.uint64))[:, None, None]
    y = np.arange(TILE_H, dtype=np.uint64)[None, :, None]
    x = np.arange(TILE_W, dtype=np.uint64)[None, None, :]

    def tri(u, p):
        return p - np.abs((u % np.uint64(2 * p)).astype(np.int64) - p)

    base = 96 + 64 * tri(x + np.uint64(37) * t, 256) // 256 + 64 * tri(y + np.uint64(53) * t, 384) // 384
    z = (np.uint64(seed) ^ (t << np.uint64(40)) ^ (y << np.uint64(20)) ^ x) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    bits = (z & np.uint64(0xFFFF)).astype(np.uint32)
    pop = np.zeros(bits.shape, np.int64)
    for i in range(16):
        pop += (bits >> i) & 1
    return np.clip(base + pop - 8, 0, 255).astype(np.uint8)


def mirror_tile(img: np.ndarray, height: int, width: int) -> np.ndarray:
    """configs[4] "synthetic-resized" (SURVEY.md 8(d) item 5): the image repeated by reflection (period 2H x 2W, so no
    seams) and cut to height x width.  Works for HxW and HxWx3."""
    h, w = img.shape[:2]
    yy = np.arange(height) % (2 * h)
    xx = np.arange(width) % (2 * w)
    yy = np.where(yy < h, yy, 2 * h - 1 - yy)
    xx = np.where(xx < w, xx, 2 * w - 1 - xx)
    return np.ascontiguousarray(img[yy][:, xx])
