"""Timing of the 16-bit path (first correct device path: the reference loops, one warp per image)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, felics_b200
from oracle import felics_oracle as fo
rng = np.random.default_rng(1)
yy, xx = np.mgrid[0:1024, 0:1024]
img = np.clip(30000 + 9000 * np.sin(xx / 97.0) * np.cos(yy / 131.0) + rng.normal(0, 300, xx.shape), 0, 65535).astype(np.uint16)
batch = np.stack([np.roll(img, s, 1) for s in range(64)])
with felics_b200.Codec(0) as c:
    fel = c.compress(img)
    t0 = time.perf_counter(); fel = c.compress(img); t1 = time.perf_counter()
    out = c.decompress(fel); t2 = time.perf_counter()
    t3 = time.perf_counter(); want = fo.compress(img); t4 = time.perf_counter()
    print(f"gray16 1024x1024: encode {img.size / (t1 - t0) / 1e6:.2f} MPixel/s, decode {img.size / (t2 - t1) / 1e6:.2f} MPixel/s, oracle encode {img.size / (t4 - t3) / 1e6:.2f} MPixel/s, "
          f"{'bit-exact' if fel == want else 'MISMATCH'}, lossless {np.array_equal(out, img)}, {8 * len(fel) / img.size:.2f} bpp")
    t0 = time.perf_counter(); arena, off = c.compress_batch(batch); t1 = time.perf_counter()
    print(f"batch of 64: encode {batch.size / (t1 - t0) / 1e6:.1f} MPixel/s")
