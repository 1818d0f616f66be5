"""Timing of the 16-bit path: device-resident encode with the per-stage breakdown, checked against the oracle.

usage: python tests/devtools/time16.py [size]      (default 4096: one size x size gray16 image per content type)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, felics_b200
from oracle import felics_oracle as fo

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(1)
yy, xx = np.mgrid[0:N, 0:N]
smooth = 30000 + 9000 * np.sin(xx / 97.0) * np.cos(yy / 131.0)
contents = {
    "sigma300": np.clip(smooth + rng.normal(0, 300, xx.shape), 0, 65535).astype(np.uint16),
    "sigma3 (8-bit-like)": np.clip(smooth / 50 + rng.normal(0, 3, xx.shape), 0, 65535).astype(np.uint16),
    "uniform noise": rng.integers(0, 65536, xx.shape, dtype=np.uint16),
}
dev = torch.device("cuda:0")
with felics_b200.Codec(0) as c:
    for name, img in contents.items():
        hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Sixteen, N, N)
        d_px = torch.from_numpy(img.view(np.int16)).to(dev)
        cap = img.size * 5 + 4096
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        for _ in range(2):
            off = c.compress_batch_device(1, d_px.data_ptr(), hdr, d_out.data_ptr(), cap)
        c.profile(True); c.profile_reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            off = c.compress_batch_device(1, d_px.data_ptr(), hdr, d_out.data_ptr(), cap)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        stages = {k: round(v[0] / reps, 3) for k, v in c.stage_times().items() if v[1]}
        c.profile(False)
        fel = d_out[: int(off[1])].cpu().numpy().tobytes()
        t2 = time.perf_counter(); want = fo.compress(img); t3 = time.perf_counter()
        print(f"{name}: gray16 {N}x{N}: encode {img.size * reps / (t1 - t0) / 1e6:.1f} MPixel/s (device resident, wall), "
              f"oracle {img.size / (t3 - t2) / 1e6:.1f} MPixel/s, {'bit-exact' if fel == want else 'MISMATCH'}, {8 * len(fel) / img.size:.2f} bpp")
        print("   stage ms:", stages)
    img = contents["sigma300"][:1024, :1024].copy()
    batch = np.stack([np.roll(img, s, 1) for s in range(64)])
    arena, off = c.compress_batch(batch)
    t0 = time.perf_counter(); arena, off = c.compress_batch(batch); t1 = time.perf_counter()
    print(f"batch of 64 1024x1024 from host memory: encode {batch.size / (t1 - t0) / 1e6:.1f} MPixel/s")
    t0 = time.perf_counter(); out = c.decompress(arena[: int(off[1])].tobytes()); t1 = time.perf_counter()
    print(f"decode of one 1024x1024: {img.size / (t1 - t0) / 1e6:.2f} MPixel/s, lossless {np.array_equal(out, batch[0])}")
