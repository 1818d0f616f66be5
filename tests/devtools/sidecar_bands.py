"""Decode time of one 8192x8192 gray8 image against the band height of its sidecar (opt-in, not the reference format)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, felics_b200
from conftest import gnat_image
img = gnat_image(8192, 8192)
with felics_b200.Codec(0) as c:
    for rows in (512, 128, 64, 32, 16, 8):
        fel, side = c.compress_with_sidecar(img, rows)
        c.decompress_with_sidecar(fel, side)
        c.profile(True)
        t0 = time.perf_counter(); out = c.decompress_with_sidecar(fel, side); wall = time.perf_counter() - t0
        st = c.stage_times(); c.profile(False)
        ms = st["decode"][0] + st["unplane"][0]
        print(f"band_rows {rows:4d}: {8192 // rows:5d} bands, sidecar {len(side) / 1e6:6.2f} MB ({100 * len(side) / len(fel):5.1f} % of the .fel), "
              f"decode {ms:7.1f} ms = {img.size / ms / 1e3:7.0f} MPixel/s (host to host {wall * 1e3:.0f} ms), lossless {np.array_equal(out, img)}", flush=True)
