import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import felics_b200
from oracle import felics_oracle as fo

rng = np.random.default_rng(21)
_ = rng.integers(0, 256, (600, 700), dtype=np.uint8)
yy, xx = np.mgrid[0:900, 0:1400]
img = np.clip(90 + ((xx * 3 + yy) // 11) % 60 + rng.integers(-1, 2, xx.shape), 0, 255).astype(np.uint8)
with felics_b200.Codec(0) as c:
    got = c.compress(img)
    recs = c.debug_last_records(img.size)
cls, k, ctx, ln, bits = fo.trace_channel(img.astype(np.int32))
glen = (recs >> 22).astype(np.int64)
bad = np.nonzero(glen != ln)[0]
print("bad pixels", len(bad), "ok bytes", got == fo.compress(img))
oor = cls.astype(np.int64) != 0
oor[:2] = False
for cval in np.unique(ctx[bad]):
    sel = np.nonzero(oor & (ctx == cval))[0]          # raster order = chain order
    rank = {int(p): r for r, p in enumerate(sel)}
    b = [rank[int(p)] for p in bad if ctx[p] == cval]
    print(f"ctx {cval}: chain len {len(sel)} bad {len(b)} ranks {b[:12]} ... win {[r // 1024 for r in b[:12]]} inwin {[r % 1024 for r in b[:12]]}")
    # runs of consecutive bad ranks
    runs = []
    for r in b:
        if runs and r <= runs[-1][1] + 8: runs[-1][1] = r
        else: runs.append([r, r])
    print("   runs", [(a, z, a // 1024, a % 1024) for a, z in runs[:10]], "n runs", len(runs))
