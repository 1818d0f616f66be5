"""GPU check of the speculative epoch walk: parity against the oracle, counters, stage times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import felics_b200
from oracle import felics_oracle as fo
from bench import gnat_image

def cases():
    rng = np.random.default_rng(21)
    yield "gnat 2048x1024", gnat_image(2048, 1024)
    yield "gnat 4096x4096", gnat_image(4096, 4096)
    yield "noise x2", rng.integers(0, 256, (600, 700), dtype=np.uint8).repeat(2, axis=1)
    yy, xx = np.mgrid[0:900, 0:1400]
    yield "sawtooth", np.clip(90 + ((xx * 3 + yy) // 11) % 60 + rng.integers(-1, 2, xx.shape), 0, 255).astype(np.uint8)
    yield "near-tied", np.clip(128 + rng.normal(0, 1.2, (1100, 1000)), 0, 255).astype(np.uint8)
    yield "rgb gnat", np.stack([gnat_image(1500, 1000, seed=s, phase=p) for s, p in ((3, 0), (4, 11), (5, 23))], axis=-1)
    if "--big" in sys.argv:
        yield "gnat 8192x8192", gnat_image(8192, 8192)

with felics_b200.Codec(0) as c:
    c.profile(True)
    for name, img in cases():
        t0 = time.time(); want = fo.compress(img); t1 = time.time()
        c.compress(img)
        c.profile_reset()
        got = c.compress(img)
        st = c.stage_times()
        cnt = c.debug_counters()
        print(f"{name}: {'OK' if got == want else 'MISMATCH'} bytes {len(got)} oracle {t1 - t0:.2f}s live {cnt[0]} flags {cnt[2]} sp tried {cnt[3]} resolved {cnt[4]} hops {cnt[5]}/{cnt[5] + cnt[6]} "
              f"spec {st['spec'][0]:.3f} walk {st['walk'][0]:.3f} total {sum(v[0] for v in st.values()):.3f} ms", flush=True)
