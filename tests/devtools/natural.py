"""GPU timing + parity on a mirror-tiled natural image (tests/golden gray8_5.3.01), 8192 x 8192."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, felics_b200
from oracle import felics_oracle as fo
from test_gpu_parity import mirror_tile
d = dict(np.load(os.path.join(ROOT, "tests", "golden", "images.npz")))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
big = mirror_tile(d["gray8_5.3.01"], n, n)
with felics_b200.Codec(0) as c:
    c.profile(True); c.compress(big); c.profile_reset()
    got = c.compress(big)
    st = c.stage_times()
    print(big.shape, "OK" if got == fo.compress(big) else "MISMATCH", {k: round(v[0], 3) for k, v in st.items() if v[0] > 0}, c.debug_counters())
