import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, felics_b200
from bench import gnat_image
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = gnat_image(n, n)
with felics_b200.Codec(0) as c:
    c.profile(True)
    c.compress(img)
    c.profile_reset()
    c.compress(img)
    print({k: round(v[0], 3) for k, v in c.stage_times().items()})
