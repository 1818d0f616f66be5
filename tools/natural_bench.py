#!/usr/bin/env python
"""Band encoder and gray batch decoder on natural content: n copies of boat.512 (image-suite/grayscale/8bit, committed in
tests/golden/images.npz), device resident.  usage: natural_bench.py [n]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import felics_b200 as fb
from oracle import felics_oracle as fo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
img = dict(np.load(Path(__file__).resolve().parents[1] / "tests/golden/images.npz"))["gray8_boat.512"]
dev = torch.device("cuda", 0)
codec = fb.Codec(device=0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
codec.set_stream(stream.cuda_stream)   # one explicit stream for torch and the codec (handle 0 would mean the codec's own)
d_in = torch.from_numpy(img).to(dev).reshape(1, -1).repeat(n, 1).contiguous()
hdr = fb._header_of(img)
cap = d_in.numel() + 4096 * n
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
for it in range(3):
    codec.profile(True)
    off = codec.compress_batch_device(n, d_in.data_ptr(), hdr, d_out.data_ptr(), cap)
    torch.cuda.synchronize(dev)
    st = codec.stage_times()
    codec.profile(False)
ms = sum(v[0] for k, v in st.items() if v[1])
want = fo.compress(img)
ok = d_out[int(off[n - 1]):int(off[n])].cpu().numpy().tobytes() == want
d_pix = torch.empty_like(d_in)
for it in range(2):
    codec.profile(True)
    status = codec.decompress_batch_device(n, d_out.data_ptr(), off, hdr, d_pix.data_ptr())
    torch.cuda.synchronize(dev)
    dst = codec.stage_times()
    codec.profile(False)
dms = dst["decode"][0] + dst["unplane"][0]
print(f"boat.512 x {n}: {8 * len(want) / img.size:.2f} bpp, encode {ms:.2f} ms = {n * img.size / ms / 1e6:.1f} GPixel/s (stages {({k: round(v[0], 2) for k, v in st.items() if v[1]})}, redone {codec.stream_redone()}), bit-exact {ok}; "
      f"decode {dms:.1f} ms = {n * img.size / dms / 1e6:.1f} GPixel/s, lossless {bool((not status.any()) and torch.equal(d_pix, d_in))}")
