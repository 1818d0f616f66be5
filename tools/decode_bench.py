#!/usr/bin/env python
"""Batch decode rate of the configs[3] tiles (device resident).  usage: decode_bench.py [tiles]   (FELICS_B200_G8_FILES / FELICS_B200_NO_G8 select the decoder)"""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import felics_b200 as fb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
codec = fb.Codec(device=0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
codec.set_stream(stream.cuda_stream)   # one explicit stream for torch and the codec (handle 0 would mean the codec's own)
px = 512 * 512
d_in = torch.empty(n * px, dtype=torch.uint8, device=dev)
codec.generate_tiles(d_in.data_ptr(), 0, n)
hdr = fb.Header(fb.ColorType.Gray, fb.PixelDepth.Eight, 512, 512)
cap = n * px * 5 // 8 + 4096 * n
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
offsets = codec.compress_batch_device(n, d_in.data_ptr(), hdr, d_out.data_ptr(), cap)
d_pix = torch.zeros(n * px, dtype=torch.uint8, device=dev)
for it in range(2):
    d_pix.zero_()
    codec.profile(True)
    status = codec.decompress_batch_device(n, d_out.data_ptr(), offsets, hdr, d_pix.data_ptr())
    torch.cuda.synchronize(dev)
    st = codec.stage_times()
    codec.profile(False)
ms = st["decode"][0] + st["unplane"][0]
ok = (not status.any()) and bool(torch.equal(d_pix, d_in))
if not ok:
    bad = (d_pix != d_in).nonzero().flatten()
    print('status any', bool(status.any()), 'mismatches', bad.numel(), 'first', bad[:8].tolist(), 'tiles', sorted(set((bad[:2000] // px).tolist()))[:10])
print(f"G8_FILES={os.environ.get('FELICS_B200_G8_FILES','auto')} NO_G8={os.environ.get('FELICS_B200_NO_G8','0')} tiles {n}: decode {ms:.2f} ms = {n * px / ms / 1e6:.2f} GPixel/s, lossless {ok}")
