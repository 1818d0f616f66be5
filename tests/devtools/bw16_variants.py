"""Walk-stage time of the 16-bit bucket walker under the FELICS_B200_BW16 experiment switches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, felics_b200

N = 4096
rng = np.random.default_rng(1)
yy, xx = np.mgrid[0:N, 0:N]
smooth = 30000 + 9000 * np.sin(xx / 97.0) * np.cos(yy / 131.0)
contents = {
    "sigma300": np.clip(smooth + rng.normal(0, 300, xx.shape), 0, 65535).astype(np.uint16),
    "sigma30": np.clip(smooth / 5 + rng.normal(0, 30, xx.shape), 0, 65535).astype(np.uint16),
    "sigma3": np.clip(smooth / 50 + rng.normal(0, 3, xx.shape), 0, 65535).astype(np.uint16),
    "noise": rng.integers(0, 65536, xx.shape, dtype=np.uint16),
}
dev = torch.device("cuda:0")
hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Sixteen, N, N)
cap = N * N * 5 + 4096
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
d_px = {k: torch.from_numpy(v.view(np.int16)).to(dev) for k, v in contents.items()}
ref = {}
for opts in sys.argv[1:] or ["0"]:
    os.environ["FELICS_B200_BW16"] = opts
    with felics_b200.Codec(0) as c:
        line = []
        for name in contents:
            off = c.compress_batch_device(1, d_px[name].data_ptr(), hdr, d_out.data_ptr(), cap)
            c.profile(True)
            for _ in range(2):
                off = c.compress_batch_device(1, d_px[name].data_ptr(), hdr, d_out.data_ptr(), cap)
            walk = c.stage_times()["walk"][0] / 2
            c.profile(False)
            digest = hash(d_out[: int(off[1])].cpu().numpy().tobytes())
            same = ref.setdefault(name, digest) == digest
            line.append(f"{name} {walk:.2f} ms{'' if same else ' DIFFERENT OUTPUT'}")
        print(f"opts {opts}: " + ", ".join(line), flush=True)
