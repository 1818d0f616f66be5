#!/usr/bin/env python
"""bench.py -- FELICS encode/decode throughput of the B200 engine (and of the CPU reference arm).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU implementation (oracle port) on host cores

Workload (BASELINE.json configs[1]): one synthetic 8192x8192 8-bit grayscale image per GPU
(generator "G-nat"; seed 2 + rank).  A step = one encode of that image.  A single
image cannot be split bit-exactly (one estimator and one bit chain per plane), so N GPUs run N
independent replicas: weak scaling, value = all ranks' pixels / max-over-ranks time.
`--workload tiles` switches to BASELINE.json configs[3]-style batches of 512x512 tiles (our own
measurements; the driver's line is the default workload).

value : whole-job encode MPixel/s with the image already resident in HBM (device entry point)
e2e   : the same through the host-memory C ABI call (felics_compress): H2D of the pixels from
        pinned host memory and D2H of the .fel bytes inside the timed region
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W2, H2 = 8192, 8192
TILE_W = TILE_H = 512


def gnat_image(width, height, sigma=3.0, seed=2, phase=0):
    """SURVEY.md 8(d) generator G-nat, produced in row bands to bound host memory."""
    rng = np.random.default_rng(seed)
    out = np.empty((height, width), np.uint8)
    x = (np.arange(width, dtype=np.float64) + phase)[None, :]
    band = 512
    for y0 in range(0, height, band):
        y = np.arange(y0, min(height, y0 + band), dtype=np.float64)[:, None]
        v = 128 + 60 * np.sin(x / 97) * np.cos(y / 131) + 40 * np.sin((x + y) / 37) + rng.normal(0, sigma, (len(y), width))
        out[y0:y0 + len(y)] = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return out


W16, H16 = 4096, 4096


def gnat16_image(width, height, seed=2):
    """G-nat scaled to 16 bits (the 16-bit path): amplitudes x 256, noise sigma 192."""
    rng = np.random.default_rng(seed)
    out = np.empty((height, width), np.uint16)
    x = np.arange(width, dtype=np.float64)[None, :]
    band = 512
    for y0 in range(0, height, band):
        y = np.arange(y0, min(height, y0 + band), dtype=np.float64)[:, None]
        v = 256 * (128 + 60 * np.sin(x / 97) * np.cos(y / 131) + 40 * np.sin((x + y) / 37)) + rng.normal(0, 192, (len(y), width))
        out[y0:y0 + len(y)] = np.clip(np.rint(v), 0, 65535).astype(np.uint16)
    return out


def tile_batch(n, first=0, seed=1):
    """SURVEY.md 8(d) config 4 integer generator (numpy twin): tile t, pixel (x, y)."""
    t = (np.arange(first, first + n, dtype=np.uint64))[:, None, None]
    y = np.arange(TILE_H, dtype=np.uint64)[None, :, None]
    x = np.arange(TILE_W, dtype=np.uint64)[None, None, :]

    def tri(u, p):
        return p - np.abs((u % (2 * p)).astype(np.int64) - p)

    base = 96 + 64 * tri(x + 37 * t, 256) // 256 + 64 * tri(y + 53 * t, 384) // 384
    z = (np.uint64(seed) ^ (t << np.uint64(40)) ^ (y << np.uint64(20)) ^ x) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    bits = (z & np.uint64(0xFFFF)).astype(np.uint32)
    pop = np.zeros(bits.shape, np.int64)
    for i in range(16):
        pop += (bits >> i) & 1
    return np.clip(base + pop - 8, 0, 255).astype(np.uint8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples taken from here on belong to the timed region."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        lines = self.lines[first:] if len(self.lines) > first else self.lines   # a very short timed region: fall back to the warm-up samples
        for line in lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel):
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (C port in oracle/) on host cores
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import felics_oracle as fo
    fo.lib()
    n = args.gpus
    cores = os.cpu_count() or 1
    threads = min(n, cores)
    if args.workload == "tiles":
        per = 64  # bounded sample: 64 tiles per replica
        imgs = [tile_batch(per, first=r * args.tiles) for r in range(n)]
        sample = f"{per} of {args.tiles} 512x512 tiles per replica, {n} replica(s), one image per thread"
        px_per_step = per * TILE_W * TILE_H * n
    elif args.workload == "gray16":
        rows = 2048
        imgs = [gnat16_image(W16, rows, seed=2 + r) for r in range(n)]
        sample = f"top {rows} rows (4096x{rows}) of each replica's 4096x4096 gray16 image, {n} replica(s), one image per thread"
        px_per_step = W16 * rows * n
    elif args.workload == "rgb":
        rows = 1024  # bounded sample: the top 1024 rows of each replica's frame (7.9 MPixel, 23.6 MSample)
        imgs = [gnat_rgb(W3, rows, rank=r) for r in range(n)]
        sample = f"top {rows} rows (7680x{rows}) of each replica's 7680x4320 RGB frame, {n} replica(s), one image per thread"
        px_per_step = W3 * rows * n
    else:
        rows = 2048  # bounded sample: the top 2048 rows of each replica's image (16.8 MPixel)
        imgs = [gnat_image(W2, rows, seed=2 + r) for r in range(n)]
        sample = f"top {rows} rows (8192x{rows}) of each replica's 8192x8192 image, {n} replica(s), one image per thread"
        px_per_step = W2 * rows * n

    def work(i):
        if args.workload == "tiles":
            fo.compress_many(imgs[i], 0, 0, TILE_W, TILE_H)
        else:
            fo.compress(imgs[i])

    def step():
        ts = [threading.Thread(target=work, args=(i,)) for i in range(n)]
        t0 = time.perf_counter()
        # ctypes releases the GIL: replicas run on separate cores (at most `cores` at once)
        for batch in range(0, n, threads):
            for t in ts[batch:batch + threads]:
                t.start()
            for t in ts[batch:batch + threads]:
                t.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = px_per_step * args.steps / total / 1e6
    line = {
        "impl": "reference", "metric": "encode MPixel/s", "value": value, "unit": "MPixel/s", "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16" if args.workload == "gray16" else "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "MPixel/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference's Rust loops (no Rust toolchain in the image); a single image is serial in the reference"},
        "e2e": {"value": value, "unit": "MPixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gnat_rgb(width, height, rank=0):
    """three G-nat planes, phases (0, 11, 23) px, seeds 3, 4, 5 (+3*rank), interleaved R, G, B."""
    return np.stack([gnat_image(width, height, seed=s + 3 * rank, phase=p) for s, p in ((3, 0), (4, 11), (5, 23))], axis=-1)


W3, H3 = 7680, 4320


def workload_name(args):
    if args.workload == "rgb":
        return "configs[2]: one synthetic 7680x4320 8-bit RGB frame (three G-nat planes, YCoCg-R inside the timed path) per GPU, encode"
    if args.workload == "gray16":
        return "SURVEY 8(f)1: one synthetic 4096x4096 gray16 image (G-nat x 256, noise sigma 192, seed 2+rank) per GPU, encode"
    if args.workload == "tiles":
        return f"configs[3]-style: {args.tiles} synthetic 512x512 gray8 tiles per GPU (integer generator, seed 1), encode"
    return "configs[1]: one synthetic 8192x8192 gray8 image (G-nat, seed 2+rank) per GPU, encode"


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import felics_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if args.workload == "tiles":
        n_img = args.tiles
        chunk = 256
        host = np.concatenate([tile_batch(min(chunk, n_img - s), first=rank * n_img + s) for s in range(0, n_img, chunk)])
        hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, TILE_W, TILE_H)
    elif args.workload == "gray16":
        n_img = 1
        host = gnat16_image(W16, H16, seed=2 + rank)[None]
        hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Sixteen, W16, H16)
    elif args.workload == "rgb":
        n_img = 1
        host = gnat_rgb(W3, H3, rank=rank)[None]
        hdr = felics_b200.Header(felics_b200.ColorType.Rgb, felics_b200.PixelDepth.Eight, W3, H3)
    else:
        n_img = 1
        host = gnat_image(W2, H2, seed=2 + rank)[None]
        hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, W2, H2)
    samples = int(host.size)
    in_bytes = int(host.nbytes)                                # S * b of SURVEY.md 8(d)
    pixels = samples // (3 if args.workload == "rgb" else 1)   # an RGB pixel is one pixel, three samples
    pin_in = torch.from_numpy(host.view(np.uint8)).pin_memory()
    d_in = pin_in.to(dev, non_blocking=True)
    cap = in_bytes + in_bytes // 2 + 4096 * n_img
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    pin_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    codec = felics_b200.Codec(device=local)
    stream = torch.cuda.current_stream(dev)
    codec.set_stream(stream.cuda_stream)
    lib = felics_b200.load_library()
    import ctypes as C
    chdr = felics_b200._c_header(hdr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def encode_device():
        return codec.compress_batch_device(n_img, d_in.data_ptr(), hdr, d_out.data_ptr(), cap)

    host_offsets = np.zeros(n_img + 1, dtype=np.uint64)

    def encode_host():
        rc = lib.felics_compress_batch(codec._h, n_img, C.c_void_p(pin_in.data_ptr()), C.byref(chdr), C.c_void_p(pin_out.data_ptr()), cap,
                                       host_offsets.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc:
            raise RuntimeError(f"felics_compress_batch failed: {rc} {lib.felics_last_error().decode()}")
        return host_offsets

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps."""
        ms = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize(dev)
            ms.append(a.elapsed_time(b))
        return ms

    sampler = ClockSampler(local)
    sampler.start()   # nvidia-smi needs a moment to come up: start before the warm-up, count samples from the timed region on
    # warm-up (also sizes the scratch buffers)
    for _ in range(max(args.warmup, 1)):
        offsets = encode_device()
    for _ in range(max(1, min(args.warmup, 2))):
        encode_host()
    fel_bytes = int(offsets[n_img])

    codec.profile(True)
    barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    ms_dev = timed(encode_device, args.steps)
    wall_dev = time.perf_counter() - t_wall0
    barrier()
    stages = codec.stage_times()
    launches = codec.total_launches()
    codec.profile(False)
    barrier()
    ms_e2e = timed(encode_host, args.steps)
    barrier()
    clocks = sampler.stop()

    # decode (reported, not the target): the same .fel decoded on the GPU, timed once per step budget
    d_pix_out = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    dec_steps = 1 if args.workload != "tiles" else min(args.steps, 3)
    codec.profile(True)
    t0 = time.perf_counter()
    lossless = True
    if args.decode:
        for _ in range(dec_steps):
            status = codec.decompress_batch_device(n_img, d_out.data_ptr(), offsets, hdr, d_pix_out.data_ptr())
        torch.cuda.synchronize(dev)
        lossless = bool((not status.any()) and torch.equal(d_pix_out, d_in.view(-1)))
    dec_wall = time.perf_counter() - t0
    dec_stage = codec.stage_times()
    codec.profile(False)

    # a file is a serial bit chain: decode throughput comes from batches.  For the single-image workloads also report the
    # same pixels cut into 512x512 tiles, encoded and decoded as one batch (rank 0 only, informational).
    batch_decode = None
    if args.decode and args.workload == "image" and rank == 0:
        tiles = np.ascontiguousarray(host[0].reshape(H2 // TILE_H, TILE_H, W2 // TILE_W, TILE_W).swapaxes(1, 2).reshape(-1, TILE_H, TILE_W))
        nt = tiles.shape[0]
        thdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, TILE_W, TILE_H)
        d_tiles = torch.from_numpy(tiles).to(dev)
        d_out_t = torch.empty(cap, dtype=torch.uint8, device=dev)   # d_out still holds the timed output (parity is checked below)
        toff = codec.compress_batch_device(nt, d_tiles.data_ptr(), thdr, d_out_t.data_ptr(), cap)
        codec.profile(True)
        tstatus = codec.decompress_batch_device(nt, d_out_t.data_ptr(), toff, thdr, d_pix_out.data_ptr())
        torch.cuda.synchronize(dev)
        tst = codec.stage_times()
        codec.profile(False)
        tms = tst["decode"][0] + tst["unplane"][0]
        batch_decode = {"tiles": int(nt), "tile": f"{TILE_W}x{TILE_H}", "value": tiles.size / (tms * 1e-3) / 1e6, "unit": "MPixel/s", "ms": tms,
                        "lossless": bool((not tstatus.any()) and torch.equal(d_pix_out, d_tiles.view(-1)))}

    # opt-in band sidecar (NOT the reference format; include/felics_b200.h): the same .fel decoded band-parallel
    sidecar_decode = None
    if args.decode and args.workload in ("image", "rgb") and rank == 0:
        encode_device()                                      # the sidecar is built from the context's last encode
        nside = C.c_size_t(0)
        lib.felics_sidecar_build(codec._h, 0, None, 0, C.byref(nside))
        side = np.empty(nside.value, dtype=np.uint8)
        lib.felics_sidecar_build(codec._h, 0, side.ctypes.data, side.size, C.byref(nside))   # sizes the staging buffer
        t0 = time.perf_counter()
        rc = lib.felics_sidecar_build(codec._h, 0, side.ctypes.data, side.size, C.byref(nside))
        build_ms = 1e3 * (time.perf_counter() - t0)
        if rc == 0:
            fel_host = d_out[:fel_bytes].cpu().numpy()
            pix_host = np.empty(in_bytes, dtype=np.uint8)
            ch2 = felics_b200._CHeader()
            codec.profile(True)
            t0 = time.perf_counter()
            rc = lib.felics_decompress_sidecar(codec._h, fel_host.ctypes.data, fel_host.size, side.ctypes.data, side.size, pix_host.ctypes.data,
                                               pix_host.size, C.byref(ch2))
            wall_ms = 1e3 * (time.perf_counter() - t0)
            sst = codec.stage_times()
            codec.profile(False)
            dev_ms = sst["decode"][0] + sst["unplane"][0]
            sidecar_decode = {"value": pixels / (dev_ms * 1e-3) / 1e6, "unit": "MPixel/s", "ms": dev_ms, "wall_ms_host_to_host": wall_ms,
                              "sidecar_bytes": int(side.size), "sidecar_build_ms": build_ms, "bands_per_plane": int(side[24:28].view(np.uint32)[0]),
                              "lossless": bool(rc == 0 and np.array_equal(pix_host, host.view(np.uint8).reshape(-1))),
                              "note": "opt-in side file, not part of the reference .fel format"}

    from felics_b200 import sharding

    def reduce_max(x):
        return sharding.reduce_scalar(x, "max", device=dev)

    def reduce_sum(x):
        return sharding.reduce_scalar(x, "sum", device=dev)

    tot_dev_ms = reduce_max(sum(ms_dev))
    tot_e2e_ms = reduce_max(sum(ms_e2e))
    all_pixels = reduce_sum(float(pixels))
    dec_ms = reduce_max(dec_stage["decode"][0] + dec_stage["unplane"][0]) / dec_steps
    ok_all = reduce_sum(0.0 if lossless else 1.0) == 0.0

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # dominant kernel = the stage with the largest device time
        enc_stages = {k: v for k, v in stages.items() if k not in ("decode", "unplane")}
        dom = max(enc_stages, key=lambda k: enc_stages[k][0])
        dom_ms, dom_launches = enc_stages[dom]
        alg_bytes = in_bytes + fel_bytes        # S*b + C, per launch: one launch covers the whole batch
        dom_avg_ms = dom_ms / max(dom_launches, 1)
        launches_per_step = max(dom_launches / max(args.steps, 1), 1.0)   # batches larger than the scratch budget run in several sub-batches
        alg_bytes_per_launch = alg_bytes / launches_per_step
        achieved = alg_bytes_per_launch / (dom_avg_ms * 1e-3) / 1e9 if dom_avg_ms > 0 else 0.0
        whole_achieved = alg_bytes * args.steps / (sum(ms_dev) * 1e-3) / 1e9

        # CPU baseline: the oracle on a bounded sample of this workload, rank 0 only
        from oracle import felics_oracle as fo
        if args.workload == "tiles":
            sample_imgs = host[:64]
            t0 = time.perf_counter()
            fo.compress_many(sample_imgs, 0, 0, TILE_W, TILE_H)
            cpu_s = time.perf_counter() - t0
            cpu_px = int(sample_imgs.size)
            sample = "first 64 tiles of rank 0's batch, 1 thread"
            want = fo.compress(host[0])
            got = d_out[: int(offsets[1])].cpu().numpy().tobytes()
        elif args.workload == "gray16":
            rows = 2048
            t0 = time.perf_counter()
            fo.compress(host[0][:rows])
            cpu_s = time.perf_counter() - t0
            cpu_px = W16 * rows
            sample = f"top {rows} rows of rank 0's image (4096x{rows} gray16), 1 thread (a single image is serial in the reference)"
            want = None
            got = d_out[:fel_bytes].cpu().numpy().tobytes()
        elif args.workload == "rgb":
            rows = 1024
            t0 = time.perf_counter()
            fo.compress(host[0][:rows])
            cpu_s = time.perf_counter() - t0
            cpu_px = W3 * rows
            sample = f"top {rows} rows of rank 0's frame (7680x{rows} RGB), 1 thread (a single image is serial in the reference)"
            want = None
            got = d_out[:fel_bytes].cpu().numpy().tobytes()
        else:
            rows = 2048
            t0 = time.perf_counter()
            fo.compress(host[0][:rows])
            cpu_s = time.perf_counter() - t0
            cpu_px = W2 * rows
            sample = f"top {rows} rows of rank 0's image (8192x{rows}), 1 thread (a single image is serial in the reference)"
            want = None
            got = d_out[:fel_bytes].cpu().numpy().tobytes()
        parity = None
        if args.verify:
            if want is None:
                want = fo.compress(host[0])
            parity = "bit-exact vs oracle" if hashlib.sha256(got).digest() == hashlib.sha256(want).digest() else "MISMATCH vs oracle"

        value = all_pixels * args.steps / (tot_dev_ms * 1e-3) / 1e6
        e2e = all_pixels * args.steps / (tot_e2e_ms * 1e-3) / 1e6
        line = {
            "metric": "encode MPixel/s", "value": value, "unit": "MPixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16" if args.workload == "gray16" else "u8", "data": "synthetic",
            "config": {"workload": workload_name(args), "l2": "flushed between timed steps (256 MiB device write); each step timed with CUDA events on the launching stream",
                       "fel_bytes_rank0": fel_bytes, "bits_per_sample": 8.0 * fel_bytes / samples,
                       "msample_per_s": value * samples / pixels},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(dom) if args.workload == "image" else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch, "kernel_ms_per_launch": dom_avg_ms, "kernel_launches_per_step": launches_per_step,
                         "whole_encode_achieved_gbs": whole_achieved, "whole_encode_frac": whole_achieved / peak},
            "cpu_baseline": {"value": cpu_px / cpu_s / 1e6, "unit": "MPixel/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": e2e, "unit": "MPixel/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": fel_bytes + 8 * (n_img + 1),
                    "ms_per_step": tot_e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "decode": ({"value": all_pixels / (dec_ms * 1e-3) / 1e6, "unit": "MPixel/s", "ms": dec_ms, "lossless": ok_all,
                        "wall_ms": 1e3 * dec_wall / dec_steps, "as_batch_of_tiles": batch_decode, "with_sidecar": sidecar_decode} if args.decode else None),
            "stages_ms_per_step": {k: v[0] / args.steps for k, v in enc_stages.items()},
            "parity": parity,
            "wall_s_timed_region": wall_dev,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["image", "rgb", "tiles", "gray16"], default="image")
    ap.add_argument("--tiles", type=int, default=2048, help="tiles per GPU for --workload tiles")
    ap.add_argument("--no-verify", dest="verify", action="store_false", help="skip the oracle parity check of the timed output")
    ap.add_argument("--no-decode", dest="decode", action="store_false", help="skip the (slow, single-stream) decode measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
