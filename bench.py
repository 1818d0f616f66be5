#!/usr/bin/env python
"""bench.py -- FELICS encode/decode throughput of the B200 engine (and of the CPU reference arm).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU implementation (oracle port) on host cores

Default workload = BASELINE.json configs[3]: a batch of 65,536 synthetic 512x512 gray8 tiles (16 GiB; integer generator of
SURVEY.md 8(d) item 4, generated ON THE DEVICE by felics_debug_generate_tiles, numpy twin felics_b200/synth.py), sharded
in contiguous ranges over the ranks (felics_b200.sharding.shard_range): STRONG scaling, the batch is the same at every N and
N = 1 holds all of it.  A step = one encode of the rank's whole shard; value = 65,536 tiles' pixels / max-over-ranks time.
The per-image sizes are gathered over NCCL (sharding.gather_sizes) -- the only exchange -- and >= 256 sampled tiles per rank
are compared with the CPU oracle.  `--workload image|rgb|gray16` select the single-image configs (configs[1], configs[2],
SURVEY 8(f)1: independent replicas per GPU, weak scaling); the default run reports them as extra keys (`other_configs`).
`--workload corpus` is configs[4]: the seven RGB images of the reference's bench/tiff_files (committed as pixels under
tests/golden/), mirror-tiled to 2048^2 and 4096^2 and replicated to 512 images per GPU (4096 at 8 GPUs), two shapes per rank.

value : whole-job encode MPixel/s with the input already resident in HBM (device entry point)
e2e   : the same through the host-memory C ABI call (felics_compress_batch): H2D of the pixels from pinned host memory and
        D2H of the .fel bytes inside the timed region
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

from felics_b200.synth import tile_batch  # noqa: E402  (numpy twin of the device tile generator)

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


def recorded_traffic(kernel, workload, units_per_launch=None):
    """dram bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload (profiles/traffic.json
    names the capture, its command and the git revision it was taken at); None when there is no capture.  Workloads whose
    launches vary in size record bytes per unit (tile) and are scaled to the units one launch of this run covers.
    Returns (bytes, source)."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None, None
    try:
        w = json.loads(p.read_text()).get(workload, {})
        cap = w.get("_capture", {})
        src = f"profiles/traffic.json[{workload}]: {cap.get('files', '?')} at git {cap.get('git', '?')}"
        per = w.get("_per_unit", {})
        if units_per_launch is not None and kernel in per:
            return per[kernel] * units_per_launch, src
        return w.get(kernel), (src if kernel in w else None)
    except Exception:
        return None, None


TOTAL_TILES = 65536
W3, H3 = 7680, 4320


def gnat_rgb(width, height, rank=0):
    """three G-nat planes, phases (0, 11, 23) px, seeds 3, 4, 5 (+3*rank), interleaved R, G, B."""
    return np.stack([gnat_image(width, height, seed=s + 3 * rank, phase=p) for s, p in ((3, 0), (4, 11), (5, 23))], axis=-1)


def workload_name(args):
    if args.workload == "rgb":
        return "configs[2]: one synthetic 7680x4320 8-bit RGB frame (three G-nat planes, YCoCg-R inside the timed path) per GPU, encode"
    if args.workload == "gray16":
        return "SURVEY 8(f)1: one synthetic 4096x4096 gray16 image (G-nat x 256, noise sigma 192, seed 2+rank) per GPU, encode"
    if args.workload == "image":
        return "configs[1]: one synthetic 8192x8192 gray8 image (G-nat, seed 2+rank) per GPU, encode"
    if args.workload == "corpus":
        return (f"configs[4]: the 7 RGB8 images of bench/tiff_files mirror-tiled to 2048x2048 and 4096x4096 (14 sources), replicated to "
                f"{args.corpus_images} images per GPU ({args.corpus_images * args.gpus} in the job; 4096 at 8 GPUs), encode")
    return (f"configs[3]: batch of {args.tiles} synthetic 512x512 gray8 tiles (integer generator, seed 1), contiguous shards over the GPUs, encode")


def cpu_info():
    model = None
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"nproc": os.cpu_count() or 1, "cpu_model": model}


def cpu_tiles_rate(fo, tiles, threads):
    """Oracle port on `threads` host threads, one image per thread at a time (ctypes releases the GIL): MPixel/s."""
    n = len(tiles)
    threads = max(1, min(threads, n))
    bounds = [n * i // threads for i in range(threads + 1)]
    ts = [threading.Thread(target=fo.compress_many, args=(tiles[bounds[i]:bounds[i + 1]], 0, 0, TILE_W, TILE_H)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return tiles.size / (time.perf_counter() - t0) / 1e6


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
    info = cpu_info()
    if args.workload == "tiles":
        # bounded sample of the batch: its first 1024 tiles per step, on all host threads (images are independent units in
        # the reference too: tests/compress.rs:15-35 loops over files)
        ns = min(1024, args.tiles)
        imgs = np.concatenate([tile_batch(min(64, ns - s0), first=s0) for s0 in range(0, ns, 64)])
        threads = min(cores, ns)
        sample = f"first {ns} of the {args.tiles} tiles per step, {threads} threads, one image per thread at a time"
        px_per_step = imgs.size

        def step():
            t0 = time.perf_counter()
            cpu_tiles_rate(fo, imgs, threads)
            return time.perf_counter() - t0
    elif args.workload == "corpus":
        src = corpus_sources()
        imgs = [im for side in CORPUS_SIDES for im in src[side]] * 2
        threads = min(cores, len(imgs))
        sample = f"the 14 source images twice per step, one image per thread at a time on {threads} threads"
        px_per_step = sum(im.shape[0] * im.shape[1] for im in imgs)

        def step():
            t0 = time.perf_counter()
            cpu_images_rate(fo, imgs, threads)
            return time.perf_counter() - t0
    else:
        # single-image configs: the whole image (a single image is serial in the reference), one replica per GPU of our arm
        if args.workload == "gray16":
            imgs = [gnat16_image(W16, H16, seed=2 + r) for r in range(n)]
            px_per_step = W16 * H16 * n
        elif args.workload == "rgb":
            imgs = [gnat_rgb(W3, H3, rank=r) for r in range(n)]
            px_per_step = W3 * H3 * n
        else:
            imgs = [gnat_image(W2, H2, seed=2 + r) for r in range(n)]
            px_per_step = W2 * H2 * n
        threads = min(n, cores)
        sample = f"the whole image of each of the {n} replica(s), one image per thread ({threads} threads)"

        def step():
            ts = [threading.Thread(target=fo.compress, args=(imgs[i],)) for i in range(n)]
            t0 = time.perf_counter()
            for b0 in range(0, n, threads):
                for t in ts[b0:b0 + threads]:
                    t.start()
                for t in ts[b0:b0 + threads]:
                    t.join()
            return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = px_per_step * args.steps / total / 1e6
    line = {
        "impl": "reference", "metric": "encode MPixel/s", "value": value, "unit": "MPixel/s", "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "tiles" else "weak",   # corpus and single images: per-GPU work is fixed
        "vs_baseline": None, "dtype": "u16" if args.workload == "gray16" else "u8", "data": "synthetic",
        "config": {"workload": workload_name(args)},
        "cpu_baseline": {"value": value, "unit": "MPixel/s", "cores": threads, "kind": "port", "sample": sample, **info,
                         "note": "C restatement of the reference's Rust loops (no Rust toolchain in the image)"},
        "e2e": {"value": value, "unit": "MPixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class Rig:
    """Device, codec and timing helpers shared by the workloads."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import felics_b200
        self.torch, self.dist, self.fb = torch, dist, felics_b200
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.codec = felics_b200.Codec(device=self.local)
        # one explicit stream for torch and for the codec: torch's default stream has handle 0, which the library reads as "use
        # your own (non-blocking) stream" -- the timing events below must sit on the stream the kernels are launched on
        self.stream = torch.cuda.Stream(self.dev)
        torch.cuda.set_stream(self.stream)
        self.codec.set_stream(self.stream.cuda_stream)
        assert self.stream.cuda_stream != 0
        self.lib = felics_b200.load_library()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps, flush=None):
        """K steps, each bracketed by CUDA events on the launching stream (optionally an L2 flush before each)."""
        torch = self.torch
        ms = []
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(self.dev)
            a.record(self.stream)
            fn()
            b.record(self.stream)
            torch.cuda.synchronize(self.dev)
            ms.append(a.elapsed_time(b))
        return ms

    def reduce(self, x, op):
        from felics_b200 import sharding
        return sharding.reduce_scalar(x, op, device=self.dev)

    def pinned(self, nbytes):
        return self.torch.empty(nbytes, dtype=self.torch.uint8).pin_memory()

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def encode_host_call(rig, n_img, pin_in, chdr, pin_out, cap, host_offsets):
    import ctypes as C
    rc = rig.lib.felics_compress_batch(rig.codec._h, n_img, C.c_void_p(pin_in.data_ptr()), C.byref(chdr), C.c_void_p(pin_out.data_ptr()), cap,
                                       host_offsets.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc:
        raise RuntimeError(f"felics_compress_batch failed: {rc} {rig.lib.felics_last_error().decode()}")


def roofline_of(stages, steps, alg_bytes_per_step, total_ms, workload, units_per_step=None):
    peak, peak_src = measured_peak_gbs()
    dom = max(stages, key=lambda k: stages[k][0])
    dom_ms, dom_launches = stages[dom]
    launches_per_step = max(dom_launches / max(steps, 1), 1.0)   # batches larger than the scratch budget run in several sub-batches
    dom_avg_ms = dom_ms / max(dom_launches, 1)
    per_launch = alg_bytes_per_step / launches_per_step
    achieved = per_launch / (dom_avg_ms * 1e-3) / 1e9 if dom_avg_ms > 0 else 0.0
    whole = alg_bytes_per_step * steps / (total_ms * 1e-3) / 1e9
    traffic, traffic_src = recorded_traffic(dom, workload, None if units_per_step is None else units_per_step / launches_per_step)
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch,
            "kernel_ms_per_launch": dom_avg_ms, "kernel_launches_per_step": launches_per_step, "whole_encode_achieved_gbs": whole,
            "whole_encode_frac": whole / peak}


# ---- configs[3]: the sharded tile batch ---------------------------------------------------------
def run_tiles(args, rig):
    import ctypes as C
    torch, fb, codec = rig.torch, rig.fb, rig.codec
    from felics_b200 import sharding
    from oracle import felics_oracle as fo

    total = args.tiles
    first, count = sharding.shard_range(total, rig.rank, rig.world)
    tile_px = TILE_W * TILE_H
    in_bytes = count * tile_px
    hdr = fb.Header(fb.ColorType.Gray, fb.PixelDepth.Eight, TILE_W, TILE_H)
    chdr = fb._c_header(hdr)
    d_in = torch.empty(max(in_bytes, 16), dtype=torch.uint8, device=rig.dev)
    codec.generate_tiles(d_in.data_ptr(), first, count)                     # on the device; numpy twin checked below
    torch.cuda.synchronize(rig.dev)
    for t in sorted({0, count // 2, count - 1}) if count else []:
        assert np.array_equal(d_in[t * tile_px:(t + 1) * tile_px].cpu().numpy().reshape(TILE_H, TILE_W), tile_batch(1, first=first + t)[0]), \
            "device generator and its numpy twin disagree"
    cap = in_bytes * 5 // 8 + 4096 * max(count, 1)
    d_out = torch.empty(cap, dtype=torch.uint8, device=rig.dev)

    def encode_device():
        return codec.compress_batch_device(count, d_in.data_ptr(), hdr, d_out.data_ptr(), cap)

    sampler = ClockSampler(rig.local)
    sampler.start()
    for _ in range(max(args.warmup, 1)):
        offsets = encode_device()
    fel_bytes = int(offsets[count])

    codec.profile(True)
    rig.barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    ms_dev = rig.timed(encode_device, args.steps)          # the shard (>= 2 GiB at N <= 8) and its output exceed the 126 MB L2
    wall_dev = time.perf_counter() - t_wall0
    rig.barrier()
    stages = codec.stage_times()
    launches = codec.total_launches()
    codec.profile(False)

    # the only exchange of the job: per-image sizes, gathered over the process group (NCCL under torchrun)
    all_sizes = sharding.gather_sizes(np.diff(offsets.astype(np.int64)))
    global_off = sharding.global_offsets(all_sizes)
    assert len(all_sizes) == total and int(global_off[first]) + fel_bytes == int(global_off[first + count])

    # parity: >= 256 sampled tiles of this rank's shard against the CPU oracle (bit-exact .fel bytes)
    mismatches = 0
    checked = 0
    if args.verify and count:
        idx = np.unique(np.linspace(0, count - 1, min(count, max(args.parity_tiles, 1))).astype(np.int64))
        for i in idx:
            got = d_out[int(offsets[i]):int(offsets[i + 1])].cpu().numpy().tobytes()
            want = fo.compress(d_in[i * tile_px:(i + 1) * tile_px].cpu().numpy().reshape(TILE_H, TILE_W))
            mismatches += got != want
            checked += 1

    # decode of the same shard as one batch (reported, not the target)
    decode = None
    if args.decode and count:
        d_pix = torch.empty(in_bytes, dtype=torch.uint8, device=rig.dev)
        for _ in range(2):   # one warm-up call (scratch allocation, kernel attributes), one timed call
            d_pix.zero_()
            codec.profile(True)
            status = codec.decompress_batch_device(count, d_out.data_ptr(), offsets, hdr, d_pix.data_ptr())
            torch.cuda.synchronize(rig.dev)
            dst = codec.stage_times()
            codec.profile(False)
        dec_ms = dst["decode"][0] + dst["unplane"][0]
        lossless = bool((not status.any()) and torch.equal(d_pix, d_in[:in_bytes]))
        del d_pix
        decode = (dec_ms, lossless)

    # end to end through the host-memory C ABI call: pinned host pixels in, pinned host arena out.  The whole shard if the
    # host can pin it, otherwise the largest power-of-two part that it can.
    e2e_tiles = count
    pin_in = pin_out = None
    while e2e_tiles:
        try:
            pin_in = rig.pinned(e2e_tiles * tile_px)
            pin_out = rig.pinned(e2e_tiles * tile_px * 5 // 8 + 4096 * e2e_tiles)
            break
        except RuntimeError:
            pin_in = pin_out = None
            e2e_tiles //= 2
    ms_e2e, e2e_fel = [0.0], 0
    if e2e_tiles:
        pin_in.copy_(d_in[:e2e_tiles * tile_px])
        host_offsets = np.zeros(e2e_tiles + 1, dtype=np.uint64)
        e2e_steps = args.steps if e2e_tiles * tile_px <= (4 << 30) else min(args.steps, 3)
        for _ in range(max(1, min(args.warmup, 3))):
            encode_host_call(rig, e2e_tiles, pin_in, chdr, pin_out, pin_out.numel(), host_offsets)
        rig.barrier()
        ms_e2e = rig.timed(lambda: encode_host_call(rig, e2e_tiles, pin_in, chdr, pin_out, pin_out.numel(), host_offsets), e2e_steps)
        rig.barrier()
        e2e_fel = int(host_offsets[e2e_tiles])
        if args.verify:
            mismatches += not np.array_equal(pin_out[:int(host_offsets[1])].numpy(), d_out[:int(offsets[1])].cpu().numpy())
            mismatches += int(host_offsets[e2e_tiles]) != int(offsets[e2e_tiles])
    clocks = sampler.stop()

    tot_dev_ms = rig.reduce(sum(ms_dev), "max")
    e2e_rate_local = e2e_tiles * tile_px * len(ms_e2e) / (sum(ms_e2e) * 1e-3) / 1e6 if e2e_tiles and sum(ms_e2e) > 0 else 0.0
    # every rank pushes its part through the same host at the same time: the job's end-to-end rate is what all ranks move
    # over the slowest rank's time, normalised to the tiles the ranks actually pushed
    e2e_px = rig.reduce(float(e2e_tiles * tile_px), "sum")
    e2e_ms = rig.reduce(sum(ms_e2e) / max(len(ms_e2e), 1), "max")
    all_pixels = rig.reduce(float(count * tile_px), "sum")
    all_fel = rig.reduce(float(fel_bytes), "sum")
    bad = rig.reduce(float(mismatches), "sum")
    n_checked = rig.reduce(float(checked), "sum")
    dec_ms = rig.reduce(decode[0], "max") if decode else None
    dec_bad = rig.reduce(0.0 if (decode is None or decode[1]) else 1.0, "sum")

    line = None
    if rig.rank == 0:
        enc_stages = {k: v for k, v in stages.items() if k not in ("decode", "unplane") and v[1]}
        roof = roofline_of(enc_stages, args.steps, in_bytes + fel_bytes, sum(ms_dev), "tiles", units_per_step=count)
        # CPU baseline on this box's host cores: the oracle port on a bounded sample of the same tiles
        info = cpu_info()
        n1 = min(count, 256)
        t1 = d_in[:n1 * tile_px].cpu().numpy().reshape(n1, TILE_H, TILE_W)
        one = cpu_tiles_rate(fo, t1, 1)
        nall = min(count, 128 * info["nproc"], 4096)
        tall = d_in[:nall * tile_px].cpu().numpy().reshape(nall, TILE_H, TILE_W)
        allc = cpu_tiles_rate(fo, tall, info["nproc"])
        value = all_pixels * args.steps / (tot_dev_ms * 1e-3) / 1e6
        e2e = e2e_px / (e2e_ms * 1e-3) / 1e6 if e2e_ms else 0.0
        line = {
            "metric": "encode MPixel/s", "value": value, "unit": "MPixel/s", "n_gpus": rig.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args), "tiles_total": total, "tiles_rank0": count, "input_bytes_total": total * tile_px,
                       "generator": "on device (felics_debug_generate_tiles), numpy twin felics_b200/synth.py checked on 3 tiles per rank",
                       "l2": "not flushed: every rank's input (>= 2 GiB) and output exceed the 126 MB L2; each step timed with CUDA events on the launching stream",
                       "fel_bytes_total": int(all_fel), "bits_per_sample": 8.0 * all_fel / max(all_pixels, 1),
                       "sizes_gathered": int(len(all_sizes)), "gather": "sharding.gather_sizes over " + ("NCCL" if rig.world > 1 else "one process")},
            "roofline": roof,
            "cpu_baseline": {"value": allc, "unit": "MPixel/s", "cores": min(info["nproc"], nall), "kind": "port",
                             "sample": f"first {nall} tiles of rank 0's shard, one image per thread at a time on all host threads",
                             "single_thread": {"value": one, "unit": "MPixel/s", "cores": 1, "sample": f"first {n1} tiles of rank 0's shard"}, **info},
            "e2e": {"value": e2e, "unit": "MPixel/s", "h2d_bytes_per_step": e2e_tiles * tile_px, "d2h_bytes_per_step": e2e_fel + 8 * (e2e_tiles + 1),
                    "ms_per_step": e2e_ms, "tiles_per_rank": e2e_tiles, "rank0_mpixel_s": e2e_rate_local,
                    "note": "felics_compress_batch on pinned host buffers; sub-batches copy in / encode / copy out on three streams"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "decode": ({"value": all_pixels / (dec_ms * 1e-3) / 1e6, "unit": "MPixel/s", "ms": dec_ms, "lossless": dec_bad == 0.0,
                        "files": total, "note": "whole shard as one batch (k_decode_g8: up to 32 files per warp, one per lane)"} if decode else None),
            "stages_ms_per_step": {k: v[0] / args.steps for k, v in enc_stages.items()},
            "parity": (None if not args.verify else (f"bit-exact vs oracle on {int(n_checked)} sampled tiles" if bad == 0 else f"MISMATCH vs oracle ({int(bad)} of {int(n_checked)})")),
            "stream_redone": codec.stream_redone(),
            "wall_s_timed_region": wall_dev,
        }
    del d_in, d_out, pin_in, pin_out
    torch.cuda.empty_cache()
    return line


# ---- single-image configs -------------------------------------------------------------------------
def single_input(workload, rank, fb):
    if workload == "gray16":
        return gnat16_image(W16, H16, seed=2 + rank)[None], fb.Header(fb.ColorType.Gray, fb.PixelDepth.Sixteen, W16, H16)
    if workload == "rgb":
        return gnat_rgb(W3, H3, rank=rank)[None], fb.Header(fb.ColorType.Rgb, fb.PixelDepth.Eight, W3, H3)
    return gnat_image(W2, H2, seed=2 + rank)[None], fb.Header(fb.ColorType.Gray, fb.PixelDepth.Eight, W2, H2)


def quick_single(workload, rig, steps=3):
    """One of the single-image configs as an extra key of the default line: device-timed encode, oracle parity."""
    torch, fb, codec = rig.torch, rig.fb, rig.codec
    from oracle import felics_oracle as fo
    host, hdr = single_input(workload, 0, fb)
    pixels = host.size // (3 if workload == "rgb" else 1)
    d_in = torch.from_numpy(host.view(np.uint8)).to(rig.dev)
    cap = host.nbytes + host.nbytes // 2 + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device=rig.dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=rig.dev)
    fn = lambda: codec.compress_batch_device(1, d_in.data_ptr(), hdr, d_out.data_ptr(), cap)  # noqa: E731
    for _ in range(3):
        offsets = fn()
    ms = rig.timed(fn, steps, flush)
    fel = int(offsets[1])
    ok = hashlib.sha256(d_out[:fel].cpu().numpy().tobytes()).digest() == hashlib.sha256(fo.compress(host[0])).digest()
    return {"ms_per_step": sum(ms) / len(ms), "mpixel_s": pixels * len(ms) / (sum(ms) * 1e-3) / 1e6, "fel_bytes": fel,
            "parity": "bit-exact vs oracle" if ok else "MISMATCH vs oracle", "l2": "flushed between timed steps"}


def run_single(args, rig):
    import ctypes as C
    torch, fb, codec = rig.torch, rig.fb, rig.codec
    rank, dev = rig.rank, rig.dev
    host, hdr = single_input(args.workload, rank, fb)
    n_img = 1
    samples = int(host.size)
    in_bytes = int(host.nbytes)                                # S * b of SURVEY.md 8(d)
    pixels = samples // (3 if args.workload == "rgb" else 1)   # an RGB pixel is one pixel, three samples
    pin_in = torch.from_numpy(host.view(np.uint8)).pin_memory()
    d_in = pin_in.to(dev, non_blocking=True)
    cap = in_bytes + in_bytes // 2 + 4096 * n_img
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    pin_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    lib = rig.lib
    chdr = fb._c_header(hdr)

    def encode_device():
        return codec.compress_batch_device(n_img, d_in.data_ptr(), hdr, d_out.data_ptr(), cap)

    host_offsets = np.zeros(n_img + 1, dtype=np.uint64)

    def encode_host():
        encode_host_call(rig, n_img, pin_in, chdr, pin_out, cap, host_offsets)

    sampler = ClockSampler(rig.local)
    sampler.start()   # nvidia-smi needs a moment to come up: start before the warm-up, count samples from the timed region on
    for _ in range(max(args.warmup, 1)):
        offsets = encode_device()
    for _ in range(max(1, min(args.warmup, 2))):
        encode_host()
    fel_bytes = int(offsets[n_img])

    codec.profile(True)
    rig.barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    ms_dev = rig.timed(encode_device, args.steps, flush)
    wall_dev = time.perf_counter() - t_wall0
    rig.barrier()
    stages = codec.stage_times()
    launches = codec.total_launches()
    codec.profile(False)
    rig.barrier()
    ms_e2e = rig.timed(encode_host, args.steps, flush)
    rig.barrier()
    clocks = sampler.stop()

    # decode (reported, not the target): the same .fel decoded on the GPU, once
    d_pix_out = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    codec.profile(True)
    t0 = time.perf_counter()
    lossless = True
    if args.decode:
        status = codec.decompress_batch_device(n_img, d_out.data_ptr(), offsets, hdr, d_pix_out.data_ptr())
        torch.cuda.synchronize(dev)
        lossless = bool((not status.any()) and torch.equal(d_pix_out, d_in.view(-1)))
    dec_wall = time.perf_counter() - t0
    dec_stage = codec.stage_times()
    codec.profile(False)

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
            ch2 = fb._CHeader()
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

    tot_dev_ms = rig.reduce(sum(ms_dev), "max")
    tot_e2e_ms = rig.reduce(sum(ms_e2e), "max")
    all_pixels = rig.reduce(float(pixels), "sum")
    dec_ms = rig.reduce(dec_stage["decode"][0] + dec_stage["unplane"][0], "max")
    ok_all = rig.reduce(0.0 if lossless else 1.0, "sum") == 0.0

    if rank != 0:
        return None
    from oracle import felics_oracle as fo
    enc_stages = {k: v for k, v in stages.items() if k not in ("decode", "unplane") and v[1]}
    roof = roofline_of(enc_stages, args.steps, in_bytes + fel_bytes, sum(ms_dev), args.workload)
    # CPU baseline: the oracle on the whole image, one thread (a single image is serial in the reference)
    t0 = time.perf_counter()
    want = fo.compress(host[0])
    cpu_s = time.perf_counter() - t0
    parity = None
    if args.verify:
        got = d_out[:fel_bytes].cpu().numpy().tobytes()
        parity = "bit-exact vs oracle" if hashlib.sha256(got).digest() == hashlib.sha256(want).digest() else "MISMATCH vs oracle"
    value = all_pixels * args.steps / (tot_dev_ms * 1e-3) / 1e6
    e2e = all_pixels * args.steps / (tot_e2e_ms * 1e-3) / 1e6
    return {
        "metric": "encode MPixel/s", "value": value, "unit": "MPixel/s", "n_gpus": rig.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16" if args.workload == "gray16" else "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "l2": "flushed between timed steps (256 MiB device write); each step timed with CUDA events on the launching stream",
                   "fel_bytes_rank0": fel_bytes, "bits_per_sample": 8.0 * fel_bytes / samples, "msample_per_s": value * samples / pixels},
        "roofline": roof,
        "cpu_baseline": {"value": pixels / cpu_s / 1e6, "unit": "MPixel/s", "cores": 1, "kind": "port",
                         "sample": "the whole image of rank 0, one thread (a single image is serial in the reference)", **cpu_info()},
        "e2e": {"value": e2e, "unit": "MPixel/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": fel_bytes + 8 * (n_img + 1),
                "ms_per_step": tot_e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "decode": ({"value": all_pixels / (dec_ms * 1e-3) / 1e6, "unit": "MPixel/s", "ms": dec_ms, "lossless": ok_all,
                    "wall_ms": 1e3 * dec_wall, "with_sidecar": sidecar_decode} if args.decode else None),
        "stages_ms_per_step": {k: v[0] / args.steps for k, v in enc_stages.items()},
        "parity": parity,
        "wall_s_timed_region": wall_dev,
    }

# ---- configs[4]: the bench corpus, mirror-tiled and replicated ---------------------------------------
CORPUS_SIDES = (2048, 4096)


def corpus_sources():
    """The 14 source images of configs[4]: tests/golden/bench_corpus.npz (the seven RGB files of the reference's
    bench/tiff_files, committed as pixels by tests/golden/make_corpus.py) mirror-tiled to 2048^2 and 4096^2."""
    from felics_b200.synth import mirror_tile
    raw = dict(np.load(ROOT / "tests" / "golden" / "bench_corpus.npz"))
    return {side: [mirror_tile(raw[n], side, side) for n in sorted(raw)] for side in CORPUS_SIDES}


def corpus_plan(first, count):
    """Images first .. first+count-1 of the job: image g is source g % 14 (g % 14 < 7: the 2048^2 version of corpus image
    g % 7, otherwise its 4096^2 version).  Returns {side: [source index of every image of that side, in job order]}."""
    plan = {side: [] for side in CORPUS_SIDES}
    for g in range(first, first + count):
        s = g % 14
        plan[CORPUS_SIDES[s // 7]].append(s % 7)
    return plan


def cpu_images_rate(fo, images, threads):
    """Oracle port, one image per thread at a time on `threads` host threads: MPixel/s over `images` (HxWx3)."""
    threads = max(1, min(threads, len(images)))
    todo = list(range(len(images)))
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                if not todo:
                    return
                i = todo.pop()
            fo.compress(images[i])
    ts = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return sum(im.shape[0] * im.shape[1] for im in images) / (time.perf_counter() - t0) / 1e6


def run_corpus(args, rig):
    import ctypes as C
    torch, fb, codec = rig.torch, rig.fb, rig.codec
    from felics_b200 import sharding
    from oracle import felics_oracle as fo

    per_rank = args.corpus_images
    total = per_rank * rig.world
    first, count = sharding.shard_range(total, rig.rank, rig.world)
    plan = corpus_plan(first, count)
    src = corpus_sources()
    groups = []   # one per shape: device batch, header, output arena
    for side in CORPUS_SIDES:
        idx = plan[side]
        if not idx:
            continue
        d_src = torch.from_numpy(np.stack(src[side])).to(rig.dev)                       # (7, side, side, 3)
        d_in = d_src[torch.tensor(idx, device=rig.dev)].contiguous()                    # replicated on the device
        del d_src
        hdr = fb.Header(fb.ColorType.Rgb, fb.PixelDepth.Eight, side, side)
        cap = d_in.numel() * 3 // 4 + 4096 * len(idx)
        groups.append({"side": side, "idx": idx, "d_in": d_in, "hdr": hdr, "cap": cap, "d_out": torch.empty(cap, dtype=torch.uint8, device=rig.dev)})
    in_bytes = sum(g["d_in"].numel() for g in groups)
    pixels = in_bytes // 3

    def encode_device():
        for g in groups:
            g["offsets"] = codec.compress_batch_device(len(g["idx"]), g["d_in"].data_ptr(), g["hdr"], g["d_out"].data_ptr(), g["cap"])

    sampler = ClockSampler(rig.local)
    sampler.start()
    for _ in range(max(args.warmup, 1)):
        encode_device()
    fel_bytes = sum(int(g["offsets"][-1]) for g in groups)
    codec.profile(True)
    rig.barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    ms_dev = rig.timed(encode_device, args.steps)      # every rank's input (>= 1 GiB) and output exceed the L2
    wall_dev = time.perf_counter() - t_wall0
    rig.barrier()
    stages = codec.stage_times()
    launches = codec.total_launches()
    codec.profile(False)

    sizes = np.concatenate([np.diff(g["offsets"].astype(np.int64)) for g in groups]) if groups else np.zeros(0, np.int64)
    all_sizes = sharding.gather_sizes(sizes)
    assert len(all_sizes) == total

    # parity: the first replica of every distinct source on this rank against the oracle, bit-exact .fel bytes
    mismatches = checked = 0
    if args.verify:
        for g in groups:
            seen = set()
            for j, s in enumerate(g["idx"]):
                if s in seen:
                    continue
                seen.add(s)
                got = g["d_out"][int(g["offsets"][j]):int(g["offsets"][j + 1])].cpu().numpy().tobytes()
                mismatches += got != fo.compress(src[g["side"]][s])
                checked += 1

    decode = None
    if args.decode:
        dec_ms, lossless = 0.0, True
        for g in groups:
            d_pix = torch.empty_like(g["d_in"])
            codec.profile(True)
            status = codec.decompress_batch_device(len(g["idx"]), g["d_out"].data_ptr(), g["offsets"], g["hdr"], d_pix.data_ptr())
            torch.cuda.synchronize(rig.dev)
            dst = codec.stage_times()
            codec.profile(False)
            dec_ms += dst["decode"][0] + dst["unplane"][0]
            lossless = lossless and (not status.any()) and bool(torch.equal(d_pix, g["d_in"]))
            del d_pix
        decode = (dec_ms, lossless)

    # end to end through the mixed-shape host call (felics_compress_batch_v): the 14 sources twice, host pixels in, host arena out
    e2e_imgs = [im for side in CORPUS_SIDES for im in src[side]] * 2
    e2e_in = sum(im.nbytes for im in e2e_imgs)
    hdrs = (fb._CHeader * len(e2e_imgs))(*[fb._c_header(fb._header_of(im)) for im in e2e_imgs])
    pins = [torch.from_numpy(im).pin_memory() for im in e2e_imgs]
    ptrs = (C.c_void_p * len(e2e_imgs))(*[p.data_ptr() for p in pins])
    pin_out = rig.pinned(e2e_in * 3 // 4)
    e2e_off = np.zeros(len(e2e_imgs) + 1, dtype=np.uint64)

    def encode_host():
        rc = rig.lib.felics_compress_batch_v(codec._h, len(e2e_imgs), ptrs, hdrs, C.c_void_p(pin_out.data_ptr()), pin_out.numel(),
                                             e2e_off.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc:
            raise RuntimeError(f"felics_compress_batch_v failed: {rc} {rig.lib.felics_last_error().decode()}")
    encode_host()
    rig.barrier()
    e2e_steps = min(args.steps, 3)
    ms_e2e = rig.timed(encode_host, e2e_steps)
    rig.barrier()
    clocks = sampler.stop()

    tot_dev_ms = rig.reduce(sum(ms_dev), "max")
    e2e_ms = rig.reduce(sum(ms_e2e) / len(ms_e2e), "max")
    all_pixels = rig.reduce(float(pixels), "sum")
    all_fel = rig.reduce(float(fel_bytes), "sum")
    bad = rig.reduce(float(mismatches), "sum")
    n_checked = rig.reduce(float(checked), "sum")
    dec_ms = rig.reduce(decode[0], "max") if decode else None
    dec_bad = rig.reduce(0.0 if (decode is None or decode[1]) else 1.0, "sum")
    if rig.rank != 0:
        return None
    enc_stages = {k: v for k, v in stages.items() if k not in ("decode", "unplane") and v[1]}
    roof = roofline_of(enc_stages, args.steps, in_bytes + fel_bytes, sum(ms_dev), "corpus")
    info = cpu_info()
    flat = [im for side in CORPUS_SIDES for im in src[side]]
    one = cpu_images_rate(fo, flat[:3], 1)
    allc = cpu_images_rate(fo, flat * 2, info["nproc"])
    value = all_pixels * args.steps / (tot_dev_ms * 1e-3) / 1e6
    e2e_px = sum(im.shape[0] * im.shape[1] for im in e2e_imgs) * rig.world
    return {
        "metric": "encode MPixel/s", "value": value, "unit": "MPixel/s", "n_gpus": rig.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "bench/tiff_files pixels (committed), mirror-tiled and replicated",
        "config": {"workload": workload_name(args), "images_total": total, "images_rank0": count, "input_bytes_rank0": in_bytes,
                   "msample_per_s": 3 * value, "fel_bytes_total": int(all_fel), "bits_per_sample": 8.0 * all_fel / max(3 * all_pixels, 1),
                   "sizes_gathered": int(len(all_sizes)), "l2": "not flushed: every rank's input and output exceed the 126 MB L2"},
        "roofline": roof,
        "cpu_baseline": {"value": allc, "unit": "MPixel/s", "cores": min(info["nproc"], 28), "kind": "port",
                         "sample": "the 14 source images twice, one image per thread at a time on all host threads",
                         "single_thread": {"value": one, "unit": "MPixel/s", "cores": 1, "sample": "the first three 2048x2048 sources"}, **info},
        "e2e": {"value": e2e_px / (e2e_ms * 1e-3) / 1e6, "unit": "MPixel/s", "h2d_bytes_per_step": e2e_in, "d2h_bytes_per_step": int(e2e_off[-1]),
                "ms_per_step": e2e_ms, "images_per_rank": len(e2e_imgs),
                "note": "felics_compress_batch_v (mixed shapes) on pinned host images: the 14 sources twice per rank"},
        "gpu_launches": int(launches), "clocks": clocks,
        "decode": ({"value": all_pixels / (dec_ms * 1e-3) / 1e6, "unit": "MPixel/s", "ms": dec_ms, "lossless": dec_bad == 0.0, "files": total} if decode else None),
        "stages_ms_per_step": {k: v[0] / args.steps for k, v in enc_stages.items()},
        "parity": (None if not args.verify else (f"bit-exact vs oracle on {int(n_checked)} images (every distinct source on every rank)" if bad == 0
                                                 else f"MISMATCH vs oracle ({int(bad)} of {int(n_checked)})")),
        "wall_s_timed_region": wall_dev,
    }


def run_ours(args):
    rig = Rig()
    if args.workload == "tiles":
        line = run_tiles(args, rig)
        if line is not None and args.extras and rig.world == 1:
            # the single-image configs as extra keys (one B200, informational; `--workload image|rgb|gray16` gives their full lines)
            extra = {}
            for wl, key in (("image", "configs[1] 8192x8192 gray8"), ("rgb", "configs[2] 7680x4320 RGB8"), ("gray16", "4096x4096 gray16")):
                try:
                    extra[key] = quick_single(wl, rig)
                except Exception as exc:   # informational only: never lose the main line
                    extra[key] = {"error": str(exc)}
            line["other_configs"] = extra
    elif args.workload == "corpus":
        line = run_corpus(args, rig)
    else:
        line = run_single(args, rig)
    if line is not None:
        print(json.dumps(line), flush=True)
    rig.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["tiles", "image", "rgb", "gray16", "corpus"], default="tiles")
    ap.add_argument("--corpus-images", type=int, default=512, help="images per GPU for --workload corpus (512 x 8 GPUs = the 4096 of configs[4])")
    ap.add_argument("--tiles", type=int, default=TOTAL_TILES, help="tiles in the whole batch (sharded over the GPUs) for --workload tiles")
    ap.add_argument("--parity-tiles", type=int, default=256, help="tiles per rank compared with the oracle")
    ap.add_argument("--no-verify", dest="verify", action="store_false", help="skip the oracle parity check of the timed output")
    ap.add_argument("--no-decode", dest="decode", action="store_false", help="skip the decode measurement")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the single-image configs reported beside the tile batch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
