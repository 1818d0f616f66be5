#!/usr/bin/env python
"""Warp instructions and stall samples of k_stream_encode per phase (source line ranges found from the `// ----` markers
of stream.cu).  usage: phase_split.py <ncu-rep> [tiles]"""
import csv, io, re, subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).parent))
import sass_lines as sl

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "felics_b200/csrc/stream.cu"


def helper_ranges(lines):
    """(first, last, phase) for the device helpers above the kernel, keyed by function name."""
    names = {"load_band": "load", "make_info": "classify", "classify_slow": "classify", "classify_lane": "classify", "load_state": "walk", "store_state": "walk",
             "cost_keys": "walk", "min6": "walk", "halve_keys": "walk", "serial_walk": "walk", "coop_walk": "walk", "make_record": "code",
             "put_clipped": "pack", "emit_fields": "pack", "info_index": "index", "cp_async16": "load", "cp_async4": "load"}
    out = []
    starts = [(i + 1, m.group(1)) for i, l in enumerate(lines) for m in [re.match(r"(?:template.*\n)?__device__ __forceinline__ \S+ (\w+)\(", l)] if m]
    for (ln, name), nxt in zip(starts, starts[1:] + [(len(lines), None)]):
        out.append((ln, nxt[0] - 1, names.get(name, "other")))
    return out


def main():
    rep = sys.argv[1]
    tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
    lines = SRC.read_text().splitlines()
    k0 = next(i + 1 for i, l in enumerate(lines) if "k_stream_encode(StreamArgs a)" in l)
    k1 = next(i + 1 for i, l in enumerate(lines) if "k_stream_scan" in l and "__global__" in l)
    marks = [(i + 1, re.sub(r"[-/ ]+", " ", l).strip().split(":")[0]) for i, l in enumerate(lines) if k0 < i + 1 < k1 and re.match(r"\s*// ---- ", l)]
    ranges = [(k0, marks[0][0] - 1, "setup/load")] + [(a, b[0] - 1, n) for (a, n), b in zip(marks, marks[1:] + [(k1, None)])]
    helpers = [r for r in helper_ranges(lines) if r[0] < k0]

    def phase(loc):
        if loc is None:
            return None
        f, l = loc
        if f == "stream.cu":
            for a, b, n in ranges + helpers:
                if a <= l <= b:
                    return n
            return None
        if f == "device_common.cuh":
            return "code" if l < 150 else "pack"
        return None

    amap = sl.line_map(str(ROOT / "felics_b200/csrc/stream.cu.o"), "k_stream_encode")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci, cs, ca = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    base, last, tot, smp = None, "setup/load", {}, {}
    for r in rows[hi + 1:]:
        if len(r) <= ci or not r[ca]:
            continue
        addr = int(r[ca], 16) if r[ca].startswith("0x") else int(r[ca])
        base = addr if base is None else base
        loc, _ = amap.get(addr - base, (None, "?"))
        p = phase(loc)
        if p in (None, "index", "other"):
            p = last
        else:
            last = p
        tot[p] = tot.get(p, 0) + int(float(r[ci] or 0))
        smp[p] = smp.get(p, 0) + int(float(r[cs] or 0))
    T, S, px = sum(tot.values()), sum(smp.values()), tiles * 262144
    for p, n in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{p:28s} {100 * n / T:6.2f}% instr {n / px:7.3f} warp-instr/px {100 * smp[p] / S:6.2f}% samples")
    print(f"total {T / px:.3f} warp-instr/px")


if __name__ == "__main__":
    main()
