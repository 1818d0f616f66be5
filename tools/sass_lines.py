#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: joins `ncu --page source --print-source sass --csv` (per-SASS-address
counters) with `nvdisasm -g` line info of the cubin.  usage: sass_lines.py <ncu-rep> <object.o> <kernel-substring> [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path


def line_map(obj, kernel):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=td, check=True, capture_output=True)
        cubin = next(Path(td).glob("*.cubin"))
        txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], check=True, capture_output=True, text=True).stdout
    amap, cur, on = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            on = kernel in ln
        if not on:
            continue
        m = re.search(r'//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return amap


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    amap = line_map(obj, kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci, cs, ca = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    base = None
    per_line = defaultdict(lambda: [0, 0])
    per_op = defaultdict(int)
    total = 0
    for r in rows[hi + 1:]:
        if len(r) <= ci or not r[ca]:
            continue
        addr = int(r[ca], 16) if r[ca].startswith("0x") else int(r[ca])
        if base is None:
            base = addr
        n = int(float(r[ci] or 0))
        smp = int(float(r[cs] or 0))
        loc, ins = amap.get(addr - base, (None, "?"))
        per_line[loc][0] += n
        per_line[loc][1] += smp
        per_op[ins.split()[0] if not ins.startswith("@") else ins.split()[1]] += n
        total += n
    tot_s = sum(v[1] for v in per_line.values())
    print(f"total warp instructions {total}, samples {tot_s}")
    src_cache = {}
    for loc, (n, smp) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            f = next((p for p in Path(obj).resolve().parent.rglob(loc[0])), None)
            if f:
                src_cache.setdefault(f, f.read_text().splitlines())
                text = src_cache[f][loc[1] - 1].strip()[:110] if loc[1] - 1 < len(src_cache[f]) else ""
        print(f"{100 * n / total:6.2f}% instr {100 * smp / max(tot_s, 1):6.2f}% samples  {loc}  {text}")
    print("--- by opcode")
    for op, n in sorted(per_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{100 * n / total:6.2f}%  {op}")


if __name__ == "__main__":
    main()
