#!/usr/bin/env python
"""SASS summary of every kernel in felics_b200/libfelics_b200.so: instruction count, the ten most frequent opcodes, and the
lines that show TMA bulk copies / mbarriers / warp matches / shared-memory atomics (UBLKCP, UTMALDG, SYNCS, MATCH, ATOMS).
usage: sass_summary.py > profiles/sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "felics_b200" / "libfelics_b200.so"
MARK = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MATCH", "ATOMS", "LDGSTS", "REDUX", "VIMNMX.U16x2", "VABSDIFF")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels, cur = {}, None
    arch = None
    for line in txt.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            kernels[cur].append(m.group(1).strip())
    print(f"# {LIB.name}: arch {arch}, {len(kernels)} kernels, {sum(len(v) for v in kernels.values())} SASS instructions")
    print("# per kernel: instructions | top opcodes | marker opcodes (count)")
    for name, ins in sorted(kernels.items(), key=lambda kv: -len(kv[1])):
        ops = Counter()
        marks = Counter()
        for i in ins:
            toks = i.split()
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ops[op.split(".")[0]] += 1
            for mk in MARK:
                if mk in i:
                    marks[op] += 1
        short = re.sub(r"\(.*", "", demangle(name).replace("(anonymous namespace)::", "").replace("felics::", "")).replace("void ", "")
        top = " ".join(f"{o}:{c}" for o, c in ops.most_common(10))
        mk = " ".join(f"{o}:{c}" for o, c in sorted(marks.items())) or "-"
        print(f"{short:44s} {len(ins):6d} | {top} | {mk}")


if __name__ == "__main__":
    main()
