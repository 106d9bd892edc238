#!/usr/bin/env python
"""Opcode histogram per kernel of libquflow_b200.so (cuobjdump -sass), the evidence for what the hot kernels execute:
DMMA (FP64 tensor MMA), UTMALDG / UBLKPF / UBLKCP (TMA and bulk copies), SYNCS (mbarrier), UCGABAR (cluster barriers),
LDGSTS (cp.async), STL/LDL (register spills).

    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "quflow_b200", "_cuda", "libquflow_b200.so")
WATCH = ["DMMA", "DFMA", "DADD", "DMUL", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTMAPF", "SYNCS", "UCGABAR", "LDGSTS", "LDS", "STS",
         "LDG", "STG", "ST.E", "LD.E", "RED", "ATOM", "SHFL", "BAR", "MEMBAR", "ERRBAR", "CCTL", "STL", "LDL", "HMMA", "UTCHMMA", "LDTM"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Za-z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            cur[op.split(".")[0]] += 1
            if op.startswith(("DMMA", "UTMALDG", "UTMASTG", "SYNCS", "UBLK", "UCGABAR", "LDG.E", "STG.E")):
                cur[op] += 1
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (static instruction counts, cuobjdump -sass, sm_100a)")
    for k, c in kernels.items():
        short = re.sub(r"\(anonymous namespace\)::", "", names[k])
        short = re.sub(r"\(.*", "", short)
        print(f"\n{short}   [{c['_total']} instructions]")
        keys = [w for w in WATCH if c.get(w)]
        print("   " + "  ".join(f"{w}={c[w]}" for w in keys))
        detail = sorted((op, n) for op, n in c.items() if "." in op)
        if detail:
            print("   " + "  ".join(f"{op}={n}" for op, n in detail))


if __name__ == "__main__":
    main()
