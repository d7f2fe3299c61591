#!/usr/bin/env python
"""Per-kernel SASS evidence from the built library (runs anywhere nvcc's cuobjdump is present; no GPU):
counts of the Blackwell-native mnemonics (B200_PROFILING.md, "What proves a Blackwell-native kernel")
and ptxas resource usage.   python profiles/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "statecatcher_b200", "csrc", "libstatecatcher_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF",
         "SYNCS", "HMMA", "FFMA2", "FMUL2", "FADD2", "MUFU", "LDG.E.128", "STG.E.128", "LDGSTS", "SHFL", "ATOMS", "RED.E"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s*(REG:\d+.*)", res):
        usage[m.group(1)] = m.group(2).strip()
    counts, total, cur = {}, {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur], total[cur] = collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if cur and m:
            op = m.group(1)
            total[cur] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[cur][w] += 1
    names = demangle(list(counts))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels (sm_100a SASS); mnemonic counts are static instruction counts")
    for k in sorted(counts, key=lambda k: names[k]):
        short = re.sub(r"\(.*", "", names[k])
        hits = ", ".join(f"{w} {n}" for w, n in counts[k].items() if n)
        print(f"{short}\n    {total[k]} instr; {usage.get(k, '')}\n    {hits or '-'}")


if __name__ == "__main__":
    sys.exit(main())
