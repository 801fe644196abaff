#!/usr/bin/env python
"""SASS evidence for libaurdf.so: per kernel, the instruction count and the mnemonics that show which hardware
paths it uses (cuobjdump -sass; runs here, no GPU).   python profiles/sass_summary.py r02 > profiles/r02_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "autourdf_b200", "libaurdf.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
WATCH = ["UBLKCP", "SYNCS", "UTMALDG", "DMMA", "FFMA2", "FADD2", "FMUL2", "VIMNMX3", "VIMNMX", "DFMA", "DADD", "DMUL", "MUFU",
         "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "ATOMS", "SHFL", "BAR", "UCGABAR", "MEMBAR", "LDL", "STL", "HMMA", "UTCHMMA", "LDTM"]
kern, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
        kern[cur]["_total"] += 1
print(f"# SASS summary of autourdf_b200/libaurdf.so ({tag}; cuobjdump -sass; architectures in the binary: {', '.join(arch)})\n")
print("Mnemonic counts per kernel (static, not executed).  UBLKCP = cp.async.bulk (TMA engine, 1-D), SYNCS = mbarrier, DMMA = FP64 tensor-core "
      "mma.sync.m8n8k4, FFMA2/FADD2/FMUL2 = packed f32x2, VIMNMX3 = 3-input integer min/max, UCGABAR = cluster barrier, LDL/STL = local memory "
      "(spills / indexed local arrays).  No UTC*MMA / LDTM / HMMA: the north star rules tensor-core GEMMs out (no dense contraction on this path).\n")
cols = [w for w in WATCH if any(c[w] for c in kern.values())]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, c in kern.items():
    print(f"| `{k[:70]}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
