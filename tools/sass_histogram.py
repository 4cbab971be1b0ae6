#!/usr/bin/env python
"""SASS opcode histogram per kernel of libffb200.so (cuobjdump -sass): how many instructions of each class a
kernel holds (static counts), and the Blackwell-relevant ones called out (conversion-pipe F2F/F2I/I2F/FRND,
fp64 D*, LDGSTS = cp.async, UBLKCP/UTMA* = bulk/tensor async copies).

    python tools/sass_histogram.py [kernel-name-regex] > profiles/r2_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "blender_flip_fluids_b200", "lib", "libffb200.so")
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = {}
hist = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
XU = ("F2F", "F2I", "I2F", "FRND", "MUFU", "F2FP", "I2FP")
print("# SASS opcode histogram per kernel (static instruction counts, `cuobjdump -sass libffb200.so`, sm_100a)\n")
print("XU = conversion / special-function pipe (F2F F2I I2F FRND MUFU), FP64 = D* opcodes, LDGSTS = cp.async, "
      "UBLKCP / UTMA* = bulk / tensor async copies (none: the MAC grids' row pitch (I+1)*4 B is not 16-byte aligned, "
      "which bulk copies require).\n")
print("| kernel | instructions | XU | FP64 | FFMA/FMUL/FADD | LDG | STG | LDS/STS | LDGSTS | UBLKCP/UTMA | ATOM/RED | BAR | top opcodes |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for (mangled, c), name in zip(hist.items(), names):
    short = re.sub(r"\(anonymous namespace\)::|ffb200::", "", name).split("(")[0]
    if pat and not pat.search(short):
        continue
    tot = sum(c.values())
    xu = sum(v for k, v in c.items() if k in XU)
    fp64 = sum(v for k, v in c.items() if k.startswith("D") and k not in ("DEPBAR",))
    fp32 = sum(v for k, v in c.items() if k in ("FFMA", "FMUL", "FADD", "FSEL", "FSETP", "FMNMX"))
    lds = c["LDS"] + c["STS"]
    bulk = sum(v for k, v in c.items() if k.startswith("UBLKCP") or k.startswith("UTMA"))
    atom = sum(v for k, v in c.items() if k.startswith("ATOM") or k.startswith("RED"))
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(6))
    print(f"| `{short}` | {tot} | {xu} | {fp64} | {fp32} | {c['LDG']} | {c['STG']} | {lds} | {c['LDGSTS']} | {bulk} | {atom} | {c['BAR']} | {top} |")
