#!/usr/bin/env python
"""SASS evidence from the in-tree build: per kernel of libvbg_sm100a.so the counts of the mnemonics that prove the Blackwell
paths (tcgen05.mma = UTCHMMA, TMA tensor loads = UTMALDG, bulk copies = UBLKCP, tcgen05.ld = LDTM, tcgen05.commit = UTCBAR,
mbarrier ops = SYNCS, cp.async = LDGSTS).  CPU-only (cuobjdump).   python scripts/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vibertgrid-pytorch_b200", "libvbg_sm100a.so")
COLS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "LDGSTS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for c in COLS:
            if re.search(r"\b" + c + r"\b|\b" + c + r"\.", line):
                counts[cur][c] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel (cuobjdump -sass vibertgrid-pytorch_b200/libvbg_sm100a.so, sm_100a); kernels with none of them omitted")
print("# " + " | ".join(f"{c:>7}" for c in COLS) + " | kernel")
for (k, c), name in sorted(zip(counts.items(), demangle), key=lambda t: t[1]):
    if sum(c.values()):
        print("  " + " | ".join(f"{c[x]:>7}" for x in COLS) + " | " + re.sub(r"\(.*", "", name)[:110])
