#!/usr/bin/env python
"""Counts of the Blackwell-specific SASS instructions per kernel of the built library (profiles/sass_summary.txt).

usage: python tools/sass_summary.py [path/to/libsmoltts_b200.so] > profiles/sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA tensor load), UBLKCP = cp.async.bulk (TMA bulk
copy), HMMA.16816 = mma.sync.m16n8k16 bf16, SYNCS = mbarrier operations, UTCBAR = tcgen05.commit.
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "smoltts_b200", "_lib", "libsmoltts_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "HMMA.16816", "SYNCS", "STL", "LDL"]
print(f"# {os.path.basename(lib)}: cuobjdump -sass, instruction counts per kernel (static)")
print(f"# {'kernel':70s} {'instr':>7s} " + " ".join(f"{w:>10s}" for w in WATCH))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk)
    c = collections.Counter()
    for op in ops:
        for w in WATCH:
            if op.startswith(w):
                c[w] += 1
    short = re.sub(r"\(.*", "", name)
    print(f"  {short:70s} {len(ops):7d} " + " ".join(f"{c[w]:10d}" for w in WATCH))
