#!/usr/bin/env python
"""Per-kernel SASS evidence of libnav24orb.so (cuobjdump -sass): TMA (UTMALDG), mbarrier (SYNCS), tensor-core (UTC*MMA / HMMA:
expected 0 — the path is integer / byte work), integer-pipe mnemonics (POPC, VABSDIFF4, IDP = DP4A/DP2A, VIMNMX3), shared
and global memory instructions.  Usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nav24_b200/libnav24orb.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = ["UTMALDG", "SYNCS", "UTC.MMA", "HMMA", "POPC", "VABSDIFF4", "IDP", "VIMNMX3", "LDS", "STS", "ATOMS", "LDG", "STG", "ATOMG/RED", "total"]
rows = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        mm = re.search(r"(\w+_kernel(?:<[^>(]*>)?)", name)
        name = mm.group(1) if mm else name[:43]
        cur = rows.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    cur["total"] += 1
    for c in COLS:
        if c == "UTC.MMA":
            hit = re.match(r"UTC[A-Z]*MMA", op)
        elif c == "ATOMG/RED":
            hit = op.startswith("ATOMG") or op.startswith("RED") or op.startswith("ATOM.")
        elif c == "IDP":
            hit = op.startswith("IDP")
        else:
            hit = op.startswith(c)
        if hit and c != "total":
            cur[c] += 1
print(f"# cuobjdump -sass {lib}: instruction counts per kernel (static SASS, sm_100a)")
print("kernel".ljust(44) + "".join(c.rjust(11) for c in COLS))
for k, v in rows.items():
    print(k[:43].ljust(44) + "".join(str(v[c]).rjust(11) for c in COLS))
