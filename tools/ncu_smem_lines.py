#!/usr/bin/env python
"""Per-source-line shared-memory wavefronts of one kernel from an .ncu-rep (needs --import-source on, -lineinfo):
wavefronts, ideal wavefronts, excess (bank conflicts).  Usage: python tools/ncu_smem_lines.py REP KERNEL_REGEX [min_pct]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = collections.OrderedDict()
hdr = None
for r in rows:
    if r and r[0] == "Line No":
        if out: break      # first launch only
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10: continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def num(k):
        try: return float(r[hdr[k]].replace(",", "") or 0)
        except (ValueError, KeyError): return 0.0
    e = out.setdefault(ln, [r[1], 0, 0, 0, 0])
    e[1] += num("L1 Wavefronts Shared"); e[2] += num("L1 Wavefronts Shared Ideal"); e[3] += num("L1 Wavefronts Shared Excessive")
    e[4] += num("Instructions Executed")
tot = sum(e[1] for e in out.values())
print(f"total shared wavefronts {tot/1e6:.1f}M ideal {sum(e[2] for e in out.values())/1e6:.1f}M excessive {sum(e[3] for e in out.values())/1e6:.1f}M; warp inst {sum(e[4] for e in out.values())/1e6:.1f}M")
for ln, e in out.items():
    if e[1] * 100 >= minpct * tot:
        print(f"{ln:5d} {100*e[1]/tot:5.1f}% wavefronts  x{e[1]/max(e[2],1):4.1f} of ideal | {e[0].strip()[:110]}")
