#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep (needs --import-source on, -lineinfo).
Usage: python tools/ncu_lines.py REP KERNEL_REGEX [min_pct]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
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
        except ValueError: return 0.0
    ie, te, smp = num("Instructions Executed"), num("Thread Instructions Executed"), num("# Samples")
    e = out.setdefault(ln, [r[1], 0, 0, 0])
    e[1] += ie; e[2] += te; e[3] += smp
tot = sum(e[1] for e in out.values()); tots = sum(e[3] for e in out.values())
print(f"total warp inst {tot/1e6:.1f}M samples {tots:.0f}")
for ln, e in out.items():
    if e[1] * 100 >= minpct * tot or e[3] * 100 >= minpct * tots:
        print(f"{ln:5d} {100*e[1]/tot:5.1f}% inst {100*e[3]/max(tots,1):5.1f}% smp  thr/inst {e[2]/max(e[1],1):4.1f} | {e[0].strip()[:110]}")
