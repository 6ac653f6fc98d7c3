#!/usr/bin/env python
"""Instruction / stall-sample share per section of fast_band_kernel. Usage: ncu_sections.py REP"""
import re, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "fast_band", "0.0"], capture_output=True, text=True).stdout
src = open('nav24_b200/csrc/orb_kernels.cu').read().split('\n')
def find(s):
    for i, l in enumerate(src):
        if s in l: return i + 1
    return None
marks = [('fast_score', 'int fast_score('), ('gt_any2', 'unsigned gt_any2('), ('mask_range_count', 'int mask_range_count('),
         ('setup', 'fast_band_kernel(const __grid_constant__ FrameGeom g'), ('stageA', 'for (int pass = 0; pass < 2; ++pass)'),
         ('nms+emit bodies', 'auto nms = [&]'), ('stageB', 'if (!dense) {'), ('stageC', 'stage C over the warp'),
         ('dense', '// dense path (rare'), ('count+emit', '// count and ordered emit, one warp per cell'), ('end', '// K3  quadtree distribution')]
marks = [(n, find(s)) for n, s in marks if find(s)]
tot = {}
for line in txt.split('\n'):
    m = re.match(r'\s*(\d+)\s+([\d.]+)% inst\s+([\d.]+)% smp', line)
    if not m: continue
    ln = int(m.group(1))
    for (n, a), (_, b) in zip(marks, marks[1:]):
        if a <= ln < b:
            t = tot.setdefault(n, [0, 0]); t[0] += float(m.group(2)); t[1] += float(m.group(3))
print(txt.split('\n')[0])
for k, v in tot.items(): print(f"{k:18s} inst {v[0]:5.1f}%  samples {v[1]:5.1f}%")
