#!/usr/bin/env python
"""Stage-by-stage comparison of one batched case against the oracle (first divergent stage per level).
Usage: python tools/gpu_debug_case.py H W nf B low seed [frame ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi
from nav24_b200.synth import sequence
from oracle import orb_oracle as oo
H, W, nf, B, low, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] == "1", int(sys.argv[6])
frames = [int(x) for x in sys.argv[7:]] or [0]
fr = sequence(H, W, seed, B, step=(2, 1), lowtex=low)
for Bsub in [1]:
    ctx = capi.OrbContext(nf)
    sub = fr[:Bsub]
    n, mono, kps, desc = ctx.detect_batch(sub)
    for f in [x for x in frames if x < Bsub]:
        o = oo.OrbOracle(nf)
        mo, ko, do = o.detect(sub[f])
        msg = []
        for l in range(8):
            if not np.array_equal(ctx.level(f, l), o.level(l)): msg.append(f"L{l}:pyr")
            rg, ro = ctx.raw_keys(f, l), o.raw(l)
            if rg.shape != ro.shape or not np.array_equal(rg, ro):
                d = ""
                if rg.shape == ro.shape:
                    bad = np.nonzero((rg != ro).any(axis=1))[0]
                    d = f"({len(bad)} of {len(ro)} differ, first {bad[:3].tolist()} gpu {rg[bad[0]].tolist()} ora {ro[bad[0]].tolist()})"
                else:
                    sg_ = set(map(tuple, rg[:, :2].tolist())); so_ = set(map(tuple, ro[:, :2].tolist()))
                    miss = sorted(so_ - sg_); extra = sorted(sg_ - so_)
                    d = f"(count {len(rg)} vs {len(ro)}; missing on GPU {miss[:12]}; extra on GPU {extra[:12]})"
                msg.append(f"L{l}:raw{d}")
            lg, lo_ = ctx.level_keypoints(f, l), o.level_kps(l)[0]
            if lg.tobytes() != lo_.tobytes(): msg.append(f"L{l}:lkp({len(lg)} vs {len(lo_)})")
        ok = mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()
        print(f"B={Bsub} frame {f}: final {'OK' if ok else 'MISMATCH'}; stages: {' '.join(msg) or 'all equal'}")
    ctx.close()
