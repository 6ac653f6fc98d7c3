#!/usr/bin/env python
"""Randomised pinning of the ORACLE against the reference's own classes (oracle/_ref, compiled unchanged from /root/reference):
random shapes, detector parameters, feature counts and image content through both detectors, and random frame pairs through
both window matchers.  CPU only (runs in the build container, where /root/reference exists).
Usage: python tools/cpu_fuzz_ref.py [seconds] [seed]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200.synth import sequence, synth  # noqa: E402
from oracle import orb_oracle as oo  # noqa: E402
from oracle import ref_lib as rl  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12)
t0 = time.time()
tot = dict(detect_cases=0, keypoints=0, detect_mismatches=0, level_mismatches=0, match_cases=0, match_mismatches=0, failures=[])
while time.time() - t0 < budget:
    W = int(rng.integers(200, 1700)); H = int(rng.integers(max(170, W // 4), min(1000, int(1.9 * W))))
    scale = float(rng.choice([1.1, 1.2, 1.2, 1.2, 1.33, 1.5, 1.6, 1.8])); nl = int(rng.integers(3, 9))
    ini = int(rng.choice([10, 20, 20, 30])); mn = int(rng.integers(3, ini + 1)); nf = int(rng.choice([100, 500, 1000, 2000, 5000]))
    if min(W, H) / scale ** (nl - 1) < 70:      # every level must hold one 35-px FAST cell (the reference divides by zero otherwise)
        continue
    # ... and at least one quadtree root: nIni = round(width / height) >= 1 at every level (taller: the reference indexes an
    # empty vector of root nodes; the library answers NAV24_E_GEOMETRY)
    if any(round((round(W / scale ** l) - 32) / (round(H / scale ** l) - 32)) < 1 for l in range(nl)):
        continue
    kind = int(rng.integers(0, 3)); seed = int(rng.integers(0, 1 << 30))
    case = dict(H=H, W=W, scale=scale, nl=nl, ini=ini, mn=mn, nf=nf, kind=kind, seed=seed)
    if os.environ.get("FUZZ_TRACE"):
        print(json.dumps(case), file=sys.stderr, flush=True)
    if kind == 2:
        fr = np.random.default_rng(seed).integers(0, 256, (2, H, W), dtype=np.uint8)
    else:
        fr = sequence(H, W, seed, 2, step=(3, 1), lowtex=kind == 1)
    r, o = rl.RefOrb(nf, scale, nl, ini, mn), oo.OrbOracle(nf, scale, nl, ini, mn)
    res = []
    for f in range(2):
        mr, kr, dr = r.detect(fr[f]); mo, ko, do = o.detect(fr[f])
        ok = mr == mo and len(kr) == len(ko) and kr.tobytes() == ko.tobytes() and np.array_equal(dr, do)
        tot["detect_cases"] += 1; tot["keypoints"] += int(len(ko))
        lv_ok = all(np.array_equal(r.level(l), o.level(l)) for l in range(nl))
        if not lv_ok:
            tot["level_mismatches"] += 1
        if not ok:
            tot["detect_mismatches"] += 1
            if len(tot["failures"]) < 20:
                tot["failures"].append(dict(case, frame=f, n=[len(kr), len(ko)]))
        res.append((ko, do))
    (k1, d1), (k2, d2) = res
    if len(k1) and len(k2):
        kw = dict(nnratio=float(rng.choice([0.6, 0.8, 0.9])), check_ori=bool(rng.integers(0, 2)))      # (the reference's window and TH_LOW are constants)
        ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
        mref = rl.match_window(k1, ud1, d1, k2, ud2, d2, W, H, **kw)
        if mref is not None:
            mora = oo.match_window(k1, ud1, d1, k2, ud2, d2, oo.grid_for(W, H), **kw)
            tot["match_cases"] += 1
            if not np.array_equal(mref, mora):
                tot["match_mismatches"] += 1
                if len(tot["failures"]) < 20:
                    tot["failures"].append(dict(case, kind_="match", kw=kw))
tot["seconds"] = round(time.time() - t0, 1)
print(json.dumps(tot))
