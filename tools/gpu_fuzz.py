#!/usr/bin/env python
"""Randomised parity run (evidence, not a unit test): random frame shapes, feature counts, textures and batch sizes through
the batched C-ABI calls, every frame compared with the CPU oracle (keypoints byte-equal, descriptor mismatches counted
against the 0.1 % budget) and every stereo-style pair with the oracle matcher on the GPU's own descriptors.
Usage: python tools/gpu_fuzz.py [seconds] [seed]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402
from nav24_b200.synth import sequence  # noqa: E402
from oracle import orb_oracle as oo  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
t0 = time.time()
tot = dict(cases=0, frames=0, keypoints=0, kp_mismatch_frames=0, desc_mismatches=0, pairs=0, match_mismatch_pairs=0, shapes=[])
while time.time() - t0 < budget:
    W = int(rng.integers(320, 1500)); H = int(rng.integers(240, min(900, int(1.4 * W))))      # (taller than 2:1 has no quadtree root: the reference divides by zero there)
    nf = int(rng.choice([500, 1000, 2000, 3000]))
    B = int(rng.choice([1, 2, 3, 16, 17, 24, 40])); low = bool(rng.integers(0, 2)); seed = int(rng.integers(0, 1 << 30))
    fr = sequence(H, W, seed, B, step=(int(rng.integers(1, 6)), int(rng.integers(0, 3))), lowtex=low)
    pairs = [(i, i + 1) for i in range(0, B - 1, 2)]
    ctx = capi.OrbContext(nf)
    try:
        if pairs:
            n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, pairs, capi.grid_for(W, H))
        else:
            n, mono, kps, desc = ctx.detect_batch(fr)
    finally:
        ctx.close()
    o = oo.OrbOracle(nf)
    check = sorted(set([0, B - 1] + [int(x) for x in rng.integers(0, B, 3)]))
    for f in check:
        mo, ko, do = o.detect(fr[f])
        ok = mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()
        tot["frames"] += 1; tot["keypoints"] += int(len(ko))
        if not ok:
            tot["kp_mismatch_frames"] += 1
            if len(tot.setdefault("failures", [])) < 40:
                kg = kps[f, :n[f]]
                d = {"H": H, "W": W, "nf": nf, "B": B, "low": low, "seed": seed, "frame": f, "mono": [int(mono[f]), int(mo)], "n": [int(n[f]), int(len(ko))]}
                if n[f] == len(ko):
                    bad = [i for i in range(len(ko)) if kg[i].tobytes() != ko[i].tobytes()]
                    d["n_bad"] = len(bad); d["first_bad"] = bad[:3]
                    if bad:
                        i = bad[0]; d["gpu"] = [float(x) for x in kg[i].tolist()[:6]]; d["ora"] = [float(x) for x in ko[i].tolist()[:6]]
                tot["failures"].append(d)
        else:
            tot["desc_mismatches"] += int((desc[f, :n[f]] != do).any(axis=1).sum())
    for q, (a, b) in enumerate(pairs[:3]):
        k1, k2 = kps[a, :n[a]], kps[b, :n[b]]
        ref = oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), desc[a, :n[a]], k2, np.stack([k2["x"], k2["y"]], 1),
                              desc[b, :n[b]], oo.grid_for(W, H))
        tot["pairs"] += 1
        if not np.array_equal(m[q, :n[a]], ref):
            tot["match_mismatch_pairs"] += 1
    tot["cases"] += 1
    if len(tot["shapes"]) < 12:
        tot["shapes"].append([H, W, nf, B, low])
tot["seconds"] = round(time.time() - t0, 1)
tot["desc_mismatch_rate"] = tot["desc_mismatches"] / max(1, tot["keypoints"])
print(json.dumps(tot))
