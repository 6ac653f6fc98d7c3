#!/usr/bin/env python
"""Fourth randomised parity run: the entry points either side of the detector — brute-force kNN-2 matcher (both norms, random
sizes incl. degenerate ones, duplicated rows for ties), camera undistortion (RadTan / KB8 with random intrinsics), the colour
ingest ring (random widths, BGR -> grey on the device, then the detector) and two-view RANSAC scoring (random hypotheses) —
each against the CPU oracle.  Usage: python tools/gpu_fuzz4.py [seconds] [seed]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402
from nav24_b200.synth import synth  # noqa: E402
from oracle import orb_oracle as oo  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
t0 = time.time()
tot = dict(bf_cases=0, bf_bad=0, ud_cases=0, ud_points=0, ud_bad_points=0, ring_cases=0, ring_bad=0, tv_cases=0, tv_bad=0, failures=[])
ctx = capi.OrbContext(600)


def fail(**kw):
    if len(tot["failures"]) < 20:
        tot["failures"].append(kw)


try:
    while time.time() - t0 < budget:
        which = int(rng.integers(0, 4))
        if which == 0:      # brute force kNN-2
            n1 = int(rng.choice([1, 2, 31, 128, 129, 700, 2100])); n2 = int(rng.choice([1, 2, 33, 127, 128, 900, 2500]))
            norm = int(rng.integers(0, 2)); ratio = float(rng.choice([0.6, 0.7, 0.8]))
            d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
            d1 = d2[rng.integers(0, n2, n1)].copy()
            d1 ^= (rng.integers(0, 256, d1.shape, dtype=np.uint8) & rng.integers(0, 256, d1.shape, dtype=np.uint8) & rng.integers(0, 256, d1.shape, dtype=np.uint8))
            if n2 > 4:
                d2[n2 // 2] = d2[n2 // 2 - 1]          # an exact tie
            ref = oo.match_bf_knn2(d1, d2, norm, ratio); got = ctx.match_bf_knn2(d1, d2, norm, ratio)
            tot["bf_cases"] += 1
            if not all(np.array_equal(r, g) for r, g in zip(ref, got)):
                tot["bf_bad"] += 1; fail(kind="bf", n1=n1, n2=n2, norm=norm, ratio=ratio)
        elif which == 1:    # undistortion
            model = int(rng.choice([capi.CAM_RADTAN, capi.CAM_KB8]))
            fx, fy = float(rng.uniform(300, 900)), float(rng.uniform(300, 900)); cx, cy = float(rng.uniform(250, 700)), float(rng.uniform(150, 400))
            D = [float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.01, 0.01)), float(rng.uniform(-0.01, 0.01))]
            n = int(rng.choice([1, 7, 500, 3000]))
            xy = np.stack([rng.uniform(0, 2 * cx, n), rng.uniform(0, 2 * cy, n)], 1).astype(np.float32)
            ref = oo.undistort(model, [fx, fy, cx, cy], D, xy)
            got = ctx.undistort_points(capi.Camera.make(model, [fx, fy, cx, cy], D), xy)
            tot["ud_cases"] += 1; tot["ud_points"] += n
            if model == capi.CAM_RADTAN:
                bad = int((got.view(np.uint32) != ref.view(np.uint32)).any(axis=1).sum())
            else:       # KB8: <= 1 float ulp
                bad = int((np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64)) > 1).any(axis=1).sum())
            tot["ud_bad_points"] += bad
            if bad:
                fail(kind="ud", model=model, K=[fx, fy, cx, cy], D=D, n=n, bad=bad)
        elif which == 2:    # colour ingest ring -> detector
            W = int(rng.integers(300, 900)); H = int(rng.integers(250, min(500, int(1.3 * W)))); nfr = int(rng.choice([1, 2, 5]))
            grey = np.stack([synth(H, W, int(rng.integers(0, 1 << 20))) for _ in range(nfr)])
            bgr = np.stack([grey, np.roll(grey, 3, axis=2), 255 - grey], axis=3).astype(np.uint8)      # three different channels
            want_grey = np.stack([oo.bgr2gray(np.ascontiguousarray(bgr[f])) for f in range(nfr)])
            ring = capi.IngestRing(ctx, W, H, 3, nfr)
            try:
                for k in range(nfr):
                    ring.slot(k)[...] = bgr[k]
                res = ring.detect_match(0, nfr, [(0, 1)] if nfr > 1 else [], capi.grid_for(W, H))
                n, mono, kps, desc = res[0], res[1], res[2], res[3]
                o = oo.OrbOracle(600)
                ok = True
                for f in range(nfr):
                    ok &= bool(np.array_equal(ctx.level(f, 0), want_grey[f]))
                    mo, ko, do = o.detect(np.ascontiguousarray(want_grey[f]))
                    ok &= bool(mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes())
                tot["ring_cases"] += 1
                if not ok:
                    tot["ring_bad"] += 1; fail(kind="ring", H=H, W=W, nfr=nfr)
            finally:
                ring.close() if hasattr(ring, "close") else None
        else:               # two-view scoring
            n = int(rng.choice([8, 50, 333, 1500])); nh = int(rng.choice([1, 17, 200]))
            x1 = rng.uniform(0, 1000, (n, 2)).astype(np.float32); x2 = (x1 + rng.normal(0, 2, (n, 2))).astype(np.float32)
            H21 = (np.eye(3, dtype=np.float32).reshape(1, 9) + rng.normal(0, 1e-3, (nh, 9))).astype(np.float32)
            H12 = np.stack([np.linalg.inv(h.reshape(3, 3).astype(np.float64)).astype(np.float32).reshape(9) for h in H21])
            F21 = rng.normal(0, 1e-3, (nh, 9)).astype(np.float32)
            r = ctx.two_view_score(x1, x2, H21, H12, F21)
            ok = True
            for h in range(nh):
                sh, ih = oo.check_homography(H21[h], H12[h], x1, x2); sf, i_f = oo.check_fundamental(F21[h], x1, x2)
                ok &= bool(np.float32(sh).tobytes() == r["score_h"][h].tobytes() and np.float32(sf).tobytes() == r["score_f"][h].tobytes())
                ok &= bool(np.array_equal(r["inliers_h"][h], ih) and np.array_equal(r["inliers_f"][h], i_f))
            tot["tv_cases"] += 1
            if not ok:
                tot["tv_bad"] += 1; fail(kind="two_view", n=n, nh=nh)
finally:
    ctx.close()
tot["seconds"] = round(time.time() - t0, 1)
print(json.dumps(tot))
