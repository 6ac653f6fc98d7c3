#!/usr/bin/env python
"""Randomised pinning of the two-view RANSAC scoring against the REFERENCE'S OWN TwoViewReconstruction
(oracle/_ref, compiled unchanged from /root/reference): random scenes (8 .. 3000 matches, planar or not, outlier share,
unmatched keypoints, pixel noise), sigma, and hypotheses from three sources — perturbed ground truth, the reference's own
8-point solvers on its own RANSAC sets, and random matrices.  Checked per hypothesis: score bits and inlier flags of the
oracle (always) and of the CUDA kernel (when a GPU is present) against CheckHomography / CheckFundamental, and the
iteration FindHomography / FindFundamental keep.
Usage: python tools/fuzz_two_view.py [seconds] [seed]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import orb_oracle as oo  # noqa: E402
from oracle import ref_lib as rl  # noqa: E402
from test_two_view import keep_loop, scene  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
ctx = None
try:
    import torch
    if torch.cuda.is_available():
        from nav24_b200 import capi
        ctx = capi.OrbContext(1000)
except Exception:
    ctx = None

tot = dict(scenes=0, hypotheses=0, oracle_mismatches=0, cuda_mismatches=0, selection_mismatches=0, nan_scores=0, gpu=ctx is not None, failures=[])
t0 = time.time()
while time.time() - t0 < budget:
    n = int(rng.choice([8, 9, 17, 64, 200, 400, 1000, 3000])); planar = bool(rng.integers(0, 2))
    outl = float(rng.choice([0.0, 0.2, 0.5])); sigma = float(rng.choice([1.0, 1.0, 0.5, 2.0, 3.3])); seed = int(rng.integers(0, 1 << 30))
    x1, x2, H21, H12, F21 = scene(seed, n, planar, outl)
    m12 = np.arange(n, dtype=np.int32)
    drop = rng.random(n) < float(rng.choice([0.0, 0.3]))
    if (~drop).sum() >= 8:
        m12[drop] = -1
    t = rl.RefTwoView(sigma=sigma)
    t.reconstruct(x1, x2, m12)
    a, b, _ = t.matches()
    Hs, His, Fs = t.hypotheses()
    R = rng.normal(0, 1, (200, 9)).astype(np.float32)
    sources = [("perturbed", H21, H12, F21), ("solver", Hs, His, Fs), ("random", R, R[::-1].copy(), R)]
    for tag, hh, hi, ff in sources:
        r = ctx.two_view_score(a, b, hh, hi, ff, sigma=sigma) if ctx is not None else None
        sh = np.zeros(200, np.float32); sf = np.zeros(200, np.float32)
        for h in range(200):
            (srh, irh), (srf, irf) = t.check_homography(hh[h], hi[h]), t.check_fundamental(ff[h])
            (soh, ioh), (sof, iof) = oo.check_homography(hh[h], hi[h], a, b, sigma=sigma), oo.check_fundamental(ff[h], a, b, sigma=sigma)
            sh[h], sf[h] = srh, srf
            nanh, nanf = bool(np.isnan(srh)), bool(np.isnan(srf))
            tot["nan_scores"] += nanh + nanf
            same = lambda p, q, isnan: (isnan and np.isnan(q)) or np.float32(p).tobytes() == np.float32(q).tobytes()  # noqa: E731
            bad_o = (not same(srh, soh, nanh)) + (not np.array_equal(irh, ioh)) + (not same(srf, sof, nanf)) + (not np.array_equal(irf, iof))
            tot["oracle_mismatches"] += bad_o
            bad_c = 0
            if r is not None:
                bad_c = ((not same(srh, r["score_h"][h], nanh)) + (not np.array_equal(irh, r["inliers_h"][h])) +
                         (not same(srf, r["score_f"][h], nanf)) + (not np.array_equal(irf, r["inliers_f"][h])))
                tot["cuda_mismatches"] += bad_c
            if (bad_o or bad_c) and len(tot["failures"]) < 5:
                tot["failures"].append(dict(seed=seed, n=n, planar=planar, outl=outl, sigma=sigma, source=tag, hyp=h))
        tot["hypotheses"] += 400
        if r is not None and (r["best_h"] != keep_loop(sh) or r["best_f"] != keep_loop(sf)):
            tot["selection_mismatches"] += 1
        if tag == "solver":      # the reference's own RANSAC loops keep what the selection loop over these scores keeps
            for find, sc, hyp in ((t.find_homography, sh, Hs), (t.find_fundamental, sf, Fs)):
                score, inl, M = find()
                i = keep_loop(sc)
                ok = (score == 0 and not inl.any()) if i < 0 else (score.tobytes() == sc[i].tobytes() and np.array_equal(M, hyp[i]))
                tot["selection_mismatches"] += not ok
    tot["scenes"] += 1
tot["seconds"] = round(time.time() - t0, 1)
if ctx is not None:
    ctx.close()
print(json.dumps(tot))
sys.exit(1 if tot["oracle_mismatches"] or tot["cuda_mismatches"] or tot["selection_mismatches"] else 0)
