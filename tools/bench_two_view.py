#!/usr/bin/env python
"""Side measurement: nav24_two_view_score — one call scores 2 x 200 RANSAC hypotheses over N matches (host
buffers in, scores + inlier masks + kept iteration out; nav24_two_view_score_kept: only the kept iteration's mask) — next to the CPU cost of the same work: CheckHomography /
CheckFundamental of the oracle port, which the reference runs as two threads (OP_2ViewReconstruction.cpp:133-134).
python tools/bench_two_view.py [N ...]  -> one JSON line per N"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from nav24_b200 import capi  # noqa: E402
from oracle import orb_oracle as oo  # noqa: E402
from test_two_view import scene  # noqa: E402

ctx = capi.OrbContext(1000)
for n in [int(a) for a in sys.argv[1:]] or [200, 2000]:
    x1, x2, H21, H12, F21 = scene(5, n)
    for _ in range(5):
        ctx.two_view_score(x1, x2, H21, H12, F21)
    reps = 100
    t0 = time.perf_counter()
    for _ in range(reps):
        r = ctx.two_view_score(x1, x2, H21, H12, F21)
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.two_view_score(x1, x2, H21, H12, F21, want_inliers=False)
    gpu_noinl_ms = (time.perf_counter() - t0) / reps * 1e3
    kept = ctx.two_view_score_kept(x1, x2, H21, H12, F21)
    assert kept["best_h"] == r["best_h"] and np.array_equal(kept["kept_inliers_f"], r["inliers_f"][r["best_f"]])
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.two_view_score_kept(x1, x2, H21, H12, F21)
    gpu_kept_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    sh = [oo.check_homography(H21[h], H12[h], x1, x2)[0] for h in range(200)]
    th = time.perf_counter() - t0
    t0 = time.perf_counter()
    sf = [oo.check_fundamental(F21[h], x1, x2)[0] for h in range(200)]
    tf = time.perf_counter() - t0
    assert np.array_equal(np.array(sh, np.float32), r["score_h"]) and np.array_equal(np.array(sf, np.float32), r["score_f"])
    print(json.dumps(dict(matches=n, hypotheses="2 x 200", gpu_call_ms=round(gpu_ms, 4), gpu_call_ms_without_inlier_masks=round(gpu_noinl_ms, 4),
                          gpu_kept_call_ms=round(gpu_kept_ms, 4), speedup_kept_vs_two_threads=round(max(th, tf) * 1e3 / gpu_kept_ms, 1),
                          cpu_check_homography_ms=round(th * 1e3, 3), cpu_check_fundamental_ms=round(tf * 1e3, 3),
                          cpu_two_threads_ms=round(max(th, tf) * 1e3, 3), speedup_vs_two_threads=round(max(th, tf) * 1e3 / gpu_ms, 1),
                          note="GPU: host buffers in / out through the ctypes binding (includes its numpy overhead); CPU: oracle port through ctypes, one thread per model")))
ctx.close()
