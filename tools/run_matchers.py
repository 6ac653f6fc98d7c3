#!/usr/bin/env python
"""Runs both matchers a few times (for an ncu capture of their kernels): windowed matching of 64 KITTI-shaped stereo pairs
(device-resident) and brute-force kNN-2 on 2000 x 2000 and 8000 x 8000 descriptors, Hamming and L2."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402
import bench  # noqa: E402

ctx = capi.OrbContext(2000)
fr = bench.make_pairs(64, seed=5)
pairs = [(2 * i, 2 * i + 1) for i in range(64)]
for _ in range(3):
    ctx.detect_match_batch(fr, pairs, capi.grid_for(bench.W, bench.H))
rng = np.random.default_rng(0)
for n in (2000, 8000):
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for norm in (0, 1):
        for _ in range(2):
            ctx.match_bf_knn2(d1, d2, norm, 0.7)
print("done")
