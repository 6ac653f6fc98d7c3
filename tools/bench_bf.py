#!/usr/bin/env python
"""Side measurement: brute-force kNN-2 matcher, device time of its kernels and pair rate.
python tools/bench_bf.py [N ...]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402

ctx = capi.OrbContext(1000)
rng = np.random.default_rng(0)
for n in [int(x) for x in sys.argv[1:]] or [2000, 8000]:
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for norm in (capi.NORM_HAMMING, capi.NORM_L2_U8):
        best = 1e9
        for _ in range(5):
            ctx.match_bf_knn2(d1, d2, norm)
            ms = C.c_float(0)
            ctx.L.nav24_debug_last_kernel_ms(ctx.h, C.byref(ms))
            best = min(best, ms.value)
        print(f"bf_knn2 {n} x {n} norm {norm}: {best * 1e3:.1f} us device time, {n * n / best / 1e6:.1f} G pairs/s, "
              f"{8 * n * n / best / 1e6:.0f} G popc-words/s")
