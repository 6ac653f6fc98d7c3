#!/usr/bin/env python
"""Side measurement: single-frame latency of nav24_orb_detect (host image in, host keypoints + descriptors out) — what a
caller of the reference's per-frame FtDt::detect sees.  python tools/bench_latency.py [H W NFEAT]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402
from nav24_b200.synth import synth  # noqa: E402

H, W, NF = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (376, 1241, 2000)
ctx = capi.OrbContext(NF)
cap = ctx.max_keypoints()
for pinned in (False, True):
    img = capi.pinned_empty((H, W)) if pinned else np.empty((H, W), np.uint8)
    img[...] = synth(H, W, 24)
    kps = capi.pinned_empty((cap,), capi.KP_DTYPE) if pinned else np.zeros(cap, capi.KP_DTYPE)
    desc = capi.pinned_empty((cap, 32)) if pinned else np.zeros((cap, 32), np.uint8)
    n = C.c_int(0)
    args = (ctx.h, img.ctypes.data_as(C.c_void_p), W, H, img.strides[0], kps.ctypes.data_as(C.c_void_p),
            desc.ctypes.data_as(C.c_void_p), cap, C.byref(n))
    for _ in range(20):
        ctx.L.nav24_orb_detect(*args)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        rc = ctx.L.nav24_orb_detect(*args)
        ts.append(time.perf_counter() - t0)
    assert rc >= 0
    ts = np.array(ts) * 1e3
    print(f"{W}x{H} / {NF}: nav24_orb_detect, {'pinned' if pinned else 'pageable'} host buffers: median {np.median(ts):.3f} ms, "
          f"p95 {np.percentile(ts, 95):.3f} ms, {n.value} keypoints")
