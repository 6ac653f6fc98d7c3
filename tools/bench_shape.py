#!/usr/bin/env python
"""Side measurement (not the bench line): device-resident detect throughput and stage times for any frame shape,
e.g. BASELINE.json configs[3] (4K, 8000 keypoints).  python tools/bench_shape.py H W NFEAT FRAMES [STEPS]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402
from nav24_b200.synth import synth  # noqa: E402

H, W, NF, F = (int(x) for x in sys.argv[1:5])
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
pitch = (W + 127) // 128 * 128
base = [synth(H, W, 24 + i) for i in range(min(F, 4))]
fr = np.zeros((F, H, pitch), np.uint8)
for f in range(F):
    fr[f, :, :W] = base[f % len(base)]
ctx = capi.OrbContext(NF)
dptr = capi.C.c_void_p()
assert ctx.L.nav24_device_alloc(fr.nbytes, capi.C.byref(dptr)) == 0
assert ctx.L.nav24_memcpy_h2d(dptr, fr.ctypes.data_as(capi.C.c_void_p), fr.nbytes) == 0
for _ in range(3):
    ctx.detect_device(dptr.value, F, W, H, pitch, pitch * H)
ctx.sync()
ctx.stage_ms_sum(reset=True)
ctx.timer_start()
for _ in range(steps):
    ctx.detect_device(dptr.value, F, W, H, pitch, pitch * H)
ms = ctx.timer_stop()
stage, calls = ctx.stage_ms_sum(reset=True)
n, mono, _, _ = ctx.fetch(F, want_data=False)
pix = 0; s = np.float32(1.0)
for _ in range(8):
    inv = np.float32(1.0) / s
    pix += int(np.rint(np.float32(W) * inv)) * int(np.rint(np.float32(H) * inv)); s = np.float32(float(s) * float(np.float32(1.2)))
raw = float(np.mean([sum(ctx.L.nav24_orb_get_raw_keys(ctx.h, f, l, None, 0) for l in range(8)) for f in range(min(F, 4))]))
pf = float(stage[0] + stage[1]) / max(calls, 1)
print(json.dumps({"shape": [H, W], "n_features": NF, "frames_per_step": F, "frames_per_sec": F * steps / (ms * 1e-3),
                  "keypoints_per_frame": float(n.mean()), "stage_ms_per_step": [float(x) / max(calls, 1) for x in stage],
                  "pyrFAST_GBps": (pix + 12 * raw) * F / (pf * 1e-3) / 1e9}))
