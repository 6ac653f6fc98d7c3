#!/usr/bin/env python
"""Third randomised parity run: ONE long-lived context driven like a front end would drive it — the shape, the feature
count (the x5 / x0.2 switch of FE_SlamMonoV), the batch size, the call (host batch, fused detect + match, single frame,
device-resident) and the matcher parameters change from call to call, so the workspace is re-planned, grown and re-used in
every order.  Sampled frames and pairs of every call are compared with the CPU oracle.
Usage: python tools/gpu_fuzz3.py [seconds] [seed]   -> one JSON line"""
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
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
SHAPES = [(376, 1241), (480, 752), (480, 640), (300, 1149), (343, 1335), (600, 460), (260, 340), (720, 1280)]
t0 = time.time()
tot = dict(calls=0, frames=0, keypoints=0, kp_mismatch_frames=0, desc_mismatches=0, pairs=0, match_mismatch_pairs=0, failures=[],
           by_call={})
nf = 1000
ctx = capi.OrbContext(nf)
oracles = {}
try:
    while time.time() - t0 < budget:
        H, W = SHAPES[int(rng.integers(0, len(SHAPES)))]
        r = rng.random()
        if r < 0.25:
            nf = int(rng.choice([300, 1000, 2000])); ctx.set_num_features(nf)
        elif r < 0.35:
            nf = min(nf * 5, 10000); ctx.set_num_features(nf)          # scaleNumFeatures(5.f)
        elif r < 0.45:
            nf = max(nf // 5, 100); ctx.set_num_features(nf)            # scaleNumFeatures(0.2f)
        B = int(rng.choice([1, 1, 2, 5, 16, 30])); seed = int(rng.integers(0, 1 << 30)); low = bool(rng.integers(0, 2))
        fr = sequence(H, W, seed, B, step=(int(rng.integers(1, 6)), int(rng.integers(0, 3))), lowtex=low)
        pairs = [(i, i + 1) for i in range(0, B - 1, 2)] + ([(0, B - 1)] if B > 2 else [])
        kw = dict(window=float(rng.choice([50.0, 100.0])), nnratio=float(rng.choice([0.6, 0.7, 0.9])), th_low=int(rng.choice([50, 70])),
                  check_ori=bool(rng.integers(0, 2)))
        call = str(rng.choice(["host", "host_match", "single", "device"]))
        m = None
        if call == "single" or B == 1:
            call = "single"
            res = [ctx.detect(fr[f]) for f in range(B)]
            n = np.array([len(x[1]) for x in res]); mono = np.array([x[0] for x in res])
            kps = [x[1] for x in res]; desc = [x[2] for x in res]
        elif call == "host" or not pairs:
            call = "host"
            n, mono, kps, desc = ctx.detect_batch(fr)
        elif call == "host_match":
            n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, pairs, capi.grid_for(W, H), **kw)
        else:
            pitch = (W + 127) // 128 * 128
            buf = np.zeros((B, H, pitch), np.uint8); buf[:, :, :W] = fr
            dptr = capi.C.c_void_p()
            assert ctx.L.nav24_device_alloc(buf.nbytes, capi.C.byref(dptr)) == 0
            assert ctx.L.nav24_memcpy_h2d(dptr, buf.ctypes.data_as(capi.C.c_void_p), buf.nbytes) == 0
            ctx.detect_match_device(dptr.value, B, W, H, pitch, pitch * H, pairs, capi.grid_for(W, H), **kw)
            ctx.sync()
            n, mono, kps, desc = ctx.fetch(B)
            m, nm = ctx.match_fetch(len(pairs))
            ctx.L.nav24_device_free(dptr)
        key = nf
        o = oracles.get(key) or oracles.setdefault(key, oo.OrbOracle(nf))
        for f in sorted(set([0, B - 1, int(rng.integers(0, B))])):
            mo, ko, do = o.detect(fr[f])
            kg = kps[f][:n[f]] if isinstance(kps, list) else kps[f, :n[f]]
            dg = desc[f][:n[f]] if isinstance(desc, list) else desc[f, :n[f]]
            ok = mono[f] == mo and n[f] == len(ko) and kg.tobytes() == ko.tobytes()
            tot["frames"] += 1; tot["keypoints"] += int(len(ko))
            if not ok:
                tot["kp_mismatch_frames"] += 1
                if len(tot["failures"]) < 20:
                    tot["failures"].append(dict(H=H, W=W, nf=nf, B=B, low=low, seed=seed, call=call, frame=f, n=[int(n[f]), len(ko)]))
            else:
                tot["desc_mismatches"] += int((dg != do).any(axis=1).sum())
        if m is not None:
            for q in sorted(set([0, len(pairs) - 1])):
                a, b = pairs[q]
                k1, k2 = kps[a, :n[a]], kps[b, :n[b]]
                ref = oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), desc[a, :n[a]], k2, np.stack([k2["x"], k2["y"]], 1),
                                      desc[b, :n[b]], oo.grid_for(W, H), **kw)
                tot["pairs"] += 1
                if not np.array_equal(m[q, :n[a]], ref):
                    tot["match_mismatch_pairs"] += 1
                    if len(tot["failures"]) < 20:
                        tot["failures"].append(dict(H=H, W=W, nf=nf, B=B, seed=seed, call=call, pair=[a, b], kw=kw))
        tot["calls"] += 1; tot["by_call"][call] = tot["by_call"].get(call, 0) + 1
finally:
    ctx.close()
tot["seconds"] = round(time.time() - t0, 1)
tot["desc_mismatch_rate"] = tot["desc_mismatches"] / max(1, tot["keypoints"])
print(json.dumps(tot))
