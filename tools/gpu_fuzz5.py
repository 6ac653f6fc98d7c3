#!/usr/bin/env python
"""Fifth randomised parity run: the chunked host pipeline — random chunk size, stream count, taper, batch size and pair lists
(pairs inside a chunk, pairs that span chunks, repeated frames) — against the same batch run as ONE chunk on one stream, and
sampled frames / pairs against the CPU oracle; then the device-resident path with random resident chunking.
Usage: python tools/gpu_fuzz5.py [seconds] [seed]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402
from nav24_b200.synth import sequence  # noqa: E402
from oracle import orb_oracle as oo  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
t0 = time.time()
tot = dict(cases=0, frames=0, pairs=0, schedule_mismatches=0, oracle_frame_mismatches=0, oracle_pair_mismatches=0, failures=[])
SHAPES = [(260, 340, 300), (300, 427, 400), (376, 621, 600)]


def run(env, fr, pairs, grid, nf, device):
    for k in ("NAV24_CHUNK_FRAMES", "NAV24_STREAMS", "NAV24_TAPER", "NAV24_RESIDENT_CHUNK"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = capi.OrbContext(nf)
    try:
        if not device:
            return ctx.detect_match_batch(fr, pairs, grid)
        B, H, W = fr.shape
        pitch = (W + 127) // 128 * 128
        buf = np.zeros((B, H, pitch), np.uint8); buf[:, :, :W] = fr
        dptr = capi.C.c_void_p()
        assert ctx.L.nav24_device_alloc(buf.nbytes, capi.C.byref(dptr)) == 0
        assert ctx.L.nav24_memcpy_h2d(dptr, buf.ctypes.data_as(capi.C.c_void_p), buf.nbytes) == 0
        for _ in range(2):
            ctx.detect_match_device(dptr.value, B, W, H, pitch, pitch * H, pairs, grid)
        ctx.sync()
        out = ctx.fetch(B) + ctx.match_fetch(len(pairs))
        ctx.L.nav24_device_free(dptr)
        return out
    finally:
        ctx.close()


while time.time() - t0 < budget:
    H, W, nf = SHAPES[int(rng.integers(0, len(SHAPES)))]
    B = int(rng.integers(20, 150)); seed = int(rng.integers(0, 1 << 30))
    fr = sequence(H, W, seed, B, step=(int(rng.integers(1, 4)), int(rng.integers(0, 2))), lowtex=bool(rng.integers(0, 2)))
    P = int(rng.integers(1, 40))
    pairs = [(int(a), int(b)) for a, b in zip(rng.integers(0, B, P), rng.integers(0, B, P))] + [(i, i + 1) for i in range(0, min(B - 1, 20), 2)]
    grid = capi.grid_for(W, H)
    device = bool(rng.integers(0, 3) == 0)
    env = ({"NAV24_RESIDENT_CHUNK": str(int(rng.choice([3, 7, 16, 33, 64])))} if device else
           {"NAV24_CHUNK_FRAMES": str(int(rng.choice([8, 16, 32, 64]))), "NAV24_STREAMS": str(int(rng.integers(1, 5))), "NAV24_TAPER": str(int(rng.integers(0, 2)))})
    n, mono, kps, desc, m, nm = run(env, fr, pairs, grid, nf, device)
    n2, mono2, kps2, desc2, m2, nm2 = run({"NAV24_CHUNK_FRAMES": "100000", "NAV24_STREAMS": "1", "NAV24_TAPER": "0", "NAV24_RESIDENT_CHUNK": "100000"}, fr, pairs, grid, nf, device)
    same = np.array_equal(n, n2) and np.array_equal(mono, mono2) and np.array_equal(nm, nm2)
    same = same and all(kps[f, :n[f]].tobytes() == kps2[f, :n[f]].tobytes() and np.array_equal(desc[f, :n[f]], desc2[f, :n[f]]) for f in range(B))
    same = same and all(np.array_equal(m[q, :n[a]], m2[q, :n[a]]) for q, (a, b) in enumerate(pairs))
    if not same:
        tot["schedule_mismatches"] += 1
        if len(tot["failures"]) < 20:
            tot["failures"].append(dict(kind="schedule", H=H, W=W, B=B, seed=seed, env=env, device=device))
    o = oo.OrbOracle(nf)
    for f in sorted(set([0, B - 1, int(rng.integers(0, B))])):
        mo, ko, do = o.detect(fr[f])
        tot["frames"] += 1
        if not (mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()):
            tot["oracle_frame_mismatches"] += 1
    for q in sorted(set([0, len(pairs) - 1, int(rng.integers(0, len(pairs)))])):
        a, b = pairs[q]
        k1, k2 = kps[a, :n[a]], kps[b, :n[b]]
        ref = oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), desc[a, :n[a]], k2, np.stack([k2["x"], k2["y"]], 1), desc[b, :n[b]],
                              oo.grid_for(W, H))
        tot["pairs"] += 1
        if not np.array_equal(m[q, :n[a]], ref):
            tot["oracle_pair_mismatches"] += 1
            if len(tot["failures"]) < 20:
                tot["failures"].append(dict(kind="pair", H=H, W=W, B=B, seed=seed, env=env, device=device, pair=[a, b]))
    tot["cases"] += 1
tot["seconds"] = round(time.time() - t0, 1)
print(json.dumps(tot))
