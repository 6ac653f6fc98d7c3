#!/usr/bin/env python
"""Second randomised parity run: detector PARAMETERS and image content instead of shapes alone — scale factor 1.1..1.6, 3..8
levels, FAST thresholds, feature counts 100..9000 (incl. the x5 mode sizes), aspect ratios up to 1 : 1.9, textures
(rectangles, low texture, pure noise, flat with a few blobs, gradients) — every case against the CPU oracle, raw FAST keys
of every level included.  Geometries the library documents as unsupported (NAV24_E_GEOMETRY) are counted and skipped.
Usage: python tools/gpu_fuzz2.py [seconds] [seed]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200 import capi  # noqa: E402
from nav24_b200.synth import synth  # noqa: E402
from oracle import orb_oracle as oo  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
t0 = time.time()
tot = dict(cases=0, skipped_geometry=0, keypoints=0, raw_level_mismatches=0, pyramid_level_mismatches=0, kp_mismatch_cases=0,
           desc_mismatches=0, failures=[])


def image(H, W, kind, seed):
    r = np.random.default_rng(seed)
    if kind == 0:
        return synth(H, W, seed)
    if kind == 1:
        return synth(H, W, seed, lowtex=True)
    if kind == 2:
        return r.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == 3:
        img = np.full((H, W), int(r.integers(0, 256)), np.uint8)
        for _ in range(int(r.integers(1, 30))):
            y, x = int(r.integers(0, H - 8)), int(r.integers(0, W - 8))
            img[y:y + int(r.integers(2, 8)), x:x + int(r.integers(2, 8))] = int(r.integers(0, 256))
        return img
    yy, xx = np.mgrid[0:H, 0:W]
    return ((xx * int(r.integers(1, 5)) + yy * int(r.integers(1, 5)) + r.integers(-4, 5, (H, W))) & 255).astype(np.uint8)


while time.time() - t0 < budget:
    if os.environ.get("FUZZ_BIG"):      # 2K .. 4K frames, feature counts up to the x5 mode of a 4K camera
        W = int(rng.integers(1500, 4100)); H = int(rng.integers(max(700, W // 4), min(2300, int(1.5 * W))))
    else:
        W = int(rng.integers(200, 1700)); H = int(rng.integers(max(160, W // 4), min(1000, int(1.9 * W))))
    scale = float(rng.choice([1.1, 1.15, 1.2, 1.25, 1.33, 1.5, 1.6])); nl = int(rng.integers(3, 9))
    ini = int(rng.choice([10, 15, 20, 30, 45])); mn = int(rng.integers(3, ini + 1))
    nf = int(rng.choice([2000, 8000, 20000, 40000] if os.environ.get("FUZZ_BIG") else [100, 400, 1000, 2000, 5000, 9000])); kind = int(rng.integers(0, 5)); seed = int(rng.integers(0, 1 << 30))
    img = image(H, W, kind, seed)
    case = dict(H=H, W=W, scale=scale, nl=nl, ini=ini, mn=mn, nf=nf, kind=kind, seed=seed)
    if os.environ.get("FUZZ_TRACE"):
        print(json.dumps(case), file=sys.stderr, flush=True)
    try:
        ctx = capi.OrbContext(nf, scale_factor=scale, n_levels=nl, ini_th_fast=ini, min_th_fast=mn,
                              raw_keys_per_kpx=250 if kind == 2 else 0)
    except capi.Nav24Error:
        tot["skipped_geometry"] += 1
        continue
    try:
        try:
            mono_g, k_g, d_g = ctx.detect(img)
        except capi.Nav24Error as e:
            if e.code in (capi.E_GEOMETRY, capi.E_OVERFLOW):      # documented limits (tiny levels, scale > ~1.85, raw-corner budget)
                tot["skipped_geometry"] += 1
                continue
            raise
        o = oo.OrbOracle(nf, scale, nl, ini, mn)
        mono_o, k_o, d_o = o.detect(img)
        bad = []
        for l in range(nl):
            if not np.array_equal(ctx.level(0, l), o.level(l)):
                tot["pyramid_level_mismatches"] += 1; bad.append(f"L{l}:pyr")
            rg, ro = ctx.raw_keys(0, l), o.raw(l)
            if rg.shape != ro.shape or not np.array_equal(rg, ro):
                tot["raw_level_mismatches"] += 1; bad.append(f"L{l}:raw")
        ok = mono_g == mono_o and len(k_g) == len(k_o) and k_g.tobytes() == k_o.tobytes()
        tot["keypoints"] += int(len(k_o))
        if not ok:
            tot["kp_mismatch_cases"] += 1; bad.append("final")
        else:
            tot["desc_mismatches"] += int((d_g != d_o).any(axis=1).sum())
        if bad and len(tot["failures"]) < 30:
            case["bad"] = bad; tot["failures"].append(case)
        tot["cases"] += 1
    finally:
        ctx.close()
tot["seconds"] = round(time.time() - t0, 1)
tot["desc_mismatch_rate"] = tot["desc_mismatches"] / max(1, tot["keypoints"])
print(json.dumps(tot))
