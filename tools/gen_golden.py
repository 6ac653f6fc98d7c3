#!/usr/bin/env python
"""Generate tests/golden/*.npz — known-answer fixtures for the ORB front end.

The reference holds no golden vectors for this path and cannot be built here,
so the fixtures come from the *real OpenCV* (cv2 4.13.0 in the build container) driven line by line
like core/operators/objDetection/OP_FtDtOrbSlam.cpp (oracle/orb_ref_cv2.py): cv2.resize,
cv2.FastFeatureDetector, cv2.GaussianBlur, cv2.fastAtan2, glibc cosf/sinf.  The quadtree and the
windowed matcher are nav24's own code (no OpenCV inside); their outputs in the fixtures come from the
C++ oracle running the real std::list / std::sort.  cv2 does not travel to the GPU box, the fixtures
do: `-m "not gpu"` tests check the C++ oracle against them, `-m gpu` tests check the CUDA path.

Each fixture stores the INPUT image(s) too, so that nothing depends on a numpy RNG stream.

    python tools/gen_golden.py          # needs cv2; rewrites tests/golden/
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nav24_b200.synth import sequence, synth  # noqa: E402
from oracle import orb_oracle as oo  # noqa: E402
from oracle.orb_ref_cv2 import OrbRefCv2  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name, H, W, nFeatures, seed, lowtex, n_frames, step
CASES = [
    ("euroc_752x480_n1000", 480, 752, 1000, 24, False, 2, (2, 1)),
    ("kitti_1241x376_n2000", 376, 1241, 2000, 24, False, 2, (11, 0)),
    ("tum_640x480_n1000_lowtex", 480, 640, 1000, 7, True, 2, (3, 1)),
    ("small_340x260_n300_lowtex", 260, 340, 300, 5, True, 1, (0, 0)),
    ("euroc_752x480_n5000", 480, 752, 5000, 25, False, 1, (0, 0)),
]


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, H, W, nf, seed, low, nfr, step in CASES:
        frames = sequence(H, W, seed, nfr, step=step, lowtex=low) if nfr > 1 else synth(H, W, seed, lowtex=low)[None]
        d = {"frames": frames, "n_features": np.int32(nf)}
        dets = []
        for f in range(nfr):
            r = OrbRefCv2(nf)
            mono, k, desc = r.detect(frames[f])
            dets.append((k, desc))
            d[f"f{f}_mono"] = np.int32(mono)
            d[f"f{f}_kps"] = k
            d[f"f{f}_desc"] = desc
            d[f"f{f}_level_sha"] = np.stack([sha(r.levels[l]) for l in range(8)])
            d[f"f{f}_blur_sha"] = np.stack([sha(b) if b is not None else np.zeros(32, np.uint8) for b in r.stage["blur"]])
            d[f"f{f}_raw_count"] = np.array([len(x) for x in r.stage["raw"]], np.int32)
            d[f"f{f}_raw_sha"] = np.stack([sha(x) for x in r.stage["raw"]])
            d[f"f{f}_level_count"] = np.array([len(x) for x in r.stage["level_kps"]], np.int32)
            if f == 0:      # one full set of per-stage arrays (small): raw keys of the coarsest two levels
                d["f0_raw_l6"] = r.stage["raw"][6]; d["f0_raw_l7"] = r.stage["raw"][7]
                d["f0_level7"] = np.ascontiguousarray(r.levels[7])
        if nfr > 1:
            (k1, d1), (k2, d2) = dets[0], dets[1]
            ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
            d["matches12"] = oo.match_window(k1, ud1, d1, k2, ud2, d2, oo.grid_for(W, H))
            d["matches12_noori"] = oo.match_window(k1, ud1, d1, k2, ud2, d2, oo.grid_for(W, H), check_ori=False)
            for norm in (0, 1):
                i0, i1, f0, f1, ps = oo.match_bf_knn2(d1[:400], d2[:500], norm, 0.7)
                d[f"bf{norm}_idx"] = np.stack([i0, i1]); d[f"bf{norm}_dist"] = np.stack([f0, f1]); d[f"bf{norm}_pass"] = ps
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name}: {sum(len(k) for k, _ in dets)} keypoints, {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
