#!/usr/bin/env python
"""tests/golden/ingest_bgr.npz: interleaved BGR frames and the grey images live cv2 makes of them
(cv2.cvtColor COLOR_BGR2GRAY, the call at core/frontEnd/FE_SlamMonoV.cpp:92-94).  Run in the build container (cv2 4.13)."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav24_b200.synth import canvas, frame_from_canvas  # noqa: E402

H, W = 260, 341          # 1023-byte BGR rows: every 4-byte alignment of a row start occurs


def colour_frame(seed, shift):
    """Three differently textured channels of one scene: a colour image whose grey version still has corners."""
    ch = [frame_from_canvas(canvas(H, W, seed + 17 * c), H, W, shift, noise_seed=seed * 31 + c) for c in range(3)]
    base = ch[0].astype(np.int32)
    bgr = np.stack([np.clip(base + (ch[1].astype(np.int32) - 128) // 2, 0, 255),
                    base,
                    np.clip(base - (ch[2].astype(np.int32) - 128) // 2, 0, 255)], axis=2).astype(np.uint8)
    return bgr


def main():
    frames = np.stack([colour_frame(5, (0, 0)), colour_frame(5, (4, 1))])
    extremes = np.zeros((1, H, W, 3), np.uint8)          # saturated colours and the rounding boundary cases
    rng = np.random.default_rng(3)
    extremes[0] = rng.choice(np.array([0, 1, 127, 128, 254, 255], np.uint8), size=(H, W, 3))
    bgr = np.concatenate([frames, extremes])
    gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in bgr])
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ingest_bgr.npz")
    np.savez_compressed(out, bgr=bgr, gray=gray, cv2_version=cv2.__version__)
    print(out, bgr.shape, gray.shape)


if __name__ == "__main__":
    main()
