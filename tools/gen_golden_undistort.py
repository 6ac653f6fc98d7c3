#!/usr/bin/env python
"""Generate tests/golden/undistort.npz — known answers for Calibration::undistort.

The reference delegates to OpenCV (cv::undistortPoints, models/PinholeRadTan.cpp:20; cv::fisheye::undistortPoints,
models/KannalaBrandt8.cpp:238) with K float, D 4 floats, R = I, P = K (models/GeometricCamera.h:62-66); the fixtures
are those very cv2 4.13.0 calls on stored inputs.  The matching case undistorts the keypoints of the EuRoC-shaped
fixture with a RadTan camera, takes the grid bounds from the undistorted image corners like
Calibration::computeImageBounds (Calibration.cpp:196-228) and stores the oracle's matchV result on them.

    python tools/gen_golden_undistort.py      # needs cv2; rewrites tests/golden/undistort.npz
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orb_oracle as oo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "undistort.npz")

# name, model, K4 (fx fy cx cy), D4, W, H
CAMS = [
    ("euroc_radtan", 1, [458.654, 457.296, 367.215, 248.375], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], 752, 480),
    ("kitti_radtan", 1, [718.856, 718.856, 607.1928, 185.2157], [-0.1, 0.05, 0.001, -0.002], 1241, 376),
    ("tum_radtan_strong", 1, [517.3, 516.5, 318.6, 255.3], [0.2624, -0.9531, -0.0054, 0.0026], 640, 480),
    ("tumvi_kb8", 2, [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504],
     [0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182], 512, 512),
    ("wide_kb8", 2, [380.0, 381.0, 320.0, 240.0], [-0.04, 0.01, -0.003, 0.0005], 640, 480),
]


def K33(k):
    return np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]], np.float32)


def cv_undistort(model, k4, d4, pts):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2)
    D = np.array(d4, np.float32).reshape(4, 1)
    eye = np.eye(3, dtype=np.float32)
    if model == 1:
        return cv2.undistortPoints(pts, K33(k4), D, R=eye, P=K33(k4)).reshape(-1, 2)
    return cv2.fisheye.undistortPoints(pts, K33(k4), D, R=eye, P=K33(k4)).reshape(-1, 2)


def main():
    rng = np.random.default_rng(11)
    d = {"names": np.array([c[0] for c in CAMS])}
    for name, model, k4, d4, W, H in CAMS:
        pts = np.stack([rng.uniform(-30, W + 30, 4000), rng.uniform(-30, H + 30, 4000)], 1).astype(np.float32)
        pts[:4] = [[0, 0], [W, 0], [0, H], [W, H]]
        pts[4] = [k4[2], k4[3]]                      # the principal point
        pts[5:2005] = np.rint(pts[5:2005])           # integral coordinates like level-0 keypoints
        d[name + "_model"] = np.int32(model)
        d[name + "_K"] = np.array(k4, np.float32); d[name + "_D"] = np.array(d4, np.float32)
        d[name + "_wh"] = np.array([W, H], np.int32)
        d[name + "_xy"] = pts
        d[name + "_ud"] = cv_undistort(model, k4, d4, pts)
    # matchV on undistorted coordinates: EuRoC-shaped frames, RadTan camera
    g = np.load(os.path.join(ROOT, "tests", "golden", "euroc_752x480_n1000.npz"))
    name, model, k4, d4, W, H = CAMS[0]
    k1, k2 = g["f0_kps"], g["f1_kps"]
    ud1 = cv_undistort(model, k4, d4, np.stack([k1["x"], k1["y"]], 1))
    ud2 = cv_undistort(model, k4, d4, np.stack([k2["x"], k2["y"]], 1))
    c = cv_undistort(model, k4, d4, np.array([[0, 0], [W, 0], [0, H], [W, H]], np.float32)).reshape(-1)
    bounds = np.array([min(c[0], c[4]), max(c[2], c[6]), min(c[1], c[3]), max(c[5], c[7])], np.float32)
    grid = oo.grid_for(W, H, tuple(float(b) for b in bounds))
    d["match_bounds"] = bounds
    d["match_ud1"] = ud1; d["match_ud2"] = ud2
    d["match_matches12"] = oo.match_window(k1, ud1, g["f0_desc"], k2, ud2, g["f1_desc"], grid)
    np.savez_compressed(OUT, **d)
    print(f"undistort.npz: {len(CAMS)} cameras, {int((d['match_matches12'] >= 0).sum())} matches on undistorted coordinates, "
          f"{os.path.getsize(OUT) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
