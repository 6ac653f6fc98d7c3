"""Calibration::undistort: the step between detect and matchV.

The reference delegates to cv::undistortPoints / cv::fisheye::undistortPoints (models/PinholeRadTan.cpp:20,
models/KannalaBrandt8.cpp:238).  tests/golden/undistort.npz holds those cv2 4.13.0 calls on stored inputs
(tools/gen_golden_undistort.py).  CPU: the oracle against the fixtures (and against cv2 itself where importable).
GPU: the CUDA path through the C ABI against the fixtures and the oracle.

Bar: RadTan bit-exact (IEEE double +,-,*,/ only).  KB8 ends in tan(): CUDA's and glibc's double tan may differ in
the last bit, which survives the rounding to float only on a float rounding boundary, so KB8 allows 1 float ulp
on at most 0.1 % of the points (observed: 0) — KB8_ULP / KB8_FRAC below.
"""
import os

import numpy as np
import pytest

from oracle import orb_oracle as oo

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "undistort.npz"))
NAMES = [str(n) for n in G["names"]]
KB8_ULP, KB8_FRAC = 1, 1e-3


def _ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64); bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7fffffff), ai); bi = np.where(bi < 0, -(bi & 0x7fffffff), bi)
    return np.abs(ai - bi)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_cv2_fixture(name):
    model = int(G[name + "_model"])
    ud = oo.undistort(model, G[name + "_K"], G[name + "_D"], G[name + "_xy"])
    assert ud.tobytes() == G[name + "_ud"].tobytes()      # bit-exact, both models (same libm as the fixture run)


def test_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for name in NAMES:
        model = int(G[name + "_model"]); k = G[name + "_K"]; d = G[name + "_D"]; W, H = G[name + "_wh"]
        pts = np.stack([rng.uniform(-40, W + 40, 3000), rng.uniform(-40, H + 40, 3000)], 1).astype(np.float32)
        K = np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]], np.float32)
        eye = np.eye(3, dtype=np.float32)
        fn = cv2.undistortPoints if model == 1 else cv2.fisheye.undistortPoints
        ref = fn(pts.reshape(-1, 1, 2), K, d.reshape(4, 1), R=eye, P=K).reshape(-1, 2)
        assert oo.undistort(model, k, d, pts).tobytes() == ref.tobytes(), name


def test_pinhole_is_identity_and_bounds():
    from nav24_b200 import capi
    xy = G["euroc_radtan_xy"]
    assert np.array_equal(oo.undistort(oo.CAM_PINHOLE, None, None, xy), xy)
    k, d = G["euroc_radtan_K"], G["euroc_radtan_D"]
    b = capi.image_bounds(lambda p: oo.undistort(1, k, d, p), 752, 480, calibrated=False)
    assert np.allclose(b, G["match_bounds"], rtol=0, atol=0)
    assert capi.image_bounds(None, 752, 480, calibrated=True) == (0.0, 752.0, 0.0, 480.0)


def test_oracle_match_on_undistorted_fixture():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "euroc_752x480_n1000.npz"))
    k1, k2 = g["f0_kps"], g["f1_kps"]
    k, d = G["euroc_radtan_K"], G["euroc_radtan_D"]
    ud1 = oo.undistort(1, k, d, np.stack([k1["x"], k1["y"]], 1)); ud2 = oo.undistort(1, k, d, np.stack([k2["x"], k2["y"]], 1))
    assert ud1.tobytes() == G["match_ud1"].tobytes() and ud2.tobytes() == G["match_ud2"].tobytes()
    grid = oo.grid_for(752, 480, tuple(float(x) for x in G["match_bounds"]))
    m = oo.match_window(k1, ud1, g["f0_desc"], k2, ud2, g["f1_desc"], grid)
    assert np.array_equal(m, G["match_matches12"]) and (m >= 0).sum() > 20
    # the distortion matters: the same frames matched on the detected coordinates give a different answer
    assert not np.array_equal(m, g["matches12"])


# ---- GPU ---------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_undistort_points(name, cuda_required):
    from nav24_b200 import capi
    model = int(G[name + "_model"])
    ctx = capi.OrbContext(300)
    try:
        cam = capi.Camera.make(model, G[name + "_K"], G[name + "_D"])
        ud = ctx.undistort_points(cam, G[name + "_xy"])
        ref = G[name + "_ud"]
        if model == capi.CAM_RADTAN:
            assert ud.tobytes() == ref.tobytes()
        else:
            u = _ulp_diff(ud, ref)
            print(f"{name}: {int((u > 0).sum())} of {u.size} coordinates differ by 1 ulp")
            assert u.max() <= KB8_ULP and (u > 0).mean() <= KB8_FRAC
        # in place + identity model
        assert np.array_equal(ctx.undistort_points(capi.Camera.make(capi.CAM_PINHOLE), G[name + "_xy"]), G[name + "_xy"])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_fused_detect_undistort_match(cuda_required):
    """detect -> undistort -> matchV with the frame staying on the device (FE_SlamMonoV.cpp:104-122)."""
    from nav24_b200 import capi
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "euroc_752x480_n1000.npz"))
    frames = np.ascontiguousarray(g["frames"])
    k, d = G["euroc_radtan_K"], G["euroc_radtan_D"]
    ctx = capi.OrbContext(int(g["n_features"]))
    try:
        cam = capi.Camera.make(capi.CAM_RADTAN, k, d)
        ctx.set_camera(cam)
        b = capi.image_bounds(lambda p: ctx.undistort_points(cam, p), 752, 480, calibrated=False)
        assert np.array_equal(np.array(b, np.float32), G["match_bounds"])
        grid = capi.grid_for(752, 480, b)
        n, mono, kps, desc, m, nm = ctx.detect_match_batch(frames, [(0, 1)], grid)
        assert kps[0, :n[0]].tobytes() == g["f0_kps"].tobytes() and kps[1, :n[1]].tobytes() == g["f1_kps"].tobytes()
        ud = ctx.fetch_undistorted(2)
        assert ud[0, :n[0]].tobytes() == G["match_ud1"].tobytes() and ud[1, :n[1]].tobytes() == G["match_ud2"].tobytes()
        exact = np.array_equal(desc[0, :n[0]], g["f0_desc"]) and np.array_equal(desc[1, :n[1]], g["f1_desc"])

        def want(frozen, ud1, ud2, ogrid):      # unconditional: the frozen answer, or the oracle on the GPU's own descriptors
            if exact:
                return frozen
            return oo.match_window(kps[0, :n[0]], ud1, desc[0, :n[0]], kps[1, :n[1]], ud2, desc[1, :n[1]], ogrid)
        assert np.array_equal(m[0, :n[0]], want(G["match_matches12"], ud[0, :n[0]], ud[1, :n[1]],
                                                oo.grid_for(752, 480, tuple(float(x) for x in b))))
        # the separate device-resident matcher sees the same coordinates
        m2, _ = ctx.match_window_frames([(0, 1)], grid)
        assert np.array_equal(m2[0, :n[0]], m[0, :n[0]])
        # back to pinhole: the plain fixture answer
        ctx.set_camera(None)
        n, mono, kps, desc, m, nm = ctx.detect_match_batch(frames, [(0, 1)], capi.grid_for(752, 480))
        xy = [np.stack([kps[f, :n[f]]["x"], kps[f, :n[f]]["y"]], 1) for f in range(2)]
        assert np.array_equal(m[0, :n[0]], want(g["matches12"], xy[0], xy[1], oo.grid_for(752, 480)))
    finally:
        ctx.close()
