"""oracle/_ref pins the oracle: the REFERENCE's own FtDtOrbSlam / FtAssocOrbSlam / FeatureGrid / FrameMonoGrid, compiled
unchanged from /root/reference (oracle/Makefile.ref; only the OpenCV containers are stand-ins, the pixel primitives are
the oracle's cv2-pinned routines), must agree with the oracle's restatement bit for bit: cell loop, quadtree order incl.
the unstable std::sort, two-ended output order, monoIndex, angles, descriptors, grid queries, matchV.

Runs wherever the library exists: it is built in the build container (where /root/reference is) and travels prebuilt."""
import numpy as np
import pytest

from nav24_b200.synth import synth, sequence
from oracle import orb_oracle as oo
from oracle import ref_lib as rl

pytestmark = pytest.mark.skipif(not rl.available(), reason="oracle/_ref not built and /root/reference absent")

DETECT_CASES = [  # H, W, nFeatures, seed, lowtex
    (480, 752, 1000, 24, False), (480, 752, 5000, 25, False), (480, 752, 200, 26, True),
    (376, 1241, 2000, 24, False), (376, 1241, 2000, 31, True), (376, 1241, 10000, 32, False),
    (480, 640, 1000, 7, True), (260, 340, 300, 5, True), (300, 900, 700, 9, False),
    # widths whose last FAST cell of level 0 is 7 / 8 px wide (a 1- / 2-px interior): the geometry behind the first bug the
    # randomised GPU runs found — pins the ORACLE's treatment of such cells to the reference's own cell loop
    (300, 1149, 2000, 1849, True), (343, 1335, 2000, 2035, True), (310, 1151, 2000, 1851, True), (300, 1149, 2000, 1850, False),
]


def _same_detect(r, o, img):
    mr, kr, dr = r.detect(img)
    mo, ko, do = o.detect(img)
    assert mr == mo
    assert len(kr) == len(ko) and kr.tobytes() == ko.tobytes()
    assert np.array_equal(dr, do)
    return mr, kr, dr


@pytest.mark.parametrize("H,W,nf,seed,low", DETECT_CASES)
def test_detect_reference_equals_oracle(H, W, nf, seed, low):
    img = synth(H, W, seed, lowtex=low)
    r, o = rl.RefOrb(nf), oo.OrbOracle(nf)
    for a, b in zip(r.tables(), o.tables()):
        assert np.array_equal(a, b)
    mono, k, _ = _same_detect(r, o, img)
    for l in range(8):
        assert np.array_equal(r.level(l), o.level(l)), f"pyramid level {l}"
    if W > 1000:
        assert mono > 0          # keypoints beyond x = 1000 fill the output from the front (OP_FtDtOrbSlam.cpp:909-918)
    assert len(k) > 0.8 * nf or low


def test_detect_4k_reference_equals_oracle():
    img = synth(2160, 3840, 24)
    _same_detect(rl.RefOrb(8000), oo.OrbOracle(8000), img)


def test_scale_num_features_modes():
    """The front end's x5 / x0.2 feature-count switch (FE_SlamMonoV.cpp:136,176,247 -> FtDt::scaleNumFeatures)."""
    img = synth(480, 752, 3)
    r, o = rl.RefOrb(1000), oo.OrbOracle(1000)
    for s, want in ((5.0, 5000), (0.2, 200), (1.0, 1000)):
        assert r.scale_num_features(s) == want
        o.set_num_features(want)
        assert np.array_equal(r.tables()[2], o.tables()[2])
        _same_detect(r, o, img)


@pytest.mark.parametrize("scale,nlevels,ini,mn", [(1.1, 6, 20, 7), (1.5, 5, 30, 10), (1.33, 4, 12, 5), (1.8, 3, 20, 7)])
def test_other_parameters(scale, nlevels, ini, mn):
    img = synth(480, 752, 77)
    r, o = rl.RefOrb(800, scale, nlevels, ini, mn), oo.OrbOracle(800, scale, nlevels, ini, mn)
    for a, b in zip(r.tables(), o.tables()):
        assert np.array_equal(a, b)
    _same_detect(r, o, img)


def test_empty_and_flat_images():
    r, o = rl.RefOrb(500), oo.OrbOracle(500)
    assert r.detect(None)[0] == -1                               # OP_FtDtOrbSlam.cpp:851-852
    flat = np.full((300, 400), 77, np.uint8)
    mono, k, d = _same_detect(r, o, flat)
    assert len(k) == 0 and mono == 0
    noise = np.random.default_rng(1).integers(0, 256, (300, 400), dtype=np.uint8)
    _same_detect(r, o, noise)


def test_quadtree_reference_equals_oracle():
    """DistributeOctTree + DivideNode + compareNodes with real std::list / std::sort on both sides: clustered keys,
    duplicated responses (first-wins ties), equal node sizes (the unstable sort decides the expansion order)."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 6000))
        w, h = int(rng.integers(60, 1300)), int(rng.integers(60, 500))
        if round(w / h) < 1:
            continue
        if trial % 3 == 0:      # clusters
            c = rng.integers(0, [w, h], (8, 2))
            xy = (c[rng.integers(0, 8, n)] + rng.integers(-12, 13, (n, 2))).clip(0, [w - 1, h - 1])
        else:
            xy = rng.integers(0, [w, h], (n, 2))
        resp = rng.integers(7, 40 if trial % 2 else 255, n)
        xyr = np.concatenate([xy, resp[:, None]], 1).astype(np.float32)
        N = int(rng.integers(1, 1500))
        a = rl.quadtree(xyr, 16, 16 + w, 16, 16 + h, N)
        b = oo.quadtree(xyr, 16, 16 + w, 16, 16 + h, N)
        assert np.array_equal(a, b), (trial, n, w, h, N)


def _ud(k, shift=0.0):
    return np.stack([k["x"], k["y"]], 1) + np.float32(shift)


def test_matcher_reference_equals_oracle_on_frames():
    for (H, W, step) in [(480, 752, (2, 1)), (376, 1241, (17, 0)), (480, 640, (3, 1))]:
        fr = sequence(H, W, 77, 4, step=step)
        o = oo.OrbOracle(1000)
        dets = [o.detect(f) for f in fr]
        for a, b in [(0, 1), (0, 3), (2, 3), (1, 1)]:
            (_, k1, d1), (_, k2, d2) = dets[a], dets[b]
            for ori in (True, False):
                ref = rl.match_window(k1, _ud(k1), d1, k2, _ud(k2), d2, W, H, check_ori=ori)
                got = oo.match_window(k1, _ud(k1), d1, k2, _ud(k2), d2, oo.grid_for(W, H), check_ori=ori)
                assert np.array_equal(ref, got)
            assert (ref >= 0).sum() > 20
        # undistorted coordinates with image bounds that are not the image rectangle (Calibration::computeImageBounds)
        bounds = (-9.5, W + 7.25, -4.75, H + 11.5)
        (_, k1, d1), (_, k2, d2) = dets[0], dets[1]
        u1 = _ud(k1) * np.float32(1.01) - np.float32(3.0); u2 = _ud(k2) * np.float32(1.01) - np.float32(3.0)
        ref = rl.match_window(k1, u1, d1, k2, u2, d2, W, H, bounds=bounds)
        got = oo.match_window(k1, u1, d1, k2, u2, d2, oo.grid_for(W, H, bounds))
        assert np.array_equal(ref, got) and (ref >= 0).sum() > 20


def test_matcher_reference_equals_oracle_adversarial():
    """Many near-identical descriptors: steals, ties, candidates rejected by vMatchedDistance."""
    rng = np.random.default_rng(3)
    n1, n2 = 600, 900
    base = rng.integers(0, 256, (6, 32), dtype=np.uint8)

    def mk(n):
        d = base[rng.integers(0, 6, n)].copy()
        f = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & \
            rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
        return d ^ f

    def kp(n):
        k = np.zeros(n, oo.KP_DTYPE)
        k["x"] = rng.uniform(0, 300, n).astype(np.float32); k["y"] = rng.uniform(0, 200, n).astype(np.float32)
        k["angle"] = rng.uniform(0, 360, n).astype(np.float32); k["octave"] = rng.integers(0, 3, n) // 2
        k["class_id"] = -1
        return k
    k1, k2, d1, d2 = kp(n1), kp(n2), mk(n1), mk(n2)
    for ratio in (0.6, 0.9, 1.0, 1.5):      # TH_LOW = 50 and the 100-px window are constants of the reference
        for ori in (True, False):
            ref = rl.match_window(k1, _ud(k1), d1, k2, _ud(k2, 0.25), d2, 300, 200, nnratio=ratio, check_ori=ori)
            got = oo.match_window(k1, _ud(k1), d1, k2, _ud(k2, 0.25), d2, oo.grid_for(300, 200), nnratio=ratio, check_ori=ori)
            assert np.array_equal(ref, got), (ratio, ori)
    assert (ref >= 0).sum() > 50


def test_grid_query_reference_equals_oracle():
    """FeatureGrid: cell of a point (round half away), query cell range (floor / ceil), ix-major iteration, strict window."""
    rng = np.random.default_rng(9)
    W, H = 752, 480
    n = 3000
    k = np.zeros(n, oo.KP_DTYPE)
    k["x"] = rng.uniform(-5, W + 5, n).astype(np.float32); k["y"] = rng.uniform(-5, H + 5, n).astype(np.float32)
    k["x"][:200] = np.round(k["x"][:200] / 5) * 5            # points on cell borders and half-cells
    k["y"][:200] = np.round(k["y"][:200] / 5) * 5
    k["octave"] = rng.integers(0, 8, n)
    for bounds in (None, (-12.25, W + 3.5, -2.0, H + 9.75)):
        for _ in range(60):
            x, y = float(rng.uniform(-50, W + 50)), float(rng.uniform(-50, H + 50))
            r = float(rng.choice([5.0, 10.0, 33.3, 100.0]))
            lo, hi = [(0, 0), (-1, -1), (2, 5), (0, -1)][int(rng.integers(0, 4))]
            a = rl.grid_query(k, _ud(k), W, H, x, y, r, lo, hi, bounds=bounds)
            b = oo.grid_query(k, _ud(k), oo.grid_for(W, H, bounds), x, y, r, lo, hi)
            assert np.array_equal(a, b), (x, y, r, lo, hi)


def test_reference_reproduces_the_golden_fixtures():
    """tests/golden/*.npz were frozen from the cv2-driven restatement (tools/gen_golden.py: every pixel primitive is a
    live cv2 4.13 call there): the reference's own code must reproduce them as well — keypoints, descriptors, monoIndex
    and the windowed matches."""
    import glob
    import os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*_n*.npz")))
    assert files
    for f in files:
        g = np.load(f)
        fr = g["frames"]
        H, W = fr.shape[1:]
        r = rl.RefOrb(int(g["n_features"]))
        det = []
        for i in range(len(fr)):
            if f"f{i}_kps" not in g:
                break
            mono, k, d = r.detect(np.ascontiguousarray(fr[i]))
            assert mono == int(g[f"f{i}_mono"]) and k.tobytes() == g[f"f{i}_kps"].tobytes() and np.array_equal(d, g[f"f{i}_desc"]), (f, i)
            det.append((k, d))
        if "matches12" in g:
            (k1, d1), (k2, d2) = det[0], det[1]
            assert np.array_equal(rl.match_window(k1, _ud(k1), d1, k2, _ud(k2), d2, W, H), g["matches12"]), f
            assert np.array_equal(rl.match_window(k1, _ud(k1), d1, k2, _ud(k2), d2, W, H, check_ori=False), g["matches12_noori"]), f
