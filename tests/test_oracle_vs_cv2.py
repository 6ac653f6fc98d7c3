"""Pin the C++ oracle against the real OpenCV (cv2) primitives and the cv2-driven restatement.

The reference holds no golden vectors for this path; what pins the oracle is
OpenCV itself (the reference delegates all pixel arithmetic to it).  Skipped where cv2 is absent.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import orb_oracle as oo
from nav24_b200.synth import synth


def _rand_img(rng, h, w, kind):
    if kind == 0:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == 1:
        return (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
    return synth(h, w, int(rng.integers(0, 1 << 30)))


@pytest.mark.parametrize("seed", range(4))
def test_resize_matches_cv2(seed):
    rng = np.random.default_rng(seed)
    for _ in range(25):
        sh, sw = int(rng.integers(20, 400)), int(rng.integers(20, 500))
        f = rng.uniform(1.05, 2.3)
        dh, dw = max(2, int(round(sh / f))), max(2, int(round(sw / rng.uniform(1.05, 2.3))))
        src = _rand_img(rng, sh, sw, int(rng.integers(0, 3)))
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(oo.resize_u8(src, dw, dh), ref), (sh, sw, dh, dw)


def test_resize_pyramid_chain_shapes():
    # the 28 level transitions of the four BASELINE shapes
    from oracle.orb_ref_cv2 import OrbRefCv2
    rng = np.random.default_rng(7)
    r = OrbRefCv2(1000)
    for (h, w) in [(480, 752), (376, 1241), (480, 640), (1080, 1920)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        r.pyramid(img)
        for l in range(1, 8):
            hh, ww = r.levels[l].shape
            assert np.array_equal(oo.resize_u8(r.levels[l - 1], ww, hh), r.levels[l]), (h, w, l)


@pytest.mark.parametrize("seed", range(3))
def test_gauss7_matches_cv2(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(12):
        h, w = int(rng.integers(8, 300)), int(rng.integers(8, 400))
        src = _rand_img(rng, h, w, int(rng.integers(0, 3)))
        ref = cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(oo.gauss7_u8(src), ref), (h, w)


@pytest.mark.parametrize("t", [7, 20, 1, 40])
def test_fast_matches_cv2(t):
    rng = np.random.default_rng(200 + t)
    det = cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    total = 0
    for it in range(30):
        h, w = int(rng.integers(7, 80)), int(rng.integers(7, 90))
        kind = int(rng.integers(0, 3))
        big = _rand_img(rng, h + 10, w + 12, kind)
        if kind == 0 and it % 2:
            big = (big // 16 * 16).astype(np.uint8)     # plateaus: ties in the NMS
        sub = big[5:5 + h, 6:6 + w]                       # non-contiguous ROI view like the cell loop
        kps = det.detect(sub)
        ref = np.array([(k.pt[0], k.pt[1], k.response) for k in kps], np.float32).reshape(-1, 3)
        got = oo.fast_u8(sub, t)
        assert got.shape == ref.shape and np.array_equal(got, ref), (h, w, kind)
        total += len(ref)
    assert total > 50


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(5)
    ys = rng.integers(-3000000, 3000000, 20000); xs = rng.integers(-3000000, 3000000, 20000)
    ys[:50] = 0; xs[50:100] = 0; ys[100:150] = xs[100:150]
    for y, x in zip(ys, xs):
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(oo.fast_atan2(y, x)), (y, x)


@pytest.mark.parametrize("norm", [0, 1])
def test_bf_knn2_matches_cv2(norm):
    rng = np.random.default_rng(9)
    d2 = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    d1 = d2[rng.integers(0, 300, 200)].copy()
    flip = rng.integers(0, 256, d1.shape, dtype=np.uint8) & rng.integers(0, 256, d1.shape, dtype=np.uint8) & \
        rng.integers(0, 256, d1.shape, dtype=np.uint8)
    d1 ^= flip
    d2[10] = d2[11]                                      # exact duplicate: lowest index wins ties
    d1[0] = d2[10]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING if norm == 0 else cv2.NORM_L2)
    knn = bf.knnMatch(d1, d2, k=2)
    i0, i1, f0, f1, ps = oo.match_bf_knn2(d1, d2, norm, 0.7)
    for q, m in enumerate(knn):
        assert (m[0].trainIdx, m[1].trainIdx) == (i0[q], i1[q])
        assert np.float32(m[0].distance) == f0[q] and np.float32(m[1].distance) == f1[q]
        assert bool(ps[q]) == bool(m[0].distance < 0.7 * m[1].distance)


@pytest.mark.parametrize("H,W,nf,seed,low", [(376, 1241, 2000, 24, False), (480, 752, 1000, 25, True),
                                            (480, 640, 5000, 3, False), (260, 340, 300, 5, True)])
def test_detect_matches_cv2_restatement(H, W, nf, seed, low):
    from oracle.orb_ref_cv2 import OrbRefCv2
    img = synth(H, W, seed, lowtex=low)
    o = oo.OrbOracle(nf); r = OrbRefCv2(nf)
    m1, k1, d1 = o.detect(img); m2, k2, d2 = r.detect(img)
    assert r.quota == list(o.tables()[2])
    assert np.array_equal(np.array(r.scale, np.float32), o.tables()[0])
    for l in range(8):
        assert np.array_equal(o.level(l), r.levels[l])
        assert np.array_equal(o.raw(l), r.stage["raw"][l])
        b = o.blurred(l)
        if b is not None:
            assert np.array_equal(b, r.stage["blur"][l])
    assert m1 == m2 and len(k1) == len(k2)
    assert k1.tobytes() == k2.tobytes()
    assert np.array_equal(d1, d2)
