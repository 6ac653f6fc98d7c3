"""SURVEY 8(f)-3 image ingest: the pinned slot ring + device-side BGR -> grey conversion.

The conversion must equal cv::cvtColor(COLOR_BGR2GRAY) (FE_SlamMonoV.cpp:92-94) bit for bit: the oracle restatement is
pinned against live cv2 where it imports and against tests/golden/ingest_bgr.npz (made by tools/gen_golden_ingest.py with
cv2 4.13) everywhere; the CUDA path is compared with both, and the keypoints / matches of colour frames must be those of
the detector run on the grey frames."""
import os

import numpy as np
import pytest

from oracle import orb_oracle as oo

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest_bgr.npz"))


def test_oracle_bgr2gray_matches_golden():
    for bgr, gray in zip(G["bgr"], G["gray"]):
        assert np.array_equal(oo.bgr2gray(bgr), gray)


def test_oracle_bgr2gray_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    for shape in ((97, 133), (260, 341), (376, 1241)):
        img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        assert np.array_equal(oo.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    # every (B, G, R) with two channels on a coarse lattice and one channel exhaustive: the rounding of the 15-bit fixed point
    v = np.arange(256, dtype=np.uint8)
    lat = np.array([0, 1, 2, 63, 64, 127, 128, 129, 200, 254, 255], np.uint8)
    for perm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
        grid = np.stack(np.meshgrid(v, lat, lat, indexing="ij"), -1).reshape(256, -1, 3)[..., list(perm)]
        grid = np.ascontiguousarray(grid)
        assert np.array_equal(oo.bgr2gray(grid), cv2.cvtColor(grid, cv2.COLOR_BGR2GRAY))


@pytest.mark.gpu
def test_ring_colour_frames(cuda_required):
    from nav24_b200 import capi
    bgr, gray = G["bgr"], G["gray"]
    n_fr, H, W = gray.shape
    nf = 300
    ctx = capi.OrbContext(nf)
    try:
        ring = capi.IngestRing(ctx, W, H, 3, n_fr + 2)
        for k in range(n_fr):
            ring.slot(k + 1)[...] = bgr[k]              # the "decoder" writes straight into the pinned slots
        grid = capi.grid_for(W, H)
        n, mono, kps, desc, m, nm = ring.detect_match(1, n_fr, [(0, 1)], grid)
        o = oo.OrbOracle(nf)
        bad = 0
        for f in range(n_fr):
            assert np.array_equal(ctx.level(f, 0), gray[f]), f"grey level 0 of frame {f} differs from cv2.cvtColor"
            mo, ko, do = o.detect(np.ascontiguousarray(gray[f]))
            assert mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()
            bad += int((desc[f, :n[f]] != do).any(axis=1).sum())
        assert bad <= 1e-3 * int(n.sum())
        k1, k2 = kps[0, :n[0]], kps[1, :n[1]]
        want = oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), desc[0, :n[0]], k2, np.stack([k2["x"], k2["y"]], 1),
                               desc[1, :n[1]], oo.grid_for(W, H))
        assert np.array_equal(m[0, :n[0]], want) and nm[0] == (want >= 0).sum() and nm[0] > 10
        # the same frames as grey through nav24_orb_detect_batch: identical results
        n2, mono2, kps2, desc2 = ctx.detect_batch(np.ascontiguousarray(gray))
        assert np.array_equal(n2, n) and np.array_equal(mono2, mono)
        for f in range(n_fr):
            assert kps2[f, :n[f]].tobytes() == kps[f, :n[f]].tobytes() and np.array_equal(desc2[f, :n[f]], desc[f, :n[f]])
        # one colour frame at a time (the camera's per-frame call; CUDA-graph replay path)
        for f in range(n_fr):
            nn, mm, kk, dd, _, _ = ring.detect_match(1 + f, 1)
            assert nn[0] == n[f] and kk[0, :nn[0]].tobytes() == kps[f, :n[f]].tobytes() and np.array_equal(dd[0, :nn[0]], desc[f, :n[f]])
        ring.close()
    finally:
        ctx.close()


@pytest.mark.gpu
def test_ring_grey_frames_and_errors(cuda_required):
    from nav24_b200 import capi
    from nav24_b200.synth import sequence
    H, W, nf = 376, 1241, 2000
    fr = sequence(H, W, 12, 4, step=(7, 0))
    ctx = capi.OrbContext(nf)
    try:
        ring = capi.IngestRing(ctx, W, H, 1, 4)
        for k in range(4):
            ring.slot(k)[...] = fr[k]
        n, mono, kps, desc, m, nm = ring.detect_match(0, 4, [(0, 1), (2, 3)], capi.grid_for(W, H))
        n2, mono2, kps2, desc2, m2, nm2 = ctx.detect_match_batch(fr, [(0, 1), (2, 3)], capi.grid_for(W, H))
        assert np.array_equal(n, n2) and np.array_equal(mono, mono2) and np.array_equal(nm, nm2)
        for f in range(4):
            assert kps[f, :n[f]].tobytes() == kps2[f, :n[f]].tobytes() and np.array_equal(desc[f, :n[f]], desc2[f, :n[f]])
        for q, a in enumerate((0, 2)):
            assert np.array_equal(m[q, :n[a]], m2[q, :n[a]])
        with pytest.raises(capi.Nav24Error) as e:
            ring.detect_match(3, 2)                       # slots 3..4 of a 4-slot ring
        assert e.value.code == capi.E_BADARG
        with pytest.raises(IndexError):
            ring.slot(4)
        ring.close()
        with pytest.raises(capi.Nav24Error) as e:
            capi.IngestRing(ctx, W, H, 2, 4)              # neither grey nor BGR
        assert e.value.code == capi.E_BADARG
    finally:
        ctx.close()
