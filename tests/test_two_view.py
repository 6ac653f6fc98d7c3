"""two-view RANSAC scoring (TwoViewReconstruction::CheckHomography / CheckFundamental,
core/operators/mapInit/OP_2ViewReconstruction.cpp:447-610) for all hypotheses in one launch.

Oracle: oracle/orb_oracle.cpp restates both functions in plain IEEE float (-ffp-contract=off).  It is pinned to the
REFERENCE'S OWN TwoViewReconstruction, compiled unchanged into oracle/_ref (oracle/Makefile.ref, ref_driver_2v.cpp):
CheckHomography / CheckFundamental are plain float code that read their matrices with at<float>(), so scores and inlier
flags must agree bit for bit on any hypothesis — perturbed ground truth or the reference's own ComputeH21 / ComputeF21
output — and the iteration FindHomography / FindFundamental keep must be the one the `currentScore > score` loop over
those scores keeps.  (cv::SVD and the cv::Mat algebra of that build are this repo's stand-ins, so the 8-point solvers'
last bits are NOT pinned to OpenCV; the solvers stay on the host and are not part of the product.)  A float64 numpy
evaluation of the same formulas bounds the oracle as well.  The CUDA path must equal the oracle AND the reference build
BIT FOR BIT: scores are sequential float sums in match order and the kernel keeps that order."""
import numpy as np
import pytest

from oracle import orb_oracle as oo


def scene(seed, n=400, planar=False, outliers=0.3):
    """Matched points of two views of a random 3-D scene (pixels, float32), 200 homography hypotheses (H21, H12) and 200
    fundamental-matrix hypotheses, perturbed around the true ones."""
    rng = np.random.default_rng(seed)
    K = np.array([[458.0, 0, 367.0], [0, 457.0, 248.0], [0, 0, 1]])
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), (np.full(n, 5.0) if planar else rng.uniform(3, 9, n))]
    ang = 0.05
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.3, 0.02, 0.05])
    x1 = (K @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
    x2 = (K @ (R @ X.T + t[:, None])).T; x2 = x2[:, :2] / x2[:, 2:]
    x1 += rng.normal(0, 0.5, x1.shape); x2 += rng.normal(0, 0.5, x2.shape)
    bad = rng.random(n) < outliers
    x2[bad] = rng.uniform(0, [752, 480], (bad.sum(), 2))
    # true homography of the plane z = 5 and the true fundamental matrix
    nrm = np.array([0, 0, 1.0]); d = 5.0
    Ht = K @ (R + np.outer(t, nrm) / d) @ np.linalg.inv(K)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Ft = np.linalg.inv(K).T @ tx @ R @ np.linalg.inv(K)
    H21, H12, F21 = [], [], []
    for i in range(200):
        s = 0.0 if i == 17 else rng.choice([1e-4, 1e-3, 1e-2, 1e-1])
        H = (Ht / Ht[2, 2]) * (1 + s * rng.normal(0, 1, (3, 3)))
        H21.append(H.astype(np.float32)); H12.append(np.linalg.inv(H.astype(np.float32)).astype(np.float32))
        F = (Ft / np.abs(Ft).max()) * (1 + s * rng.normal(0, 1, (3, 3)))
        F21.append(F.astype(np.float32))
    return (x1.astype(np.float32), x2.astype(np.float32), np.stack(H21).reshape(-1, 9), np.stack(H12).reshape(-1, 9),
            np.stack(F21).reshape(-1, 9))


def test_oracle_against_float64():
    """The float restatement against the same formulas in float64 numpy: scores agree to float rounding, inlier flags may
    differ only where a chi-square sits on the threshold."""
    x1, x2, H21, H12, F21 = scene(1)
    u1, v1, u2, v2 = [a.astype(np.float64) for a in (x1[:, 0], x1[:, 1], x2[:, 0], x2[:, 1])]
    for h in (17, 3, 120):
        s, inl = oo.check_homography(H21[h], H12[h], x1, x2)
        A, Ai = H21[h].astype(np.float64).reshape(3, 3), H12[h].astype(np.float64).reshape(3, 3)
        w = Ai[2, 0] * u2 + Ai[2, 1] * v2 + Ai[2, 2]
        c1 = (u1 - (Ai[0, 0] * u2 + Ai[0, 1] * v2 + Ai[0, 2]) / w) ** 2 + (v1 - (Ai[1, 0] * u2 + Ai[1, 1] * v2 + Ai[1, 2]) / w) ** 2
        w = A[2, 0] * u1 + A[2, 1] * v1 + A[2, 2]
        c2 = (u2 - (A[0, 0] * u1 + A[0, 1] * v1 + A[0, 2]) / w) ** 2 + (v2 - (A[1, 0] * u1 + A[1, 1] * v1 + A[1, 2]) / w) ** 2
        ref = np.where(c1 <= 5.991, 5.991 - c1, 0).sum() + np.where(c2 <= 5.991, 5.991 - c2, 0).sum()
        assert abs(float(s) - ref) <= 2e-3 * max(1.0, ref)
        assert (inl != ((c1 <= 5.991) & (c2 <= 5.991))).sum() <= 2
        s, inl = oo.check_fundamental(F21[h], x1, x2)
        F = F21[h].astype(np.float64).reshape(3, 3)
        a2, b2, c2_ = [F[r, 0] * u1 + F[r, 1] * v1 + F[r, 2] for r in range(3)]
        d1 = (a2 * u2 + b2 * v2 + c2_) ** 2 / (a2 * a2 + b2 * b2)
        a1, b1, c1_ = [F[0, c] * u2 + F[1, c] * v2 + F[2, c] for c in range(3)]
        d2 = (a1 * u1 + b1 * v1 + c1_) ** 2 / (a1 * a1 + b1 * b1)
        ref = np.where(d1 <= 3.841, 5.991 - d1, 0).sum() + np.where(d2 <= 3.841, 5.991 - d2, 0).sum()
        assert abs(float(s) - ref) <= 2e-3 * max(1.0, ref)
        assert (inl != ((d1 <= 3.841) & (d2 <= 3.841))).sum() <= 2
    # the unperturbed hypothesis of the planar scene explains most matches
    x1, x2, H21, H12, F21 = scene(2, planar=True)
    s, inl = oo.check_homography(H21[17], H12[17], x1, x2)
    assert inl.sum() > 0.5 * len(x1)


def keep_loop(sc):
    """The reference's `if (currentScore > score)` selection (OP_2ViewReconstruction.cpp:307, :358): first strict maximum
    above 0; NaN scores are never kept."""
    b, best = -1, np.float32(0)
    for i, v in enumerate(sc):
        if v > best:
            b, best = i, v
    return b


def ref_session(seed, n, planar, unmatched=0.0, sigma=1.0):
    """The reference's TwoViewReconstruction after Reconstruct() on a scene: (session, xy1, xy2 in match order)."""
    from oracle import ref_lib as rl
    x1, x2, H21, H12, F21 = scene(seed, n, planar)
    rng = np.random.default_rng(seed + 100)
    # the second view holds extra, shuffled keypoints; some of the first view's stay unmatched (-1)
    perm = rng.permutation(n + 7)
    k2 = np.zeros((n + 7, 2), np.float32)
    k2[perm[:n]] = x2
    k2[perm[n:]] = rng.uniform(0, 480, (7, 2))
    m12 = perm[:n].astype(np.int32)
    m12[rng.random(n) < unmatched] = -1
    t = rl.RefTwoView(sigma=sigma)
    t.reconstruct(x1, k2, m12)
    a, b, sets = t.matches()
    keep = m12 >= 0
    assert t.n == keep.sum() and np.array_equal(a, x1[keep]) and np.array_equal(b, x2[keep])
    assert all(len(set(r)) == 8 for r in sets.tolist()) and sets.min() >= 0 and sets.max() < t.n
    return t, a, b, (H21, H12, F21)


ref_required = pytest.mark.skipif("not __import__('oracle.ref_lib', fromlist=['x']).available()",
                                  reason="oracle/_ref is not built and /root/reference is not here")


@ref_required
@pytest.mark.parametrize("seed,n,planar,unmatched,sigma", [(1, 400, False, 0.0, 1.0), (2, 400, True, 0.2, 1.0), (3, 8, False, 0.0, 1.0),
                                                           (4, 2500, False, 0.1, 1.0), (5, 300, True, 0.0, 2.5), (6, 64, False, 0.5, 0.7)])
def test_oracle_equals_reference_build(seed, n, planar, unmatched, sigma):
    """oracle == the reference's own CheckHomography / CheckFundamental, bit for bit, on perturbed ground-truth hypotheses
    and on the hypotheses of the reference's own 8-point solvers; FindHomography / FindFundamental keep the iteration the
    selection loop over the oracle's scores keeps."""
    t, a, b, (H21, H12, F21) = ref_session(seed, n, planar, unmatched, sigma)
    Hs, His, Fs = t.hypotheses()
    for tag, (hh, hi, ff) in (("perturbed", (H21, H12, F21)), ("solver", (Hs, His, Fs))):
        sh = np.zeros(200, np.float32); sf = np.zeros(200, np.float32)
        for h in range(200):
            sr, ir = t.check_homography(hh[h], hi[h])
            sh[h], io = oo.check_homography(hh[h], hi[h], a, b, sigma=sigma)
            assert sr.tobytes() == sh[h].tobytes() and np.array_equal(ir, io), f"{tag} homography {h}"
            sr, ir = t.check_fundamental(ff[h])
            sf[h], io = oo.check_fundamental(ff[h], a, b, sigma=sigma)
            assert sr.tobytes() == sf[h].tobytes() and np.array_equal(ir, io), f"{tag} fundamental {h}"
    # sh / sf now hold the solver hypotheses' scores: the reference's RANSAC loops must keep the same iteration
    for find, sc, hyp, check in ((t.find_homography, sh, Hs, lambda i: oo.check_homography(Hs[i], His[i], a, b, sigma=sigma)),
                                 (t.find_fundamental, sf, Fs, lambda i: oo.check_fundamental(Fs[i], a, b, sigma=sigma))):
        score, inl, M = find()
        i = keep_loop(sc)
        if i < 0:
            assert score == 0 and not inl.any()
        else:
            assert score.tobytes() == sc[i].tobytes() and np.array_equal(M, hyp[i]) and np.array_equal(inl, check(i)[1])


@ref_required
def test_reference_build_solvers_are_sane():
    """The stand-in cv::SVD / cv::Mat algebra of the reference build (oracle/ref_shim/opencv2/core_algebra.hpp) is not part
    of what is pinned, but the hypotheses it produces must be real ones: on exact correspondences of a plane the
    reference's FindHomography explains every match, on exact correspondences of a 3-D scene its FindFundamental does,
    and H21 * H12 of every iteration is the identity."""
    from oracle import ref_lib as rl
    rng = np.random.default_rng(3)
    K = np.array([[458.0, 0, 367.0], [0, 457.0, 248.0], [0, 0, 1]])
    ang = 0.05
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]); tv = np.array([0.3, 0.02, 0.05])
    for planar in (True, False):
        n = 300
        X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), (np.full(n, 5.0) if planar else rng.uniform(3, 9, n))]
        x1 = (K @ X.T).T; x1 = (x1[:, :2] / x1[:, 2:]).astype(np.float32)
        x2 = (K @ (R @ X.T + tv[:, None])).T; x2 = (x2[:, :2] / x2[:, 2:]).astype(np.float32)
        t = rl.RefTwoView()
        t.reconstruct(x1, x2, np.arange(n, dtype=np.int32))
        Hs, His, Fs = t.hypotheses()
        if planar:
            score, inl, H = t.find_homography()
            assert inl.all() and score > 0.95 * 2 * 5.991 * n
            P = np.einsum("nij,njk->nik", Hs.reshape(-1, 3, 3).astype(np.float64), His.reshape(-1, 3, 3).astype(np.float64))
            assert np.abs(P - np.eye(3)).max() < 1e-3
            p = (H.reshape(3, 3).astype(np.float64) @ np.c_[x1, np.ones(n)].T).T
            assert np.abs(p[:, :2] / p[:, 2:] - x2).max() < 0.05                 # pixels
        else:
            score, inl, F = t.find_fundamental()
            assert inl.all() and score > 0.95 * 2 * 5.991 * n
            F = F.reshape(3, 3).astype(np.float64)
            assert abs(np.linalg.det(F / np.abs(F).max())) < 1e-6                # rank 2 (:440-444)
            e = np.einsum("ni,ij,nj->n", np.c_[x2, np.ones(n)], F, np.c_[x1, np.ones(n)])
            assert np.abs(e).max() / np.abs(F).max() < 0.5


@ref_required
def test_oracle_equals_reference_build_degenerate():
    """All-zero and non-finite hypotheses: 0/0 makes every chi-square NaN, `NaN > th` is false, so both sides count every
    match as an inlier of a NaN score — and the selection loop never keeps it."""
    t, a, b, _ = ref_session(7, 60, False)
    Z = np.zeros(9, np.float32)
    for M in (Z, np.full(9, np.inf, np.float32), np.full(9, np.nan, np.float32)):
        sr, ir = t.check_homography(M, M); so, io = oo.check_homography(M, M, a, b)
        assert np.isnan(sr) and np.isnan(so) and np.array_equal(ir, io)
        sr, ir = t.check_fundamental(M); so, io = oo.check_fundamental(M, a, b)
        assert np.isnan(sr) and np.isnan(so) and np.array_equal(ir, io)
    assert keep_loop(np.array([np.nan, 1.0, np.nan, 0.5], np.float32)) == 1
    # a hypothesis that explains nothing: finite, score 0, no inliers
    H = np.array([1, 0, 1e4, 0, 1, 1e4, 0, 0, 1], np.float32); Hi = np.array([1, 0, -1e4, 0, 1, -1e4, 0, 0, 1], np.float32)
    sr, ir = t.check_homography(H, Hi); so, io = oo.check_homography(H, Hi, a, b)
    assert sr == 0 and so == 0 and not ir.any() and not io.any()


@ref_required
@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,planar,unmatched,sigma", [(11, 400, False, 0.1, 1.0), (12, 1500, True, 0.0, 1.0), (13, 8, False, 0.0, 1.0),
                                                           (14, 777, False, 0.3, 2.0)])
def test_cuda_equals_reference_two_view(seed, n, planar, unmatched, sigma, cuda_required):
    """CUDA == the reference build with no restatement in between: the device scores every hypothesis of the reference's
    own solvers; scores, inlier masks and the kept iteration equal CheckHomography / CheckFundamental / FindHomography /
    FindFundamental of the reference's TwoViewReconstruction (the library travels to the GPU box prebuilt)."""
    from nav24_b200 import capi
    t, a, b, _ = ref_session(seed, n, planar, unmatched, sigma)
    Hs, His, Fs = t.hypotheses()
    ctx = capi.OrbContext(1000)
    try:
        r = ctx.two_view_score(a, b, Hs, His, Fs, sigma=sigma)
        for h in range(200):
            sr, ir = t.check_homography(Hs[h], His[h])
            assert r["score_h"][h].tobytes() == sr.tobytes() and np.array_equal(r["inliers_h"][h], ir), f"homography {h}"
            sr, ir = t.check_fundamental(Fs[h])
            assert r["score_f"][h].tobytes() == sr.tobytes() and np.array_equal(r["inliers_f"][h], ir), f"fundamental {h}"
        for find, best, sc, inl, hyp in ((t.find_homography, r["best_h"], r["score_h"], r["inliers_h"], Hs),
                                         (t.find_fundamental, r["best_f"], r["score_f"], r["inliers_f"], Fs)):
            score, mask, M = find()
            if best < 0:
                assert score == 0 and not mask.any()
            else:
                assert score.tobytes() == sc[best].tobytes() and np.array_equal(mask, inl[best]) and np.array_equal(M, hyp[best])
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,planar", [(1, 400, False), (2, 400, True), (3, 1, False), (4, 2500, False), (5, 1024, True), (6, 1025, False)])
def test_cuda_scores_equal_oracle(seed, n, planar, cuda_required):
    from nav24_b200 import capi
    x1, x2, H21, H12, F21 = scene(seed, n, planar)
    ctx = capi.OrbContext(1000)
    try:
        r = ctx.two_view_score(x1, x2, H21, H12, F21)
        sh = np.zeros(200, np.float32); sf = np.zeros(200, np.float32)
        for h in range(200):
            sh[h], ih = oo.check_homography(H21[h], H12[h], x1, x2)
            sf[h], i_f = oo.check_fundamental(F21[h], x1, x2)
            assert np.array_equal(r["inliers_h"][h], ih), f"homography inliers of hypothesis {h}"
            assert np.array_equal(r["inliers_f"][h], i_f), f"fundamental inliers of hypothesis {h}"
        assert r["score_h"].tobytes() == sh.tobytes(), "homography scores are not bit-equal"
        assert r["score_f"].tobytes() == sf.tobytes(), "fundamental scores are not bit-equal"

        def keep(sc):      # the reference's `if (currentScore > score)` loop
            b, best = -1, np.float32(0)
            for i, v in enumerate(sc):
                if v > best:
                    b, best = i, v
            return b
        assert r["best_h"] == keep(sh) and r["best_f"] == keep(sf)
        # one model at a time, no inlier masks
        r2 = ctx.two_view_score(x1, x2, F21=F21, want_inliers=False)
        assert r2["score_h"] is None and r2["score_f"].tobytes() == sf.tobytes() and r2["best_f"] == r["best_f"]
        r3 = ctx.two_view_score(x1, x2, H21, H12, want_inliers=False)
        assert r3["score_f"] is None and r3["score_h"].tobytes() == sh.tobytes()
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,planar", [(21, 400, False), (22, 1500, True), (23, 1, False), (24, 8, False)])
def test_cuda_kept_only_call_equals_full_call(seed, n, planar, cuda_required):
    """nav24_two_view_score_kept: the scores of all iterations, the kept iteration and only its inlier mask — identical to
    the full call's row `best` (and to the oracle), for both models, one model at a time, and when nothing is kept."""
    from nav24_b200 import capi
    x1, x2, H21, H12, F21 = scene(seed, n, planar)
    ctx = capi.OrbContext(1000)
    try:
        full = ctx.two_view_score(x1, x2, H21, H12, F21)
        kept = ctx.two_view_score_kept(x1, x2, H21, H12, F21)
        assert kept["score_h"].tobytes() == full["score_h"].tobytes() and kept["score_f"].tobytes() == full["score_f"].tobytes()
        assert kept["best_h"] == full["best_h"] and kept["best_f"] == full["best_f"]
        for tag, b in (("h", full["best_h"]), ("f", full["best_f"])):
            want = full["inliers_" + tag][b] if b >= 0 else np.zeros(n, np.uint8)
            assert np.array_equal(kept["kept_inliers_" + tag], want), tag
        if full["best_h"] >= 0:
            assert np.array_equal(kept["kept_inliers_h"], oo.check_homography(H21[full["best_h"]], H12[full["best_h"]], x1, x2)[1])
        kf = ctx.two_view_score_kept(x1, x2, F21=F21)
        assert kf["score_h"] is None and kf["best_f"] == full["best_f"] and np.array_equal(kf["kept_inliers_f"], kept["kept_inliers_f"])
        # hypotheses that explain nothing: no iteration is kept, the masks come back all zero
        far = np.tile(np.array([1, 0, 1e5, 0, 1, 1e5, 0, 0, 1], np.float32), (3, 1)); farI = np.tile(np.array([1, 0, -1e5, 0, 1, -1e5, 0, 0, 1], np.float32), (3, 1))
        none = ctx.two_view_score_kept(x1, x2, far, farI, None)
        assert none["best_h"] == -1 and not none["score_h"].any() and not none["kept_inliers_h"].any()
        empty = ctx.two_view_score_kept(x1[:0], x2[:0], H21, H12, F21)
        assert empty["best_h"] == -1 and empty["best_f"] == -1 and len(empty["kept_inliers_h"]) == 0
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_two_view_edge_cases(cuda_required):
    from nav24_b200 import capi
    x1, x2, H21, H12, F21 = scene(9, 50)
    ctx = capi.OrbContext(1000)
    try:
        r = ctx.two_view_score(x1[:0], x2[:0], H21, H12, F21)                # no matches: all scores 0, nothing kept
        assert not r["score_h"].any() and not r["score_f"].any() and r["best_h"] == -1 and r["best_f"] == -1
        with pytest.raises(capi.Nav24Error) as e:
            ctx.two_view_score(x1, x2, H21, None, None)                       # H21 without H12
        assert e.value.code == capi.E_BADARG
        with pytest.raises(capi.Nav24Error):
            ctx.two_view_score(x1, x2, H21, H12, F21, sigma=0.0)
        # matches on real detections: frame 0 vs frame 1 of a shifted sequence, pure translation => the identity-plus-shift
        # homography explains the matches
        from nav24_b200.synth import sequence
        fr = sequence(480, 752, 5, 2, step=(6, 2))
        n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, [(0, 1)], capi.grid_for(752, 480))
        i1 = np.nonzero(m[0, :n[0]] >= 0)[0]; i2 = m[0, i1]
        p1 = np.stack([kps[0, i1]["x"], kps[0, i1]["y"]], 1); p2 = np.stack([kps[1, i2]["x"], kps[1, i2]["y"]], 1)
        H = np.array([[1, 0, -6], [0, 1, -2], [0, 0, 1]], np.float32)
        r = ctx.two_view_score(p1, p2, H.reshape(1, 9), np.linalg.inv(H).astype(np.float32).reshape(1, 9))
        s, inl = oo.check_homography(H.reshape(9), np.linalg.inv(H).astype(np.float32).reshape(9), p1, p2)
        assert r["score_h"][0].tobytes() == np.float32(s).tobytes() and np.array_equal(r["inliers_h"][0], inl)
        assert inl.sum() > 0.8 * len(p1)
    finally:
        ctx.close()
