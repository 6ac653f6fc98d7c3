"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded frames.

Bit-exact: pyramid pixels, blurred pixels, raw FAST corners (coordinates, scores, order), retained keypoint
sets and their order, orientation, final two-ended output order, match indices.  Descriptors: bit-exact
except where a 1-ulp cosf/sinf difference flips a rounded sample coordinate; mismatching descriptors are
counted and must stay <= 0.1 % (BASELINE.json north_star).
"""
import numpy as np
import pytest

from nav24_b200 import capi
from nav24_b200.synth import synth, sequence
from oracle import orb_oracle as oo

pytestmark = pytest.mark.gpu

DESC_TOL = 1e-3   # fraction of descriptors allowed to differ (north_star: <= 0.1 %)

CASES = [  # H, W, nFeatures, seed, lowtex
    (480, 752, 1000, 24, False),      # config 1 (EuRoC)
    (480, 752, 5000, 25, False),      # config 1, the front end's x5 init mode
    (480, 752, 200, 26, True),        # x0.2 mode, low texture
    (376, 1241, 2000, 24, False),     # config 2 (KITTI): monoIndex > 0
    (376, 1241, 2000, 31, True),
    (480, 640, 1000, 7, False),       # config 3 (TUM)
    (260, 340, 300, 5, True),         # near the minimum size
    (300, 900, 700, 9, False),        # 3:1 panorama, nIni = 3
    (2160, 3840, 8000, 24, False),    # config 4 (4K stress): 108 x 60 cells at level 0, 18 FAST segments per cell row
    (376, 1241, 10000, 33, False),    # config 2 in the front end's x5 init mode (FE_SlamMonoV.cpp:247): 2172 nodes at level 0
    (2160, 3840, 40000, 25, False),   # config 4 in the x5 mode: 8687 level-0 nodes
    (2160, 3840, 60000, 26, False),   # 13030 level-0 nodes: more sort records than the 12000 the CTA sorts in shared memory -> the global-memory sort path
]


def _oracle_matches_on(kps, desc, n, a, b, grid, **kw):
    """The oracle matcher on the GPU's OWN keypoints and descriptors of frames a, b: a valid expectation even when a
    descriptor differs from the oracle's inside the 0.1 % tolerance, so match assertions never have to be gated on
    descriptor equality (keypoints are asserted byte-equal separately)."""
    k1, k2 = kps[a, :n[a]], kps[b, :n[b]]
    return oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), desc[a, :n[a]], k2, np.stack([k2["x"], k2["y"]], 1),
                           desc[b, :n[b]], grid, **kw)


@pytest.fixture(scope="module")
def ctxs(cuda_required):
    cache = {}
    yield cache
    for c in cache.values():
        c.close()


def _ctx(cache, nf):
    if nf not in cache:
        cache[nf] = capi.OrbContext(nf)
    return cache[nf]


def _compare_detect(ctx, img, nf, check_stages=True):
    o = oo.OrbOracle(nf)
    mono_o, k_o, d_o = o.detect(img)
    mono_g, k_g, d_g = ctx.detect(img)
    if check_stages:
        s_g, i_g, q_g = ctx.tables(); s_o, i_o, q_o, _ = o.tables()
        assert np.array_equal(s_g, s_o) and np.array_equal(i_g, i_o) and np.array_equal(q_g, q_o)
        for l in range(8):
            assert np.array_equal(ctx.level(0, l), o.level(l)), f"pyramid level {l}"
            assert np.array_equal(ctx.raw_keys(0, l), o.raw(l)), f"raw FAST keys level {l}"
            b = o.blurred(l)
            if b is not None:
                assert np.array_equal(ctx.level(0, l, blurred=True), b), f"blurred level {l}"
            lk_o, _ = o.level_kps(l)
            lk_g = ctx.level_keypoints(0, l)
            assert lk_g.tobytes() == lk_o.tobytes(), f"level keypoints {l}"
    assert mono_g == mono_o
    assert len(k_g) == len(k_o)
    assert k_g.tobytes() == k_o.tobytes()
    bad = int((d_g != d_o).any(axis=1).sum())
    assert bad <= DESC_TOL * max(1, len(k_o)), f"{bad}/{len(k_o)} descriptors differ"
    return bad, len(k_o), (mono_g, k_g, d_g), (mono_o, k_o, d_o)


@pytest.mark.parametrize("H,W,nf,seed,low", CASES)
def test_detect_parity(ctxs, H, W, nf, seed, low):
    img = synth(H, W, seed, lowtex=low)
    bad, n, _, _ = _compare_detect(_ctx(ctxs, nf), img, nf)
    print(f"descriptor mismatches {bad}/{n}")


@pytest.mark.parametrize("W", [384, 385, 386, 387, 389, 390, 513, 514, 517, 518, 640 + 2, 640 + 5])
def test_blur_right_edge_geometries(ctxs, W):
    """Every (width mod 4) and the widths whose last word lands in lane 0/1 of a 128-px blur tile."""
    ctx = _ctx(ctxs, 300)
    img = synth(264, W, 1000 + W)
    o = oo.OrbOracle(300); o.detect(img)
    ctx.detect(img)
    for l in range(8):
        b = o.blurred(l)
        if b is not None:
            assert np.array_equal(ctx.level(0, l, blurred=True), b), f"blurred level {l} of width {W}"
        assert np.array_equal(ctx.level(0, l), o.level(l)), f"pyramid level {l} of width {W}"


GEOMS = [(280, 384), (281, 385), (283, 511), (420, 512), (421, 513), (419, 640), (560, 641), (561, 767), (559, 768),
         (300, 1024), (287, 1025), (480, 320), (423, 387), (562, 899)]


@pytest.mark.parametrize("H,W", GEOMS)
def test_tile_edge_geometries(ctxs, H, W):
    """Heights around multiples of the 140-row blur CTA tile / 32-row resize tile and widths around multiples of 128:
    every stage output of every level against the oracle (the TMA boxes of resize, blur, FAST and describe reach over
    the image edges there; zero fill, REFLECT_101 patches and the host-computed box sizes must all agree)."""
    _compare_detect(_ctx(ctxs, 400), synth(H, W, 5000 + H + W, lowtex=(H + W) % 3 == 0), 400)


@pytest.mark.parametrize("H,W,low", [(300, 1149, True), (343, 1335, True), (310, 1151, True), (300, 1149, False), (290, 1150, False)])
def test_last_fast_segment_of_one_narrow_cell(ctxs, H, W, low):
    """Widths whose last FAST segment of level 0 is ONE cell with a 1- or 2-px interior (31 cells of 37 px over 1117 px: the last
    cell is 7 px wide; found by tools/gpu_fuzz.py): the segment has a single word column, whose index / column split had
    no 32-bit magic divisor, so only the first row of that cell was scanned.  Raw FAST keys of every level against the oracle."""
    img = synth(H, W, 700 + W, lowtex=low)
    _compare_detect(_ctx(ctxs, 2000), img, 2000)


@pytest.mark.parametrize("scale,nlevels,ini,mn", [(1.1, 6, 20, 7), (1.5, 5, 30, 10), (1.33, 4, 12, 5), (1.8, 3, 20, 7)])
def test_other_pyramid_parameters(scale, nlevels, ini, mn, cuda_required):
    """Scale factor, level count and FAST thresholds other than the defaults (OP_FtDt.cpp:31-48 reads them from YAML):
    the resize tile geometry (rows per warp, TMA box sizes) is derived from the scale factor."""
    img = synth(480, 752, 77)
    ctx = capi.OrbContext(800, scale_factor=scale, n_levels=nlevels, ini_th_fast=ini, min_th_fast=mn)
    try:
        o = oo.OrbOracle(800, scale, nlevels, ini, mn)
        mono_o, k_o, d_o = o.detect(img)
        mono_g, k_g, d_g = ctx.detect(img)
        for l in range(nlevels):
            assert np.array_equal(ctx.level(0, l), o.level(l)), f"pyramid level {l}"
            assert np.array_equal(ctx.raw_keys(0, l), o.raw(l)), f"raw FAST keys level {l}"
        assert mono_g == mono_o and k_g.tobytes() == k_o.tobytes()
        assert int((d_g != d_o).any(axis=1).sum()) <= DESC_TOL * max(1, len(k_o))
    finally:
        ctx.close()


@pytest.mark.parametrize("H,W,scale,nl", [(909, 762, 1.6, 5), (300, 475, 1.6, 4), (360, 371, 1.5, 4), (400, 697, 1.33, 6)])
def test_resize_lanes_beyond_a_narrow_level(cuda_required, H, W, scale, nl):
    """Level widths that are not a multiple of the pixels a thread owns, at large scale factors (found by tools/gpu_fuzz2.py:
    scale 1.6, 762 px -> ... -> 116 px): the lanes beyond the level used to place their source window up to three source steps
    right of the last active lane's, past the last row of the shared-memory tile (an illegal address that killed the
    context).  Every level against the oracle."""
    img = np.random.default_rng(H * W).integers(0, 256, (H, W), dtype=np.uint8)
    ctx = capi.OrbContext(100, scale_factor=scale, n_levels=nl, ini_th_fast=30, min_th_fast=30, raw_keys_per_kpx=250)
    try:
        o = oo.OrbOracle(100, scale, nl, 30, 30)
        mono_o, k_o, d_o = o.detect(img)
        mono_g, k_g, d_g = ctx.detect(img)
        for l in range(nl):
            assert np.array_equal(ctx.level(0, l), o.level(l)), f"pyramid level {l}"
            assert np.array_equal(ctx.raw_keys(0, l), o.raw(l)), f"raw FAST keys level {l}"
        assert mono_g == mono_o and k_g.tobytes() == k_o.tobytes()
        assert int((d_g != d_o).any(axis=1).sum()) <= DESC_TOL * max(1, len(k_o))
    finally:
        ctx.close()


def test_detect_strided_input_and_reuse(ctxs):
    ctx = _ctx(ctxs, 1000)
    big = synth(500, 800, 3)
    view = big[10:490, 20:772]                  # non-contiguous rows, odd base alignment
    _compare_detect(ctx, view, 1000)
    _compare_detect(ctx, synth(480, 640, 4), 1000)       # shape change on the same context
    ctx.set_num_features(5000)                           # scaleNumFeatures(5.f) of the front end
    _compare_detect(ctx, synth(480, 640, 4), 5000)
    ctx.set_num_features(1000)


def test_flat_and_noise_images(ctxs):
    ctx = _ctx(ctxs, 1000)
    flat = np.full((480, 640), 93, np.uint8)
    mono, k, d = ctx.detect(flat)
    assert len(k) == 0 and mono == 0
    o = oo.OrbOracle(1000); assert len(o.detect(flat)[1]) == 0
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, (300, 400), dtype=np.uint8)        # corners everywhere: buffer stress
    c2 = capi.OrbContext(1000, raw_keys_per_kpx=250)
    try:
        _compare_detect(c2, noise, 1000)
    finally:
        c2.close()
    with pytest.raises(capi.Nav24Error) as e:                       # default capacity must fail loudly, not truncate
        c3 = capi.OrbContext(1000, raw_keys_per_kpx=5)
        try:
            c3.detect(noise)
        finally:
            c3.close()
    assert e.value.code == capi.E_OVERFLOW


def test_error_codes(ctxs):
    ctx = _ctx(ctxs, 1000)
    with pytest.raises(capi.Nav24Error) as e:
        ctx.detect(np.zeros((100, 120), np.uint8))                  # too small for 8 levels
    assert e.value.code == capi.E_GEOMETRY
    n = capi.C.c_int(0)
    rc = ctx.L.nav24_orb_detect(ctx.h, None, 0, 0, 0, None, None, 0, capi.C.byref(n))
    assert rc == capi.E_BADARG                                      # empty image -> -1 like the reference


def test_batch_equals_single(ctxs):
    ctx = _ctx(ctxs, 2000)
    frames = np.stack([synth(376, 1241, 100 + i, lowtex=(i % 2 == 1)) for i in range(5)])
    n, mono, kps, desc = ctx.detect_batch(frames)
    o = oo.OrbOracle(2000)
    for f in range(5):
        mo, ko, do = o.detect(frames[f])
        assert mono[f] == mo and n[f] == len(ko)
        assert kps[f, :n[f]].tobytes() == ko.tobytes()
        assert int((desc[f, :n[f]] != do).any(axis=1).sum()) <= DESC_TOL * len(ko)


@pytest.mark.parametrize("H,W,nf", [(480, 752, 1000), (243, 427, 500), (376, 1241, 2000)])
def test_batch_launch_paths_equal_single_frame_paths(ctxs, H, W, nf):
    """Batches of >= 16 frames take launch paths a single frame never sees: FAST launched per tile-height group, the
    <= 128-px remainder column of a wide pyramid level through the four-pixel resize kernel, the output order as its own
    kernel.  Every frame of such a batch must equal the same frame detected alone (which takes the other paths), and
    sampled frames must equal the oracle, level by level."""
    ctx = _ctx(ctxs, nf)
    frames = np.stack([synth(H, W, 300 + i, lowtex=(i % 5 == 4)) for i in range(18)])
    n, mono, kps, desc = ctx.detect_batch(frames)
    lv = {f: [ctx.level(f, l) for l in range(8)] for f in (0, 17)}
    raw = {f: [ctx.raw_keys(f, l) for l in range(8)] for f in (0, 17)}
    for f in range(len(frames)):
        m1, k1, d1 = ctx.detect(frames[f])
        assert mono[f] == m1 and n[f] == len(k1)
        assert kps[f, :n[f]].tobytes() == k1.tobytes() and np.array_equal(desc[f, :n[f]], d1)
    o = oo.OrbOracle(nf)
    for f in (0, 17):
        mo, ko, do = o.detect(frames[f])
        assert mono[f] == mo and kps[f, :n[f]].tobytes() == ko.tobytes()
        assert int((desc[f, :n[f]] != do).any(axis=1).sum()) <= DESC_TOL * len(ko)
        for l in range(8):
            assert np.array_equal(lv[f][l], o.level(l)), f"pyramid level {l} of frame {f}"
            assert np.array_equal(raw[f][l], o.raw(l)), f"raw FAST keys level {l} of frame {f}"


@pytest.mark.parametrize("H,W", [(240, 512), (300, 641), (260, 769), (333, 897), (250, 1025), (301, 1153), (256, 1281), (270, 1409),
                                 (480, 307), (600, 460)])
def test_batch_launch_paths_geometry_sweep(cuda_required, H, W):
    """Level widths around the 128 / 256-column boundaries of the resize kernels and cell heights that put the levels into
    different FAST launch groups: a batch of 16 (batch launch paths) must equal the single-frame call (single-launch paths)
    byte for byte, pyramid levels and raw FAST keys included; one frame per shape is also checked against the oracle."""
    nf = 700
    ctx = capi.OrbContext(nf)
    try:
        img = synth(H, W, 500 + W, lowtex=(W % 2 == 0))
        frames = np.stack([img] * 16)
        n, mono, kps, desc = ctx.detect_batch(frames)
        lv = [ctx.level(15, l) for l in range(8)]
        raw = [ctx.raw_keys(15, l) for l in range(8)]
        assert (n == n[0]).all() and (mono == mono[0]).all()
        for f in range(1, 16):
            assert kps[f, :n[f]].tobytes() == kps[0, :n[0]].tobytes() and np.array_equal(desc[f, :n[f]], desc[0, :n[0]])
        m1, k1, d1 = ctx.detect(img)
        assert mono[0] == m1 and kps[0, :n[0]].tobytes() == k1.tobytes() and np.array_equal(desc[0, :n[0]], d1)
        for l in range(8):
            assert np.array_equal(lv[l], ctx.level(0, l)), f"pyramid level {l}"
            assert np.array_equal(raw[l], ctx.raw_keys(0, l)), f"raw FAST keys level {l}"
        o = oo.OrbOracle(nf)
        mo, ko, do = o.detect(img)
        assert m1 == mo and k1.tobytes() == ko.tobytes()
        assert int((d1 != do).any(axis=1).sum()) <= DESC_TOL * max(1, len(ko))
    finally:
        ctx.close()


def test_window_matcher_parity(ctxs):
    ctx = _ctx(ctxs, 1000)
    for (H, W, step) in [(480, 752, (2, 1)), (376, 1241, (17, 0)), (480, 640, (3, 1))]:
        fr = sequence(H, W, 77, 4, step=step)
        o = oo.OrbOracle(1000)
        dets = [o.detect(f) for f in fr]
        grid_o = oo.grid_for(W, H); grid_g = capi.grid_for(W, H)
        for a, b in [(0, 1), (0, 3), (2, 3)]:
            (_, k1, d1), (_, k2, d2) = dets[a], dets[b]
            ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
            for ori in (True, False):
                ref = oo.match_window(k1, ud1, d1, k2, ud2, d2, grid_o, check_ori=ori)
                got, nm = ctx.match_window(k1, ud1, d1, k2, ud2, d2, grid_g, check_ori=ori)
                assert np.array_equal(got, ref)
                assert nm == int((ref >= 0).sum())
            assert (ref >= 0).sum() > 20


def test_config1_sequence_100_frames(ctxs):
    """BASELINE.json configs[0] as SURVEY 8(d) spells it: 752x480, 1000 keypoints, 100 frames (frame t shifted by
    (2t, t) px), every frame against the oracle, frame 0 matched against frames 1, 10, 50, 99."""
    H, W, nf = 480, 752, 1000
    fr = sequence(H, W, 24, 100, step=(2, 1))
    ctx = _ctx(ctxs, nf)
    pairs = [(0, 1), (0, 10), (0, 50), (0, 99)]
    n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, pairs, capi.grid_for(W, H))
    o = oo.OrbOracle(nf)
    bad = 0
    for f in range(len(fr)):
        mo, ko, do = o.detect(fr[f])
        assert mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes(), f
        bad += int((desc[f, :n[f]] != do).any(axis=1).sum())
    assert bad <= DESC_TOL * int(n.sum())
    for q, (a, b) in enumerate(pairs):
        want = _oracle_matches_on(kps, desc, n, a, b, oo.grid_for(W, H))
        assert np.array_equal(m[q, :n[a]], want) and nm[q] == (want >= 0).sum()
    assert nm[0] > 100


def test_cuda_equals_reference_build(ctxs):
    """The CUDA path against the REFERENCE's own classes (oracle/_ref, compiled unchanged from /root/reference and
    carried to this box prebuilt) with no restatement in between: detect and matchV on a KITTI-shaped stereo pair."""
    from oracle import ref_lib as rl
    if not rl.available():
        pytest.skip("oracle/_ref/libnav24_ref.so did not travel to this box")
    H, W, nf = 376, 1241, 2000
    fr = sequence(H, W, 42, 2, step=(11, 0))
    ctx = _ctx(ctxs, nf)
    n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, [(0, 1)], capi.grid_for(W, H))
    r = rl.RefOrb(nf)
    bad = 0
    for f in range(2):
        mr, kr, dr = r.detect(fr[f])
        assert mono[f] == mr and n[f] == len(kr) and kps[f, :n[f]].tobytes() == kr.tobytes()
        bad += int((desc[f, :n[f]] != dr).any(axis=1).sum())
    assert bad <= DESC_TOL * int(n.sum())
    k1, k2 = kps[0, :n[0]], kps[1, :n[1]]
    want = rl.match_window(k1, np.stack([k1["x"], k1["y"]], 1), desc[0, :n[0]], k2, np.stack([k2["x"], k2["y"]], 1),
                           desc[1, :n[1]], W, H)
    assert np.array_equal(m[0, :n[0]], want) and nm[0] == (want >= 0).sum() and nm[0] > 100


def test_window_matcher_adversarial(ctxs):
    """Many near-identical descriptors: steals, ties and the >32-candidate slow path."""
    ctx = _ctx(ctxs, 1000)
    rng = np.random.default_rng(3)
    n1, n2 = 600, 900
    base = rng.integers(0, 256, (6, 32), dtype=np.uint8)
    def mk(n):
        d = base[rng.integers(0, 6, n)].copy()
        flips = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & \
            rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
        return d ^ flips
    def kp(n):
        k = np.zeros(n, capi.KP_DTYPE)
        k["x"] = rng.uniform(0, 300, n).astype(np.float32); k["y"] = rng.uniform(0, 200, n).astype(np.float32)
        k["angle"] = rng.uniform(0, 360, n).astype(np.float32); k["octave"] = rng.integers(0, 3, n) // 2
        k["class_id"] = -1
        return k
    k1, k2, d1, d2 = kp(n1), kp(n2), mk(n1), mk(n2)
    ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1) + np.float32(0.25)
    for (th, ratio) in [(50, 0.6), (90, 0.9), (120, 1.0), (256, 1.5)]:
        ref = oo.match_window(k1, ud1, d1, k2, ud2, d2, oo.grid_for(300, 200), th_low=th, nnratio=ratio)
        got, _ = ctx.match_window(k1, ud1, d1, k2, ud2, d2, capi.grid_for(300, 200), th_low=th, nnratio=ratio)
        assert np.array_equal(got, ref), (th, ratio)


def test_device_resident_match_frames(ctxs):
    ctx = _ctx(ctxs, 2000)
    H, W = 376, 1241
    fr = sequence(H, W, 5, 3, step=(9, 0))
    n, mono, kps, desc = ctx.detect_batch(fr)
    m, nm = ctx.match_window_frames([(0, 1), (0, 2), (1, 2)], capi.grid_for(W, H))
    for p, (a, b) in enumerate([(0, 1), (0, 2), (1, 2)]):
        k1, k2 = kps[a, :n[a]], kps[b, :n[b]]
        ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
        ref = oo.match_window(k1, ud1, desc[a, :n[a]], k2, ud2, desc[b, :n[b]], oo.grid_for(W, H))
        assert np.array_equal(m[p, :n[a]], ref)
        assert nm[p] == (ref >= 0).sum() and nm[p] > 50


def test_large_resident_batch_runs_as_two_halves(monkeypatch, cuda_required):
    """Device-resident batches of >= 512 frames are cut into two halves on two streams (capi.cu, detect_match_device).
    The results must equal the same batch run as one chunk, pairs that span the halves included, and sampled frames / pairs
    must equal the oracle."""
    H, W, nf, B = 260, 340, 300, 514
    base = sequence(H, W, 77, 8, step=(4, 1))
    fr = np.ascontiguousarray(np.stack([base[(f * 5) % 8] if f % 3 else synth(H, W, 400 + f % 11) for f in range(B)]))
    pitch = (W + 127) // 128 * 128
    buf = np.zeros((B, H, pitch), np.uint8)
    buf[:, :, :W] = fr
    pairs = [(2 * i, 2 * i + 1) for i in range(B // 2)] + [(256, 257), (255, 258), (0, 513), (256, 255)]
    grid = capi.grid_for(W, H)
    res = {}
    for name, env in (("halves", None), ("one", "100000")):
        if env:
            monkeypatch.setenv("NAV24_RESIDENT_CHUNK", env)
        ctx = capi.OrbContext(nf)
        try:
            dptr = capi.C.c_void_p()
            assert ctx.L.nav24_device_alloc(buf.nbytes, capi.C.byref(dptr)) == 0
            assert ctx.L.nav24_memcpy_h2d(dptr, buf.ctypes.data_as(capi.C.c_void_p), buf.nbytes) == 0
            for _ in range(2):      # twice: the second call runs ahead of the first one's tail on the second stream
                ctx.detect_match_device(dptr.value, B, W, H, pitch, pitch * H, pairs, grid)
            ctx.sync()
            res[name] = ctx.fetch(B) + ctx.match_fetch(len(pairs))
            ctx.L.nav24_device_free(dptr)
        finally:
            ctx.close()
    (n, mono, kps, desc, m, nm), (n2, mono2, kps2, desc2, m2, nm2) = res["halves"], res["one"]
    assert np.array_equal(n, n2) and np.array_equal(mono, mono2) and np.array_equal(nm, nm2)
    for f in range(B):
        assert kps[f, :n[f]].tobytes() == kps2[f, :n[f]].tobytes() and np.array_equal(desc[f, :n[f]], desc2[f, :n[f]])
    for q, (a, b) in enumerate(pairs):
        assert np.array_equal(m[q, :n[a]], m2[q, :n[a]])
    o = oo.OrbOracle(nf)
    for f in (0, 255, 256, 257, 513):
        mo, ko, do = o.detect(fr[f])
        assert mono[f] == mo and kps[f, :n[f]].tobytes() == ko.tobytes()
        assert int((desc[f, :n[f]] != do).any(axis=1).sum()) <= DESC_TOL * max(1, len(ko))
    for q in (128, B // 2, B // 2 + 1, B // 2 + 2, B // 2 + 3):
        a, b = pairs[q]
        assert np.array_equal(m[q, :n[a]], _oracle_matches_on(kps, desc, n, a, b, oo.grid_for(W, H)))


@pytest.mark.parametrize("norm", [0, 1])
def test_bf_knn2_parity(ctxs, norm):
    ctx = _ctx(ctxs, 1000)
    rng = np.random.default_rng(11)
    d2 = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    d1 = d2[rng.integers(0, 1500, 1100)].copy()
    d1 ^= rng.integers(0, 256, d1.shape, dtype=np.uint8) & rng.integers(0, 256, d1.shape, dtype=np.uint8) & \
        rng.integers(0, 256, d1.shape, dtype=np.uint8)
    d2[40] = d2[41]; d1[0] = d2[40]
    ref = oo.match_bf_knn2(d1, d2, norm, 0.7)
    got = ctx.match_bf_knn2(d1, d2, norm, 0.7)
    for r, g in zip(ref, got):
        assert np.array_equal(r, g)
    # degenerate sizes
    got = ctx.match_bf_knn2(d1[:3], d2[:1], norm, 0.7)
    ref = oo.match_bf_knn2(d1[:3], d2[:1], norm, 0.7)
    for r, g in zip(ref, got):
        assert np.array_equal(r, g)


def test_fused_detect_match_chunked(ctxs, monkeypatch):
    """nav24_orb_detect_match_batch / _device: chunks on two streams, pairs inside a chunk and pairs spanning chunks."""
    H, W, nf = 376, 1241, 2000
    fr = sequence(H, W, 9, 10, step=(6, 0))
    pairs = [(0, 1), (2, 3), (4, 5), (6, 7), (8, 9), (1, 2), (0, 9), (3, 3)]
    o = oo.OrbOracle(nf)
    ref = [o.detect(f) for f in fr]
    grid_o = oo.grid_for(W, H)
    mref = []
    for a, b in pairs:
        (_, k1, d1), (_, k2, d2) = ref[a], ref[b]
        mref.append(oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), d1, k2, np.stack([k2["x"], k2["y"]], 1), d2, grid_o))
    for chunk in ("4", "2", "64", "3"):
        monkeypatch.setenv("NAV24_CHUNK_FRAMES", chunk)
        monkeypatch.setenv("NAV24_RESIDENT_CHUNK", chunk)
        ctx = capi.OrbContext(nf)
        try:
            n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, pairs, capi.grid_for(W, H))
            exact = True
            for f in range(len(fr)):
                mo, ko, do = ref[f]
                assert mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()
                exact &= np.array_equal(desc[f, :n[f]], do)
            for q, (a, b) in enumerate(pairs):
                want = mref[q] if exact else _oracle_matches_on(kps, desc, n, a, b, grid_o)
                assert np.array_equal(m[q, :n[a]], want), (chunk, q)
                assert nm[q] == (want >= 0).sum()
            # device-resident form, twice back to back (asynchronous), then fetch
            padded = np.zeros((len(fr), H, 1248), np.uint8); padded[:, :, :W] = fr
            dptr = capi.C.c_void_p()
            assert ctx.L.nav24_device_alloc(padded.nbytes, capi.C.byref(dptr)) == 0
            assert ctx.L.nav24_memcpy_h2d(dptr, padded.ctypes.data_as(capi.C.c_void_p), padded.nbytes) == 0
            for _ in range(2):
                ctx.detect_match_device(dptr.value, len(fr), W, H, 1248, 1248 * H, pairs, capi.grid_for(W, H))
            n2, mono2, kps2, desc2 = ctx.fetch(len(fr))
            m2, nm2 = ctx.match_fetch(len(pairs))
            assert np.array_equal(n2, n) and np.array_equal(mono2, mono)
            for f in range(len(fr)):
                assert kps2[f, :n[f]].tobytes() == kps[f, :n[f]].tobytes() and np.array_equal(desc2[f, :n[f]], desc[f, :n[f]])
            assert np.array_equal(nm2, nm)
            for q, (a, b) in enumerate(pairs):
                assert np.array_equal(m2[q, :n[a]], m[q, :n[a]])
            ctx.L.nav24_device_free(dptr)
        finally:
            ctx.close()


def test_async_device_calls_with_changing_pairs_and_features(monkeypatch, cuda_required):
    """Back-to-back ASYNCHRONOUS device-resident calls whose pair counts, pair lists and feature counts change between
    calls (the front end toggles scaleNumFeatures(5) / (0.2), FE_SlamMonoV.cpp:136,176,247): the pair-table slots are fixed
    size and guarded per slot, and the extra streams may not run ahead across a change of geometry."""
    H, W = 376, 1241
    fr = sequence(H, W, 61, 8, step=(5, 0))
    padded = np.zeros((len(fr), H, 1248), np.uint8); padded[:, :, :W] = fr
    monkeypatch.setenv("NAV24_RESIDENT_CHUNK", "3")          # several chunks on several streams: the run-ahead path
    ctx = capi.OrbContext(2000)
    try:
        dptr = capi.C.c_void_p()
        assert ctx.L.nav24_device_alloc(padded.nbytes, capi.C.byref(dptr)) == 0
        assert ctx.L.nav24_memcpy_h2d(dptr, padded.ctypes.data_as(capi.C.c_void_p), padded.nbytes) == 0
        grid = capi.grid_for(W, H)
        calls = [(2000, [(0, 1), (2, 3), (4, 5), (6, 7), (1, 2)]), (2000, [(7, 0), (3, 4)]), (400, [(0, 7), (1, 6), (2, 5), (3, 4), (4, 3), (5, 2), (6, 1)]),
                 (2000, [(5, 6)]), (10000, [(0, 1), (1, 2), (2, 3)])]
        for nf, pairs in calls:                                # enqueue everything without a host synchronisation in between
            ctx.set_num_features(nf)
            ctx.detect_match_device(dptr.value, len(fr), W, H, 1248, 1248 * H, pairs, grid)
        nf, pairs = calls[-1]
        n, mono, kps, desc = ctx.fetch(len(fr))
        m, nm = ctx.match_fetch(len(pairs))
        o = oo.OrbOracle(nf)
        for f in (0, 1, 2, 3):
            mo, ko, do = o.detect(fr[f])
            assert mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()
        for q, (a, b) in enumerate(pairs):
            want = _oracle_matches_on(kps, desc, n, a, b, oo.grid_for(W, H))
            assert np.array_equal(m[q, :n[a]], want) and nm[q] == (want >= 0).sum()
        # the same sequence of calls with a synchronisation after each gives the same final answer
        for nf2, pairs2 in calls:
            ctx.set_num_features(nf2)
            ctx.detect_match_device(dptr.value, len(fr), W, H, 1248, 1248 * H, pairs2, grid)
            ctx.sync()
        n2, mono2, kps2, desc2 = ctx.fetch(len(fr))
        m2, nm2 = ctx.match_fetch(len(pairs))
        assert np.array_equal(n2, n) and np.array_equal(nm2, nm)
        for q, (a, b) in enumerate(pairs):
            assert np.array_equal(m2[q, :n[a]], m[q, :n[a]])          # (entries beyond a frame's keypoint count are not written)
        for f in range(len(fr)):
            assert kps2[f, :n[f]].tobytes() == kps[f, :n[f]].tobytes()
        ctx.L.nav24_device_free(dptr)
    finally:
        ctx.close()


def test_stage_timers_only_after_detect_device(ctxs):
    """nav24_orb_stage_ms covers nav24_orb_detect_device only: after any other detect call it must say so instead of
    returning the timings of an older call."""
    ctx = capi.OrbContext(500)
    try:
        img = synth(260, 340, 3)
        ctx.detect(img)
        with pytest.raises(capi.Nav24Error) as e:
            ctx.stage_ms()
        assert e.value.code == capi.E_BADARG
        padded = np.zeros((1, 260, 352), np.uint8); padded[0, :, :340] = img
        dptr = capi.C.c_void_p()
        assert ctx.L.nav24_device_alloc(padded.nbytes, capi.C.byref(dptr)) == 0
        assert ctx.L.nav24_memcpy_h2d(dptr, padded.ctypes.data_as(capi.C.c_void_p), padded.nbytes) == 0
        ctx.detect_device(dptr.value, 1, 340, 260, 352, 352 * 260)
        ms = ctx.stage_ms()
        assert (ms > 0).all() and ms[4] >= ms[:4].sum() * 0.9
        ctx.detect(img)
        with pytest.raises(capi.Nav24Error):
            ctx.stage_ms()
        ctx.L.nav24_device_free(dptr)
    finally:
        ctx.close()


def test_batch_schedule_invariance_at_bench_scale(ctxs, monkeypatch):
    """A bench-sized batch (96 KITTI frames, 48 stereo pairs + pairs that span chunks): the result must not depend on
    how the host pipeline cuts it (tapered chunks on three streams vs one uniform chunk), and sampled frames / pairs
    must equal the oracle."""
    import threading
    H, W, nf = 376, 1241, 2000
    fr = sequence(H, W, 31, 96, step=(5, 1))
    pairs = [(2 * i, 2 * i + 1) for i in range(48)] + [(0, 95), (7, 40), (30, 31)]
    res = {}
    for name, env in (("taper", {"NAV24_CHUNK_FRAMES": "16", "NAV24_STREAMS": "3", "NAV24_TAPER": "1"}),
                      ("uniform", {"NAV24_CHUNK_FRAMES": "96", "NAV24_STREAMS": "1", "NAV24_TAPER": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = capi.OrbContext(nf)
        try:
            res[name] = ctx.detect_match_batch(fr, pairs, capi.grid_for(W, H))
        finally:
            ctx.close()
    (n, mono, kps, desc, m, nm), (n2, mono2, kps2, desc2, m2, nm2) = res["taper"], res["uniform"]
    assert np.array_equal(n, n2) and np.array_equal(mono, mono2) and np.array_equal(nm, nm2)
    for f in range(len(fr)):
        assert kps[f, :n[f]].tobytes() == kps2[f, :n[f]].tobytes() and np.array_equal(desc[f, :n[f]], desc2[f, :n[f]])
    for q, (a, b) in enumerate(pairs):
        assert np.array_equal(m[q, :n[a]], m2[q, :n[a]])
    o = oo.OrbOracle(nf)
    ref = {f: o.detect(fr[f]) for f in (0, 7, 40, 95)}
    for f, (mo, ko, do) in ref.items():
        assert mono[f] == mo and kps[f, :n[f]].tobytes() == ko.tobytes()
    for q, (a, b) in ((48, (0, 95)), (49, (7, 40))):
        assert np.array_equal(m[q, :n[a]], _oracle_matches_on(kps, desc, n, a, b, oo.grid_for(W, H)))

    # two contexts driven from two host threads at once (bench.py's e2e loop) give the same answers
    out = [None, None]

    def work(i):
        c = capi.OrbContext(nf)
        try:
            for _ in range(2):
                out[i] = c.detect_match_batch(fr[:32], pairs[:16], capi.grid_for(W, H))
        finally:
            c.close()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for i in range(2):
        ni, _, ki, di, mi, nmi = out[i]
        assert np.array_equal(ni, n[:32]) and np.array_equal(nmi, nm[:16])
        for f in range(32):
            assert ki[f, :ni[f]].tobytes() == kps[f, :n[f]].tobytes() and np.array_equal(di[f, :ni[f]], desc[f, :n[f]])


def test_4k_pair_detect_and_match(cuda_required):
    """BASELINE.json configs[3] shape: 3840x2160, 8000 keypoints per image, windowed matching of two shifted frames
    (8008 keypoints per frame, 384 x 216 grid cells)."""
    H, W, nf = 2160, 3840, 8000
    fr = sequence(H, W, 24, 2, step=(9, 2))
    ctx = capi.OrbContext(nf)
    try:
        n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, [(0, 1)], capi.grid_for(W, H))
        o = oo.OrbOracle(nf)
        ref = [o.detect(f) for f in fr]
        exact = True
        for f in range(2):
            mo, ko, do = ref[f]
            assert mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()
            bad = int((desc[f, :n[f]] != do).any(axis=1).sum())
            assert bad <= DESC_TOL * n[f]
            exact &= bad == 0
        (_, k1, d1), (_, k2, d2) = ref
        mref = oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), d1, k2, np.stack([k2["x"], k2["y"]], 1), d2, oo.grid_for(W, H))
        assert (mref >= 0).sum() > 200
        assert np.array_equal(m[0, :n[0]], mref if exact else _oracle_matches_on(kps, desc, n, 0, 1, oo.grid_for(W, H)))
    finally:
        ctx.close()


def test_two_devices_in_one_process(cuda_required):
    """One context per GPU inside ONE process (SURVEY 8(b) threading row): both devices give the oracle's answer, also when
    driven from two host threads at once.  Skipped on a one-GPU box."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    H, W, nf = 376, 1241, 2000
    fr = sequence(H, W, 11, 6, step=(7, 0))
    pairs = [(0, 1), (2, 3), (4, 5)]
    o = oo.OrbOracle(nf)
    ref = [o.detect(f) for f in fr]
    out = [None, None]

    def work(dev):
        c = capi.OrbContext(nf, device=dev)
        try:
            for _ in range(3):
                out[dev] = c.detect_match_batch(fr, pairs, capi.grid_for(W, H))
        finally:
            c.close()
    ths = [threading.Thread(target=work, args=(d,)) for d in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for dev in range(2):
        n, mono, kps, desc, m, nm = out[dev]
        for f in range(len(fr)):
            mo, ko, do = ref[f]
            assert mono[f] == mo and kps[f, :n[f]].tobytes() == ko.tobytes()
        assert np.array_equal(out[dev][4], out[0][4]) and np.array_equal(out[dev][5], out[0][5])


def test_warp_sort_equals_std_sort(ctxs):
    """The warp-parallel introsort must leave the exact permutation libstdc++'s std::sort leaves (ties included)."""
    ctx = _ctx(ctxs, 1000)
    rng = np.random.default_rng(2024)
    # odd sizes run the one-warp version, even sizes the whole-CTA version (launch_debug_sort)
    sizes = [1, 2, 3, 15, 16, 17, 18, 31, 32, 33, 34, 47, 48, 64, 100, 217, 218, 434, 435, 1000, 1737, 1738, 4096, 4097, 8687, 8688]
    for n in sizes:
        for kind in range(6):
            if kind == 0:   cnt = rng.integers(2, 6, n); ulx = rng.integers(0, 40, n) * 31       # tie-heavy, like the quadtree
            elif kind == 1: cnt = np.full(n, 3); ulx = np.zeros(n, np.int64)                       # all equivalent
            elif kind == 2: cnt = np.arange(n) // 7 + 2; ulx = np.zeros(n, np.int64)               # sorted runs of ties
            elif kind == 3: cnt = (n - np.arange(n)) // 3 + 2; ulx = rng.integers(0, 3, n)         # descending
            elif kind == 4: cnt = rng.integers(2, 2000, n); ulx = rng.integers(0, 8192, n)         # mostly distinct
            else:                                                                                  # median-of-3 killer -> heap sort
                k = n // 2; v = np.zeros(n, np.int64)
                for i in range(1, k + 1):
                    if i % 2: v[i - 1] = i; v[i] = k + i
                    v[k + i - 1] = 2 * i
                cnt = v + 2; ulx = np.zeros(n, np.int64)
            cnt = np.asarray(cnt, np.int64); ulx = np.asarray(ulx, np.int64)
            ref = oo.sort_sized(cnt.astype(np.int32), ulx.astype(np.int32))
            got = ctx.debug_sort((cnt * 8192 + ulx).astype(np.uint32))
            assert np.array_equal(got, ref), (n, kind)
