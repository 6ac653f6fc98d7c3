"""The C++ host operators (nav24_b200/host/nav24_ops.hpp: FtDtOrbB200 / FtAssocB200, the mirror of the reference's
OP::FtDt / OP::FtAssoc interface) driven like FE_SlamMonoV drives the reference's, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from nav24_b200.synth import sequence
from oracle import orb_oracle as oo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_ops")


def build_host_test():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_host_ops.cpp")
    lib = os.path.join(ROOT, "nav24_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", BIN, src, "-L" + lib, "-lnav24orb",
                           "-Wl,-rpath," + lib])
    return BIN


def test_host_layer_compiles_and_refuses_without_gpu(tmp_path):
    build_host_test()
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("CUDA device present; covered by the gpu test")
    fr = sequence(260, 340, 3, 1)
    inp = tmp_path / "in.raw"
    with open(inp, "wb") as f:
        f.write(np.array([1, 260, 340, 300], np.int32).tobytes()); f.write(fr.tobytes())
    r = subprocess.run([BIN, str(inp), str(tmp_path / "out.raw")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr       # fails loudly, never falls back


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,nf,scale", [(480, 752, 1000, None), (480, 752, 1000, 5.0), (376, 1241, 2000, None)])
def test_host_ops_match_oracle(tmp_path, cuda_required, H, W, nf, scale):
    if not os.path.exists(BIN):
        build_host_test()
    n = 3
    fr = sequence(H, W, 17, n, step=(4, 1))
    inp, out = tmp_path / "in.raw", tmp_path / "out.raw"
    with open(inp, "wb") as f:
        f.write(np.array([n, H, W, nf], np.int32).tobytes()); f.write(fr.tobytes())
    args = [BIN, str(inp), str(out)] + ([str(scale)] if scale else [])
    subprocess.check_call(args)
    buf = open(out, "rb").read()
    pos = 0

    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(buf, dtype, count, pos); pos += a.nbytes
        return a
    o = oo.OrbOracle(int(scale * nf) if scale else nf)
    ref = []
    for f in range(n):
        mono, nobs = take(np.int32, 2)
        k = take(oo.KP_DTYPE, nobs); d = take(np.uint8, nobs * 32).reshape(nobs, 32)
        mo, ko, do = o.detect(fr[f])
        assert mono == mo and nobs == len(ko)
        assert k.tobytes() == ko.tobytes()
        assert int((d != do).any(axis=1).sum()) <= 1e-3 * nobs
        ref.append((ko, d.copy(), np.array_equal(d, do)))      # the program's OWN descriptors feed the oracle matcher below
    m01 = None
    for f in range(1, n):
        n1 = int(take(np.int32, 1)[0]); m = take(np.int32, n1)
        if f == 1:
            m01 = m.copy()
        (k1, d1, e1), (k2, d2, e2) = ref[0], ref[f]
        ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
        mref = oo.match_window(k1, ud1, d1, k2, ud2, d2, oo.grid_for(W, H))
        assert np.array_equal(m, mref)
        assert (m >= 0).sum() > 20
    assert int(take(np.int32, 1)[0]) == -1          # empty image -> -1, like the reference
    # CalibrationB200: bounds, undistorted points and matchV on them against the oracle (RadTan, EuRoC cam0 numbers)
    K4 = [458.654, 457.296, 367.215, 248.375]; D4 = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]
    from nav24_b200 import capi
    b = take(np.float32, 4)
    bref = capi.image_bounds(lambda p: oo.undistort(oo.CAM_RADTAN, K4, D4, p), W, H, calibrated=False)
    assert np.array_equal(b, np.array(bref, np.float32))
    uds = []
    for f in range(2):
        ko = ref[f][0]
        ud = take(np.float32, 2 * len(ko)).reshape(-1, 2)
        assert ud.tobytes() == oo.undistort(oo.CAM_RADTAN, K4, D4, np.stack([ko["x"], ko["y"]], 1)).tobytes()
        uds.append(ud)
    n1 = int(take(np.int32, 1)[0]); m = take(np.int32, n1)
    (k1, d1, e1), (k2, d2, e2) = ref[0], ref[1]
    mref = oo.match_window(k1, uds[0], d1, k2, uds[1], d2, oo.grid_for(W, H, tuple(float(x) for x in b)))
    assert np.array_equal(m, mref)
    assert int(take(np.int32, 1)[0]) == 1           # the struct-of-arrays store gives identical observations and matches
    assert int(take(np.int32, 1)[0]) == 1           # the ingest ring (pinned slots) gives identical observations
    # two-view scoring of the frame 0 / frame 1 matches: scores bit-equal to the oracle's CheckHomography / CheckFundamental
    ok, nm = take(np.int32, 2)
    assert ok == 1
    H21 = take(np.float32, 18).reshape(2, 9); H12 = take(np.float32, 18).reshape(2, 9); F21 = take(np.float32, 18).reshape(2, 9)
    sH = take(np.float32, 2); sF = take(np.float32, 2)
    bestH, bestF = take(np.int32, 2)
    inH = take(np.uint8, 2 * nm).reshape(2, nm); inF = take(np.uint8, 2 * nm).reshape(2, nm)
    (k1, _, _), (k2, _, _) = ref[0], ref[1]
    i1 = np.nonzero(m01 >= 0)[0]; i2 = m01[i1]
    assert len(i1) == nm
    xy1 = np.stack([k1["x"][i1], k1["y"][i1]], 1); xy2 = np.stack([k2["x"][i2], k2["y"][i2]], 1)
    for hyp in range(2):
        s, inl = oo.check_homography(H21[hyp], H12[hyp], xy1, xy2)
        assert np.float32(s).tobytes() == sH[hyp].tobytes() and np.array_equal(inl, inH[hyp])
        s, inl = oo.check_fundamental(F21[hyp], xy1, xy2)
        assert np.float32(s).tobytes() == sF[hyp].tobytes() and np.array_equal(inl, inF[hyp])
    assert bestH == int(np.argmax(sH)) and inH[0].sum() > 0.8 * nm and bestF == int(np.argmax(sF))
    assert int(take(np.int32, 1)[0]) == 1           # TwoViewScorerB200::scoreKept == score (scores, kept iteration, its mask)
