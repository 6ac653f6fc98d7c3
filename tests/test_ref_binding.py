"""The drop-in boundary in the reference's own types.  nav24_b200/host/ref_binding holds the two classes a
nav24 maintainer adds — FtDtOrbB200 : OP::FtDt (core/operators/objDetection/OP_FtDt.hpp:14-29) and FtAssocB200 :
OP::FtAssoc (core/operators/objAssoc/OP_FtAssoc.hpp:15-22).  oracle/Makefile.ref compiles them against the REFERENCE'S
OWN HEADERS (unchanged, from /root/reference; only OpenCV / glog / Eigen are container stand-ins) and links them with the
reference's own FtDtOrbSlam / FtAssocOrbSlam / FrameMonoGrid / FeatureGrid / KeyPoint2D (oracle/_ref/libnav24_ref.so)
and with libnav24orb.so.  tests/cpp/test_ref_binding.cpp then drives both detectors and both matchers through the
reference's interfaces in the order FE_SlamMonoV::handleImageMsg does (FE_SlamMonoV.cpp:96-122) and compares the
reference's objects: every KeyPoint2D (bit for bit), monoIndex, matchV vectors, the MatchedObs stored on frame 2, and
the error paths (-1 on an empty image, {} without a grid frame).  No restatement, no ctypes, no oracle in between."""
import json
import os
import subprocess

import numpy as np
import pytest

from nav24_b200.synth import sequence
from oracle import ref_lib

needs_binary = pytest.mark.skipif("not __import__('oracle.ref_lib', fromlist=['x']).build_binding_test()",
                                  reason="oracle/_ref/test_ref_binding is not built and /root/reference is not here")


def write_input(path, fr, nf):
    n, H, W = fr.shape
    with open(path, "wb") as f:
        f.write(np.array([n, H, W, nf], np.int32).tobytes()); f.write(np.ascontiguousarray(fr).tobytes())


@needs_binary
def test_binding_compiles_against_reference_headers_and_refuses_without_gpu(tmp_path):
    """The compile + link against the reference tree is the CPU half of the check; without a CUDA device the binding
    fails loudly (exit 3) instead of falling back to the reference's CPU classes it is linked next to."""
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("CUDA device present; covered by the gpu test")
    write_input(tmp_path / "in.raw", sequence(260, 340, 3, 2), 300)
    r = subprocess.run([ref_lib.BINDING_BIN, str(tmp_path / "in.raw")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr[-300:])


@needs_binary
@pytest.mark.gpu
@pytest.mark.parametrize("H,W,nf,scale,n", [(480, 752, 1000, None, 4), (376, 1241, 2000, None, 4), (376, 1241, 2000, 5.0, 3),
                                            (480, 640, 1000, 0.2, 3),
                                            (480, 752, 1000, None, 100)])      # BASELINE.json configs[0]: a 100-frame EuRoC-shaped mono sequence, match(first, t)
def test_binding_equals_reference_classes(tmp_path, cuda_required, H, W, nf, scale, n):
    write_input(tmp_path / "in.raw", sequence(H, W, 23, n, step=(5, 2)), nf)
    args = [ref_lib.BINDING_BIN, str(tmp_path / "in.raw")] + ([str(scale)] if scale else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.stdout.strip(), r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0, d
    assert d["frames"] == n and d["num_features"] == (int(scale * nf) if scale else nf)
    assert d["keypoint_mismatches"] == 0 and d["mono_index_mismatches"] == 0
    assert d["descriptor_mismatches"] <= 1e-3 * d["keypoints"]
    assert d["match_mismatches"] == 0 and d["match_mismatches_vs_all_reference"] == 0 and d["stored_matchedobs_mismatches"] == 0
    assert d["edge_cases_ok"] is True
    assert d["keypoints"] > 0.8 * n * d["num_features"] and d["matches"] > 20 * (n - 1)


def write_two_view_input(path, seed, n, planar, unmatched):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_two_view import scene
    x1, x2, _, _, _ = scene(seed, n, planar)
    rng = np.random.default_rng(seed + 7)
    perm = rng.permutation(n + 5)
    k2 = np.zeros((n + 5, 2), np.float32)
    k2[perm[:n]] = x2
    k2[perm[n:]] = rng.uniform(0, 480, (5, 2))
    m12 = perm[:n].astype(np.int32)
    m12[rng.random(n) < unmatched] = -1
    with open(path, "wb") as f:
        f.write(np.array([n, n + 5], np.int32).tobytes()); f.write(x1.tobytes()); f.write(k2.tobytes()); f.write(m12.tobytes())
    return int((m12 >= 0).sum())


@needs_binary
def test_two_view_binding_refuses_without_gpu(tmp_path):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("CUDA device present; covered by the gpu test")
    write_two_view_input(tmp_path / "tv.raw", 1, 100, False, 0.0)
    r = subprocess.run([ref_lib.BINDING_BIN, "--two-view", str(tmp_path / "tv.raw")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr[-300:])


@needs_binary
@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,planar,unmatched", [(31, 400, False, 0.1), (32, 1200, True, 0.0), (33, 9, False, 0.0), (34, 250, True, 0.4)])
def test_two_view_binding_equals_reference_class(tmp_path, cuda_required, seed, n, planar, unmatched):
    """TwoViewScoringB200.hpp in the reference's own types: all 2 x 200 hypotheses of the reference's own solvers scored on
    the device == its CheckHomography / CheckFundamental, and the (score, inliers, matrix) handed back in place of
    FindHomography / FindFundamental == what those return, in one process."""
    nm = write_two_view_input(tmp_path / "tv.raw", seed, n, planar, unmatched)
    r = subprocess.run([ref_lib.BINDING_BIN, "--two-view", str(tmp_path / "tv.raw")], capture_output=True, text=True, timeout=300)
    assert r.stdout.strip(), r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0, d
    assert d["matches"] == nm and d["iterations"] == 200
    assert d["score_mismatches"] == 0 and d["inlier_mismatches"] == 0 and d["kept_result_mismatches"] == 0
    assert d["best_f"] >= 0 and d["score_f"] > 0
