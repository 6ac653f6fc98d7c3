"""ctypes binding of oracle/_ref/libnav24_ref.so: the REFERENCE's own FtDtOrbSlam / FtAssocOrbSlam / FeatureGrid /
FrameMonoGrid, compiled unchanged from /root/reference by oracle/Makefile.ref (test infrastructure; the product never
loads it).  The library is built in the build container (where /root/reference exists) and travels to the GPU box as a
prebuilt file; `available()` is False where neither the file nor the reference tree exists."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import orb_oracle as oo

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libnav24_ref.so")
REF_ROOT = os.environ.get("NAV24_REFERENCE", "/root/reference")
KP_DTYPE = oo.KP_DTYPE


def build(force=False):
    """Builds the library when the reference tree is present; returns its path or None."""
    have_ref = os.path.exists(os.path.join(REF_ROOT, "core", "operators", "objDetection", "OP_FtDtOrbSlam.cpp"))
    if have_ref:
        oo.build()
        src = [os.path.join(_HERE, f) for f in ("ref_driver.cpp", "ref_driver_2v.cpp", "ref_access_2v.hpp", "ref_link_stubs.cpp", "Makefile.ref",
                                                "ref_shim/opencv2/core.hpp", "ref_shim/opencv2/core_algebra.hpp",
                                                "_build/liborb_oracle.so")]
        stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
        if force or stale:
            subprocess.check_call(["make", "-C", _HERE, "-f", "Makefile.ref", "-B", "REF=" + REF_ROOT, "_ref/libnav24_ref.so"],
                                  stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


def available():
    return build() is not None


BINDING_BIN = os.path.join(_HERE, "_ref", "test_ref_binding")


def build_binding_test():
    """tests/cpp/test_ref_binding.cpp: the nav24-side binding (nav24_b200/host/ref_binding) compiled against the reference's
    own headers, linked to the reference's own classes and to libnav24orb.so.  Built where the reference tree exists;
    elsewhere the prebuilt binary is used.  Returns its path or None."""
    product = os.path.join(os.path.dirname(_HERE), "nav24_b200", "libnav24orb.so")      # linked, never loaded by this module
    if build() is not None and os.path.exists(product) and os.path.exists(os.path.join(REF_ROOT, "core", "operators", "objDetection", "OP_FtDt.hpp")):
        subprocess.check_call(["make", "-C", _HERE, "-f", "Makefile.ref", "REF=" + REF_ROOT, "_ref/test_ref_binding"],
                              stdout=subprocess.DEVNULL)
    return BINDING_BIN if os.path.exists(BINDING_BIN) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libnav24_ref.so is not built and /root/reference is not here")
        oo.lib()                      # liborb_oracle.so (the cv2-pinned primitives the container shim forwards to)
        L = C.CDLL(so)
        i32p = C.POINTER(C.c_int)
        L.ref_orb_create.restype = C.c_void_p
        L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_orb_destroy.argtypes = [C.c_void_p]
        L.ref_orb_scale_num_features.argtypes = [C.c_void_p, C.c_float]
        L.ref_orb_get_num_features.restype = C.c_int
        L.ref_orb_get_num_features.argtypes = [C.c_void_p]
        L.ref_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ref_orb_detect.restype = C.c_int
        L.ref_orb_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, i32p]
        L.ref_orb_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.ref_orb_get_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_quadtree.restype = C.c_int
        L.ref_quadtree.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 5 + [C.c_void_p, C.c_int]
        L.ref_match_window.restype = C.c_int
        L.ref_match_window.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.ref_grid_query.restype = C.c_int
        L.ref_grid_query.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                     C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
        u8p = C.c_void_p
        L.ref_2v_create.restype = C.c_void_p
        L.ref_2v_create.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.ref_2v_destroy.argtypes = [C.c_void_p]
        L.ref_2v_reconstruct.restype = C.c_int
        L.ref_2v_reconstruct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, i32p]
        L.ref_2v_get_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_2v_check_h.restype = C.c_float
        L.ref_2v_check_h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, u8p]
        L.ref_2v_check_f.restype = C.c_float
        L.ref_2v_check_f.argtypes = [C.c_void_p, C.c_void_p, C.c_float, u8p]
        L.ref_2v_hypotheses.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_2v_find_h.argtypes = [C.c_void_p, u8p, C.c_void_p, C.c_void_p]
        L.ref_2v_find_f.argtypes = [C.c_void_p, u8p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefOrb:
    """NAV24::OP::FtDtOrbSlam itself."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.L.ref_orb_create(nfeatures, scale, nlevels, ini_th, min_th))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_orb_destroy(self.h)
            self.h = None

    def scale_num_features(self, s):          # FtDt::scaleNumFeatures (OP_FtDt.hpp:22)
        self.L.ref_orb_scale_num_features(self.h, float(s))
        return self.L.ref_orb_get_num_features(self.h)

    def tables(self):
        s = np.zeros(self.nlevels, np.float32); i = np.zeros(self.nlevels, np.float32)
        q = np.zeros(self.nlevels, np.int32); u = np.zeros(16, np.int32)
        self.L.ref_orb_tables(self.h, _p(s), _p(i), _p(q), _p(u))
        return s, i, q, u

    def detect(self, img):
        """FtDtOrbSlam::detect: (mono_index, keypoints, descriptors) in the frame's observation order."""
        H, W = (img.shape if img is not None else (0, 0))
        cap = 4 * max(self.L.ref_orb_get_num_features(self.h), 100) + 64 * self.nlevels
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = self.L.ref_orb_detect(self.h, _p(img) if img is not None else None, W, H, img.strides[0] if img is not None else 0,
                                     _p(kps), _p(desc), cap, C.byref(n))
        assert n.value <= cap
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.ref_orb_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.ref_orb_get_level(self.h, l, _p(out))
        return out


def quadtree(xyr, minX, maxX, minY, maxY, N):
    """FtDtOrbSlam::DistributeOctTree: indices of the kept keys in the reference's output order."""
    xyr = np.ascontiguousarray(xyr, np.float32)
    kept = np.zeros(max(16, N + 8 + len(xyr)), np.int32)
    n = lib().ref_quadtree(_p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(kept), len(kept))
    return kept[:n].copy()


def match_window(k1, ud1, d1, k2, ud2, d2, W, H, bounds=None, nnratio=0.6, check_ori=True):
    """FeatureGrid::setImageBounds(Size(W, H), bounds) + FtAssocOrbSlam::matchV(frame1, frame2).  W <= 0 keeps the grid
    configuration of the previous call (the reference configures it once per process; bench.py's worker threads set it
    once and then match concurrently)."""
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2)
    ud1 = np.ascontiguousarray(ud1, np.float32); ud2 = np.ascontiguousarray(ud2, np.float32)
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    b = np.asarray(bounds or (0.0, float(W), 0.0, float(H)), np.float32)
    m = np.full(max(1, len(k1)), -1, np.int32)
    nm = lib().ref_match_window(_p(k1), _p(ud1), _p(d1), len(k1), _p(k2), _p(ud2), _p(d2), len(k2), W, H, _p(b),
                                nnratio, int(check_ori), _p(m))
    assert nm >= 0, "FtAssocOrbSlam::match(f1, f2) stored a different match count than matchV returned"
    return m[:len(k1)]


def grid_query(k2, ud2, W, H, x, y, r, min_level, max_level, bounds=None):
    k2 = np.ascontiguousarray(k2); ud2 = np.ascontiguousarray(ud2, np.float32)
    b = np.asarray(bounds or (0.0, float(W), 0.0, float(H)), np.float32)
    out = np.zeros(max(1, len(k2)), np.int32)
    n = lib().ref_grid_query(_p(k2), _p(ud2), len(k2), W, H, _p(b), x, y, r, min_level, max_level, _p(out), len(out))
    return out[:n].copy()


class RefTwoView:
    """NAV24::OP::TwoViewReconstruction itself (core/operators/mapInit/OP_2ViewReconstruction.cpp), SURVEY 8(f)-4.
    `reconstruct` runs the public entry point (match list, RANSAC sets, FindHomography / FindFundamental in two threads,
    model selection); afterwards the private scoring members can be called on the match list it built."""

    def __init__(self, K=None, sigma=1.0, iterations=200):
        self.L = lib()
        K = np.ascontiguousarray(K if K is not None else [[458.0, 0, 367.0], [0, 457.0, 248.0], [0, 0, 1]], np.float32)
        self.iterations = iterations
        self.sigma = float(sigma)
        self.n = 0
        self.h = C.c_void_p(self.L.ref_2v_create(_p(K), self.sigma, iterations))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_2v_destroy(self.h)
            self.h = None

    def reconstruct(self, xy1, xy2, matches12):
        """Reconstruct(vKeys1, vKeys2, vMatches12, ...): matches12[i] = index into xy2 or -1.  Needs >= 8 matches (the
        reference draws 8 of them per RANSAC set without checking).  Returns its bool."""
        xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
        m = np.ascontiguousarray(matches12, np.int32)
        assert len(m) == len(xy1) and (m >= 0).sum() >= 8
        n = C.c_int(0)
        ok = self.L.ref_2v_reconstruct(self.h, _p(xy1), len(xy1), _p(xy2), len(xy2), _p(m), C.byref(n))
        self.n = n.value
        return bool(ok)

    def matches(self):
        """(xy1, xy2, sets): the matched points in mvMatches12 order and the 8-point sets of every iteration."""
        a = np.zeros((self.n, 2), np.float32); b = np.zeros((self.n, 2), np.float32)
        s = np.zeros((self.iterations, 8), np.int32)
        self.L.ref_2v_get_matches(self.h, _p(a), _p(b), _p(s))
        return a, b, s

    def check_homography(self, H21, H12, sigma=None):
        H21 = np.ascontiguousarray(H21, np.float32).reshape(9); H12 = np.ascontiguousarray(H12, np.float32).reshape(9)
        inl = np.zeros(max(1, self.n), np.uint8)
        s = self.L.ref_2v_check_h(self.h, _p(H21), _p(H12), self.sigma if sigma is None else sigma, _p(inl))
        return np.float32(s), inl[:self.n]

    def check_fundamental(self, F21, sigma=None):
        F21 = np.ascontiguousarray(F21, np.float32).reshape(9)
        inl = np.zeros(max(1, self.n), np.uint8)
        s = self.L.ref_2v_check_f(self.h, _p(F21), self.sigma if sigma is None else sigma, _p(inl))
        return np.float32(s), inl[:self.n]

    def hypotheses(self):
        """(H21, H12, F21), iterations x 9 each: T2inv*Hn*T1, its inverse and T2t*Fn*T1 of every RANSAC iteration, from the
        reference's own Normalize / ComputeH21 / ComputeF21."""
        H21 = np.zeros((self.iterations, 9), np.float32); H12 = np.zeros_like(H21); F21 = np.zeros_like(H21)
        self.L.ref_2v_hypotheses(self.h, _p(H21), _p(H12), _p(F21))
        return H21, H12, F21

    def find_homography(self):
        """FindHomography: (score, inliers, H21) its selection loop keeps."""
        inl = np.zeros(max(1, self.n), np.uint8); sc = np.zeros(1, np.float32); H = np.zeros(9, np.float32)
        self.L.ref_2v_find_h(self.h, _p(inl), _p(sc), _p(H))
        return sc[0], inl[:self.n], H

    def find_fundamental(self):
        inl = np.zeros(max(1, self.n), np.uint8); sc = np.zeros(1, np.float32); F = np.zeros(9, np.float32)
        self.L.ref_2v_find_f(self.h, _p(inl), _p(sc), _p(F))
        return sc[0], inl[:self.n], F
