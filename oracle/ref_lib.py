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
        src = [os.path.join(_HERE, f) for f in ("ref_driver.cpp", "ref_link_stubs.cpp", "Makefile.ref",
                                                "ref_shim/opencv2/core.hpp", "_build/liborb_oracle.so")]
        stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
        if force or stale:
            subprocess.check_call(["make", "-C", _HERE, "-f", "Makefile.ref", "-B", "REF=" + REF_ROOT], stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


def available():
    return build() is not None


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
