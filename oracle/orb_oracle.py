"""ctypes binding of oracle/orb_oracle.cpp (CPU oracle; test infrastructure, never a product path)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


class Grid(C.Structure):
    _fields_ = [("cols", C.c_int), ("rows", C.c_int), ("minX", C.c_float), ("maxX", C.c_float),
                ("minY", C.c_float), ("maxY", C.c_float)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "pattern_31.inc", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, f32p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_num_features.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_tables.argtypes = [C.c_void_p, f32p, f32p, i32p, i32p]
        L.orc_detect.restype = C.c_int
        L.orc_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                 C.c_int, i32p]
        L.orc_detect_count.restype = C.c_int
        L.orc_detect_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.orc_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.orc_get_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_get_blurred.restype = C.c_int
        L.orc_get_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_raw_count.restype = C.c_int
        L.orc_raw_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_raw.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_level_kp_count.restype = C.c_int
        L.orc_level_kp_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_level_kps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_resize_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.orc_gauss7_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_fast_u8.restype = C.c_int
        L.orc_fast_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_bgr2gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_quadtree.restype = C.c_int
        L.orc_quadtree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_sort_sized.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_match_window.restype = C.c_int
        L.orc_match_window.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.POINTER(Grid), C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orc_grid_query.restype = C.c_int
        L.orc_grid_query.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Grid), C.c_float, C.c_float, C.c_float,
                                     C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_match_bf_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_check_homography.restype = C.c_float
        L.orc_check_homography.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.orc_check_fundamental.restype = C.c_float
        L.orc_check_fundamental.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_undistort_radtan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_undistort_kb8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def grid_for(W, H, bounds=None):
    """FeatureGrid::setImageBounds (FeatureGrid.cpp:100-113) for a pinhole camera."""
    b = bounds or (0.0, float(W), 0.0, float(H))
    return Grid(W // 10, H // 10, b[0], b[1], b[2], b[3])


class OrbOracle:
    """CPU restatement of OP::FtDtOrbSlam (detector)."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.L.orc_create(nfeatures, scale, nlevels, ini_th, min_th))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_num_features(self, n):
        self.nfeatures = n
        self.L.orc_set_num_features(self.h, n)

    def tables(self):
        s = np.zeros(self.nlevels, np.float32); i = np.zeros(self.nlevels, np.float32)
        q = np.zeros(self.nlevels, np.int32); u = np.zeros(16, np.int32)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int)
        self.L.orc_get_tables(self.h, s.ctypes.data_as(f32p), i.ctypes.data_as(f32p), q.ctypes.data_as(i32p),
                              u.ctypes.data_as(i32p))
        return s, i, q, u

    def detect(self, img, cap=None):
        """Returns (mono_index, keypoints[KP_DTYPE], descriptors[n,32])."""
        assert img.dtype == np.uint8 and img.ndim == 2
        H, W = img.shape
        cap = cap or (self.nfeatures * 2 + 64 * self.nlevels)
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = self.L.orc_detect(self.h, _p(img), W, H, img.strides[0], _p(kps), _p(desc), cap, C.byref(n))
        if mono == -3:
            return self.detect(img, cap=n.value)
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def detect_count(self, img):
        H, W = img.shape
        return self.L.orc_detect_count(self.h, _p(img), W, H, img.strides[0])

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.orc_get_level(self.h, l, _p(out))
        return out

    def blurred(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        return out if self.L.orc_get_blurred(self.h, l, _p(out)) else None

    def raw(self, l):
        n = self.L.orc_raw_count(self.h, l)
        out = np.zeros((n, 3), np.float32)
        if n:
            self.L.orc_get_raw(self.h, l, _p(out))
        return out

    def level_kps(self, l):
        n = self.L.orc_level_kp_count(self.h, l)
        k = np.zeros(n, KP_DTYPE); d = np.zeros((n, 32), np.uint8)
        if n:
            self.L.orc_get_level_kps(self.h, l, _p(k), _p(d))
        return k, d


def resize_u8(src, dw, dh):
    src = np.ascontiguousarray(src)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def gauss7_u8(src):
    src = np.ascontiguousarray(src)
    dst = np.zeros_like(src)
    lib().orc_gauss7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast_u8(img, t):
    """cv::FAST(img, t, nms=True) restatement; accepts non-contiguous row-strided views."""
    assert img.strides[1] == 1
    cap = max(16, (img.shape[0] * img.shape[1]) // 4 + 16)
    out = np.zeros((cap, 3), np.float32)
    n = lib().orc_fast_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], t, _p(out), cap)
    return out[:n].copy()


def bgr2gray(bgr):
    """cv::cvtColor(COLOR_BGR2GRAY) restatement on an [H, W, 3] uint8 image."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    out = np.zeros(bgr.shape[:2], np.uint8)
    lib().orc_bgr2gray(_p(bgr), bgr.shape[1], bgr.shape[0], bgr.strides[0], _p(out), out.strides[0])
    return out


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def quadtree(xyr, minX, maxX, minY, maxY, N):
    xyr = np.ascontiguousarray(xyr, np.float32)
    kept = np.zeros(max(16, N + 8 + len(xyr)), np.int32)
    n = lib().orc_quadtree(_p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(kept), len(kept))
    if n < 0:
        raise ValueError("degenerate quadtree geometry")
    return kept[:n].copy()


def sort_sized(count, ulx):
    count = np.ascontiguousarray(count, np.int32); ulx = np.ascontiguousarray(ulx, np.int32)
    perm = np.zeros(len(count), np.int32)
    lib().orc_sort_sized(_p(count), _p(ulx), len(count), _p(perm))
    return perm


def match_window(k1, ud1, d1, k2, ud2, d2, grid, window=100.0, nnratio=0.6, th_low=50, check_ori=True):
    """FtAssocOrbSlam::matchV restatement. Returns matches12 (int32[n1])."""
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2)
    ud1 = np.ascontiguousarray(ud1, np.float32); ud2 = np.ascontiguousarray(ud2, np.float32)
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    m = np.full(max(1, len(k1)), -1, np.int32)
    lib().orc_match_window(_p(k1), _p(ud1), _p(d1), len(k1), _p(k2), _p(ud2), _p(d2), len(k2), C.byref(grid),
                           window, nnratio, th_low, int(check_ori), _p(m))
    return m[:len(k1)]


def grid_query(k2, ud2, grid, x, y, r, min_level, max_level):
    k2 = np.ascontiguousarray(k2); ud2 = np.ascontiguousarray(ud2, np.float32)
    out = np.zeros(max(1, len(k2)), np.int32)
    n = lib().orc_grid_query(_p(k2), _p(ud2), len(k2), C.byref(grid), x, y, r, min_level, max_level, _p(out), len(out))
    return out[:n].copy()


def match_bf_knn2(d1, d2, norm=0, ratio=0.7):
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    n1 = len(d1)
    i0 = np.zeros(n1, np.int32); i1 = np.zeros(n1, np.int32)
    f0 = np.zeros(n1, np.float32); f1 = np.zeros(n1, np.float32); ps = np.zeros(n1, np.uint8)
    lib().orc_match_bf_knn2(_p(d1), n1, _p(d2), len(d2), norm, ratio, _p(i0), _p(i1), _p(f0), _p(f1), _p(ps))
    return i0, i1, f0, f1, ps


CAM_PINHOLE, CAM_RADTAN, CAM_KB8 = 0, 1, 2


def undistort(model, K4, D4, xy):
    """Calibration::undistort restatement: model 0 identity, 1 cv::undistortPoints (RadTan, P = K), 2 cv::fisheye (KB8)."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    if model == CAM_PINHOLE:
        return xy.copy()
    K4 = np.ascontiguousarray(K4, np.float32); D4 = np.ascontiguousarray(D4, np.float32)
    out = np.zeros_like(xy)
    fn = lib().orc_undistort_radtan if model == CAM_RADTAN else lib().orc_undistort_kb8
    fn(_p(K4), _p(D4), _p(xy), len(xy), _p(out))
    return out


def check_homography(H21, H12, xy1, xy2, sigma=1.0, th=5.991):
    """TwoViewReconstruction::CheckHomography restatement: (score float32, inliers uint8[n])."""
    H21 = np.ascontiguousarray(H21, np.float32); H12 = np.ascontiguousarray(H12, np.float32)
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    inl = np.zeros(max(1, len(xy1)), np.uint8)
    s = lib().orc_check_homography(_p(H21), _p(H12), _p(xy1), _p(xy2), len(xy1), sigma, th, _p(inl))
    return np.float32(s), inl[:len(xy1)]


def check_fundamental(F21, xy1, xy2, sigma=1.0, th=3.841, th_score=5.991):
    """TwoViewReconstruction::CheckFundamental restatement: (score float32, inliers uint8[n])."""
    F21 = np.ascontiguousarray(F21, np.float32)
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    inl = np.zeros(max(1, len(xy1)), np.uint8)
    s = lib().orc_check_fundamental(_p(F21), _p(xy1), _p(xy2), len(xy1), sigma, th, th_score, _p(inl))
    return np.float32(s), inl[:len(xy1)]
