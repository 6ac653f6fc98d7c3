"""ctypes binding of libnav24orb.so — the C ABI declared in include/nav24_orb.h.

This is plumbing for tests and bench.py; the product is the shared library.  There is no CPU
fallback: if the library is missing this module raises, and without a CUDA device
nav24_orb_create fails with NAV24_E_CUDA.
"""
import ctypes as C
import os
import re
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NAV24_LIB") or os.path.join(_HERE, "libnav24orb.so")     # NAV24_LIB: kernel-variant experiments
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nav24_orb.h")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

OK, E_BADARG, E_GEOMETRY, E_CAPACITY, E_OVERFLOW, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4, -5, -6
NORM_HAMMING, NORM_L2_U8 = 0, 1
CAM_PINHOLE, CAM_RADTAN, CAM_KB8 = 0, 1, 2


class Params(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("scale_factor", C.c_float), ("n_levels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("raw_keys_per_kpx", C.c_int32)]


class GridCfg(C.Structure):
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("min_x", C.c_float), ("max_x", C.c_float),
                ("min_y", C.c_float), ("max_y", C.c_float)]


class Camera(C.Structure):
    """nav24_camera: model + K (fx fy cx cy) + 4 distortion coefficients, all float like the reference's cv::Mat."""
    _fields_ = [("model", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("d", C.c_float * 4)]

    @staticmethod
    def make(model, K4=(1, 1, 0, 0), D4=(0, 0, 0, 0)):
        return Camera(model, K4[0], K4[1], K4[2], K4[3], (C.c_float * 4)(*D4))


class Nav24Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nav24 error {code}: {msg}")
        self.code = code


def declared_symbols():
    """Every function name declared in include/nav24_orb.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nav24_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.nav24_abi_version.restype = C.c_int
    L.nav24_orb_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.nav24_orb_destroy.argtypes = [vp]; L.nav24_orb_destroy.restype = None
    L.nav24_orb_set_num_features.argtypes = [vp, C.c_int]
    L.nav24_orb_get_num_features.argtypes = [vp]
    L.nav24_orb_get_tables.argtypes = [vp, vp, vp, vp]
    L.nav24_last_error_string.argtypes = [vp]; L.nav24_last_error_string.restype = C.c_char_p
    L.nav24_orb_detect.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, C.c_int, ip]
    L.nav24_orb_detect_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, C.c_int, vp, vp]
    L.nav24_orb_detect_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t]
    L.nav24_orb_fetch.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.nav24_orb_detect_match_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, C.c_int, vp, vp,
                                               C.c_int, vp, C.POINTER(GridCfg), C.c_float, C.c_float, C.c_int, C.c_int, vp,
                                               C.c_int, vp]
    L.nav24_orb_detect_match_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int, vp,
                                                C.POINTER(GridCfg), C.c_float, C.c_float, C.c_int, C.c_int]
    L.nav24_match_fetch.argtypes = [vp, vp, C.c_int, vp]
    L.nav24_orb_fetch_range.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    L.nav24_match_fetch_range.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.nav24_orb_sync.argtypes = [vp]
    L.nav24_orb_max_keypoints.argtypes = [vp]
    L.nav24_orb_get_level.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, ip, ip]
    L.nav24_orb_get_raw_keys.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
    L.nav24_orb_get_level_keypoints.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
    L.nav24_orb_stage_ms.argtypes = [vp, vp]
    L.nav24_orb_launch_count.argtypes = [vp]; L.nav24_orb_launch_count.restype = C.c_longlong
    L.nav24_orb_stage_ms_sum.argtypes = [vp, vp, ip, C.c_int]
    L.nav24_orb_timer_start.argtypes = [vp]
    L.nav24_orb_timer_stop.argtypes = [vp, vp]
    L.nav24_match_window.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.POINTER(GridCfg), C.c_float, C.c_float,
                                     C.c_int, C.c_int, vp]
    L.nav24_match_window_batch.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(GridCfg), C.c_float,
                                           C.c_float, C.c_int, C.c_int, vp, vp]
    L.nav24_match_window_frames.argtypes = [vp, C.c_int, vp, C.POINTER(GridCfg), C.c_float, C.c_float, C.c_int, C.c_int, vp,
                                            C.c_int, vp]
    L.nav24_undistort_points.argtypes = [vp, C.POINTER(Camera), vp, C.c_int, vp]
    L.nav24_orb_set_camera.argtypes = [vp, C.POINTER(Camera)]
    L.nav24_orb_fetch_undistorted.argtypes = [vp, vp, C.c_int]
    L.nav24_match_bf_knn2.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, vp, vp, vp, vp, vp]
    L.nav24_debug_sort_u32.argtypes = [vp, vp, C.c_int, vp]
    L.nav24_debug_last_kernel_ms.argtypes = [vp, vp]
    L.nav24_ingest_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.nav24_ingest_destroy.argtypes = [vp]; L.nav24_ingest_destroy.restype = None
    L.nav24_ingest_slot.argtypes = [vp, C.c_int]; L.nav24_ingest_slot.restype = vp
    L.nav24_ingest_slot_bytes.argtypes = [vp]; L.nav24_ingest_slot_bytes.restype = C.c_size_t
    L.nav24_ingest_detect.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    L.nav24_ingest_detect_match.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.POINTER(GridCfg),
                                            C.c_float, C.c_float, C.c_int, C.c_int, vp, C.c_int, vp]
    L.nav24_two_view_score.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp,
                                       vp, vp, ip, ip]
    L.nav24_two_view_score_kept.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp,
                                       vp, vp, ip, ip]
    L.nav24_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.nav24_host_free.argtypes = [vp]
    L.nav24_device_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.nav24_device_free.argtypes = [vp]
    L.nav24_memcpy_h2d.argtypes = [vp, vp, C.c_size_t]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def grid_for(W, H, bounds=None):
    """FeatureGrid::setImageBounds (FeatureGrid.cpp:100-113) for a pinhole camera."""
    b = bounds or (0.0, float(W), 0.0, float(H))
    return GridCfg(W // 10, H // 10, b[0], b[1], b[2], b[3])


def image_bounds(undistort, W, H, calibrated):
    """Calibration::computeImageBounds (Calibration.cpp:196-228): the undistorted image corners for a distorted camera,
    the image rectangle otherwise.  undistort: callable xy[n,2] -> ud[n,2].  Returns (minX, maxX, minY, maxY)."""
    if calibrated:
        return (0.0, float(W), 0.0, float(H))
    c = undistort(np.array([[0, 0], [W, 0], [0, H], [W, H]], np.float32)).reshape(-1)
    return (float(min(c[0], c[4])), float(max(c[2], c[6])), float(min(c[1], c[3])), float(max(c[5], c[7])))


def pinned_empty(shape, dtype=np.uint8):
    """numpy array backed by pinned host memory (cudaHostAlloc); keep the returned array alive."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = C.c_void_p()
    rc = lib().nav24_host_alloc(max(n, 1), C.byref(ptr))
    if rc != OK:
        raise Nav24Error(rc, "cudaHostAlloc failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(ptr.value)
    weakref.finalize(buf, lib().nav24_host_free, C.c_void_p(ptr.value))      # released when the last view of it is gone
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr


class OrbContext:
    """Thin owner of a nav24_orb* handle."""

    def __init__(self, n_features=1000, scale_factor=1.2, n_levels=8, ini_th_fast=20, min_th_fast=7, device=0,
                 raw_keys_per_kpx=0):
        self.L = lib()
        self.n_levels = n_levels
        self.h = C.c_void_p()
        prm = Params(n_features, scale_factor, n_levels, ini_th_fast, min_th_fast, raw_keys_per_kpx)
        rc = self.L.nav24_orb_create(C.byref(prm), device, C.byref(self.h))
        if rc != OK:
            self.h = None
            raise Nav24Error(rc, "nav24_orb_create failed (a CUDA device is required; there is no CPU fallback)")

    def close(self):
        if getattr(self, "h", None):
            self.L.nav24_orb_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise Nav24Error(rc, self.L.nav24_last_error_string(self.h).decode())
        return rc

    def set_num_features(self, n):
        self._check(self.L.nav24_orb_set_num_features(self.h, n))

    def num_features(self):
        return self.L.nav24_orb_get_num_features(self.h)

    def tables(self):
        s = np.zeros(self.n_levels, np.float32); i = np.zeros(self.n_levels, np.float32); q = np.zeros(self.n_levels, np.int32)
        self._check(self.L.nav24_orb_get_tables(self.h, _p(s), _p(i), _p(q)))
        return s, i, q

    def max_keypoints(self):
        return self.L.nav24_orb_max_keypoints(self.h)

    def detect(self, img, cap=None):
        """One host image -> (mono_index, kps[KP_DTYPE], desc[n,32])."""
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        H, W = img.shape
        cap = cap or self.max_keypoints()
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rc = self.L.nav24_orb_detect(self.h, _p(img), W, H, img.strides[0], _p(kps), _p(desc), cap, C.byref(n))
        if rc == E_CAPACITY:
            return self.detect(img, cap=n.value)
        self._check(rc)
        return rc, kps[:n.value].copy(), desc[:n.value].copy()

    def detect_batch(self, frames, cap=None, kps=None, desc=None):
        """frames: uint8 [B,H,W] (C-contiguous rows). Returns (n[B], mono[B], kps[B,cap], desc[B,cap,32])."""
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.strides[2] == 1
        B, H, W = frames.shape
        cap = cap or self.max_keypoints()
        if kps is None:
            kps = np.zeros((B, cap), KP_DTYPE)
        if desc is None:
            desc = np.zeros((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32); mono = np.zeros(B, np.int32)
        rc = self.L.nav24_orb_detect_batch(self.h, _p(frames), B, W, H, frames.strides[1], frames.strides[0], _p(kps), _p(desc),
                                           cap, _p(n), _p(mono))
        if rc == E_CAPACITY:
            return self.detect_batch(frames, cap=int(n.max()))
        self._check(rc)
        return n, mono, kps, desc

    def detect_match_batch(self, frames, pairs, grid, cap=None, kps=None, desc=None, matches=None, window=100.0, nnratio=0.6,
                           th_low=50, check_ori=True):
        """Fused detect + window matching on host frames. Returns (n, mono, kps, desc, matches12[P,cap], n_matches[P])."""
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.strides[2] == 1
        B, H, W = frames.shape
        cap = cap or self.max_keypoints()
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        P = len(pairs)
        kps = np.zeros((B, cap), KP_DTYPE) if kps is None else kps
        desc = np.zeros((B, cap, 32), np.uint8) if desc is None else desc
        mcap = max(cap, self.max_keypoints())
        matches = np.full((max(P, 1), mcap), -1, np.int32) if matches is None else matches
        n = np.zeros(B, np.int32); mono = np.zeros(B, np.int32); nm = np.zeros(max(P, 1), np.int32)
        rc = self.L.nav24_orb_detect_match_batch(self.h, _p(frames), B, W, H, frames.strides[1], frames.strides[0], _p(kps),
                                                 _p(desc), cap, _p(n), _p(mono), P, _p(pairs), C.byref(grid), window, nnratio,
                                                 th_low, int(check_ori), _p(matches), matches.shape[1], _p(nm))
        if rc == E_CAPACITY and cap < int(n.max()):
            return self.detect_match_batch(frames, pairs, grid, cap=int(n.max()), window=window, nnratio=nnratio,
                                           th_low=th_low, check_ori=check_ori)
        self._check(rc)
        return n, mono, kps, desc, matches[:P], nm[:P]

    def detect_match_device(self, dptr, B, W, H, stride, frame_stride, pairs, grid, window=100.0, nnratio=0.6, th_low=50,
                            check_ori=True):
        """Enqueue only (device-resident frames); fetch with fetch() / match_fetch()."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        self._check(self.L.nav24_orb_detect_match_device(self.h, C.c_void_p(dptr), B, W, H, stride, frame_stride, len(pairs),
                                                         _p(pairs), C.byref(grid), window, nnratio, th_low, int(check_ori)))

    def match_fetch(self, P, want_matches=True, out=None):
        cap = self.max_keypoints()
        m = out if out is not None else (np.full((P, cap), -1, np.int32) if want_matches else None)
        nm = np.zeros(P, np.int32)
        self._check(self.L.nav24_match_fetch(self.h, _p(m), cap, _p(nm)))
        return m, nm

    def detect_device(self, dptr, B, W, H, stride, frame_stride):
        self._check(self.L.nav24_orb_detect_device(self.h, C.c_void_p(dptr), B, W, H, stride, frame_stride))

    def fetch(self, B, cap=None, want_data=True):
        cap = cap or self.max_keypoints()
        n = np.zeros(B, np.int32); mono = np.zeros(B, np.int32)
        kps = np.zeros((B, cap), KP_DTYPE) if want_data else None
        desc = np.zeros((B, cap, 32), np.uint8) if want_data else None
        self._check(self.L.nav24_orb_fetch(self.h, _p(kps), _p(desc), cap, _p(n), _p(mono)))
        return n, mono, kps, desc

    def fetch_range(self, f0, nf, cap=None):
        """Results of frames [f0, f0+nf) of the last (device-resident) detect call."""
        cap = cap or self.max_keypoints()
        n = np.zeros(nf, np.int32); mono = np.zeros(nf, np.int32)
        kps = np.zeros((nf, cap), KP_DTYPE); desc = np.zeros((nf, cap, 32), np.uint8)
        self._check(self.L.nav24_orb_fetch_range(self.h, f0, nf, _p(kps), _p(desc), cap, _p(n), _p(mono)))
        return n, mono, kps, desc

    def match_fetch_range(self, p0, npairs):
        cap = self.max_keypoints()
        m = np.full((npairs, cap), -1, np.int32); nm = np.zeros(npairs, np.int32)
        self._check(self.L.nav24_match_fetch_range(self.h, p0, npairs, _p(m), cap, _p(nm)))
        return m, nm

    def sync(self):
        self._check(self.L.nav24_orb_sync(self.h))

    def level(self, frame, level, blurred=False):
        w, h = C.c_int(), C.c_int()
        self._check(self.L.nav24_orb_get_level(self.h, frame, level, int(blurred), None, 0, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self._check(self.L.nav24_orb_get_level(self.h, frame, level, int(blurred), _p(out), out.strides[0], C.byref(w), C.byref(h)))
        return out

    def raw_keys(self, frame, level):
        n = self._check(self.L.nav24_orb_get_raw_keys(self.h, frame, level, None, 0))
        out = np.zeros((n, 3), np.float32)
        if n:
            self._check(self.L.nav24_orb_get_raw_keys(self.h, frame, level, _p(out), n))
        return out

    def level_keypoints(self, frame, level):
        n = self._check(self.L.nav24_orb_get_level_keypoints(self.h, frame, level, None, 0))
        out = np.zeros(n, KP_DTYPE)
        if n:
            self._check(self.L.nav24_orb_get_level_keypoints(self.h, frame, level, _p(out), n))
        return out

    def stage_ms(self):
        ms = np.zeros(5, np.float32)
        self._check(self.L.nav24_orb_stage_ms(self.h, _p(ms)))
        return ms

    def stage_ms_sum(self, reset=True):
        ms = np.zeros(5, np.float32); calls = C.c_int(0)
        self._check(self.L.nav24_orb_stage_ms_sum(self.h, _p(ms), C.byref(calls), int(reset)))
        return ms, calls.value

    def timer_start(self):
        self._check(self.L.nav24_orb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self.L.nav24_orb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.L.nav24_orb_launch_count(self.h))

    # ---- camera ----
    def undistort_points(self, cam, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros_like(xy)
        self._check(self.L.nav24_undistort_points(self.h, C.byref(cam), _p(xy), len(xy), _p(out)))
        return out

    def set_camera(self, cam):
        self._check(self.L.nav24_orb_set_camera(self.h, C.byref(cam) if cam is not None else None))

    def fetch_undistorted(self, B, cap=None):
        cap = cap or self.max_keypoints()
        ud = np.zeros((B, cap, 2), np.float32)
        self._check(self.L.nav24_orb_fetch_undistorted(self.h, _p(ud), cap))
        return ud

    # ---- matchers ----
    def match_window(self, k1, ud1, d1, k2, ud2, d2, grid, window=100.0, nnratio=0.6, th_low=50, check_ori=True):
        k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2)
        ud1 = np.ascontiguousarray(ud1, np.float32); ud2 = np.ascontiguousarray(ud2, np.float32)
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        m = np.full(max(1, len(k1)), -1, np.int32)
        nm = self._check(self.L.nav24_match_window(self.h, _p(k1), _p(ud1), _p(d1), len(k1), _p(k2), _p(ud2), _p(d2), len(k2),
                                                   C.byref(grid), window, nnratio, th_low, int(check_ori), _p(m)))
        return m[:len(k1)], nm

    def match_window_frames_async(self, pairs, grid, window=100.0, nnratio=0.6, th_low=50, check_ori=True):
        """Enqueue only; results stay on the device (no host synchronisation)."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        self._check(self.L.nav24_match_window_frames(self.h, len(pairs), _p(pairs), C.byref(grid), window, nnratio, th_low,
                                                     int(check_ori), None, 0, None))

    def match_window_frames(self, pairs, grid, window=100.0, nnratio=0.6, th_low=50, check_ori=True, want_matches=True,
                            out=None):
        """out: optional preallocated int32 [P, max_keypoints] (e.g. pinned) receiving matches12."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        P = len(pairs)
        cap = self.max_keypoints()
        if out is not None:
            assert out.dtype == np.int32 and out.shape == (P, cap)
        m = out if out is not None else (np.full((P, cap), -1, np.int32) if want_matches else None)
        nm = np.zeros(P, np.int32)
        self._check(self.L.nav24_match_window_frames(self.h, P, _p(pairs), C.byref(grid), window, nnratio, th_low, int(check_ori),
                                                     _p(m), cap, _p(nm)))
        return m, nm

    def two_view_score(self, xy1, xy2, H21=None, H12=None, F21=None, sigma=1.0, th_h=5.991, th_f=3.841, th_score=5.991,
                       want_inliers=True):
        """CheckHomography / CheckFundamental of every hypothesis.  Returns dict(score_h, score_f, inliers_h, inliers_f,
        best_h, best_f) (entries of a skipped model are None)."""
        xy1 = np.ascontiguousarray(xy1, np.float32).reshape(-1, 2); xy2 = np.ascontiguousarray(xy2, np.float32).reshape(-1, 2)
        n = len(xy1)
        H21 = None if H21 is None else np.ascontiguousarray(H21, np.float32).reshape(-1, 9)
        H12 = None if H12 is None else np.ascontiguousarray(H12, np.float32).reshape(-1, 9)
        F21 = None if F21 is None else np.ascontiguousarray(F21, np.float32).reshape(-1, 9)
        nh = len(H21) if H21 is not None else len(F21)
        sh = np.zeros(nh, np.float32) if H21 is not None else None
        sf = np.zeros(nh, np.float32) if F21 is not None else None
        ih = np.zeros((nh, n), np.uint8) if (H21 is not None and want_inliers) else None
        i_f = np.zeros((nh, n), np.uint8) if (F21 is not None and want_inliers) else None
        bh, bf = C.c_int(-1), C.c_int(-1)
        self._check(self.L.nav24_two_view_score(self.h, _p(xy1), _p(xy2), n, _p(H21), _p(H12), _p(F21), nh, sigma, th_h, th_f,
                                                th_score, _p(sh), _p(sf), _p(ih), _p(i_f), C.byref(bh), C.byref(bf)))
        return {"score_h": sh, "score_f": sf, "inliers_h": ih, "inliers_f": i_f, "best_h": bh.value, "best_f": bf.value}

    def two_view_score_kept(self, xy1, xy2, H21=None, H12=None, F21=None, sigma=1.0, th_h=5.991, th_f=3.841, th_score=5.991):
        """What FindHomography / FindFundamental hand back: all scores, the kept iteration and only ITS inlier mask.
        Returns dict(score_h, score_f, kept_inliers_h, kept_inliers_f, best_h, best_f)."""
        xy1 = np.ascontiguousarray(xy1, np.float32).reshape(-1, 2); xy2 = np.ascontiguousarray(xy2, np.float32).reshape(-1, 2)
        n = len(xy1)
        H21 = None if H21 is None else np.ascontiguousarray(H21, np.float32).reshape(-1, 9)
        H12 = None if H12 is None else np.ascontiguousarray(H12, np.float32).reshape(-1, 9)
        F21 = None if F21 is None else np.ascontiguousarray(F21, np.float32).reshape(-1, 9)
        nh = len(H21) if H21 is not None else len(F21)
        sh = np.zeros(nh, np.float32) if H21 is not None else None
        sf = np.zeros(nh, np.float32) if F21 is not None else None
        kh = np.full(max(n, 1), 255, np.uint8) if H21 is not None else None      # (255: the call must overwrite every entry)
        kf = np.full(max(n, 1), 255, np.uint8) if F21 is not None else None
        bh, bf = C.c_int(-1), C.c_int(-1)
        self._check(self.L.nav24_two_view_score_kept(self.h, _p(xy1), _p(xy2), n, _p(H21), _p(H12), _p(F21), nh, sigma, th_h, th_f,
                                                     th_score, _p(sh), _p(sf), _p(kh), _p(kf), C.byref(bh), C.byref(bf)))
        return {"score_h": sh, "score_f": sf, "kept_inliers_h": None if kh is None else kh[:n], "kept_inliers_f": None if kf is None else kf[:n],
                "best_h": bh.value, "best_f": bf.value}

    def debug_sort(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        perm = np.zeros(len(keys), np.int32)
        self._check(self.L.nav24_debug_sort_u32(self.h, _p(keys), len(keys), _p(perm)))
        return perm

    def match_bf_knn2(self, d1, d2, norm=NORM_HAMMING, ratio=0.7):
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        n1 = len(d1)
        i0 = np.zeros(n1, np.int32); i1 = np.zeros(n1, np.int32)
        f0 = np.zeros(n1, np.float32); f1 = np.zeros(n1, np.float32); ps = np.zeros(n1, np.uint8)
        self._check(self.L.nav24_match_bf_knn2(self.h, _p(d1), n1, _p(d2), len(d2), norm, ratio, _p(i0), _p(i1), _p(f0), _p(f1), _p(ps)))
        return i0, i1, f0, f1, ps


class IngestRing:
    """nav24_ingest: pinned slots the camera decodes into (grey or interleaved BGR), detect without a second host copy."""

    def __init__(self, ctx, width, height, channels, n_slots):
        self.ctx, self.w, self.h, self.ch, self.n = ctx, width, height, channels, n_slots
        self.h_ = C.c_void_p()
        ctx._check(ctx.L.nav24_ingest_create(ctx.h, width, height, channels, n_slots, C.byref(self.h_)))

    def close(self):
        if getattr(self, "h_", None):
            self.ctx.L.nav24_ingest_destroy(self.h_)
            self.h_ = None

    __del__ = close

    def slot(self, k):
        """numpy view of slot k: [H, W] (grey) or [H, W, 3] (BGR)."""
        ptr = self.ctx.L.nav24_ingest_slot(self.h_, k)
        if not ptr:
            raise IndexError("slot outside the ring")
        nb = self.ctx.L.nav24_ingest_slot_bytes(self.h_)
        buf = (C.c_uint8 * nb).from_address(ptr)
        shape = (self.h, self.w) if self.ch == 1 else (self.h, self.w, 3)
        return np.frombuffer(buf, np.uint8).reshape(shape)

    def detect_match(self, first, n_frames, pairs=(), grid=None, window=100.0, nnratio=0.6, th_low=50, check_ori=True):
        ctx = self.ctx
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        P = len(pairs)
        # (the workspace must know the shape before max_keypoints is final: a first call may report E_CAPACITY)
        cap = ctx.max_keypoints()
        for _ in range(2):
            kps = np.zeros((n_frames, cap), KP_DTYPE); desc = np.zeros((n_frames, cap, 32), np.uint8)
            n = np.zeros(n_frames, np.int32); mono = np.zeros(n_frames, np.int32)
            m = np.full((max(P, 1), max(cap, ctx.max_keypoints())), -1, np.int32); nm = np.zeros(max(P, 1), np.int32)
            rc = ctx.L.nav24_ingest_detect_match(self.h_, first, n_frames, _p(kps), _p(desc), cap, _p(n), _p(mono), P,
                                                 _p(pairs) if P else None, C.byref(grid) if grid is not None else None,
                                                 window, nnratio, th_low, int(check_ori), _p(m) if P else None, m.shape[1],
                                                 _p(nm) if P else None)
            if rc == E_CAPACITY:
                cap = max(ctx.max_keypoints(), int(n.max()) if n.max() > 0 else cap)
                continue
            break
        ctx._check(rc)
        return n, mono, kps, desc, m[:P], nm[:P]
