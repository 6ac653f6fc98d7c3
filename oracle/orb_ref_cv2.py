"""cv2-driven restatement of nav24's ORB detector (oracle pinning; test infrastructure only).

Follows core/operators/objDetection/OP_FtDtOrbSlam.cpp line by line but takes every OpenCV
primitive from the real library (cv2 4.13.0 here): cv2.resize / copyMakeBorder (:936-960),
cv2.FastFeatureDetector per cell (:751-818), cv2.GaussianBlur (:890-891), cv2.fastAtan2 (:43).
The quadtree (:502-725) is nav24's own code, not OpenCV, so it is taken from the C++ oracle
(real std::list / std::sort); cosf/sinf come from glibc through ctypes, as in the reference.
Used to (a) pin oracle/orb_oracle.cpp against OpenCV and (b) generate tests/golden/.
"""
import ctypes
import ctypes.util

import cv2
import numpy as np

from . import orb_oracle as oo

_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.cosf.restype = ctypes.c_float; _libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float; _libm.sinf.argtypes = [ctypes.c_float]
_libm.lrintf.restype = ctypes.c_long; _libm.lrintf.argtypes = [ctypes.c_float]

EDGE = 19
F32 = np.float32


def cv_round(v):
    return int(_libm.lrintf(float(F32(v))))


class OrbRefCv2:
    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels, self.ini_th, self.min_th = nlevels, ini_th, min_th
        # tables: plain float/double arithmetic as in the ctor (:441-500)
        sf = float(F32(scale))  # double member initialised from a float
        self.scale = [F32(1.0)]
        for _ in range(1, nlevels):
            self.scale.append(F32(float(self.scale[-1]) * sf))
        self.inv = [F32(1.0) / s for s in self.scale]
        self.sf = sf
        self.set_num_features(nfeatures)
        self.umax = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
        self.pattern = np.array(open(oo._HERE + "/pattern_31.inc").read().split("\n", 2)[2].replace("\n", "")
                                .rstrip(",").split(","), dtype=np.int32).reshape(512, 2)
        self.fast_ini = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self.fast_min = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

    def set_num_features(self, n):
        self.nfeatures = n
        factor = F32(1.0 / self.sf)
        per = F32(F32(n) * (F32(1) - factor)) / (F32(1) - F32(float(factor) ** self.nlevels))
        per = F32(per)
        self.quota, s = [], 0
        for _ in range(self.nlevels - 1):
            q = cv_round(per); self.quota.append(q); s += q
            per = F32(per * factor)
        self.quota.append(max(n - s, 0))

    def pyramid(self, image):
        self.levels = []
        for l in range(self.nlevels):
            w = cv_round(F32(image.shape[1]) * self.inv[l]); h = cv_round(F32(image.shape[0]) * self.inv[l])
            temp = np.zeros((h + 2 * EDGE, w + 2 * EDGE), np.uint8)
            roi = temp[EDGE:EDGE + h, EDGE:EDGE + w]
            if l:
                cv2.resize(self.levels[l - 1], (w, h), dst=roi, fx=0, fy=0, interpolation=cv2.INTER_LINEAR)
                cv2.copyMakeBorder(roi, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101 + cv2.BORDER_ISOLATED, dst=temp)
            else:
                cv2.copyMakeBorder(image, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101, dst=temp)
            self.levels.append(roi)

    def raw_keys(self, l):
        img = self.levels[l]
        minB = EDGE - 3
        maxBX, maxBY = img.shape[1] - EDGE + 3, img.shape[0] - EDGE + 3
        width, height = F32(maxBX - minB), F32(maxBY - minB)
        nCols, nRows = int(width / F32(35)), int(height / F32(35))
        wCell, hCell = int(np.ceil(width / F32(nCols))), int(np.ceil(height / F32(nRows)))
        out = []
        for i in range(nRows):
            iniY = minB + i * hCell; maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minB + j * wCell; maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                cell = img[iniY:maxY, iniX:maxX]
                kps = self.fast_ini.detect(cell)
                if not kps:
                    kps = self.fast_min.detect(cell)
                for k in kps:
                    out.append((k.pt[0] + j * wCell, k.pt[1] + i * hCell, k.response))
        return np.array(out, np.float32).reshape(-1, 3), (minB, maxBX, minB, maxBY)

    def ic_angle(self, img, x, y):
        cx, cy = cv_round(x), cv_round(y)
        m01 = m10 = 0
        row = img[cy].astype(np.int64)
        for u in range(-15, 16):
            m10 += u * int(row[cx + u])
        for v in range(1, 16):
            d = self.umax[v]
            p = img[cy + v, cx - d:cx + d + 1].astype(np.int64); m = img[cy - v, cx - d:cx + d + 1].astype(np.int64)
            us = np.arange(-d, d + 1)
            m01 += v * int((p - m).sum()); m10 += int((us * (p + m)).sum())
        return F32(cv2.fastAtan2(float(m01), float(m10)))

    def describe(self, blurred, x, y, angle):
        ang = F32(F32(angle) * F32(np.pi / F32(180.0)))
        a, b = F32(_libm.cosf(float(ang))), F32(_libm.sinf(float(ang)))
        cy, cx = cv_round(y), cv_round(x)
        px = self.pattern[:, 0].astype(F32); py = self.pattern[:, 1].astype(F32)
        ry = (px * b).astype(F32) + (py * a).astype(F32)
        rx = (px * a).astype(F32) - (py * b).astype(F32)
        iy = np.rint(ry.astype(F32)).astype(np.int64); ix = np.rint(rx.astype(F32)).astype(np.int64)
        v = blurred[cy + iy, cx + ix].astype(np.int32)
        bits = (v[0::2] < v[1::2]).astype(np.uint8).reshape(32, 8)
        return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)

    def detect(self, image):
        """Returns (mono_index, kps[KP_DTYPE], desc) and keeps per-stage results in self.stage."""
        self.pyramid(image)
        per_level = []
        self.stage = {"raw": [], "kept": [], "blur": []}
        for l in range(self.nlevels):
            raw, (minX, maxX, minY, maxY) = self.raw_keys(l)
            kept = oo.quadtree(raw, minX, maxX, minY, maxY, self.quota[l])
            self.stage["raw"].append(raw); self.stage["kept"].append(kept)
            ks = np.zeros(len(kept), oo.KP_DTYPE)
            ks["x"] = raw[kept, 0] + F32(minX); ks["y"] = raw[kept, 1] + F32(minY)
            ks["response"] = raw[kept, 2]; ks["octave"] = l; ks["class_id"] = -1
            ks["size"] = F32(int(F32(31) * self.scale[l]))
            for i in range(len(ks)):
                ks["angle"][i] = self.ic_angle(self.levels[l], ks["x"][i], ks["y"][i])
            per_level.append(ks)
        n = sum(len(k) for k in per_level)
        outK = np.zeros(n, oo.KP_DTYPE); outD = np.zeros((n, 32), np.uint8)
        mono, stereo = 0, n - 1
        self.stage["level_kps"] = per_level
        self.stage["level_desc"] = []
        for l, ks in enumerate(per_level):
            if len(ks) == 0:
                self.stage["blur"].append(None); self.stage["level_desc"].append(np.zeros((0, 32), np.uint8))
                continue
            work = self.levels[l].copy()
            work = cv2.GaussianBlur(work, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            self.stage["blur"].append(work)
            D = np.zeros((len(ks), 32), np.uint8)
            for i in range(len(ks)):
                D[i] = self.describe(work, ks["x"][i], ks["y"][i], ks["angle"][i])
            self.stage["level_desc"].append(D)
            for i in range(len(ks)):
                k = ks[i].copy()
                if l:
                    k["x"] = F32(k["x"] * self.scale[l]); k["y"] = F32(k["y"] * self.scale[l])
                if 0 <= k["x"] <= 1000:
                    dst = stereo; stereo -= 1
                else:
                    dst = mono; mono += 1
                outK[dst] = k; outD[dst] = D[i]
        return mono, outK, outD
