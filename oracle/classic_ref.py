"""ctypes wrapper of oracle/classic_ref.c (GFTT + pyramidal LK restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libclassicref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            import subprocess

            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_LIB_PATH)
        _lib.classicref_good_features.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def min_eigen(img: np.ndarray, block: int = 21) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros(img.shape, np.float32)
    lib().classicref_min_eigen(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_int(block), _p(out))
    return out


def good_features(img: np.ndarray, max_corners=400, quality=0.01, min_distance=7.0, block=21) -> np.ndarray:
    """cv2.goodFeaturesToTrack(...) -> float32 [K,2] (x, y)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    xy = np.zeros((max(max_corners, 1) if max_corners > 0 else img.size, 2), np.float32)
    k = lib().classicref_good_features(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_int(max_corners),
                                       C.c_double(quality), C.c_double(min_distance), C.c_int(block), _p(xy))
    return xy[:k].copy()


def pyr_lk(prev_img, next_img, pts, win=31, max_level=3, max_iter=50, eps=0.01, exact=True):
    """cv2.calcOpticalFlowPyrLK(prev, next, pts, None, winSize=(win,win), maxLevel, (EPS|COUNT, max_iter, eps)) -> (next_pts, status).
    exact=True (default; what the CUDA tracker does since round 2) sums the window terms in the lane order of cv2's SIMD
    code: bit-exact positions and status against the wheel.  exact=False sums them serially (<= 2e-3 px off; kept to show
    the difference, see classic_ref.c)."""
    a = np.ascontiguousarray(prev_img, dtype=np.uint8)
    b = np.ascontiguousarray(next_img, dtype=np.uint8)
    p = np.ascontiguousarray(np.asarray(pts, dtype=np.float32).reshape(-1, 2))
    n = p.shape[0]
    out = np.zeros((n, 2), np.float32)
    st = np.zeros((n,), np.uint8)
    (lib().classicref_pyr_lk_exact if exact else lib().classicref_pyr_lk)(_p(a), _p(b), C.c_int(a.shape[0]), C.c_int(a.shape[1]), _p(p), C.c_int(n), C.c_int(win), C.c_int(max_level),
                            C.c_int(max_iter), C.c_double(eps), _p(out), _p(st))
    return out, st
