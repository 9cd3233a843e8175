"""ctypes wrapper of oracle/dis_ref.c (plain-C restatement of cv2.DISOpticalFlow).
TEST INFRASTRUCTURE ONLY.  Build with `make -C oracle` (done by __graft_entry__.build())."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdisref.so")
_lib = None


class DisParams(C.Structure):
    _fields_ = [
        ("finest_scale", C.c_int), ("patch_size", C.c_int), ("patch_stride", C.c_int), ("gd_iter", C.c_int),
        ("vr_iter", C.c_int), ("vr_alpha", C.c_float), ("vr_delta", C.c_float), ("vr_gamma", C.c_float),
        ("vr_epsilon", C.c_float), ("use_mean_norm", C.c_int), ("use_spatial", C.c_int), ("sor_iter", C.c_int),
        ("omega", C.c_float), ("reserved", C.c_int),
    ]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            import subprocess

            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_LIB_PATH)
        _lib.disref_calc.restype = C.c_int
        _lib.disref_coarsest_scale.restype = C.c_int
    return _lib


def default_params(**over) -> DisParams:
    p = DisParams()
    lib().disref_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def select_scales(h: int, w: int, finest_state: int = 2):
    """(finest, coarsest) pyramid levels cv2's calc() works on for an h x w frame when the backend object
    holds `finest_state` (autoSelectPatchSizeAndScales for small frames); None where cv2 raises."""
    fi, co = C.c_int(), C.c_int()
    rc = lib().disref_select_scales(C.c_int(h), C.c_int(w), C.byref(default_params(finest_scale=finest_state)), C.byref(fi), C.byref(co))
    return None if rc else (fi.value, co.value)


def calc(i0: np.ndarray, i1: np.ndarray, params: DisParams | None = None) -> np.ndarray:
    """uint8 [h,w] pair -> float32 [h,w,2] flow, the reference's DIS configuration by default: what the
    FIRST calc() of a freshly configured backend object returns (see Backend for the later ones)."""
    i0 = np.ascontiguousarray(i0, dtype=np.uint8)
    i1 = np.ascontiguousarray(i1, dtype=np.uint8)
    h, w = i0.shape
    params = params or default_params()
    flow = np.zeros((h, w, 2), np.float32)
    rc = lib().disref_calc(_p(i0), _p(i1), C.c_int(h), C.c_int(w), C.byref(params), _p(flow))
    if rc == -1:
        raise ValueError("cv2 raises on this size: the input image must have either width or height >= 12")
    if rc != 0:
        raise ValueError("coarsest pyramid level smaller than one patch: cv2 reads out of bounds on this size")
    return flow


class Backend:
    """The stateful object of nodes/video_stabilizer_flow.py:76-87 / :312: one per clip, calc() per pair.
    On frames too small for finest scale 2 cv2 rewrites the object's finest scale during the first calc(),
    so later pairs may run on other levels than the first one (90x50: 2..0, then 1..0)."""

    def __init__(self, **over):
        self.params = default_params(**over)

    def calc(self, i0: np.ndarray, i1: np.ndarray) -> np.ndarray:
        flow = calc(i0, i1, self.params)
        sel = select_scales(i0.shape[0], i0.shape[1], self.params.finest_scale)
        self.params.finest_scale = sel[0]
        return flow


def resize_area_u8(src: np.ndarray, dsize) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.zeros((dsize[1], dsize[0]), np.uint8)
    lib().disref_resize_area_u8(_p(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), _p(dst), C.c_int(dsize[1]), C.c_int(dsize[0]))
    return dst


def resize_linear_f32(src: np.ndarray, dsize, mode: int = 0) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.float32)
    cn = 1 if src.ndim == 2 else src.shape[2]
    shape = (dsize[1], dsize[0]) if src.ndim == 2 else (dsize[1], dsize[0], cn)
    dst = np.zeros(shape, np.float32)
    lib().disref_resize_linear_f32(_p(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), C.c_int(cn), _p(dst),
                                   C.c_int(dsize[1]), C.c_int(dsize[0]), C.c_int(mode))
    return dst


def variational_refinement(i0, i1, u, v, alpha=20.0, delta=5.0, gamma=10.0, epsilon=0.01, fp_iter=5, sor_iter=5, omega=1.6):
    """cv2.VariationalRefinement.calcUV restated; i0/i1 any real dtype [h,w]; returns (u', v')."""
    f0 = np.ascontiguousarray(i0, dtype=np.float32)
    f1 = np.ascontiguousarray(i1, dtype=np.float32)
    u = np.array(u, dtype=np.float32, copy=True, order="C")
    v = np.array(v, dtype=np.float32, copy=True, order="C")
    h, w = f0.shape
    lib().disref_variational_refinement(_p(f0), _p(f1), C.c_int(h), C.c_int(w), _p(u), _p(v), C.c_float(alpha), C.c_float(delta),
                                        C.c_float(gamma), C.c_float(epsilon), C.c_int(fp_iter), C.c_int(sor_iter), C.c_float(omega))
    return u, v
