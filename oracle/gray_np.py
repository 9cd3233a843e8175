"""numpy restatement of the estimation pre-pass (K1 + K2).  TEST INFRASTRUCTURE ONLY.

Follows nodes/stabilizer_utils.py:236-242 (_make_gray), :248-268 (_working_estimation_size)
and :271-276 (_make_gray_for_estimation) of the reference, with the cv2 4.13.0.92 arithmetic
they call into restated from black-box probes:
  cvtColor(RGB2GRAY) f32 : fma(B, .114f, fma(R, .299f, G*.587f))   (bit-exact vs the wheel)
  resize INTER_AREA      : x2 -> (a+b+c+d+2)>>2; integer KxL -> rint(sum * (1.f/(K*L)));
                           otherwise cv::computeResizeAreaTab + float32 accumulation in
                           table order.
Pinned against live cv2 in tests/test_oracle_gray.py.
"""
from __future__ import annotations

import math

import numpy as np


def working_size(width: int, height: int, max_side: int = 960):
    longest = max(int(width), int(height))
    if longest <= max_side:
        return None
    scale = max_side / float(longest)
    w = max(1, int(round(width * scale)))
    h = max(1, int(round(height * scale)))
    if w >= width or h >= height:
        return None
    return w, h


def _f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)


def luma_f32(rgb: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(rgb, COLOR_RGB2GRAY) for float32 input; products of two f32 are exact in f64,
    so rounding the f64 sum once to f32 reproduces a fused multiply-add."""
    r, g, b = (rgb[..., i].astype(np.float64) for i in range(3))
    cr, cg, cb = (np.float64(np.float32(v)) for v in (0.299, 0.587, 0.114))
    t = _f32(g * cg)
    t = _f32(r * cr + t)
    return (b * cb + t).astype(np.float32)


def gray_u8(rgb: np.ndarray) -> np.ndarray:
    g = luma_f32(np.asarray(rgb, dtype=np.float32))
    return np.clip(g * np.float32(255.0), 0, 255).astype(np.uint8)


def area_table(ssize: int, dsize: int):
    scale = 1.0 / (float(dsize) / ssize)
    tab = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = math.ceil(f1), math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - f1) / cell)))
        for s in range(s1, s2):
            tab.append((d, s, np.float32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
    return tab


def resize_area_u8(src: np.ndarray, dsize) -> np.ndarray:
    """cv2.resize(src_u8, (dw, dh), interpolation=INTER_AREA) for down-scaling."""
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()
    sx, sy = 1.0 / (float(dw) / sw), 1.0 / (float(dh) / sh)
    ix, iy = int(round(sx)), int(round(sy))
    if abs(sx - ix) < np.finfo(np.float64).eps and abs(sy - iy) < np.finfo(np.float64).eps:
        acc = src[: dh * iy, : dw * ix].astype(np.int64).reshape(dh, iy, dw, ix).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((acc + 2) >> 2).astype(np.uint8)
        val = acc.astype(np.float32) * np.float32(1.0 / (ix * iy))
        return np.clip(np.rint(val), 0, 255).astype(np.uint8)
    s = src.astype(np.float32)
    buf = np.zeros((sh, dw), np.float32)
    for d, si, a in area_table(sw, dw):
        buf[:, d] = buf[:, d] + s[:, si] * a
    out = np.zeros((dh, dw), np.float32)
    seen = set()
    for d, si, b in area_table(sh, dh):
        if d in seen:
            out[d] = out[d] + b * buf[si]
        else:
            out[d] = b * buf[si]
            seen.add(d)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def gray_for_estimation(rgb: np.ndarray, work_size=None) -> np.ndarray:
    g = gray_u8(rgb)
    if work_size is None:
        return g
    return resize_area_u8(g, work_size)
