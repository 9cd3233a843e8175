"""Host-side (O(N) scalars / 3x3 matrices) half of the stabilizer path, in float64 numpy.

Mirrors the behaviour -- names, argument meaning, numeric types -- of the helpers in the
reference's ``nodes/stabilizer_utils.py`` that sit between the estimation kernels and the warp
kernel.  None of this touches pixels; pixels only ever live in HBM (see ``pipeline.py``).

Reference anchors (file:line in the reference repository):
  working size            nodes/stabilizer_utils.py:248-268
  rescale to full         nodes/stabilizer_utils.py:279-297
  matrix <-> params       nodes/stabilizer_utils.py:300-358
  box smoothing           nodes/stabilizer_utils.py:361-383
  expand transform        nodes/stabilizer_utils.py:386-406
  padding colour parser   nodes/stabilizer_utils.py:843-873
  warp meta builder       nodes/stabilizer_utils.py:876-896
  bounding boxes / ratio  nodes/stabilizer_utils.py:1010-1052
"""
from __future__ import annotations

import contextlib
import gc
import math
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

ESTIMATION_MAX_SIDE = 960
DEFAULT_PADDING_RGB = (127, 127, 127)
TRANSFORM_MODES = ("translation", "similarity", "perspective")
FRAMING_MODES = ("crop", "crop_and_pad", "expand")
MODE_LADDER = {
    "perspective": ("perspective", "similarity", "translation"),
    "similarity": ("similarity", "translation"),
    "translation": ("translation",),
}


class _GcPause:
    """Handle of a gc_paused() block: collect_young() runs the generation-0 pass the block has been holding back."""

    def __init__(self, active: bool):
        self.active = active

    def collect_young(self) -> None:
        if self.active:
            gc.collect(0)


@contextlib.contextmanager
def gc_paused():
    """Cyclic GC off while a call builds its result.  `meta` is tens of thousands of acyclic lists and
    dicts per clip; allocating them with the collector on triggers a generation-0 pass every 700
    containers, promotes the survivors and soon a full collection over every object of the process
    (tens of ms with torch loaded) -- measured 8 ms median / 125 ms worst per call for a 968-frame meta
    against 4 ms with the collector paused.  Reference counting still frees everything as usual.
    The young objects of the call still have to be looked at once: switching the collector back on makes the very
    next allocation run that generation-0 pass (0.15 ms for a 121-frame meta) -- after the GPU has finished, on the
    critical path.  The driver therefore calls collect_young() on the yielded handle when the meta tree is built and
    the resampler is still running.  A host application that wants its collector left alone sets VSTAB_GC_PAUSE=0."""
    import os

    if os.environ.get("VSTAB_GC_PAUSE", "1") == "0":
        yield _GcPause(False)
        return
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield _GcPause(was_enabled)
    finally:
        if was_enabled:
            gc.enable()


def working_estimation_size(width: int, height: int, max_side: int = ESTIMATION_MAX_SIDE):
    """(w, h) of the estimation image, or None when the frame is used as is."""
    longest = max(int(width), int(height))
    if longest <= max_side:
        return None
    ratio = max_side / float(longest)
    w = max(1, int(round(width * ratio)))
    h = max(1, int(round(height * ratio)))
    if w >= width or h >= height:
        return None
    return w, h


def rescale_transform_to_full(matrix, source_size, working_size) -> np.ndarray:
    """S^-1 @ M @ S in float64 -> float32, S = diag(work/src)."""
    kx = working_size[0] / float(source_size[0])
    ky = working_size[1] / float(source_size[1])
    down = np.diag([kx, ky, 1.0]).astype(np.float64)
    up = np.diag([1.0 / kx, 1.0 / ky, 1.0]).astype(np.float64)
    return (up @ np.asarray(matrix).astype(np.float64) @ down).astype(np.float32)


def rescale_transforms_to_full(matrices: np.ndarray, source_size, working_size) -> np.ndarray:
    """Stacked variant of rescale_transform_to_full ([P,3,3] float32 in, float32 out).  The two
    diagonal products have a single non-zero term per element, so this is bit-identical."""
    kx = working_size[0] / float(source_size[0])
    ky = working_size[1] / float(source_size[1])
    down = np.array([kx, ky, 1.0], dtype=np.float64)
    up = np.array([1.0 / kx, 1.0 / ky, 1.0], dtype=np.float64)
    # (up @ M) @ down with diagonal factors: element (i, j) is (up_i * M_ij) * down_j plus exact zeros
    return ((np.asarray(matrices).astype(np.float64) * up[None, :, None]) * down[None, None, :]).astype(np.float32)


def matrix_to_params(matrix, base_mode: str) -> np.ndarray:
    m = matrix
    if base_mode == "translation":
        return np.array([m[0, 2], m[1, 2]], dtype=np.float64)
    if base_mode == "similarity":
        a, c = m[0, 0], m[1, 0]
        mag = math.sqrt(max(a * a + c * c, 1e-10))
        return np.array([m[0, 2], m[1, 2], math.atan2(c, a), math.log(mag)], dtype=np.float64)
    return np.array(
        [m[0, 0] - 1.0, m[0, 1], m[0, 2], m[1, 0], m[1, 1] - 1.0, m[1, 2], m[2, 0], m[2, 1]],
        dtype=np.float64,
    )


def params_to_matrix(params, base_mode: str) -> np.ndarray:
    p = params
    if base_mode == "translation":
        rows = [[1.0, 0.0, p[0]], [0.0, 1.0, p[1]], [0.0, 0.0, 1.0]]
    elif base_mode == "similarity":
        k = math.exp(p[3])
        cs, sn = math.cos(p[2]), math.sin(p[2])
        rows = [[k * cs, -k * sn, p[0]], [k * sn, k * cs, p[1]], [0.0, 0.0, 1.0]]
    else:
        rows = [[p[0] + 1.0, p[1], p[2]], [p[3], p[4] + 1.0, p[5]], [p[6], p[7], 1.0]]
    return np.array(rows, dtype=np.float32)


_MATH = {"atan2": math.atan2, "log": math.log, "exp": math.exp, "cos": math.cos, "sin": math.sin}


def _libm(op: str, a, b=None) -> np.ndarray:
    """Element-wise libm over float64 arrays, the bits of Python's math.<op> (what the reference calls per frame).
    Through libvstab's host helper when the library is built (one C loop instead of a Python call per element: the
    O(frames) part of the trajectory solve every rank repeats); through map(math.<op>) otherwise."""
    try:
        from . import _native

        return _native.host_libm(op, a, b)
    except Exception:  # library not built (host-only checkouts): the slow, equally exact route
        fn = _MATH[op]
        a = np.asarray(a, dtype=np.float64)
        if b is None:
            return np.fromiter(map(fn, a.tolist()), dtype=np.float64, count=a.size)
        return np.fromiter(map(fn, a.tolist(), np.asarray(b, dtype=np.float64).tolist()), dtype=np.float64, count=a.size)


def matrices_to_params(matrices: np.ndarray, base_mode: str) -> np.ndarray:
    """Stacked matrix_to_params: [P,3,3] float32 -> [P,K] float64, same bits as the per-matrix
    version (float32 products for a*a + c*c, libm atan2 / log per element)."""
    m = np.asarray(matrices)
    n = m.shape[0]
    if base_mode == "translation":
        return np.stack([m[:, 0, 2], m[:, 1, 2]], axis=1).astype(np.float64)
    if base_mode == "similarity":
        a, c = m[:, 0, 0], m[:, 1, 0]
        mag2 = a * a + c * c  # float32 arithmetic when the matrices are float32, like numpy scalars
        out = np.empty((n, 4), dtype=np.float64)
        out[:, 0] = m[:, 0, 2]
        out[:, 1] = m[:, 1, 2]
        # libm per element through map(): no interpreter frame per item.  float32 -> Python float is exact,
        # so atan2 sees the same doubles as math.atan2(np.float32, np.float32) does in the reference.
        out[:, 2] = _libm("atan2", c.astype(np.float64), a.astype(np.float64))
        # max(v, 1e-10) keeps v when v > 1e-10 compared in v's own dtype (NEP 50: the Python float is weak)
        floor = mag2.dtype.type(1e-10)
        clamped = np.where(mag2 > floor, mag2.astype(np.float64), 1e-10)
        out[:, 3] = _libm("log", np.sqrt(clamped))  # sqrt is IEEE-exact
        return out
    one = m.dtype.type(1.0)
    return np.stack(
        [m[:, 0, 0] - one, m[:, 0, 1], m[:, 0, 2], m[:, 1, 0], m[:, 1, 1] - one, m[:, 1, 2], m[:, 2, 0], m[:, 2, 1]], axis=1
    ).astype(np.float64)


def params_to_matrices(params: np.ndarray, base_mode: str) -> np.ndarray:
    """Stacked params_to_matrix: [N,K] float64 -> [N,3,3] float32 (libm exp / cos / sin per element)."""
    p = np.asarray(params, dtype=np.float64)
    n = p.shape[0]
    out = np.zeros((n, 3, 3), dtype=np.float64)
    out[:, 2, 2] = 1.0
    if base_mode == "translation":
        out[:, 0, 0] = 1.0
        out[:, 1, 1] = 1.0
        out[:, 0, 2] = p[:, 0]
        out[:, 1, 2] = p[:, 1]
    elif base_mode == "similarity":
        ang = np.ascontiguousarray(p[:, 2])
        k = _libm("exp", p[:, 3])
        cs = _libm("cos", ang)
        sn = _libm("sin", ang)
        out[:, 0, 0] = k * cs
        out[:, 0, 1] = -k * sn
        out[:, 1, 0] = k * sn
        out[:, 1, 1] = k * cs
        out[:, 0, 2] = p[:, 0]
        out[:, 1, 2] = p[:, 1]
    else:
        out[:, 0, 0] = p[:, 0] + 1.0
        out[:, 0, 1] = p[:, 1]
        out[:, 0, 2] = p[:, 2]
        out[:, 1, 0] = p[:, 3]
        out[:, 1, 1] = p[:, 4] + 1.0
        out[:, 1, 2] = p[:, 5]
        out[:, 2, 0] = p[:, 6]
        out[:, 2, 1] = p[:, 7]
    return out.astype(np.float32)


def left_multiply(shift: np.ndarray, matrices: np.ndarray) -> np.ndarray:
    """[shift @ m for m in matrices] as one batched matmul (numpy runs the same 3x3 BLAS product per
    matrix, so the result has the same bits as the reference's per-frame `translate_matrix @ mat`)."""
    return np.matmul(np.asarray(shift)[None], np.asarray(matrices))


def native_trajectory(cands, mode: str, source_size, working_size):
    """Acceptance test + float32 transforms at full size + cumulative path through libvstab's host helper
    (vstab_host_trajectory): (matrices [P,3,3] float32, path [P+1,K] float64), or None when a pair falls back to another
    model or the library is not built -- the caller then takes the numpy route above, which computes the same bits."""
    try:
        from . import _native

        return _native.host_trajectory(cands.raw_words(), cands.detected, cands.min_points, mode, source_size, working_size)
    except (OSError, RuntimeError, AttributeError):  # library not built (host-only checkouts)
        return None


def native_framing(diffs: np.ndarray, mode: str, width: int, height: int):
    """params_to_matrices + compute_bounding_boxes + inner / outer rectangle in one C pass (vstab_host_framing):
    (apply [N,3,3] float32, mins, maxs, box[9]) or None without the library."""
    try:
        from . import _native

        return _native.host_framing(diffs, mode, width, height)
    except (OSError, RuntimeError, AttributeError):  # library not built (host-only checkouts)
        return None


_TARGET_WINDOW_OK: Dict[int, bool] = {}


def numpy_target(path: np.ndarray, strength: float, smooth: float, fps: float, camera_lock: bool):
    """(target_path, diffs) the way the reference forms them (flow.py:351-374), in numpy."""
    target = np.zeros_like(path) if camera_lock else path + strength * (smooth_path(path, smooth, fps) - path)
    return target, target - path


def native_target(path: np.ndarray, strength: float, smooth: float, fps: float, camera_lock: bool):
    """numpy_target through libvstab's host helper (vstab_host_target) when that is known to produce numpy's bits: the box
    filter is np.convolve, whose summation order is numpy's own loop up to 11 taps and the BLAS's beyond.  The first
    call with a given window compares the helper with numpy on a test path; a window that disagrees (another numpy
    build) stays on numpy for the rest of the process.  None = use numpy_target."""
    try:
        from . import _native

        smooth_c = clip01(smooth)
        window = 0 if (camera_lock or smooth_c <= 0.0 or len(path) <= 2) else smoothing_window(smooth_c, fps)
        if window > 11:
            return None
        ok = True if window == 0 else _TARGET_WINDOW_OK.get(window)  # no filter: element-wise operations only
        if ok is None:
            probe = np.cumsum(np.random.default_rng(window).normal(0.0, 3.0, (40, 4)), axis=0)
            want = numpy_target(probe, 0.7, smooth, fps, False)
            got = _native.host_target(probe, window, 0.7, False)
            ok = got is not None and got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes()
            _TARGET_WINDOW_OK[window] = ok
        if not ok:
            return None
        return _native.host_target(path, window, strength, camera_lock)
    except (OSError, RuntimeError, AttributeError):  # library not built (host-only checkouts)
        return None


def translate_matrices(matrices: np.ndarray, off_x: float, off_y: float, affine: bool = False) -> np.ndarray:
    """[[1, 0, off_x], [0, 1, off_y], [0, 0, 1]] (float32) @ m for every m: the recentring / expand shift of the framing
    modes.  Affine stacks go through vstab_host_shift (one inexact addition per element, so independent of the BLAS);
    anything else through the float32 matmul the reference runs."""
    if affine:
        try:
            from . import _native

            out = _native.host_shift(np.ascontiguousarray(matrices, dtype=np.float32), off_x, off_y)
            if out is not None:
                return out
        except (OSError, RuntimeError, AttributeError):  # library not built (host-only checkouts)
            pass
    shift = np.array([[1.0, 0.0, off_x], [0.0, 1.0, off_y], [0.0, 0.0, 1.0]], dtype=np.float32)
    return left_multiply(shift, matrices)


def clip01(value) -> float:
    """float(np.clip(value, 0.0, 1.0)) without the 5 us of a numpy call on a scalar (NaN stays NaN either way)."""
    value = float(value)
    return min(max(value, 0.0), 1.0)


def smoothing_window(smooth: float, fps: float) -> int:
    fps = float(max(1.0, fps))
    seconds = 3.0 / 16.0 + smooth * (13.0 / 16.0 - 3.0 / 16.0)
    window = max(3, int(round(seconds * fps)))
    return window + 1 if window % 2 == 0 else window


def smooth_path(path: np.ndarray, smooth: float, fps: float) -> np.ndarray:
    """Edge-padded odd box filter along time, per parameter column (float64)."""
    smooth = float(np.clip(smooth, 0.0, 1.0))
    if smooth <= 0.0 or len(path) <= 2:
        return path.copy()
    window = smoothing_window(smooth, fps)
    half = window // 2
    taps = np.ones(window, dtype=np.float64) / float(window)
    n, cols = path.shape
    # np.pad(mode="edge") written out for all columns at once (it costs 35 us per call): one contiguous row per column
    padded = np.empty((cols, n + 2 * half), dtype=path.dtype)
    padded[:, half : half + n] = path.T
    padded[:, :half] = path[0][:, None]
    padded[:, half + n :] = path[n - 1][:, None]
    out = np.empty_like(path)
    for col in range(cols):
        out[:, col] = np.convolve(padded[col], taps, mode="valid")
    return out


def compute_bounding_boxes(matrices: Sequence[np.ndarray], width: int, height: int):
    corners = np.array(
        [[0.0, float(width), 0.0, float(width)], [0.0, 0.0, float(height), float(height)], [1.0, 1.0, 1.0, 1.0]],
        dtype=np.float64,
    )
    # one 3x3 @ 3x4 product per matrix, evaluated term by term in the order of a dot product so the
    # stacked form returns the same bits as the reference's per-matrix `matrix @ corners`
    stacked_f32 = isinstance(matrices, np.ndarray) and matrices.ndim == 3 and matrices.dtype == np.float32
    if not stacked_f32 and any(np.asarray(x).dtype != np.float32 for x in matrices):
        # float64 @ float64 goes through BLAS (fused multiply-adds): keep the per-matrix product
        lo, hi = [], []
        for mat in matrices:
            q = mat @ corners
            q /= q[2, :]
            lo.append([q[0].min(), q[1].min()])
            hi.append([q[0].max(), q[1].max()])
        return np.array(lo), np.array(hi)
    # float32 @ float64 promotes the float32 operand and multiplies without FMA (verified bit-equal)
    m = (matrices if stacked_f32 else np.stack([np.asarray(x) for x in matrices], axis=0)).astype(np.float64)
    q = m[:, :, 0:1] * corners[0][None, None, :] + m[:, :, 1:2] * corners[1][None, None, :] + m[:, :, 2:3] * corners[2][None, None, :]
    xy = q[:, :2] / q[:, 2:3]  # the reference divides all three rows; the third one (w / w) is never read
    # min / max over the four corners as three element-wise operations (a reduction over an axis of 4 costs more)
    a, b, c, d = xy[:, :, 0], xy[:, :, 1], xy[:, :, 2], xy[:, :, 3]
    return np.minimum(np.minimum(a, b), np.minimum(c, d)), np.maximum(np.maximum(a, b), np.maximum(c, d))


def inner_rectangle(mins, maxs):
    """(x0, y0, x1, y1) common to all bounding boxes: max of the minima, min of the maxima (two reductions)."""
    lo, hi = np.max(mins, axis=0), np.min(maxs, axis=0)
    return lo[0], lo[1], hi[0], hi[1]


def min_content_ratio(mins, maxs, width: int, height: int, inner=None) -> float:
    x0, y0, x1, y1 = inner if inner is not None else inner_rectangle(mins, maxs)
    iw = max(0.0, x1 - x0)
    ih = max(0.0, y1 - y0)
    if iw <= 0.0 or ih <= 0.0:
        return 1e-6
    return max(1e-6, min(iw / width, ih / height))


def prepare_expand_transform(mins, maxs):
    x_lo, y_lo = float(np.min(mins[:, 0])), float(np.min(mins[:, 1]))
    x_hi, y_hi = float(np.max(maxs[:, 0])), float(np.max(maxs[:, 1]))
    shift = np.array([[1.0, 0.0, -x_lo], [0.0, 1.0, -y_lo], [0.0, 0.0, 1.0]], dtype=np.float32)
    return shift, (max(int(math.ceil(x_hi - x_lo)), 1), max(int(math.ceil(y_hi - y_lo)), 1))


def parse_padding_color(value) -> Tuple[int, int, int]:
    """'#RRGGBB' | '#RGB' | 'r,g,b' | 'r/g/b' | 0xRRGGBB int -> (r, g, b); junk -> (127,127,127)."""
    if isinstance(value, str):
        text = value.strip()
        if "," in text or "/" in text:
            try:
                nums = [int(tok) for tok in text.replace("/", ",").replace(" ", ",").split(",") if tok != ""]
            except (TypeError, ValueError):
                return DEFAULT_PADDING_RGB
            if len(nums) == 1:
                nums = nums * 3
            if len(nums) != 3:
                return DEFAULT_PADDING_RGB
            return tuple(int(np.clip(v, 0, 255)) for v in nums)
        digits = text.removeprefix("#")
        if len(digits) == 3:
            digits = "".join(ch * 2 for ch in digits)
        if len(digits) != 6:
            return DEFAULT_PADDING_RGB
        try:
            packed = int(digits, 16)
        except (TypeError, ValueError):
            return DEFAULT_PADDING_RGB
    else:
        try:
            packed = int(value)
        except (TypeError, ValueError):
            return DEFAULT_PADDING_RGB
    packed = int(np.clip(packed, 0, 0xFFFFFF))
    return (packed >> 16) & 0xFF, (packed >> 8) & 0xFF, packed & 0xFF


def border_value(padding_rgb) -> Tuple[float, float, float]:
    """padding colour as float32/255, the borderValue of every warp (flow.py:547-548)."""
    v = np.array(padding_rgb, dtype=np.float32) / 255.0
    return tuple(float(x) for x in v)


def build_stabilization_warp_meta(*, source_size, output_size, framing_mode, applied_matrices, first_index: int = 0) -> Dict[str, Any]:
    return {
        "source_size": [int(source_size[0]), int(source_size[1])],
        "output_size": [int(output_size[0]), int(output_size[1])],
        "framing_mode": framing_mode,
        "matrix_convention": "source_to_stabilized",
        "per_frame": [
            {"index": i, "applied_matrix": m}
            for i, m in enumerate(np.asarray(applied_matrices, dtype=np.float32).reshape(-1, 3, 3).tolist(), first_index)
        ],
    }


def resolve_fps_for_stabilizer(frame_rate, context_fps):
    """fps rule of Flow/Classic (flow.py:230-238): widget wins, then container fps, then 16."""

    def ok(v):
        return isinstance(v, (int, float)) and np.isfinite(v) and v > 0.0

    candidate = frame_rate if ok(frame_rate) else (context_fps if ok(context_fps) else 16.0)
    effective = float(max(1.0, candidate))
    requested = float(frame_rate) if isinstance(frame_rate, (int, float)) and frame_rate > 0.0 else None
    return effective, requested


def padded_fraction(pad_count: int, pixels: int) -> float:
    """float(mask.mean()) of a 0/1 float32 mask: numpy sums exactly (< 2^24), divides in f32."""
    return float(np.float32(pad_count) / np.float32(pixels))


def padded_fractions(pad_counts, pixels: int) -> List[float]:
    """padded_fraction for every frame of a clip in one array operation."""
    return (np.asarray(pad_counts).astype(np.float32) / np.float32(pixels)).astype(np.float64).tolist()


# reference-compatible private names (the reference's tests and scripts import these)
_working_estimation_size = working_estimation_size
_rescale_transform_to_full = rescale_transform_to_full
_matrix_to_params = matrix_to_params
_params_to_matrix = params_to_matrix
_smooth_path = smooth_path
_compute_bounding_boxes = compute_bounding_boxes
_min_content_ratio = min_content_ratio
_prepare_expand_transform = prepare_expand_transform
_parse_padding_color = parse_padding_color
_build_stabilization_warp_meta = build_stabilization_warp_meta
