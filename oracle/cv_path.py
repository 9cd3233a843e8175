"""The reference's OpenCV path re-stated over the real cv2 calls.  TEST INFRASTRUCTURE ONLY.

This is the "reference CPU implementation of the path": the same sequence of cv2 / numpy
operations, per frame and per pair, as
  nodes/stabilizer_utils.py:96-197   input adaptation (per-frame max() range test, views)
  nodes/stabilizer_utils.py:236-276  gray + INTER_AREA working image
  nodes/video_stabilizer_flow.py:76-87, :133-210     DIS + grid sampling + fit ladder
  nodes/video_stabilizer_classic.py:69-160           GFTT + pyramidal LK + fit ladder
  nodes/video_stabilizer_flow.py:213-640             _stabilize_frames (crop_and_pad / expand)
  nodes/motion_apply.py:75-202, :297-429             Motion Apply warp / blur
written from the behaviour documented in SURVEY.md, not copied.  It is used (a) by the parity
tests as the end-to-end checker where cv2 is importable and (b) by bench.py as the timed CPU
baseline / `--impl reference` arm on the GPU box, where /root/reference does not exist.  In the
build container tests/test_cv_path_vs_reference.py checks it against the unmodified reference.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Sequence, Tuple

import cv2
import numpy as np

LADDER = {
    "perspective": ("perspective", "similarity", "translation"),
    "similarity": ("similarity", "translation"),
    "translation": ("translation",),
}


# ------------------------------------------------------------------------------ adapters ----

def normalize_frames(frames) -> List[np.ndarray]:
    """Per-frame float32 RGB 0..1 arrays; float32 contiguous 0..1 input stays a view."""
    out = []
    for f in frames:
        arr = f.detach().cpu().numpy() if hasattr(f, "detach") else np.asarray(f)
        if arr.dtype == np.uint8:
            arr = arr.astype(np.float32)
            arr /= 255.0
        elif arr.size and float(arr.max()) > 1.5:
            arr = arr.astype(np.float32)
            arr /= 255.0
        elif arr.dtype != np.float32 or not arr.flags["C_CONTIGUOUS"]:
            arr = np.ascontiguousarray(arr, dtype=np.float32)
        if arr.shape[2] == 1:
            arr = np.repeat(arr, 3, axis=2)
        elif arr.shape[2] > 3:
            arr = arr[..., :3]
        out.append(arr)
    return out


def working_size(width: int, height: int, max_side: int = 960):
    longest = max(int(width), int(height))
    if longest <= max_side:
        return None
    s = max_side / float(longest)
    w, h = max(1, int(round(width * s))), max(1, int(round(height * s)))
    return None if (w >= width or h >= height) else (w, h)


def gray_for_estimation(frame: np.ndarray, ws) -> np.ndarray:
    g = cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY)
    g = np.clip(g * 255.0, 0, 255).astype(np.uint8)
    return g if ws is None else cv2.resize(g, ws, interpolation=cv2.INTER_AREA)


# ---------------------------------------------------------------------------- estimators ----

def make_dis():
    dis = cv2.DISOpticalFlow.create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
    dis.setFinestScale(2)
    dis.setPatchSize(8)
    dis.setPatchStride(4)
    dis.setUseSpatialPropagation(True)
    return dis


def _fit_ladder(prev, curr, n_total, requested, with_residual):
    eye = np.eye(3, dtype=np.float32)
    for mode in LADDER[requested]:
        if mode == "perspective" and len(prev) >= 4:
            H, inl = cv2.findHomography(prev, curr, method=cv2.RANSAC, ransacReprojThreshold=2.5, maxIters=2000, confidence=0.992)
            if H is not None and inl is not None:
                conf = float(inl.sum()) / float(len(prev))
                if conf >= 0.15:
                    res = float(np.abs((prev @ H[:2, :2].T + H[:2, 2]) - curr).mean()) if with_residual else None
                    return H.astype(np.float32), "perspective", conf, res
        elif mode == "similarity" and len(prev) >= 3:
            A, inl = cv2.estimateAffinePartial2D(prev, curr, method=cv2.RANSAC, ransacReprojThreshold=2.0, maxIters=2000, confidence=0.992)
            if A is not None:
                conf = float(inl.sum()) / float(len(prev)) if inl is not None else 0.0
                if conf >= 0.1:
                    m = np.vstack([A, np.array([0.0, 0.0, 1.0], dtype=np.float32)])
                    res = float(np.abs((prev @ A[:, :2].T + A[:, 2]) - curr).mean()) if with_residual else None
                    return m.astype(np.float32), "similarity", conf, res
        elif mode == "translation":
            shifts = (curr - prev).reshape(-1, 2)
            delta = np.median(shifts, axis=0).reshape(-1).astype(np.float32)
            tx, ty = float(delta[0]), float(delta[1])
            m = np.array([[1.0, 0.0, tx], [0.0, 1.0, ty], [0.0, 0.0, 1.0]], dtype=np.float32)
            conf = float(len(prev)) / float(n_total)
            res = float(np.abs((prev + np.array([tx, ty], dtype=np.float32)) - curr).mean()) if with_residual else None
            return m, "translation", conf, res
    return eye, "translation", 0.0, (0.0 if with_residual else None)


def estimate_flow_pair(dis, prev_gray, curr_gray, requested, step: int = 8):
    flow = dis.calc(prev_gray, curr_gray, None)
    h, w = prev_gray.shape
    gy, gx = np.meshgrid(np.arange(0, h, step, dtype=np.int32), np.arange(0, w, step, dtype=np.int32), indexing="ij")
    prev = np.stack([gx.ravel(), gy.ravel()], axis=1).astype(np.float32)
    curr = prev + flow[gy, gx].reshape(-1, 2)
    ok = np.isfinite(curr).all(axis=1)
    pv, cv = prev[ok], curr[ok]
    if len(pv) < 12:
        return np.eye(3, dtype=np.float32), "translation", 0.0, 0.0
    return _fit_ladder(pv, cv, len(prev), requested, True)


def estimate_classic_pair(prev_gray, curr_gray, requested):
    feats = cv2.goodFeaturesToTrack(prev_gray, maxCorners=400, qualityLevel=0.01, minDistance=7, blockSize=21, mask=None)
    if feats is None or len(feats) < 12:
        return np.eye(3, dtype=np.float32), "translation", 0.0, None
    nxt, status, _ = cv2.calcOpticalFlowPyrLK(
        prev_gray, curr_gray, feats, None, winSize=(31, 31), maxLevel=3,
        criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 50, 0.01),
    )
    status = status.reshape(-1)
    gp, gc = feats[status == 1].reshape(-1, 2), nxt[status == 1].reshape(-1, 2)
    if len(gp) < 8:
        return np.eye(3, dtype=np.float32), "translation", 0.0, None
    return _fit_ladder(gp, gc, len(feats), requested, False)


# ------------------------------------------------------------------------------ host math ----

def rescale_to_full(m, src_size, ws):
    sx, sy = ws[0] / float(src_size[0]), ws[1] / float(src_size[1])
    S = np.array([[sx, 0, 0], [0, sy, 0], [0, 0, 1.0]])
    Si = np.array([[1.0 / sx, 0, 0], [0, 1.0 / sy, 0], [0, 0, 1.0]])
    return (Si @ m.astype(np.float64) @ S).astype(np.float32)


def to_params(m, mode):
    if mode == "translation":
        return np.array([m[0, 2], m[1, 2]], dtype=np.float64)
    if mode == "similarity":
        a, c = m[0, 0], m[1, 0]
        return np.array([m[0, 2], m[1, 2], math.atan2(c, a), math.log(math.sqrt(max(a * a + c * c, 1e-10)))], dtype=np.float64)
    return np.array([m[0, 0] - 1.0, m[0, 1], m[0, 2], m[1, 0], m[1, 1] - 1.0, m[1, 2], m[2, 0], m[2, 1]], dtype=np.float64)


def to_matrix(p, mode):
    if mode == "translation":
        return np.array([[1.0, 0.0, p[0]], [0.0, 1.0, p[1]], [0.0, 0.0, 1.0]], dtype=np.float32)
    if mode == "similarity":
        k, c, s = math.exp(p[3]), math.cos(p[2]), math.sin(p[2])
        return np.array([[k * c, -k * s, p[0]], [k * s, k * c, p[1]], [0.0, 0.0, 1.0]], dtype=np.float32)
    return np.array([[p[0] + 1.0, p[1], p[2]], [p[3], p[4] + 1.0, p[5]], [p[6], p[7], 1.0]], dtype=np.float32)


def box_smooth(path, smooth, fps):
    smooth = float(np.clip(smooth, 0.0, 1.0))
    if smooth <= 0.0 or len(path) <= 2:
        return path.copy()
    fps = float(max(1.0, fps))
    win = max(3, int(round((3.0 / 16.0 + smooth * (13.0 / 16.0 - 3.0 / 16.0)) * fps)))
    if win % 2 == 0:
        win += 1
    pad = win // 2
    k = np.ones(win, dtype=np.float64) / float(win)
    out = np.zeros_like(path)
    for d in range(path.shape[1]):
        out[:, d] = np.convolve(np.pad(path[:, d], (pad, pad), mode="edge"), k, mode="valid")
    return out


def bboxes(mats, w, h):
    corners = np.array([[0.0, 0.0, 1.0], [w, 0.0, 1.0], [0.0, h, 1.0], [w, h, 1.0]], dtype=np.float64).T
    lo, hi = [], []
    for m in mats:
        q = m @ corners
        q /= q[2, :]
        lo.append([q[0].min(), q[1].min()])
        hi.append([q[0].max(), q[1].max()])
    return np.array(lo), np.array(hi)


# -------------------------------------------------------------------------------- drivers ----

def stabilize(frames_in, node: str, framing: str, mode: str, camera_lock: bool, strength: float, smooth: float,
              keep_fov: float, padding_rgb, fps: float) -> Tuple[np.ndarray, np.ndarray, Dict[str, Any]]:
    """Flow ('flow') or Classic ('classic') stabilizer, crop_and_pad / expand framing."""
    if framing not in ("crop_and_pad", "expand"):
        raise NotImplementedError("the oracle restates crop_and_pad and expand framing")
    frames = normalize_frames(frames_in)
    n = len(frames)
    h, w = frames[0].shape[:2]
    fps_eff = float(max(1.0, fps if (isinstance(fps, (int, float)) and np.isfinite(fps) and fps > 0) else 16.0))
    ws = working_size(w, h)
    grays = [gray_for_estimation(f, ws) for f in frames]
    dis = make_dis() if node == "flow" else None
    active = mode
    mats, confs, resids, modes, deltas = [], [], [], [], []
    for i in range(1, n):
        if node == "flow":
            m, used, conf, res = estimate_flow_pair(dis, grays[i - 1], grays[i], active)
        else:
            m, used, conf, res = estimate_classic_pair(grays[i - 1], grays[i], active)
        if used != active:
            active = used
        if ws is not None:
            m = rescale_to_full(m, (w, h), ws)
        mats.append(m); confs.append(conf); resids.append(res); modes.append(used)
        deltas.append(to_params(m, mode))
    del grays
    path = np.zeros((n, deltas[0].shape[0]), dtype=np.float64)
    for i, d in enumerate(deltas, start=1):
        path[i] = path[i - 1] + d
    strength = float(np.clip(strength, 0.0, 1.0))
    smooth = float(np.clip(smooth, 0.0, 1.0))
    if camera_lock:
        smooth = max(smooth, 0.85)
        target = np.zeros_like(path)
    else:
        target = path + strength * (box_smooth(path, smooth, fps_eff) - path)
    diffs = target - path
    apply_m = [to_matrix(d, mode) for d in diffs]
    lo, hi = bboxes(apply_m, w, h)
    out_size = (w, h)
    framing_meta: Dict[str, Any] = {"mode": framing}
    if framing == "crop_and_pad":
        x0, y0 = float(np.max(lo[:, 0])), float(np.max(lo[:, 1]))
        x1, y1 = float(np.min(hi[:, 0])), float(np.min(hi[:, 1]))
        ox, oy = w * 0.5 - (x0 + x1) * 0.5, h * 0.5 - (y0 + y1) * 0.5
        T = np.array([[1.0, 0.0, ox], [0.0, 1.0, oy], [0.0, 0.0, 1.0]], dtype=np.float32)
        final = [T @ m for m in apply_m]
        framing_meta["center_offset"] = [ox, oy]
    else:
        x_min, y_min = float(np.min(lo[:, 0])), float(np.min(lo[:, 1]))
        x_max, y_max = float(np.max(hi[:, 0])), float(np.max(hi[:, 1]))
        T = np.array([[1.0, 0.0, -x_min], [0.0, 1.0, -y_min], [0.0, 0.0, 1.0]], dtype=np.float32)
        out_size = (max(int(math.ceil(x_max - x_min)), 1), max(int(math.ceil(y_max - y_min)), 1))
        final = [T @ m for m in apply_m]
        framing_meta["expanded_size"] = list(out_size)
    border = (np.array(padding_rgb, dtype=np.float32) / 255.0).tolist()
    ow, oh = out_size
    out_frames = np.empty((n, oh, ow, 3), dtype=np.float32)
    out_masks = np.empty((n, oh, ow, 1), dtype=np.float32)
    ratios = []
    for i, m in enumerate(final):
        warped = cv2.warpPerspective(frames[i], m, out_size, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=border)
        out_frames[i] = warped.astype(np.float32)
        frames[i] = None
        content = cv2.warpPerspective(np.ones((h, w), dtype=np.float32), m, out_size, flags=cv2.INTER_NEAREST,
                                      borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
        mask = 1.0 - (content > 0.5).astype(np.float32)
        mask[mask < 1e-3] = 0.0
        ratios.append(float(mask.mean()))
        out_masks[i, ..., 0] = mask
    meta = {
        "frames": n,
        "transform_mode_applied": active,
        "framing": framing_meta,
        "final_matrices": [np.asarray(m, dtype=np.float32) for m in final],
        "per_transition": [
            {"index": i, "mode": modes[i], "confidence": confs[i], "residual": resids[i], "matrix": mats[i]} for i in range(n - 1)
        ],
        "path": path,
        "target_path": target,
        "padding_fraction_mean": float(np.mean(ratios)),
        "padding_fraction_max": float(np.max(ratios)),
        "output_size": out_size,
    }
    return out_frames, out_masks, meta


def blur_samples(mats, idx, blur, count):
    if len(mats) <= 1:
        return [mats[idx]]
    base = np.asarray(mats[idx], dtype=np.float64)
    delta = (np.asarray(mats[idx + 1], dtype=np.float64) - base) if idx < len(mats) - 1 else (base - np.asarray(mats[idx - 1], dtype=np.float64))
    return [base + delta * t for t in np.linspace(0.0, float(blur), int(count), dtype=np.float64)]


def apply_motion(frames_in, matrices: Sequence[np.ndarray], input_size, output_size, padding_rgb, framing="crop_and_pad",
                 interpolation="bilinear", motion_blur=0.0, samples=9):
    """Motion Apply for crop_and_pad / expand (matrices = motion_meta per_frame, float64)."""
    frames = normalize_frames(frames_in)
    mats = [np.asarray(m, dtype=np.float64) for m in matrices]
    w, h = input_size
    out_size = tuple(output_size)
    if framing == "expand":
        lo, hi = bboxes(mats, w, h)
        x_min, y_min = float(np.min(lo[:, 0])), float(np.min(lo[:, 1]))
        x_max, y_max = float(np.max(hi[:, 0])), float(np.max(hi[:, 1]))
        T = np.array([[1.0, 0.0, -x_min], [0.0, 1.0, -y_min], [0.0, 0.0, 1.0]], dtype=np.float32)
        out_size = (max(int(math.ceil(x_max - x_min)), 1), max(int(math.ceil(y_max - y_min)), 1))
        mats = [T @ m for m in mats]
    elif framing not in ("crop_and_pad", "pad"):
        raise NotImplementedError(framing)
    flag = cv2.INTER_LINEAR if interpolation == "bilinear" else cv2.INTER_CUBIC
    border = (np.array(padding_rgb, dtype=np.float32) / 255.0).tolist()
    n = len(mats)
    ow, oh = out_size
    out_f = np.empty((n, oh, ow, 3), dtype=np.float32)
    out_m = np.zeros((n, oh, ow, 1), dtype=np.float32)
    ones = np.ones((h, w), dtype=np.float32)
    blur = float(np.clip(motion_blur, 0.0, 1.0))
    S = int(np.clip(samples, 3, 33))
    for i in range(n):
        if blur <= 0.0:
            m32 = np.asarray(mats[i], dtype=np.float32)
            out_f[i] = cv2.warpPerspective(frames[i], m32, out_size, flags=flag, borderMode=cv2.BORDER_CONSTANT, borderValue=border).astype(np.float32)
            content = cv2.warpPerspective(ones, m32, out_size, flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
            mask = 1.0 - (content > 0.5).astype(np.float32)
            mask[mask < 1e-3] = 0.0
            out_m[i, ..., 0] = mask
        else:
            acc = np.zeros((oh, ow, 3), dtype=np.float32)
            cov = np.zeros((oh, ow), dtype=np.float32)
            for m in blur_samples(mats, i, blur, S):
                m32 = np.asarray(m, dtype=np.float32)
                acc += cv2.warpPerspective(frames[i], m32, out_size, flags=flag, borderMode=cv2.BORDER_CONSTANT, borderValue=border).astype(np.float32)
                c = cv2.warpPerspective(ones, m32, out_size, flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
                cov += (c > 0.5).astype(np.float32)
            out_f[i] = acc / float(S)
            mask = 1.0 - cov / float(S)
            mask[mask < 1e-3] = 0.0
            out_m[i, ..., 0] = mask
    return out_f, out_m, out_size
