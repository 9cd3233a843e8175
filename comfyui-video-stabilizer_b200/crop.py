"""`crop` framing: keep_fov scale search + padding-free refinement.

Behavioural mirror of the reference's crop branch: nodes/video_stabilizer_flow.py:386-470 (and the
identical block in video_stabilizer_classic.py), nodes/stabilizer_utils.py:518-746
(_compute_crop_with_keep_fov_parametric), :749-837 (_refine_no_padding_crop) and :448-504
(_largest_aspect_ratio_rectangle).

Where the reference warps a ones image per frame with cv2 (INTER_NEAREST), dilates / erodes it and
scans it with numpy, this module asks the GPU for
  * the bounding box of the 3x3-closed coverage of every frame  (vstab_coverage_bbox)
  * the AND of the coverage of all frames                         (vstab_common_coverage)
and keeps the O(iterations) scalar search and the single-mask integral-image search on the host.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _native, hostmath as hm


def _scaled_matrices(base_mode: str, deltas: np.ndarray, scale: float) -> np.ndarray:
    scale = float(np.clip(scale, 0.0, 1.0))
    return hm.params_to_matrices(np.asarray(deltas) * scale, base_mode)


def _evaluate_bbox_only(base_mode, deltas, width, height, scale, safety_margin_px):
    """One step of the keep_fov search: intersection of the warped frame bounds at `scale`."""
    mats = _scaled_matrices(base_mode, deltas, scale)
    mins, maxs = hm.compute_bounding_boxes(mats, width, height)
    x0, y0 = float(np.max(mins[:, 0])), float(np.max(mins[:, 1]))
    x1, y1 = float(np.min(maxs[:, 0])), float(np.min(maxs[:, 1]))
    safe_w, safe_h = max(0.0, x1 - x0), max(0.0, y1 - y0)
    margin = min(safety_margin_px, safe_w * 0.25, safe_h * 0.25)
    sx0, sy0 = x0 + margin, y0 + margin
    safe_w, safe_h = max(0.0, safe_w - 2.0 * margin), max(0.0, safe_h - 2.0 * margin)
    if safe_w <= 0.0 or safe_h <= 0.0:
        return 0.0, {"scale": scale, "pre_crop": mats, "final": mats, "crop_origin": [0.0, 0.0],
                     "crop_size": [float(width), float(height)], "has_overlap": False}
    ratio = min(1.0, safe_w / width, safe_h / height)
    cw, ch = width * ratio, height * ratio
    cx0, cy0 = sx0 + (safe_w - cw) * 0.5, sy0 + (safe_h - ch) * 0.5
    k = width / cw
    crop_matrix = np.array([[k, 0.0, -k * cx0], [0.0, k, -k * cy0], [0.0, 0.0, 1.0]], dtype=np.float32)
    return ratio, {"scale": scale, "pre_crop": mats, "final": hm.left_multiply(crop_matrix, mats),
                   "crop_origin": [cx0, cy0], "crop_size": [cw, ch], "has_overlap": True}


def _min_closed_content_ratio(device, final: np.ndarray, width: int, height: int) -> float:
    """finalize_with_masks: min over frames of the closed-coverage bounding box ratio."""
    h = _native.get_handle(device)
    fwd = torch.from_numpy(np.ascontiguousarray(np.asarray(final, dtype=np.float32).reshape(-1, 9))).to(device)
    box = h.coverage_bbox(fwd, (width, height), (width, height)).cpu().numpy().astype(np.int64)
    ratio = 1.0
    for xmin, ymin, xmax, ymax in box:
        if xmax < 0:
            r = 0.0
        else:
            r = min(float(max(1, xmax - xmin + 1)) / width, float(max(1, ymax - ymin + 1)) / height)
        if r < ratio:
            ratio = r
    return float(ratio)


def compute_crop_with_keep_fov(device, base_mode, deltas, width, height, keep_fov_target, safety_margin_px,
                               max_iterations: int = 18, interrupt_check: Optional[Callable[[], None]] = None):
    """-> (final, pre_crop, ratio_final, status, note, scale, crop_origin, crop_size)."""
    target = float(np.clip(keep_fov_target, 0.0, 1.0))
    eps = 1e-4

    def finish(raw, status, note, scale):
        if interrupt_check is not None:
            interrupt_check()
        ratio_final = _min_closed_content_ratio(device, raw["final"], width, height)
        return raw["final"], raw["pre_crop"], ratio_final, status, note, scale, list(raw["crop_origin"]), list(raw["crop_size"])

    ratio_full, raw_full = _evaluate_bbox_only(base_mode, deltas, width, height, 1.0, safety_margin_px)
    if target <= eps:
        if raw_full["has_overlap"]:
            return finish(raw_full, "disabled", None, 1.0)
        _, raw = _evaluate_bbox_only(base_mode, deltas, width, height, 0.0, safety_margin_px)
        return finish(raw, "disabled", "No common crop region at full stabilization; stabilization was disabled.", 0.0)
    if ratio_full >= target - eps:
        return finish(raw_full, "met", None, 1.0)
    low, high, best = 0.0, 1.0, None
    for _ in range(max_iterations):
        mid = 0.5 * (low + high)
        ratio_mid, raw_mid = _evaluate_bbox_only(base_mode, deltas, width, height, mid, safety_margin_px)
        if ratio_mid >= target - eps:
            best, low = raw_mid, mid
        else:
            high = mid
    if best is None:
        _, raw_zero = _evaluate_bbox_only(base_mode, deltas, width, height, 0.0, safety_margin_px)
        note = f"keep_fov target {target:.3f} could not be satisfied even with zero stabilisation."
        return finish(raw_zero, "failed", note, 0.0)
    final, pre, ratio_final, _, _, _, origin, size = finish(best, "met", None, float(best["scale"]))
    status = "met" if ratio_final >= target - eps else "clamped"
    note = None
    scale_best = float(best["scale"])
    if status == "clamped":
        note = f"keep_fov target {target:.3f} reduced to {ratio_final:.3f} at stabilisation scale {scale_best:.3f}."
    return final, pre, ratio_final, status, note, scale_best, origin, size


def _erode3(mask: np.ndarray, k: int) -> np.ndarray:
    """cv2.erode with a (2k+1)^2 rectangle; out-of-image neighbours are ignored."""
    h, w = mask.shape
    pad = np.pad(mask.astype(bool), k, mode="constant", constant_values=True)
    out = np.ones((h, w), dtype=bool)
    for dy in range(2 * k + 1):
        for dx in range(2 * k + 1):
            out &= pad[dy : dy + h, dx : dx + w]
    return out


def largest_aspect_ratio_rectangle(mask: np.ndarray, target_width: int, target_height: int):
    """Largest all-valid crop with the target aspect ratio: binary search on the crop height over an
    integral image; the centred position wins when it is valid, else the first valid one."""
    if target_width <= 0 or target_height <= 0:
        return None
    height, width = mask.shape
    aspect = float(target_width) / float(target_height)
    integral = np.zeros((height + 1, width + 1), dtype=np.float64)
    integral[1:, 1:] = np.cumsum(np.cumsum((mask > 0).astype(np.float64), axis=0), axis=1)

    def find_fit(ch: int):
        cw = int(math.ceil(aspect * ch))
        if ch <= 0 or ch > height or cw > width:
            return None
        sums = integral[ch:, cw:] - integral[:-ch, cw:] - integral[ch:, :-cw] + integral[:-ch, :-cw]
        ok = sums == cw * ch
        if not np.any(ok):
            return None
        y0 = int(np.clip(round((height - ch) * 0.5), 0, ok.shape[0] - 1))
        x0 = int(np.clip(round((width - cw) * 0.5), 0, ok.shape[1] - 1))
        if not ok[y0, x0]:
            y0, x0 = np.unravel_index(int(np.argmax(ok)), ok.shape)
        return int(x0), int(y0)

    low, high = 1, min(height, int(math.floor(width / aspect)))
    best = None
    while low <= high:
        ch = (low + high) // 2
        loc = find_fit(ch)
        if loc is None:
            high = ch - 1
        else:
            best = (loc[0], loc[1], ch)
            low = ch + 1
    if best is None:
        return None
    x0, y0, ch = best
    return float(x0), float(y0), aspect * ch, float(ch)


def refine_no_padding_crop(device, final: np.ndarray, width: int, height: int, safety_shrink_px: int = 1):
    """-> (refined matrices, crop_origin, crop_size, keep_fov_effective)."""
    h = _native.get_handle(device)
    fwd = torch.from_numpy(np.ascontiguousarray(np.asarray(final, dtype=np.float32).reshape(-1, 9))).to(device)
    common = h.common_coverage(fwd, (width, height), (width, height)).cpu().numpy() > 0
    if safety_shrink_px > 0:
        common = _erode3(common, safety_shrink_px)
    if not common.any():
        return np.asarray(final), [0.0, 0.0], [float(width), float(height)], 0.0
    rect = largest_aspect_ratio_rectangle(common, width, height)
    if rect is None:
        return np.asarray(final), [0.0, 0.0], [float(width), float(height)], 0.0
    x0, y0, cw, ch = rect
    k = width / cw
    crop_matrix = np.array([[k, 0.0, -k * x0], [0.0, k, -k * y0], [0.0, 0.0, 1.0]], dtype=np.float32)
    return hm.left_multiply(crop_matrix, np.asarray(final)), [x0, y0], [cw, ch], 1.0


def solve_crop_framing(context, base_mode, delta_full, path, target_path, keep_fov_clamped, transform_mode, camera_lock,
                       strength, smooth, fps_requested, fps_effective, padding_rgb, flow_keys, is_flow, attach, progress,
                       check, output, shard=None):
    """The crop branch of _stabilize_frames.  Returns either a finished StabilizationResult (the
    keep_fov ~= 1 bypass) or (final_matrices, apply_matrices, framing-meta additions, scale).

    shard: frame-range shard of the clip (sharding.FrameShard) or None.  `context` then holds only the rank's
    load range (own frames + halo); every clip-wide quantity comes from `path` / `delta_full`, which all ranks
    hold in full after the candidate all-gather."""
    from .stabilizer_core import StabilizationResult

    width, height = context.width, context.height
    n = len(context) if shard is None else shard.total_frames
    if keep_fov_clamped >= 0.9999:
        # per-frame lists: the rank's share of the clip (clip-wide indices), like stabilizer_core does
        m_lo, m_hi = (0, n) if shard is None else shard.meta_frame_range
        meta = {
            "frames": n,
            "note": "keep_fov~=1.0 in crop mode; returning original frames.",
            "transform_mode_requested": transform_mode,
            "transform_mode_applied": "identity",
            "camera_lock": camera_lock,
            "strength": strength,
            "strength_effective": 0.0,
            "smooth": smooth,
            "fps_requested": fps_requested,
            "fps_effective": fps_effective,
            "framing": {
                "mode": "crop",
                "input_size": [width, height],
                "keep_fov_requested": keep_fov_clamped,
                "keep_fov_effective": 1.0,
                "min_content_ratio": 1.0,
                "padding_color_rgb": [int(c) for c in padding_rgb],
                "stabilization_scale": 0.0,
            },
            "keep_fov_applied": False,
            **flow_keys,
            "stabilization_warp": hm.build_stabilization_warp_meta(
                source_size=(width, height), output_size=(width, height), framing_mode="crop",
                applied_matrices=[np.eye(3, dtype=np.float32) for _ in range(m_hi - m_lo)], first_index=m_lo,
            ),
            "estimated_motion": {
                "per_transition": [],
                "path": path[m_lo:m_hi].tolist(),
                "target_path": target_path[m_lo:m_hi].tolist(),
                "target_path_effective": path[m_lo:m_hi].tolist(),
            },
            "padding_fraction_mean": 0.0,
            "padding_fraction_max": 0.0,
        }
        progress.finish()
        frames, masks = (context if shard is None else shard.owned_context(context)).untouched(output)
        if shard is None or (m_lo, m_hi) == (0, n):
            meta = attach(meta)
        else:
            from .motion_meta import applied_motion_meta_from_matrices

            if m_hi > m_lo:
                try:
                    meta["motion_meta"] = applied_motion_meta_from_matrices(
                        np.tile(np.eye(3, dtype=np.float32), (m_hi - m_lo, 1, 1)), source_size=(width, height),
                        output_size=(width, height), fps=fps_effective, source="estimated_flow" if is_flow else "estimated_classic",
                        first_index=m_lo,
                    )
                except (KeyError, TypeError, ValueError, np.linalg.LinAlgError):
                    pass
            meta["shard"] = {"rank": shard.rank, "world": shard.world, "frame_range": list(shard.frame_range),
                             "meta_frame_range": [m_lo, m_hi]}
        return StabilizationResult(frames, masks, meta)

    safety_margin_px = max(0.5, 0.02 * max(width, height))
    final, pre_crop, _ratio, status, note, scale, _origin, _size = compute_crop_with_keep_fov(
        context.device, base_mode, delta_full, width, height, keep_fov_clamped, safety_margin_px, interrupt_check=check,
    )
    refined, crop_origin, crop_size, keep_fov_effective = refine_no_padding_crop(context.device, final, width, height, 1)
    crop_meta: Dict[str, Any] = {
        "keep_fov_status": status,
        "keep_fov_effective": keep_fov_effective,
        "crop_origin": crop_origin,
        "crop_size": crop_size,
        "actual_content_ratio": keep_fov_effective,
        "stabilization_scale": float(scale),
    }
    if note:
        crop_meta["_keep_fov_note"] = note
    return np.asarray(refined, dtype=np.float32), np.asarray(pre_crop, dtype=np.float32), crop_meta, scale
