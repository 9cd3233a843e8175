"""Motion Apply engine on the fused CUDA resampler.

Behavioural mirror of the reference's ``nodes/motion_apply.py``: ``apply_motion`` (:297-429)
with its helpers ``_resolve_motion_for_context`` (:45-67), ``_validate_context`` (:32-42),
``_blurred_matrix_samples`` (:125-134), ``_common_valid_mask`` / ``_center_crop_matrix_from_common``
(:205-285) and ``_expand_matrices`` (:288-294).  The per-frame cv2.warpPerspective calls
(image + INTER_NEAREST ones) and the float32 shutter accumulate of ``_warp_with_matrices`` /
``_warp_with_motion_blur`` (:75-202) are ONE launch of ``vstab_warp_fused``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, Literal, Optional, Tuple

import numpy as np
import torch

from . import _native
from .hostmath import border_value, compute_bounding_boxes, gc_paused, prepare_expand_transform
from .motion_meta import MotionMeta, motion_meta_from_stabilization_warp, resolve_motion_meta
from .pipeline import VideoContext, fused_warp

ProgressCallback = Callable[[], None]


@dataclass
class MotionApplyResult:
    frames: Any  # [N,H',W',3] float32 (numpy view of a pinned CPU tensor, or CUDA tensor)
    masks: Any   # [N,H',W',1] float32
    meta: Dict[str, Any]


def _check_interpolation(interpolation: str) -> str:
    if interpolation not in ("bilinear", "bicubic"):
        raise ValueError(f"Unsupported interpolation {interpolation!r}; expected 'bilinear' or 'bicubic'.")
    return interpolation


def _validate_context(context: VideoContext, motion: MotionMeta) -> None:
    if (context.width, context.height) != motion.input_size:
        raise ValueError(
            "Input frames must match motion_meta.input_size "
            f"{motion.input_size}, got {(context.width, context.height)}."
        )
    if len(context) != motion.frame_count:
        raise ValueError(
            "Frame count mismatch: "
            f"got {len(context)} frame(s), metadata has {motion.frame_count} matrix entry/entries."
        )


def _resolve_motion_for_context(meta: Dict[str, Any], context: VideoContext) -> MotionMeta:
    """Pick the block whose input_size matches the frames, else the inverted legacy warp."""
    if not isinstance(meta, dict):
        return resolve_motion_meta(meta)
    size = (context.width, context.height)
    block = meta.get("motion_meta")
    if isinstance(block, dict):
        motion = resolve_motion_meta({"motion_meta": block})
        if motion.input_size == size:
            return motion
    warp_meta = meta.get("stabilization_warp")
    if isinstance(warp_meta, dict):
        fps = float(block.get("fps", 16.0)) if isinstance(block, dict) else 16.0
        inverse = motion_meta_from_stabilization_warp(warp_meta, fps=fps, source="legacy_stabilization")
        if inverse is not None:
            motion = resolve_motion_meta({"motion_meta": inverse})
            if motion.input_size == size:
                return motion
    return resolve_motion_meta(meta)


def sample_matrices(matrices, motion_blur: float, samples: int) -> np.ndarray:
    """[N,S,9] float32 forward matrices: the shutter samples of every frame (f64 lerp -> f32)."""
    n = len(matrices)
    base = np.stack([np.asarray(m, dtype=np.float64) for m in matrices], axis=0)
    if motion_blur <= 0.0 or samples <= 1 or n <= 1:
        return base.astype(np.float32).reshape(n, 1, 9)
    delta = np.empty_like(base)
    delta[:-1] = base[1:] - base[:-1]
    delta[-1] = base[-1] - base[-2]
    ts = np.linspace(0.0, float(motion_blur), int(samples), dtype=np.float64)
    out = base[:, None] + delta[:, None] * ts[None, :, None, None]
    return out.astype(np.float32).reshape(n, int(samples), 9)


def _tick(cb: Optional[ProgressCallback], count: int) -> None:
    if cb is not None:
        for _ in range(count):
            cb()


def _run_warp(context, matrices, output_size, interpolation, padding_rgb, motion_blur, samples,
              masks_zero, progress_callback, output):
    n = len(matrices)
    blur_on = motion_blur > 0.0 and samples > 1
    fwd = sample_matrices(matrices, motion_blur if blur_on else 0.0, samples)
    frames, masks, _ = fused_warp(
        context, fwd, output_size, interpolation, border_value(padding_rgb),
        want_mask=not masks_zero, output=output,
    )
    s_nominal = int(np.clip(samples, 3, 33)) if blur_on else 1
    if blur_on and n <= 1:
        # Reference quirk (motion_apply.py:126-127 with :195): a single frame yields ONE sample
        # but is still divided by the nominal sample count.
        frames /= float(s_nominal)
        if masks is not None:
            masks.copy_(1.0 - (1.0 - masks) / float(s_nominal))
            masks[masks < 1e-3] = 0.0
    _tick(progress_callback, n * (s_nominal if n > 1 or not blur_on else 1))
    if masks is None:
        oh, ow = frames.shape[1], frames.shape[2]
        masks = torch.zeros((n, oh, ow), dtype=torch.float32, device=frames.device)
    return frames, masks


def common_valid_mask(context: VideoContext, input_size, output_size, matrices, progress_callback=None) -> np.ndarray:
    """AND of the INTER_NEAREST coverage of every matrix (bool [H',W'])."""
    h = _native.get_handle(context.device)
    fwd = np.stack([np.asarray(m, dtype=np.float32).reshape(9) for m in matrices], axis=0)
    fwd_t = torch.from_numpy(np.ascontiguousarray(fwd)).to(context.device)
    common = h.common_coverage(fwd_t, input_size, output_size).cpu().numpy().astype(bool)
    _tick(progress_callback, len(matrices))
    return common


def center_crop_matrix_from_common(common: np.ndarray, output_size) -> Optional[np.ndarray]:
    """Smallest centred zoom (binary search, <=4x) whose crop window is all-valid."""
    ow, oh = output_size
    cx, cy = (ow - 1) * 0.5, (oh - 1) * 0.5
    aspect = ow / float(oh)

    def window(scale: float):
        cw = max(1.0, ow / scale)
        ch = cw / aspect
        if ch > oh:
            ch = oh / scale
            cw = ch * aspect
        return cw, ch

    def fits(scale: float) -> bool:
        cw, ch = window(scale)
        x0, y0 = int(np.ceil(cx - cw * 0.5)), int(np.ceil(cy - ch * 0.5))
        x1, y1 = int(np.floor(cx + cw * 0.5)), int(np.floor(cy + ch * 0.5))
        if x0 < 0 or y0 < 0 or x1 >= ow or y1 >= oh or x1 <= x0 or y1 <= y0:
            return False
        return bool(common[y0 : y1 + 1, x0 : x1 + 1].all())

    lo, hi = 0.0, 1.0
    if not fits(1.0):
        while hi <= 4.0 and not fits(hi):
            hi *= 1.25
        if hi > 4.0:
            return None
    for _ in range(32):
        mid = max((lo + hi) * 0.5, 1.0)
        if fits(mid):
            hi = mid
        else:
            lo = mid
    scale = float(hi)
    cw = ow / scale
    ch = cw / aspect
    if ch > oh:
        ch = oh / scale
        cw = ch * aspect
    x0, y0 = cx - cw * 0.5, cy - ch * 0.5
    return np.array([[scale, 0.0, -scale * x0], [0.0, scale, -scale * y0], [0.0, 0.0, 1.0]], dtype=np.float64)


def expand_matrices(matrices, input_size):
    mins, maxs = compute_bounding_boxes(matrices, input_size[0], input_size[1])
    shift, out_size = prepare_expand_transform(mins, maxs)
    return [shift @ m for m in matrices], out_size


def apply_motion(
    context: VideoContext,
    meta: Dict[str, Any],
    padding_rgb: Tuple[int, int, int],
    *,
    framing_mode: str = "crop_and_pad",
    interpolation: str = "bilinear",
    motion_blur: float = 0.0,
    motion_blur_samples: int = 9,
    progress_callback: Optional[ProgressCallback] = None,
    output: Literal["host", "device"] = "host",
) -> MotionApplyResult:
    with gc_paused():  # per-frame matrices and meta are acyclic; see hostmath.gc_paused
        return _apply_motion(context, meta, padding_rgb, framing_mode, interpolation, motion_blur, motion_blur_samples,
                             progress_callback, output)


def _apply_motion(context, meta, padding_rgb, framing_mode, interpolation, motion_blur, motion_blur_samples, progress_callback,
                  output) -> MotionApplyResult:
    motion = _resolve_motion_for_context(meta, context)
    _validate_context(context, motion)

    matrices = [t.matrix for t in motion.per_frame]
    output_size = motion.output_size
    _check_interpolation(interpolation)
    result_meta = dict(meta)
    requested = "crop_and_pad" if framing_mode == "pad" else framing_mode
    effective = requested
    motion_blur = float(np.clip(motion_blur, 0.0, 1.0))
    motion_blur_samples = int(np.clip(motion_blur_samples, 3, 33))
    masks_zero = False

    if requested == "crop_and_pad":
        pass
    elif requested == "crop":
        common = common_valid_mask(context, motion.input_size, output_size, matrices, progress_callback)
        crop = center_crop_matrix_from_common(common, output_size)
        if crop is None:
            result_meta["framing_fallback"] = "crop_and_pad"
            effective = "crop_and_pad"
        else:
            matrices = [crop @ m for m in matrices]
            masks_zero = True
    elif requested == "expand":
        matrices, output_size = expand_matrices(matrices, motion.input_size)
    else:
        raise ValueError(f"Unsupported framing_mode {framing_mode!r}; expected 'crop_and_pad', 'crop', or 'expand'.")

    frames, masks = _run_warp(
        context, matrices, output_size, interpolation, padding_rgb, motion_blur, motion_blur_samples,
        masks_zero, progress_callback, output,
    )
    result_meta["motion_apply"] = {
        "input_size": [int(motion.input_size[0]), int(motion.input_size[1])],
        "output_size": [int(output_size[0]), int(output_size[1])],
        "framing_mode": effective,
        "interpolation": interpolation,
        "motion_blur": motion_blur,
        "motion_blur_samples": motion_blur_samples,
        "source": motion.source,
    }
    if output == "host":
        return MotionApplyResult(frames.numpy(), masks.numpy()[..., None], result_meta)
    return MotionApplyResult(frames, masks[..., None], result_meta)
