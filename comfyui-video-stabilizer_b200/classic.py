"""Video Stabilizer Classic driver: Shi-Tomasi corners + pyramidal Lucas-Kanade on the GPU.

Mirrors nodes/video_stabilizer_classic.py of the reference: ``_estimate_motion_pair`` (:69-160:
cv2.goodFeaturesToTrack(400, 0.01, 7, blockSize=21) -> cv2.calcOpticalFlowPyrLK(31x31, 3 levels,
50 iterations / eps 0.01) -> status filter -> fit ladder without residual) and
``_stabilize_frames`` (:163-567, body shared with Flow in stabilizer_core.py).
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple

import numpy as np
import torch

from . import _native, pipeline
from .flow import mode_mask_for
from .pipeline import VideoContext
from .stabilizer_core import DeviceCandidates, PairCandidates, StabilizationResult, stabilize_frames as _core

MAX_CORNERS = 400


def estimate_candidates(context: VideoContext, work_w: int, work_h: int, requested_mode: str,
                        out_raw: Optional[torch.Tensor] = None) -> DeviceCandidates:
    """K1/K2 -> K5 (corners of frame i) -> K6 (track into frame i+1) -> K7-K9 candidates per pair; enqueued only,
    the table stays in HBM (out_raw: see flow.estimate_candidates)."""
    h = _native.get_handle(context.device)
    gray = pipeline.gray_working(context, (work_w, work_h))
    prev, curr, detected = h.gftt_lk(gray, MAX_CORNERS)  # [P,400,2] each, NaN rows = not found / lost
    raw = h.fit_points(prev, curr, mode_mask_for(requested_mode), out=out_raw)
    return DeviceCandidates(raw, min_points=8, detected=detected)


def stabilize_frames(
    context: VideoContext,
    framing_mode: str,
    transform_mode: str,
    camera_lock: bool,
    strength: float,
    smooth: float,
    keep_fov: float,
    padding_rgb: Tuple[int, int, int],
    frame_rate: float,
    *,
    progress_bar: Any = None,
    interrupt_check: Optional[Callable[[], None]] = None,
    output: str = "host",
    shard=None,
) -> StabilizationResult:
    estimator = shard.wrap_estimator(estimate_candidates) if shard is not None else estimate_candidates
    return _core(
        context, framing_mode, transform_mode, camera_lock, strength, smooth, keep_fov, padding_rgb, frame_rate,
        estimator=estimator, flavour="classic", progress_bar=progress_bar, interrupt_check=interrupt_check,
        output=output, shard=shard,
    )


_stabilize_frames = stabilize_frames
