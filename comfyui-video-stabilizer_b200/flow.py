"""Video Stabilizer Flow driver: DIS dense optical flow -> candidate models, all on the GPU.

Mirrors nodes/video_stabilizer_flow.py of the reference: ``_create_flow_backend`` (:76-87, the DIS
configuration is compiled into libvstab's dis.cu), ``_estimate_motion_flow`` (:133-210) and
``_stabilize_frames`` (:213-640, body shared with Classic in stabilizer_core.py).  The TV-L1 and
phase-correlation fallbacks of the reference (:77-80, :110-130) are unreachable with the cv2 wheel
it pins (no cv2.optflow) and are not provided: this path has exactly one backend.
"""
from __future__ import annotations

import functools
from typing import Any, Callable, Optional, Tuple

import numpy as np
import torch

from . import _native, pipeline
from .pipeline import VideoContext
from .stabilizer_core import DeviceCandidates, PairCandidates, StabilizationResult, stabilize_frames as _core

SAMPLE_STEP = 8  # flow.py:138 sample_step


def mode_mask_for(requested: str) -> int:
    """Candidates the ladder can reach from `requested` (flow.py:156-160)."""
    bits = {"translation": 0b001, "similarity": 0b011, "perspective": 0b111}
    return bits[requested]


def estimate_candidates(context: VideoContext, work_w: int, work_h: int, requested_mode: str,
                        first_pair: int = 0, last_pair: Optional[int] = None, clip_pair_offset: int = 0,
                        out_raw: Optional[torch.Tensor] = None) -> DeviceCandidates:
    """K1/K2 -> K3 -> K4/K7-K9 for pairs [first_pair, last_pair) of the clip held by `context`
    (pair i = frames i, i+1).  Everything is enqueued and stays on the device: the [P,3] result table is returned
    as it sits in HBM (out_raw: where the fit kernels write it -- a frame-range shard passes its all-gather send buffer).
    clip_pair_offset: clip-wide index of the context's pair 0 (a frame-range shard holds a slice of the
    clip; on small frames cv2's DIS treats the first pair of a clip differently, see vstab_dis_flow_at)."""
    h = _native.get_handle(context.device)
    n = len(context)
    last_pair = n - 1 if last_pair is None else last_pair
    gray = pipeline.gray_working(context, (work_w, work_h), first_pair, last_pair + 1)
    _, grid = h.dis_flow(gray, want_flow=False, grid_step=SAMPLE_STEP, first_pair=clip_pair_offset + first_pair)
    raw = h.fit_grid(grid, SAMPLE_STEP, mode_mask_for(requested_mode), out=out_raw)
    return DeviceCandidates(raw, min_points=12)


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
    if shard is not None:
        estimator = shard.wrap_estimator(functools.partial(estimate_candidates, clip_pair_offset=shard.pair_range[0]))
    else:
        estimator = estimate_candidates
    return _core(
        context, framing_mode, transform_mode, camera_lock, strength, smooth, keep_fov, padding_rgb, frame_rate,
        estimator=estimator, flavour="flow", progress_bar=progress_bar, interrupt_check=interrupt_check,
        output=output, shard=shard,
    )


_stabilize_frames = stabilize_frames  # reference-compatible private name
