"""Legacy inverse stabilization.

What the reference's ``_apply_inverse_stabilization`` does (nodes/stabilizer_utils.py:929-1007): every frame of an
edited, stabilized clip goes back to the source canvas through the inverse of the ``applied_matrix`` that the
stabilizer recorded for it in ``meta["stabilization_warp"]`` (convention source_to_stabilized), bilinear, with the
padding mask of the pixels that have no source.  The per-frame ``cv2.warpPerspective`` pair of :975-995 is ONE fused
launch per chunk here (pipeline.fused_warp); the block checks are motion_meta's validators (same ``ValueError``
messages as the reference, owner ``stabilization_warp``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Literal, Tuple

import numpy as np

from .hostmath import border_value
from .motion_meta import _matrix_entries, _size_pair
from .pipeline import VideoContext, fused_warp

OWNER = "stabilization_warp"


@dataclass
class InverseStabilizationResult:
    frames: Any  # [N,H,W,3] float32 on the source canvas
    masks: Any   # [N,H,W,1] float32, 1 = no source pixel
    meta: Dict[str, Any]


def _warp_block(meta: Any) -> Dict[str, Any]:
    if not isinstance(meta, dict):
        raise ValueError("meta must be a dictionary containing stabilization_warp.")
    block = meta.get(OWNER)
    if not isinstance(block, dict):
        raise ValueError("meta.stabilization_warp is required for inverse stabilization.")
    convention = block.get("matrix_convention")
    if convention != "source_to_stabilized":
        raise ValueError(f"{OWNER}.matrix_convention must be 'source_to_stabilized' for inverse stabilization, got {convention!r}.")
    return block


def apply_inverse_stabilization(context: VideoContext, meta: Dict[str, Any], padding_rgb: Tuple[int, int, int], *,
                                output: Literal["host", "device"] = "host") -> InverseStabilizationResult:
    block = _warp_block(meta)
    canvas = _size_pair(OWNER, block, "source_size")       # where the frames go back to
    stabilized = _size_pair(OWNER, block, "output_size")   # what the frames must be now
    have = (context.width, context.height)
    if have != stabilized:
        raise ValueError(f"Input frames must match stabilization_warp.output_size {stabilized}, got {have}.")
    entries = block.get("per_frame")
    if not isinstance(entries, list):
        raise ValueError(f"{OWNER}.per_frame must be a list.")
    if len(entries) != len(context):
        raise ValueError(f"Frame count mismatch: got {len(context)} frame(s), metadata has {len(entries)} matrix entry/entries.")
    forward = _matrix_entries(OWNER, entries, "applied_matrix", require_finite=False)
    # stabilized -> source: float64 inverse, cast to float32 at the warp call like the reference (:972-975)
    back = np.linalg.inv(np.stack(forward)).astype(np.float32).reshape(-1, 1, 9) if forward else np.zeros((0, 1, 9), np.float32)
    border = border_value(padding_rgb)
    if context.channels == 1:  # :966 -- the clip itself is RGB by the time it is resampled
        border = (float(np.mean(np.array(padding_rgb, dtype=np.float32) / 255.0)),) * 3
    frames, masks, _ = fused_warp(context, back, canvas, "bilinear", border, want_mask=True, output=output)
    out_meta = dict(meta)
    out_meta["inverse_stabilization"] = {
        "source_size": list(canvas),
        "input_size": list(stabilized),
        "output_size": list(canvas),
        "matrix_convention": "stabilized_to_source",
        "source_matrix_convention": block.get("matrix_convention"),
        "framing_mode": block.get("framing_mode"),
        "note": "Restores original motion/canvas; pixels discarded by crop framing cannot be recovered.",
    }
    if output == "host":
        return InverseStabilizationResult(frames.numpy(), masks.numpy()[..., None], out_meta)
    return InverseStabilizationResult(frames, masks[..., None], out_meta)


_apply_inverse_stabilization = apply_inverse_stabilization  # reference-compatible private name
