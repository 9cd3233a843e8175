"""Legacy inverse stabilization.

Mirrors ``_apply_inverse_stabilization`` of the reference (nodes/stabilizer_utils.py:929-1007): every frame of an
edited, stabilized clip goes back to the source canvas through the inverse of the ``applied_matrix`` the
stabilizer recorded for it (``stabilization_warp``, convention source_to_stabilized), bilinear, with the padding
mask of the pixels that have no source.  Same validation order and ``ValueError`` messages.  The per-frame
``cv2.warpPerspective`` pair of :975-995 is ONE fused launch per chunk here (pipeline.fused_warp).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Literal, Tuple

import numpy as np

from .hostmath import border_value
from .pipeline import VideoContext, fused_warp


@dataclass
class InverseStabilizationResult:
    frames: Any  # [N,H,W,3] float32 on the source canvas
    masks: Any   # [N,H,W,1] float32, 1 = no source pixel
    meta: Dict[str, Any]


def _read_size_pair(meta: Dict[str, Any], key: str) -> Tuple[int, int]:
    value = meta.get(key)
    if not isinstance(value, (list, tuple)) or len(value) != 2:
        raise ValueError(f"stabilization_warp.{key} must be [width, height].")
    try:
        width, height = int(value[0]), int(value[1])
    except (TypeError, ValueError) as exc:
        raise ValueError(f"stabilization_warp.{key} must contain integer width/height.") from exc
    if width <= 0 or height <= 0:
        raise ValueError(f"stabilization_warp.{key} must contain positive width/height.")
    return width, height


def _read_applied_matrix(entry: Any, expected_index: int) -> np.ndarray:
    if not isinstance(entry, dict):
        raise ValueError(f"stabilization_warp.per_frame[{expected_index}] must be an object.")
    if entry.get("index") != expected_index:
        raise ValueError(
            f"stabilization_warp.per_frame[{expected_index}].index must be {expected_index}, got {entry.get('index')!r}."
        )
    if "applied_matrix" not in entry:
        raise ValueError(f"stabilization_warp.per_frame[{expected_index}].applied_matrix is missing.")
    matrix = np.asarray(entry["applied_matrix"], dtype=np.float64)
    if matrix.shape != (3, 3):
        raise ValueError(f"stabilization_warp.per_frame[{expected_index}].applied_matrix must be 3x3.")
    return matrix


def apply_inverse_stabilization(context: VideoContext, meta: Dict[str, Any], padding_rgb: Tuple[int, int, int], *,
                                output: Literal["host", "device"] = "host") -> InverseStabilizationResult:
    if not isinstance(meta, dict):
        raise ValueError("meta must be a dictionary containing stabilization_warp.")
    warp_meta = meta.get("stabilization_warp")
    if not isinstance(warp_meta, dict):
        raise ValueError("meta.stabilization_warp is required for inverse stabilization.")
    if warp_meta.get("matrix_convention") != "source_to_stabilized":
        raise ValueError(
            "stabilization_warp.matrix_convention must be 'source_to_stabilized' "
            f"for inverse stabilization, got {warp_meta.get('matrix_convention')!r}."
        )
    source_size = _read_size_pair(warp_meta, "source_size")
    output_size = _read_size_pair(warp_meta, "output_size")
    if (context.width, context.height) != output_size:
        raise ValueError(
            f"Input frames must match stabilization_warp.output_size {output_size}, got {(context.width, context.height)}."
        )
    per_frame = warp_meta.get("per_frame")
    if not isinstance(per_frame, list):
        raise ValueError("stabilization_warp.per_frame must be a list.")
    if len(per_frame) != len(context):
        raise ValueError(
            f"Frame count mismatch: got {len(context)} frame(s), metadata has {len(per_frame)} matrix entry/entries."
        )
    inverses = []
    for idx, entry in enumerate(per_frame):
        matrix = _read_applied_matrix(entry, idx)
        try:
            inverses.append(np.linalg.inv(matrix).astype(np.float32))
        except np.linalg.LinAlgError as exc:
            raise ValueError(f"stabilization_warp.per_frame[{idx}].applied_matrix is not invertible.") from exc
    border = border_value(padding_rgb)
    if context.channels == 1:  # stabilizer_utils.py:966 (the clip itself is RGB by the time it is resampled)
        border = (float(np.mean(np.array(padding_rgb, dtype=np.float32) / 255.0)),) * 3
    fwd = np.stack(inverses, axis=0).reshape(-1, 1, 9) if inverses else np.zeros((0, 1, 9), np.float32)
    frames, masks, _ = fused_warp(context, fwd, source_size, "bilinear", border, want_mask=True, output=output)
    result_meta = dict(meta)
    result_meta["inverse_stabilization"] = {
        "source_size": [int(source_size[0]), int(source_size[1])],
        "input_size": [int(output_size[0]), int(output_size[1])],
        "output_size": [int(source_size[0]), int(source_size[1])],
        "matrix_convention": "stabilized_to_source",
        "source_matrix_convention": warp_meta.get("matrix_convention"),
        "framing_mode": warp_meta.get("framing_mode"),
        "note": "Restores original motion/canvas; pixels discarded by crop framing cannot be recovered.",
    }
    if output == "host":
        return InverseStabilizationResult(frames.numpy(), masks.numpy()[..., None], result_meta)
    return InverseStabilizationResult(frames, masks[..., None], result_meta)


_apply_inverse_stabilization = apply_inverse_stabilization  # reference-compatible private name
