"""motion_meta v2 JSON contract (host side, no pixels).

Behavioural mirror of the reference's ``nodes/motion_meta.py`` (validate :62-100, build
:123-152, from / applied-from stabilization_warp :155-220, resolve :223-235): the block this
module emits is part of the drop-in contract, and the messages of the ``ValueError``s it raises
are what callers of the reference see.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class FrameTransform:
    index: int
    matrix: np.ndarray


@dataclass(frozen=True)
class MotionMeta:
    source: str
    frame_count: int
    fps: float
    input_size: Tuple[int, int]
    output_size: Tuple[int, int]
    per_frame: List[FrameTransform]
    generator: Optional[Dict[str, Any]] = None


def _size_pair(owner: str, block: Dict[str, Any], key: str) -> Tuple[int, int]:
    raw = block.get(key)
    if not isinstance(raw, (list, tuple)) or len(raw) != 2:
        raise ValueError(f"{owner}.{key} must be [width, height].")
    try:
        w, h = int(raw[0]), int(raw[1])
    except (TypeError, ValueError) as exc:
        raise ValueError(f"{owner}.{key} must contain integer width/height.") from exc
    if w <= 0 or h <= 0:
        raise ValueError(f"{owner}.{key} must contain positive width/height.")
    return w, h


def _matrix_entry(owner: str, entry: Any, position: int, key: str, require_finite: bool = True) -> np.ndarray:
    where = f"{owner}.per_frame[{position}]"
    if not isinstance(entry, dict):
        raise ValueError(f"{where} must be an object.")
    if entry.get("index") != position:
        raise ValueError(f"{where}.index must be {position}, got {entry.get('index')!r}.")
    if key not in entry:
        raise ValueError(f"{where}.{key} is missing.")
    m = np.asarray(entry[key], dtype=np.float64)
    if m.shape != (3, 3):
        raise ValueError(f"{where}.{key} must be 3x3.")
    if require_finite and not np.isfinite(m).all():
        raise ValueError(f"{where}.{key} must contain finite numbers.")
    try:
        np.linalg.inv(m)
    except np.linalg.LinAlgError as exc:
        raise ValueError(f"{where}.{key} is not invertible.") from exc
    return m


def _matrix_entries(owner: str, entries: list, key: str, require_finite: bool = True) -> List[np.ndarray]:
    """All per_frame matrices, validated.  Vectorised happy path (one stacked isfinite + inv);
    any irregularity re-runs the per-entry checks so the first offending entry is reported with
    exactly the reference's message.  require_finite=False: the legacy inverse helper of the reference
    (stabilizer_utils.py:898-927) has no finiteness check, only shape and invertibility."""
    try:
        if all(isinstance(e, dict) and e.get("index") == i and key in e for i, e in enumerate(entries)):
            stack = np.asarray([e[key] for e in entries], dtype=np.float64)
            if stack.shape == (len(entries), 3, 3) and (not require_finite or np.isfinite(stack).all()):
                np.linalg.inv(stack)
                return list(stack)
    except (ValueError, TypeError, np.linalg.LinAlgError):
        pass
    return [_matrix_entry(owner, e, i, key, require_finite) for i, e in enumerate(entries)]


def validate_motion_meta(block: Dict[str, Any]) -> None:
    if not isinstance(block, dict):
        raise ValueError("motion_meta must be an object.")
    if block.get("version") != 2:
        raise ValueError(f"motion_meta.version must be 2, got {block.get('version')!r}.")
    if block.get("matrix_convention") != "input_to_output":
        raise ValueError(
            "motion_meta.matrix_convention must be 'input_to_output', " f"got {block.get('matrix_convention')!r}."
        )
    source = block.get("source")
    if not isinstance(source, str) or not source:
        raise ValueError("motion_meta.source must be a non-empty string.")
    try:
        count = int(block.get("frame_count"))
    except (TypeError, ValueError) as exc:
        raise ValueError("motion_meta.frame_count must be an integer.") from exc
    if count < 0:
        raise ValueError("motion_meta.frame_count must be non-negative.")
    try:
        fps = float(block.get("fps"))
    except (TypeError, ValueError) as exc:
        raise ValueError("motion_meta.fps must be a positive number.") from exc
    if not np.isfinite(fps) or fps <= 0.0:
        raise ValueError("motion_meta.fps must be a positive number.")
    _size_pair("motion_meta", block, "input_size")
    _size_pair("motion_meta", block, "output_size")
    entries = block.get("per_frame")
    if not isinstance(entries, list):
        raise ValueError("motion_meta.per_frame must be a list.")
    if len(entries) != count:
        raise ValueError(
            "motion_meta.frame_count mismatch: " f"frame_count is {count}, per_frame has {len(entries)} entry/entries."
        )
    _matrix_entries("motion_meta", entries, "matrix")
    if source == "generated_shake" and not isinstance(block.get("generator"), dict):
        raise ValueError("motion_meta.generator is required when source is 'generated_shake'.")


def _parse_block(block: Dict[str, Any]) -> MotionMeta:
    validate_motion_meta(block)
    gen = block.get("generator")
    return MotionMeta(
        source=str(block["source"]),
        frame_count=int(block["frame_count"]),
        fps=float(block["fps"]),
        input_size=_size_pair("motion_meta", block, "input_size"),
        output_size=_size_pair("motion_meta", block, "output_size"),
        per_frame=[
            FrameTransform(index=i, matrix=np.asarray(e["matrix"], dtype=np.float64))
            for i, e in enumerate(block["per_frame"])
        ],
        generator=dict(gen) if isinstance(gen, dict) else None,
    )


def build_motion_meta_v2(
    *,
    source: str,
    frame_count: int,
    fps: float,
    input_size: Tuple[int, int],
    output_size: Tuple[int, int],
    matrices: Sequence[np.ndarray],
    generator: Optional[Dict[str, Any]] = None,
) -> Dict[str, Any]:
    block: Dict[str, Any] = {
        "version": 2,
        "source": source,
        "frame_count": int(frame_count),
        "fps": float(fps),
        "input_size": [int(input_size[0]), int(input_size[1])],
        "output_size": [int(output_size[0]), int(output_size[1])],
        "matrix_convention": "input_to_output",
        "per_frame": [
            {"index": int(i), "matrix": np.asarray(m, dtype=np.float64).tolist()} for i, m in enumerate(matrices)
        ],
    }
    if generator is not None:
        block["generator"] = dict(generator)
    validate_motion_meta(block)
    return block


def _warp_header(warp_meta: Dict[str, Any]):
    if not isinstance(warp_meta, dict):
        raise ValueError("stabilization_warp must be an object.")
    if warp_meta.get("matrix_convention") != "source_to_stabilized":
        raise ValueError(
            "stabilization_warp.matrix_convention must be 'source_to_stabilized', "
            f"got {warp_meta.get('matrix_convention')!r}."
        )
    src = _size_pair("stabilization_warp", warp_meta, "source_size")
    out = _size_pair("stabilization_warp", warp_meta, "output_size")
    entries = warp_meta.get("per_frame")
    if not isinstance(entries, list):
        raise ValueError("stabilization_warp.per_frame must be a list.")
    return src, out, entries


def motion_meta_from_stabilization_warp(warp_meta, fps: float, source: str):
    """Inverse (stabilized -> source) motion block, or None if a matrix is singular."""
    src, out, entries = _warp_header(warp_meta)
    inverses = []
    for i, entry in enumerate(entries):
        m = _matrix_entry("stabilization_warp", entry, i, "applied_matrix")
        try:
            inverses.append(np.linalg.inv(m))
        except np.linalg.LinAlgError:
            return None
    return build_motion_meta_v2(
        source=source, frame_count=len(inverses), fps=fps, input_size=out, output_size=src, matrices=inverses
    )


def applied_motion_meta_from_stabilization_warp(warp_meta, fps: float, source: str):
    """Forward (source -> stabilized) motion block: the matrices exactly as applied."""
    src, out, entries = _warp_header(warp_meta)
    applied = _matrix_entries("stabilization_warp", entries, "applied_matrix")
    return build_motion_meta_v2(
        source=source, frame_count=len(applied), fps=fps, input_size=src, output_size=out, matrices=applied
    )


def applied_motion_meta_from_matrices(matrices, *, source_size, output_size, fps: float, source: str, first_index: int = 0) -> Dict[str, Any]:
    """Same block as applied_motion_meta_from_stabilization_warp(build_stabilization_warp_meta(...)) for
    matrices that are already a float32 numpy stack: the checks (finite, invertible) run once on the
    stack instead of once per nested-list entry.  Raises the same ValueErrors."""
    stack = np.asarray(matrices, dtype=np.float32).reshape(-1, 3, 3).astype(np.float64)
    src = _size_pair("stabilization_warp", {"source_size": list(source_size)}, "source_size")
    out = _size_pair("stabilization_warp", {"output_size": list(output_size)}, "output_size")
    fps = float(fps)
    if not np.isfinite(fps) or fps <= 0.0:
        raise ValueError("motion_meta.fps must be a positive number.")
    if not isinstance(source, str) or not source:
        raise ValueError("motion_meta.source must be a non-empty string.")
    for i in np.flatnonzero(~np.isfinite(stack).all(axis=(1, 2))):
        raise ValueError(f"stabilization_warp.per_frame[{int(i) + first_index}].applied_matrix must contain finite numbers.")
    # np.linalg.inv fails on an exactly zero LU pivot; such a matrix has a determinant at rounding level.
    # Only matrices whose cofactor determinant is tiny against their scale go through the LAPACK check
    # (a per-matrix call costs microseconds, which adds up on long clips).
    a, b, c = stack[:, 0, 0], stack[:, 0, 1], stack[:, 0, 2]
    d, e, f = stack[:, 1, 0], stack[:, 1, 1], stack[:, 1, 2]
    g, h2, k = stack[:, 2, 0], stack[:, 2, 1], stack[:, 2, 2]
    det = a * (e * k - f * h2) - b * (d * k - f * g) + c * (d * h2 - e * g)
    scale = np.abs(stack).reshape(-1, 9).max(axis=1)
    for i in np.flatnonzero(~(np.abs(det) > 1e-9 * scale * scale * scale)):
        try:
            np.linalg.inv(stack[i])
        except np.linalg.LinAlgError as exc:
            raise ValueError(f"stabilization_warp.per_frame[{int(i) + first_index}].applied_matrix is not invertible.") from exc
    return {
        "version": 2,
        "source": source,
        "frame_count": int(stack.shape[0]),
        "fps": fps,
        "input_size": [int(src[0]), int(src[1])],
        "output_size": [int(out[0]), int(out[1])],
        "matrix_convention": "input_to_output",
        "per_frame": [{"index": i, "matrix": m} for i, m in enumerate(stack.tolist(), first_index)],
    }


def resolve_motion_meta(meta: Dict[str, Any]) -> MotionMeta:
    if not isinstance(meta, dict):
        raise ValueError("meta must be a dictionary containing motion_meta or stabilization_warp.")
    block = meta.get("motion_meta")
    if isinstance(block, dict):
        return _parse_block(block)
    warp_meta = meta.get("stabilization_warp")
    if isinstance(warp_meta, dict):
        inverse = motion_meta_from_stabilization_warp(warp_meta, fps=16.0, source="legacy_stabilization")
        if inverse is None:
            raise ValueError("stabilization_warp contains a non-invertible applied_matrix.")
        return _parse_block(inverse)
    raise ValueError("meta must contain motion_meta or stabilization_warp.")
