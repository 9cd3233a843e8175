"""vstab-b200: the data-parallel hot path of ComfyUI-Video-Stabilizer on B200 (sm_100a).

The directory name carries a hyphen because it is meant to be dropped into ComfyUI's
``custom_nodes/`` (ComfyUI imports custom nodes by path).  From plain Python use
``vstab_loader.load()`` at the repository root, which registers this package as ``vstab_b200``.

Public surface (same names / argument meaning as the reference's node layer):
  nodes.VideoStabilizerFlow / VideoStabilizerClassic / VideoStabilizerMotionApply   (ComfyUI V3)
  flow.stabilize_frames, classic.stabilize_frames, motion_apply.apply_motion         (drivers)
  pipeline.normalize_video_input / reconstruct_video / convert_masks_for_output      (adapters)
  _native.Handle                                                                     (C ABI)
"""
from __future__ import annotations

__version__ = "0.1.0"

from . import _native, hostmath, motion_meta  # noqa: F401  (no GPU needed to import)


async def comfy_entrypoint():
    """ComfyUI V3 extension entry point (reference: __init__.py:37-39)."""
    from .nodes import VideoStabilizerB200Extension

    return VideoStabilizerB200Extension()
