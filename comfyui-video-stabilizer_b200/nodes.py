"""ComfyUI V3 node layer: the drop-in boundary.

Same node ids, display names, socket names / order / defaults / ranges and outputs as the
reference's node classes (nodes/video_stabilizer_flow.py:643-763,
nodes/video_stabilizer_classic.py:570-691, nodes/video_stabilizer_motion_apply.py:29-129; pinned by
the reference's scripts/check_node_schema.py:11-94), so saved workflows keep loading.  ``execute``
adapts the IMAGE tensor into HBM and calls the CUDA drivers; outputs are CPU tensors like the
reference's (IMAGE float32 [N,H',W',3], MASK float32 [N,H',W'], JSON dict).
"""
from __future__ import annotations

from typing import Any

from comfy_api.latest import ComfyExtension, io
from comfy.utils import ProgressBar

try:  # interrupt polling is optional exactly like in the reference (flow.py:24-27)
    import comfy.model_management as model_management
except ImportError:  # pragma: no cover
    model_management = None

from . import classic, flow, motion_apply
from .motion_meta import resolve_motion_meta
from .hostmath import parse_padding_color
from .pipeline import convert_masks_for_output, normalize_video_input, reconstruct_video

JSONType = io.Custom("JSON")
BLUR_QUALITY_SAMPLES = {"Draft": 5, "Standard": 9, "High": 17, "Ultra": 33}

_TOOLTIPS = {
    "flow": {
        "framing_mode": "Choose how borders produced by stabilization are handled.",
        "transform_mode": "Select the geometric model fitted to the optical flow.",
        "camera_lock": "Aggressively pull the motion curve toward a locked tripod-like solution.",
        "strength": "Removal gain (0 keeps original motion, 1 removes it using the smoothed motion curve).",
        "smooth": "Temporal smoothing amount applied to the motion curve before removal.",
    },
    "classic": {
        "framing_mode": "Choose how to handle borders produced by stabilization.",
        "transform_mode": "Select the geometric model used to estimate camera motion.",
        "camera_lock": "Treat the shot as tripod-like by aggressively damping motion.",
        "strength": "Removal gain (0 keeps original motion, 1 removes it based on smoothing).",
        "smooth": "Temporal smoothing amount applied to the estimated motion path.",
    },
}


def _check_interrupt() -> None:
    if model_management is not None:
        model_management.throw_exception_if_processing_interrupted()


def _stabilizer_inputs(kind: str):
    tip = _TOOLTIPS[kind]
    return [
        io.Image.Input("frames", display_name="Frames"),
        io.Float.Input("frame_rate", default=16.0, min=1.0, step=0.1, display_name="Input FPS",
                       tooltip="Frame rate in frames per second used to scale smoothing window."),
        io.Combo.Input("framing_mode", options=["crop", "crop_and_pad", "expand"], default="crop_and_pad",
                       display_name="Framing Mode", tooltip=tip["framing_mode"]),
        io.Combo.Input("transform_mode", options=["translation", "similarity", "perspective"], default="similarity",
                       display_name="Transform Mode", tooltip=tip["transform_mode"]),
        io.Boolean.Input("camera_lock", default=False, display_name="Camera Lock", tooltip=tip["camera_lock"]),
        io.Float.Input("strength", default=0.7, min=0.0, max=1.0, step=0.05, display_name="Strength",
                       tooltip=tip["strength"], display_mode=io.NumberDisplay.slider),
        io.Float.Input("smooth", default=0.5, min=0.0, max=1.0, step=0.05, display_name="Smooth",
                       tooltip=tip["smooth"], display_mode=io.NumberDisplay.slider),
        io.Float.Input("keep_fov", default=0.6, min=0.0, max=1.0, step=0.05, display_name="Keep FOV",
                       tooltip=("[Crop only] How much of the original FOV to preserve (1.0 = no zoom, 0.0 = maximum zoom). "
                                "Ignored when framing_mode is crop_and_pad or expand."),
                       display_mode=io.NumberDisplay.slider),
        io.Color.Input("padding_color", default="#7F7F7F", display_name="Padding Color",
                       tooltip="HEX padding color applied in crop_and_pad / expand (e.g. #404040)."),
    ]


def _stabilizer_outputs():
    return [
        io.Image.Output("frames_stabilized", display_name="Stabilized Frames"),
        io.Mask.Output("padding_mask", display_name="Padding Mask"),
        JSONType.Output("meta", display_name="Motion Meta"),
    ]


def _run_stabilizer(driver, frames, frame_rate, framing_mode, transform_mode, camera_lock, strength, smooth, keep_fov,
                    padding_color):
    context = normalize_video_input(frames, defer_range=True)  # the range rule rides on the luma kernel's read
    total = len(context)
    bar = ProgressBar(max(0, total - 1) + total)
    result = driver.stabilize_frames(
        context, framing_mode, transform_mode, camera_lock, strength, smooth, keep_fov,
        parse_padding_color(padding_color), frame_rate, progress_bar=bar, interrupt_check=_check_interrupt,
    )
    return io.NodeOutput(reconstruct_video(result.frames, context), convert_masks_for_output(result.masks), result.meta)


class VideoStabilizerFlow(io.ComfyNode):
    """Dense optical flow (DIS) stabilizer on the CUDA hot path."""

    @classmethod
    def define_schema(cls) -> io.Schema:
        schema = io.Schema(
            node_id="video_stabilizer_flow",
            display_name="Video Stabilizer Flow",
            category="Video/Stabilization",
            description=("Video stabilization using dense optical flow with configurable transforms and framing, "
                         "emitting stabilized frames, a padding mask, and motion diagnostics (B200 CUDA path)."),
        )
        schema.inputs = _stabilizer_inputs("flow")
        schema.outputs = _stabilizer_outputs()
        return schema

    @classmethod
    def execute(cls, frames: Any, frame_rate: float, framing_mode: str, transform_mode: str, camera_lock: bool,
                strength: float, smooth: float, keep_fov: float, padding_color: str) -> io.NodeOutput:
        return _run_stabilizer(flow, frames, frame_rate, framing_mode, transform_mode, camera_lock, strength, smooth,
                               keep_fov, padding_color)


class VideoStabilizerClassic(io.ComfyNode):
    """Sparse feature tracking (GFTT + pyramidal LK) stabilizer on the CUDA hot path."""

    @classmethod
    def define_schema(cls) -> io.Schema:
        schema = io.Schema(
            node_id="video_stabilizer_classic",
            display_name="Video Stabilizer Classic",
            category="Video/Stabilization",
            description=("Video stabilization with classic feature tracking, configurable transforms and framing, "
                         "emitting stabilized frames, a padding mask, and motion diagnostics (B200 CUDA path)."),
        )
        schema.inputs = _stabilizer_inputs("classic")
        schema.outputs = _stabilizer_outputs()
        return schema

    @classmethod
    def execute(cls, frames: Any, frame_rate: float, framing_mode: str, transform_mode: str, camera_lock: bool,
                strength: float, smooth: float, keep_fov: float, padding_color: str) -> io.NodeOutput:
        return _run_stabilizer(classic, frames, frame_rate, framing_mode, transform_mode, camera_lock, strength, smooth,
                               keep_fov, padding_color)


class VideoStabilizerMotionApply(io.ComfyNode):
    """Apply motion_meta matrices to a video sequence with the fused CUDA resampler."""

    @classmethod
    def define_schema(cls) -> io.Schema:
        schema = io.Schema(
            node_id="video_stabilizer_motion_apply",
            display_name="Video Stabilizer Motion Apply",
            category="Video/Stabilization",
            description="Applies motion metadata to frames and emits a padding mask.",
        )
        schema.inputs = [
            io.Image.Input("frames", display_name="Frames"),
            JSONType.Input("motion_meta", display_name="Motion Meta"),
            io.Combo.Input("framing_mode", options=["crop_and_pad", "crop", "expand"], default="crop_and_pad",
                           display_name="Framing Mode"),
            io.Combo.Input("interpolation", options=["bilinear", "bicubic"], default="bilinear",
                           display_name="Interpolation"),
            io.Color.Input("padding_color", default="#7F7F7F", display_name="Padding Color",
                           tooltip="HEX padding color used where warping exposes empty pixels."),
            io.Float.Input("motion_blur", default=0.0, min=0.0, max=1.0, step=0.05, display_name="Motion Blur",
                           tooltip="Shutter fraction for matrix-sampled motion blur. 0 disables blur.",
                           display_mode=io.NumberDisplay.slider),
            io.Combo.Input("motion_blur_quality", options=list(BLUR_QUALITY_SAMPLES.keys()), default="Standard",
                           display_name="Blur Quality",
                           tooltip="Draft is faster. High and Ultra average more shutter samples for smoother blur."),
        ]
        schema.outputs = [
            io.Image.Output("frames", display_name="Frames"),
            io.Mask.Output("padding_mask", display_name="Padding Mask"),
            JSONType.Output("meta", display_name="Meta"),
        ]
        return schema

    @classmethod
    def execute(cls, frames: Any, motion_meta: dict, framing_mode: str, interpolation: str, padding_color: str,
                motion_blur: float, motion_blur_quality: str) -> io.NodeOutput:
        context = normalize_video_input(frames)
        quality = motion_blur_quality if motion_blur_quality in BLUR_QUALITY_SAMPLES else "Standard"
        samples = BLUR_QUALITY_SAMPLES[quality]
        n = len(context)
        per_frame = int(max(3, min(33, samples))) if motion_blur > 0.0 else 1
        total = max(n * per_frame + (n if framing_mode == "crop" else 0), 1)
        bar = ProgressBar(total)
        done = [0]

        def tick() -> None:
            done[0] += 1
            bar.update_absolute(min(done[0], total), total)

        result = motion_apply.apply_motion(
            context, motion_meta, parse_padding_color(padding_color), framing_mode=framing_mode,
            interpolation=interpolation, motion_blur=motion_blur, motion_blur_samples=samples, progress_callback=tick,
        )
        result.meta.setdefault("motion_apply", {})["motion_blur_quality"] = quality
        bar.update_absolute(total, total)
        return io.NodeOutput(reconstruct_video(result.frames, context), convert_masks_for_output(result.masks), result.meta)


class VideoStabilizerInverse(io.ComfyNode):
    """Deprecated in the reference in favour of Motion Apply, kept for saved workflows
    (nodes/video_stabilizer_inverse.py:26-100): restores edited stabilized frames to the source canvas."""

    @classmethod
    def define_schema(cls) -> io.Schema:
        schema = io.Schema(
            node_id="video_stabilizer_inverse",
            display_name="Video Stabilizer Inverse",
            category="Video/Stabilization",
            description=("Deprecated: use Video Stabilizer Motion Apply. Restores stabilized frames to the "
                         "original canvas using stabilization metadata, and emits a padding mask for areas "
                         "without source pixels."),
            is_deprecated=True,
        )
        schema.inputs = [
            io.Image.Input("frames", display_name="Frames"),
            JSONType.Input("meta", display_name="Meta"),
            io.Color.Input("padding_color", default="#7F7F7F", display_name="Padding Color",
                           tooltip="HEX padding color used where inverse warping exposes empty pixels."),
        ]
        schema.outputs = [
            io.Image.Output("frames_restored", display_name="Restored Frames"),
            io.Mask.Output("padding_mask", display_name="Padding Mask"),
            JSONType.Output("meta", display_name="Meta"),
        ]
        return schema

    @classmethod
    def execute(cls, frames: Any, meta: dict, padding_color: str) -> io.NodeOutput:
        context = normalize_video_input(frames)
        inverse_meta = dict(meta)
        inverse_meta.pop("motion_meta", None)  # force the legacy path: the inverted stabilization_warp
        motion = resolve_motion_meta(inverse_meta)
        result = motion_apply.apply_motion(context, inverse_meta, parse_padding_color(padding_color),
                                           framing_mode="crop_and_pad", interpolation="bilinear")
        if isinstance(meta, dict) and isinstance(meta.get("motion_meta"), dict):
            result.meta["motion_meta"] = meta["motion_meta"]
        result.meta.pop("motion_apply", None)
        result.meta["inverse_stabilization"] = {
            "source_size": [int(motion.output_size[0]), int(motion.output_size[1])],
            "input_size": [int(motion.input_size[0]), int(motion.input_size[1])],
            "output_size": [int(motion.output_size[0]), int(motion.output_size[1])],
            "matrix_convention": "stabilized_to_source",
            "source_matrix_convention": "source_to_stabilized",
            "framing_mode": meta.get("stabilization_warp", {}).get("framing_mode") if isinstance(meta, dict) else None,
            "note": "Restores original motion/canvas; pixels discarded by crop framing cannot be recovered.",
        }
        return io.NodeOutput(reconstruct_video(result.frames, context), convert_masks_for_output(result.masks), result.meta)


class VideoStabilizerB200Extension(ComfyExtension):
    async def get_node_list(self) -> list[type[io.ComfyNode]]:
        return [VideoStabilizerClassic, VideoStabilizerFlow, VideoStabilizerMotionApply, VideoStabilizerInverse]
