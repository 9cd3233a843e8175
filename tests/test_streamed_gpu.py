"""Clips that do not fit in HBM are streamed through the device twice:
the results must be the ones of the resident path, bit for bit."""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu


def _flow(frames_t, case, monkeypatch, limit_mb, chunk_bytes=None):
    from vstab_b200 import flow, pipeline

    if limit_mb is None:
        monkeypatch.delenv("VSTAB_RESIDENT_LIMIT_MB", raising=False)
    else:
        monkeypatch.setenv("VSTAB_RESIDENT_LIMIT_MB", str(limit_mb))
    if chunk_bytes is not None:
        monkeypatch.setattr(pipeline, "CHUNK_BYTES", chunk_bytes)
    ctx = pipeline.normalize_video_input(frames_t)
    res = flow.stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"],
                                case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"])
    return ctx, res


@pytest.mark.parametrize("as_uint8", [False, True])
def test_streamed_flow_equals_resident(monkeypatch, as_uint8):
    case = [c for c in cases.STABILIZER_CASES if c["node"] == "flow"][0]
    frames = cases.make_frames(case)
    if as_uint8:
        frames_t = torch.from_numpy(np.clip(np.rint(frames * 255.0), 0, 255).astype(np.uint8))
    else:
        frames_t = torch.from_numpy(frames)
    ctx_r, want = _flow(frames_t, case, monkeypatch, None)
    assert not ctx_r.streamed
    # three frames per chunk: several uploads in both passes, chunk boundaries inside the clip
    per_frame = frames.shape[1] * frames.shape[2] * 3 * 4
    ctx_s, got = _flow(frames_t, case, monkeypatch, 0, chunk_bytes=3 * per_frame)
    assert ctx_s.streamed and ctx_s.frames is None
    assert np.array_equal(got.frames, want.frames)
    assert np.array_equal(got.masks, want.masks)
    assert got.meta == want.meta


def test_streamed_motion_apply_and_bypass(monkeypatch):
    from vstab_b200 import flow, motion_apply, pipeline

    case = [c for c in cases.STABILIZER_CASES if c["node"] == "flow"][0]
    frames = cases.make_frames(case)
    frames_t = torch.from_numpy(frames)
    _, stab = _flow(frames_t, case, monkeypatch, None)
    meta = stab.meta
    want = motion_apply.apply_motion(pipeline.normalize_video_input(frames_t), meta, (10, 20, 30), framing_mode="crop_and_pad",
                                     interpolation="bicubic", motion_blur=0.5, motion_blur_samples=5)
    per_frame = frames.shape[1] * frames.shape[2] * 3 * 4
    monkeypatch.setenv("VSTAB_RESIDENT_LIMIT_MB", "0")
    monkeypatch.setattr(pipeline, "CHUNK_BYTES", 2 * per_frame)
    ctx = pipeline.normalize_video_input(frames_t)
    assert ctx.streamed
    got = motion_apply.apply_motion(ctx, meta, (10, 20, 30), framing_mode="crop_and_pad", interpolation="bicubic", motion_blur=0.5,
                                    motion_blur_samples=5)
    assert np.array_equal(got.frames, want.frames)
    assert np.array_equal(got.masks, want.masks)
    # keep_fov bypass hands the (normalised) input back
    res = flow.stabilize_frames(ctx, "crop", case["mode"], False, 0.7, 0.5, 1.0, (127, 127, 127), 16.0)
    assert np.array_equal(res.frames, frames)
    assert float(np.abs(res.masks).max()) == 0.0
    with pytest.raises(Exception):
        flow.stabilize_frames(ctx, "crop_and_pad", case["mode"], False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device")
