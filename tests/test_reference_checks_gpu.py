"""The reference's own script-level checks, re-run against the CUDA path:
scripts/check_motion_meta.py:289-415 (identity apply, blur==0 baseline path, expand canvas, blur
determinism, progress tick counts, crop -> crop_and_pad fallback),
scripts/check_crop_aspect_ratio.py:123-161 (Motion Apply replay of a stabilizer's meta must equal the
stabilizer's own frames and masks) and scripts/compare_refactor_behavior.py:289-324 (input adapters)."""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu


def _frames(n, w=32, h=24, seed=0):
    return np.random.default_rng(seed).random((n, h, w, 3), dtype=np.float32)


def _meta(mats, w=32, h=24):
    from vstab_b200.motion_meta import build_motion_meta_v2

    return {"motion_meta": build_motion_meta_v2(source="generated_shake", frame_count=len(mats), fps=16.0, input_size=(w, h),
                                                output_size=(w, h), matrices=mats, generator={"node": "test"})}


def _ctx(frames):
    from vstab_b200 import pipeline

    return pipeline.normalize_video_input(torch.from_numpy(np.ascontiguousarray(frames)))


def test_identity_blur_zero_expand_determinism_and_ticks():
    from vstab_b200.motion_apply import apply_motion

    frames = _frames(3)
    eye = [np.eye(3) for _ in range(3)]
    base = apply_motion(_ctx(frames), _meta(eye), (127, 127, 127))
    ident = apply_motion(_ctx(frames), _meta(eye), (127, 127, 127), motion_blur=0.0)
    assert np.allclose(ident.frames, frames, atol=1e-6) and float(ident.masks.max()) == 0.0
    assert np.array_equal(ident.frames, base.frames) and np.array_equal(ident.masks, base.masks)

    shift = np.array([[1.0, 0.0, 6.0], [0.0, 1.0, -4.0], [0.0, 0.0, 1.0]])
    exp = apply_motion(_ctx(frames), _meta([np.eye(3), shift, np.linalg.inv(shift)]), (127, 127, 127), framing_mode="expand")
    assert exp.frames.shape[1] > 24 and exp.frames.shape[2] > 32 and exp.meta["motion_apply"]["framing_mode"] == "expand"

    blur_m = np.array([[1.0, 0.0, 2.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    bm = _meta([np.eye(3), blur_m, blur_m @ blur_m])
    a = apply_motion(_ctx(frames), bm, (127, 127, 127), motion_blur=0.5, motion_blur_samples=7)
    b = apply_motion(_ctx(frames), bm, (127, 127, 127), motion_blur=0.5, motion_blur_samples=7)
    assert np.array_equal(a.frames, b.frames) and np.array_equal(a.masks, b.masks)
    assert a.meta["motion_apply"]["motion_blur"] == 0.5 and a.meta["motion_apply"]["motion_blur_samples"] == 7

    ticks = [0]

    def tick():
        ticks[0] += 1

    apply_motion(_ctx(frames), bm, (127, 127, 127), motion_blur=0.5, motion_blur_samples=7, progress_callback=tick)
    assert ticks[0] == 3 * 7
    ticks[0] = 0
    apply_motion(_ctx(frames), bm, (127, 127, 127), framing_mode="crop", motion_blur=0.5, motion_blur_samples=7, progress_callback=tick)
    assert ticks[0] == 3 + 3 * 7

    far = np.array([[1.0, 0.0, 60.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    fb = apply_motion(_ctx(frames), _meta([np.eye(3), far, np.eye(3)]), (127, 127, 127), framing_mode="crop")
    assert fb.meta.get("framing_fallback") == "crop_and_pad" and fb.meta["motion_apply"]["framing_mode"] == "crop_and_pad"


def test_errors_match_the_reference_messages():
    from vstab_b200.motion_apply import apply_motion

    frames = _frames(3)
    with pytest.raises(ValueError, match="Frame count mismatch"):
        apply_motion(_ctx(frames), _meta([np.eye(3)] * 2), (0, 0, 0))
    with pytest.raises(ValueError, match="Input frames must match motion_meta.input_size"):
        apply_motion(_ctx(frames), _meta([np.eye(3)] * 3, w=40), (0, 0, 0))
    with pytest.raises(ValueError, match="Unsupported interpolation"):
        apply_motion(_ctx(frames), _meta([np.eye(3)] * 3), (0, 0, 0), interpolation="lanczos")
    with pytest.raises(ValueError, match="Unsupported framing_mode"):
        apply_motion(_ctx(frames), _meta([np.eye(3)] * 3), (0, 0, 0), framing_mode="stretch")


@pytest.mark.parametrize("framing,mode", [("expand", "similarity"), ("crop_and_pad", "translation")])
def test_motion_apply_replays_the_stabilizer_exactly(framing, mode):
    """Replaying a stabilizer's emitted meta through Motion Apply must reproduce its frames and masks bit for bit."""
    from vstab_b200 import flow
    from vstab_b200.motion_apply import apply_motion

    frames = cases.make_frames(dict(n=6, w=416, h=240, seed=61, frames="texture"))
    direct = flow.stabilize_frames(_ctx(frames), framing, mode, False, 1.0, 0.5, 0.6, (127, 127, 127), 24.0)
    replay = apply_motion(_ctx(frames), direct.meta, (127, 127, 127), framing_mode="crop_and_pad", interpolation="bilinear")
    assert replay.frames.shape == direct.frames.shape
    assert np.array_equal(replay.frames, direct.frames) and np.array_equal(replay.masks, direct.masks)


def test_input_adapters_agree():
    """list / batch tensor / dict / uint8 / 0..255 float / 1-channel inputs all reach the same clip in HBM."""
    from vstab_b200 import pipeline

    frames = _frames(4, 48, 32, seed=3)
    want = pipeline.normalize_video_input(torch.from_numpy(frames)).frames.cpu().numpy()
    as_list = pipeline.normalize_video_input([f for f in frames]).frames.cpu().numpy()
    as_dict = pipeline.normalize_video_input({"frames": torch.from_numpy(frames), "fps": 24.0})
    assert np.array_equal(want, as_list) and np.array_equal(want, as_dict.frames.cpu().numpy()) and as_dict.fps == 24.0
    scaled = pipeline.normalize_video_input(torch.from_numpy(frames * np.float32(255.0))).frames.cpu().numpy()
    assert np.allclose(scaled, want, atol=1e-6)
    u8 = (frames * 255).astype(np.uint8)
    got_u8 = pipeline.normalize_video_input(torch.from_numpy(u8)).frames.cpu().numpy()
    assert np.array_equal(got_u8, u8.astype(np.float32) / np.float32(255.0))
    chw = pipeline.normalize_video_input([np.moveaxis(f, -1, 0) for f in frames]).frames.cpu().numpy()
    assert np.array_equal(chw, want)
    gray = pipeline.normalize_video_input(torch.from_numpy(frames[..., :1].copy())).frames.cpu().numpy()
    assert gray.shape[-1] == 3 and np.array_equal(gray[..., 0], gray[..., 2])
    with pytest.raises(ValueError, match="empty"):
        pipeline.normalize_video_input([])


def _check_script_frames(width=121, height=73, count=6):
    """scripts/check_crop_aspect_ratio.py:58-79 make_synthetic_frames (cv2 only draws the test card)."""
    cv2 = pytest.importorskip("cv2")
    yy, xx = np.mgrid[0:height, 0:width]
    base = np.zeros((height, width, 3), dtype=np.float32)
    base[..., 0] = xx / max(width - 1, 1)
    base[..., 1] = yy / max(height - 1, 1)
    base[..., 2] = ((xx // 9 + yy // 7) % 2).astype(np.float32)
    cv2.rectangle(base, (12, 10), (44, 34), (1.0, 0.2, 0.1), -1)
    cv2.circle(base, (width - 22, height - 16), 9, (0.1, 0.9, 0.3), -1)
    out = []
    for index in range(count):
        matrix = np.array([[1.0, 0.0, index * 1.1], [0.0, 1.0, -index * 0.35]], dtype=np.float32)
        out.append(cv2.warpAffine(base, matrix, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT))
    return np.stack(out)


@pytest.mark.parametrize("node", ["classic", "flow"])
@pytest.mark.parametrize("mode", ["translation", "similarity"])
@pytest.mark.parametrize("keep_fov", [0.0, 0.6])
def test_check_crop_aspect_ratio_grid(node, mode, keep_fov):
    """scripts/check_crop_aspect_ratio.py:82-120 + :236-241, the reference's own grid on its own 121x73 frames:
    crop never pads, the crop keeps the frame's aspect ratio, the applied scale is uniform."""
    from vstab_b200 import classic, flow

    frames = _check_script_frames()
    driver = flow if node == "flow" else classic
    res = driver.stabilize_frames(_ctx(frames), "crop", mode, False, 1.0, 0.5, keep_fov, (127, 127, 127), 24.0)
    assert len(res.frames) == len(frames) and len(res.masks) == len(frames)
    for mask in res.masks:
        assert tuple(mask.shape) == (73, 121, 1) and float(np.max(mask)) == 0.0
    cw, ch = res.meta["framing"]["crop_size"]
    assert abs(cw / ch - 121 / 73) <= 1e-6
    for entry in res.meta["stabilization_warp"]["per_frame"]:
        m = np.asarray(entry["applied_matrix"], dtype=np.float64)
        assert abs(m[0, 0] - m[1, 1]) <= 1e-6


@pytest.mark.parametrize("node", ["classic", "flow"])
@pytest.mark.parametrize("framing", ["expand", "crop_and_pad"])
def test_check_script_replay(node, framing):
    """scripts/check_crop_aspect_ratio.py:123-161 + :242-243 on the same frames: Motion Apply replays the
    stabilizer's meta to the stabilizer's own frames and masks, bit for bit."""
    from vstab_b200 import classic, flow
    from vstab_b200.motion_apply import apply_motion

    frames = _check_script_frames()
    driver = flow if node == "flow" else classic
    direct = driver.stabilize_frames(_ctx(frames), framing, "similarity", False, 1.0, 0.5, 0.6, (127, 127, 127), 24.0)
    replay = apply_motion(_ctx(frames), direct.meta, (127, 127, 127), framing_mode="crop_and_pad", interpolation="bilinear")
    assert replay.frames.shape == direct.frames.shape
    assert np.array_equal(replay.frames, direct.frames) and np.array_equal(replay.masks, direct.masks)


@pytest.mark.parametrize("node", ["classic", "flow"])
@pytest.mark.parametrize("scenario", cases.AB_SCENARIOS, ids=[s[0] for s in cases.AB_SCENARIOS])
def test_reference_ab_gate(node, scenario):
    """scripts/compare_refactor_behavior.py:352-377, the reference's own gate between two revisions of itself, with
    the unmodified reference as "base" (tests/golden/ab_73x45*, scripts/make_golden.py --only ab) and the CUDA path
    as "head", on the script's own 8 x 73x45 clip.  The script's tolerance is 2e-5 on pixels, masks and meta.
    Flow meets it (DIS is bit-exact; what remains is float32 rounding of the fitted matrices, which can move a
    pixel across one of cv2's 1/32-px quantisation steps: at most 0.1 % of the pixels may differ, by one step on
    an edge).  Classic: corners and tracks carry cv2's bits since round 2 (window sums in cv2's lane order), so it is held
    to the same gate."""
    import json
    import os

    from tests import parity
    from tests.conftest import GOLDEN_DIR
    from vstab_b200 import classic, flow

    name, framing, mode, keep_fov = scenario
    gold = np.load(os.path.join(GOLDEN_DIR, "ab_73x45.npz"))
    with open(os.path.join(GOLDEN_DIR, "ab_73x45_meta.json")) as fh:
        gmeta = json.load(fh)[f"{node}.{name}"]
    a = cases.AB_ARGS
    driver = flow if node == "flow" else classic
    res = driver.stabilize_frames(_ctx(gold["input"]), framing, mode, a["camera_lock"], a["strength"], a["smooth"], keep_fov,
                                  a["padding_rgb"], a["fps"])
    want_f, want_m = gold[f"{node}.{name}.frames"], gold[f"{node}.{name}.masks"]
    got_f, got_m = np.asarray(res.frames), np.asarray(res.masks)
    assert got_f.shape == want_f.shape and got_m.shape == want_m.shape
    assert 0.0 <= float(got_m.min()) <= float(got_m.max()) <= 1.0
    meta = json.loads(json.dumps(res.meta))
    assert meta["transform_mode_applied"] == gmeta["transform_mode_applied"]
    err = np.abs(got_f - want_f)
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
        assert mine["mode"] == ref["mode"]
        parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}")
    parity.compare_nested(gmeta, meta, "meta", atol=2e-5, rtol=2e-5)
    assert float((err > 2e-5).mean()) <= 1e-3 and float(err.max()) <= 0.04, (float((err > 2e-5).mean()), float(err.max()))
    assert float((got_m != want_m).mean()) <= 1e-3


@pytest.mark.parametrize("name", ["expand", "crop"])
def test_inverse_stabilization_on_the_check_script_scenario(name):
    """scripts/check_inverse_stabilization.py:134-181 on the CUDA path; golden = the reference's own helper
    (tests/golden/inverse_73x45*, scripts/make_golden.py --only inverse)."""
    import json
    import os

    from tests.conftest import GOLDEN_DIR
    from vstab_b200 import inverse

    gold = np.load(os.path.join(GOLDEN_DIR, "inverse_73x45.npz"))
    with open(os.path.join(GOLDEN_DIR, "inverse_73x45_meta.json")) as fh:
        metas = json.load(fh)
    res = inverse.apply_inverse_stabilization(_ctx(gold[f"{name}.input"]), metas[f"{name}.in"], (127, 127, 127))
    assert res.frames.shape == gold[f"{name}.frames"].shape and res.frames.dtype == np.float32 and res.masks.dtype == np.float32
    assert float(np.abs(res.frames - gold[f"{name}.frames"]).max()) <= 1e-6  # bilinear: same bits as cv2 in practice
    assert np.array_equal(res.masks, gold[f"{name}.masks"])
    inv = res.meta["inverse_stabilization"]
    assert inv == metas[f"{name}.helper_out"]["inverse_stabilization"] and inv["matrix_convention"] == "stabilized_to_source"
    if name == "expand":
        err = np.abs(res.frames - gold["source"])
        assert float(np.quantile(err, 0.99)) <= 0.3 and float(err.mean()) <= 0.035
    else:
        assert float(res.masks.max()) > 0.0
