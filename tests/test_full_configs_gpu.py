"""The BASELINE.json configurations AS QUOTED -- full clip lengths -- on the CUDA path against summaries of the
unmodified reference's output (tests/golden/full_*, scripts/make_golden.py --only full):

  cfg2  Video Stabilizer Flow DIS, similarity, crop_and_pad, strength 0.7 smooth 0.5, 121 x 1920x1080
  cfg3  Video Stabilizer Classic (GFTT + LK), translation and similarity, 241 x 1280x720
  cfg4  Motion Apply bicubic, expand, motion_blur 0.5 at Ultra quality (33 samples), 121 x 1920x1080
  cfg5  Video Stabilizer Flow DIS, perspective, camera_lock at 3840x2160 (6 frames: the 2000-frame clip is 199 GB)

(cfg1, Shake Generator -> Motion Apply on 81 x 832x480, is tests/test_warp_gpu.py::test_apply_motion_matches_reference_golden.)
Per-frame float64 sums of frames and masks, three 48x64 patches of both, and the whole meta tree at the tolerance of
the reference's own A/B gate (scripts/compare_refactor_behavior.py:37-38).  bench.py additionally compares EVERY pixel
of the cfg2 clip with the reference run on the GPU box (`parity` block).
"""
import json
import os

import numpy as np
import pytest
import torch

from tests import cases, parity
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _gold(name):
    gold = np.load(os.path.join(GOLDEN_DIR, f"full_{name}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"full_{name}_meta.json")) as fh:
        return gold, json.load(fh)


def _run(case):
    from vstab_b200 import classic, flow, motion_apply, pipeline

    frames = cases.make_frames(case)
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
    del frames
    if case["kind"] == "apply":
        with open(os.path.join(GOLDEN_DIR, f"full_{case['name']}_motion_meta.json")) as fh:
            meta_in = json.load(fh)
        return motion_apply.apply_motion(ctx, meta_in, case["padding_rgb"], framing_mode=case["framing"], interpolation=case["interp"],
                                         motion_blur=case["blur"], motion_blur_samples=case["samples"])
    driver = flow if case["node"] == "flow" else classic
    return driver.stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                   case["keep_fov"], case["padding_rgb"], case["fps"])


# (meta tolerance, patch tolerance, relative tolerance of the per-frame sums); 0.0 = identical bits expected:
#   Flow similarity and Motion Apply reproduce cv2's arithmetic exactly (DIS, fit inputs, bilinear / bicubic, blur sum);
#   Classic and the perspective fit agree with cv2 to ~1e-9 in double, which can move a float32 matrix entry by one ulp.
TOL = {
    "cfg2_flow_1080p_121": (2e-5, 1e-6, 1e-9),
    "cfg3_classic_sim_720p_241": (2e-5, 1e-3, 1e-6),
    "cfg3_classic_trans_720p_241": (2e-5, 1e-3, 1e-6),
    "cfg4_apply_1080p_121": (2e-5, 0.0, 0.0),
    "cfg5_flow_persp_lock_4k_6": (2e-4, 1e-3, 1e-5),
}


@pytest.mark.parametrize("case", cases.FULL_CASES, ids=[c["name"] for c in cases.FULL_CASES])
def test_baseline_config_as_quoted(case):
    name = case["name"]
    if not os.path.exists(os.path.join(GOLDEN_DIR, f"full_{name}.npz")):
        pytest.skip("golden not generated")
    gold, gmeta = _gold(name)
    res = _run(case)
    meta_tol, patch_tol, sum_rtol = TOL[name]
    frames, masks = np.asarray(res.frames), np.asarray(res.masks)
    assert tuple(frames.shape) == tuple(gold["shape"])
    meta = json.loads(json.dumps(res.meta))
    if case["kind"] == "stab":
        assert meta["transform_mode_applied"] == gmeta["transform_mode_applied"] and meta["frames"] == case["n"]
        for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
            assert mine["mode"] == ref["mode"]
            parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}", scale_px=case["w"] / 1920.0 if case["w"] > 1920 else 1.0)
    parity.compare_nested(gmeta, meta, "meta", atol=meta_tol, rtol=meta_tol)
    fs = frames.reshape(frames.shape[0], -1).astype(np.float64).sum(axis=1)
    ms = masks.reshape(masks.shape[0], -1).astype(np.float64).sum(axis=1)
    if sum_rtol == 0.0:
        assert np.array_equal(fs, gold["frame_sum"]) and np.array_equal(ms, gold["mask_sum"])
    else:
        assert np.allclose(fs, gold["frame_sum"], rtol=sum_rtol, atol=0.0), float(np.abs(fs / gold["frame_sum"] - 1).max())
        assert float(np.abs(ms - gold["mask_sum"]).max()) <= (0.0 if patch_tol <= 1e-6 else 64.0)
    for k in range(3):
        f, y, x, hh, ww = gold[f"patch{k}_at"]
        err = float(np.abs(frames[f, y:y + hh, x:x + ww] - gold[f"patch{k}"]).max())
        assert err <= patch_tol, (k, err)
        merr = float(np.abs(masks[f, y:y + hh, x:x + ww, 0] - gold[f"mpatch{k}"]).max())
        assert merr == 0.0, (k, merr)
