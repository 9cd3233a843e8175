"""Video Stabilizer Flow on the CUDA path vs outputs of the real reference (tests/golden/stab_flow_*)."""
import json
import os

import numpy as np
import pytest
import torch

from tests import cases, parity
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

FLOW_CASES = [c for c in cases.STABILIZER_CASES if c["node"] == "flow"]


def _run(case):
    from vstab_b200 import flow, pipeline

    frames = cases.make_frames(case)
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
    return flow.stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"],
                                 case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"])


@pytest.mark.parametrize("case", FLOW_CASES, ids=[c["name"] for c in FLOW_CASES])
def test_flow_matches_reference_golden(case):
    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    res = _run(case)
    meta = res.meta
    # per-pair estimated transforms: north_star tolerance (0.05 px / 0.01 deg / 1e-4), full-res units
    worst = [0.0, 0.0, 0.0]
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
        assert mine["mode"] == ref["mode"]
        d = parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}")
        worst = [max(a, b) for a, b in zip(worst, d)]
        assert abs(mine["confidence"] - ref["confidence"]) <= 1e-3
        assert abs(mine["residual"] - ref["residual"]) <= 1e-3
    # the CUDA DIS is bit-exact, so in practice the transforms agree to float32 rounding
    assert worst[0] <= 1e-4 and worst[1] <= 1e-5 and worst[2] <= 1e-6, worst
    # applied matrices and the whole meta tree (same key set, 2e-5 like the reference's own A/B script)
    for mine, ref in zip(meta["stabilization_warp"]["per_frame"], gmeta["stabilization_warp"]["per_frame"]):
        parity.assert_transform_close(mine["applied_matrix"], ref["applied_matrix"], f"frame {ref['index']}")
    parity.compare_nested(gmeta, json.loads(json.dumps(meta)), "meta", atol=2e-4, rtol=2e-4)
    # pixels + mask
    assert tuple(res.frames.shape) == tuple(gold["shape"])
    f, y, x, hh, ww = gold["patch0_at"]
    err = float(np.abs(res.frames[f, y:y + hh, x:x + ww] - gold["patch0"]).max())
    assert err <= parity.TOL_PIXEL["bilinear"], err
    assert float(np.abs(res.masks[f, y:y + hh, x:x + ww, 0] - gold["mpatch0"]).max()) == 0.0
    fs = res.frames.reshape(res.frames.shape[0], -1).astype(np.float64).sum(axis=1)
    assert np.allclose(fs, gold["frame_sum"], rtol=1e-5)
    ms = res.masks.reshape(res.masks.shape[0], -1).astype(np.float64).sum(axis=1)
    assert np.abs(ms - gold["mask_sum"]).max() <= 64  # identical matrices => identical masks; allow rounding-level edges


def test_flow_single_frame_bypass():
    from vstab_b200 import flow, pipeline

    frames = np.random.default_rng(0).random((1, 96, 160, 3), dtype=np.float32)
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
    res = flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0)
    assert np.array_equal(res.frames, frames) and float(res.masks.max()) == 0.0
    assert res.meta["note"].startswith("Single-frame") and "motion_meta" in res.meta
