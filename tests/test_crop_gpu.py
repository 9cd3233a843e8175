"""`crop` framing (keep_fov search + no-padding refinement) on the CUDA path vs reference goldens."""
import json
import os

import numpy as np
import pytest
import torch

from tests import cases, parity
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", cases.CROP_CASES, ids=[c["name"] for c in cases.CROP_CASES])
def test_crop_matches_reference_golden(case):
    from vstab_b200 import classic, flow, pipeline

    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    frames = cases.make_frames(case)
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
    driver = flow if case["node"] == "flow" else classic
    res = driver.stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                  case["keep_fov"], case["padding_rgb"], case["fps"])
    tol = 2e-4 if case["node"] == "flow" else 5e-3
    parity.compare_nested(gmeta, json.loads(json.dumps(res.meta)), "meta", atol=tol, rtol=tol)
    for mine, ref in zip(res.meta["stabilization_warp"]["per_frame"], gmeta["stabilization_warp"]["per_frame"]):
        parity.assert_transform_close(mine["applied_matrix"], ref["applied_matrix"], f"frame {ref['index']}")
    assert tuple(res.frames.shape) == tuple(gold["shape"])
    assert float(res.masks.max()) == 0.0  # crop mode never pads (scripts/check_crop_aspect_ratio.py:100-103)
    f, y, x, hh, ww = gold["patch0_at"]
    pix_tol = 1e-3 if case["node"] == "flow" else 5e-3
    assert float(np.abs(res.frames[f, y:y + hh, x:x + ww] - gold["patch0"]).max()) <= pix_tol
    if "crop_size" in res.meta["framing"]:
        cw, ch = res.meta["framing"]["crop_size"]
        assert abs(cw / ch - case["w"] / case["h"]) <= 1e-6  # aspect ratio preserved


def test_aspect_rectangle_solver_matches_reference_rules():
    """largest_aspect_ratio_rectangle on a hand-made mask: centred crop preferred, exact aspect."""
    from vstab_b200.crop import largest_aspect_ratio_rectangle

    mask = np.ones((90, 160), dtype=np.uint8)
    mask[:7] = 0
    mask[:, 150:] = 0
    x0, y0, cw, ch = largest_aspect_ratio_rectangle(mask, 160, 90)
    assert abs(cw / ch - 160 / 90) < 1e-9 and y0 >= 7 and x0 + cw <= 150.0 + 1e-9
    assert largest_aspect_ratio_rectangle(np.zeros((9, 16), np.uint8), 16, 9) is None
