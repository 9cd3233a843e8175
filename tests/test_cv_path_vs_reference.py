"""oracle/cv_path.py (restated drivers over real cv2) vs the UNMODIFIED reference (build container)."""
import numpy as np
import pytest

from oracle import cv_path
from tests import cases

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("node,mode,framing,lock", [
    ("flow", "similarity", "crop_and_pad", False), ("flow", "translation", "expand", False),
    ("flow", "perspective", "crop_and_pad", True), ("classic", "similarity", "crop_and_pad", False),
    ("classic", "translation", "expand", False),
])
def test_stabilize_equals_reference(reference_nodes, node, mode, framing, lock):
    case = dict(n=6, w=832, h=480, seed=41, frames="texture", perspective=(mode == "perspective"))
    frames = cases.make_frames(case)
    mod = reference_nodes.video_stabilizer_flow if node == "flow" else reference_nodes.video_stabilizer_classic
    ctx = reference_nodes.stabilizer_utils._normalize_video_input([f for f in frames])
    ref = mod._stabilize_frames(ctx, framing, mode, lock, 0.7, 0.5, 0.6, (127, 127, 127), 16.0)
    out_f, out_m, meta = cv_path.stabilize([f for f in frames], node, framing, mode, lock, 0.7, 0.5, 0.6, (127, 127, 127), 16.0)
    assert np.array_equal(out_f, np.asarray(ref.frames)) and np.array_equal(out_m, np.asarray(ref.masks))
    assert meta["transform_mode_applied"] == ref.meta["transform_mode_applied"]
    assert np.array_equal(meta["path"], np.asarray(ref.meta["estimated_motion"]["path"]))
    assert meta["padding_fraction_mean"] == ref.meta["padding_fraction_mean"]


def test_apply_motion_equals_reference(reference_nodes):
    case = dict(n=5, w=160, h=96, seed=3, meta="shake", amount=2.0)
    frames = np.random.default_rng(3).random((5, 96, 160, 3), dtype=np.float32)
    meta = cases.make_motion_meta(case, reference_nodes)
    mats = [np.asarray(e["matrix"]) for e in meta["motion_meta"]["per_frame"]]
    for framing, interp, blur, s in [("crop_and_pad", "bilinear", 0.0, 9), ("expand", "bicubic", 0.5, 5)]:
        ctx = reference_nodes.stabilizer_utils._normalize_video_input([f for f in frames])
        ref = reference_nodes.motion_apply.apply_motion(ctx, meta, (127, 127, 127), framing_mode=framing, interpolation=interp,
                                                       motion_blur=blur, motion_blur_samples=s)
        f, m, _ = cv_path.apply_motion([x for x in frames], mats, (160, 96), (160, 96), (127, 127, 127), framing, interp, blur, s)
        assert np.array_equal(f, ref.frames) and np.array_equal(m, ref.masks)
