"""Frames below ~91 px: the sizes the reference's own script-level checks run the stabilizers on
(scripts/compare_refactor_behavior.py:220 -> 73x45, scripts/check_crop_aspect_ratio.py:165 -> 121x73).
cv2's DIS selects its pyramid levels itself there, computes down to full resolution and leaves the new
finest scale on the backend object, so pair 0 of a clip can run on other levels than the later pairs
(include/vstab.h, vstab_dis_flow_at).  CUDA path vs the C oracle (pinned to live cv2 in
tests/test_oracle_dis.py) and vs goldens produced by the unmodified reference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dis_ref
from tests import cases, parity
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

SIZES = [(73, 45), (121, 73), (90, 50), (40, 24), (30, 100), (64, 48), (20, 9), (12, 12), (89, 33)]


def _gray_clip(w, h, n, seed):
    import synth
    from oracle.gray_np import gray_u8

    base = synth.base_texture(seed, w, h).numpy()
    clip = synth.render_clip_numpy(base, synth.shake_matrices(n, seed + 1, w, h, amount=2.0), w, h)
    return np.stack([gray_u8(f) for f in clip])


@pytest.mark.parametrize("size", SIZES, ids=[f"{w}x{h}" for w, h in SIZES])
def test_dis_small_frames_bit_exact_and_stateful(handle, size):
    w, h = size
    gray = _gray_clip(w, h, 4, 100 + w + h)
    g = torch.from_numpy(gray).cuda()
    flow, grid = handle.dis_flow(g, want_flow=True, grid_step=8)
    later, _ = handle.dis_flow(g, want_flow=True, grid_step=8, first_pair=5)  # a shard in the middle of a clip
    torch.cuda.synchronize()
    flow, grid, later = flow.cpu().numpy(), grid.cpu().numpy(), later.cpu().numpy()
    backend = dis_ref.Backend()          # fresh object: pair 0 is its first calc()
    for p in range(3):
        want = backend.calc(gray[p], gray[p + 1])
        assert np.array_equal(flow[p], want), (p, float(np.abs(flow[p] - want).max()))
        assert np.array_equal(grid[p], want[::8, ::8])
    for p in range(3):                    # the same (now rewritten) object: what pairs 5, 6, 7 of a clip meet
        want = backend.calc(gray[p], gray[p + 1])
        assert np.array_equal(later[p], want), (p, float(np.abs(later[p] - want).max()))


@pytest.mark.parametrize("size", [(100, 30), (200, 31), (10, 10), (7, 64)])
def test_dis_rejects_what_cv2_cannot_do(handle, size):
    """cv2 raises below 12 px and reads outside its coarsest level (usually a crash) on e.g. 100x30."""
    from vstab_b200._native import VstabNativeError

    w, h = size
    g = torch.zeros((2, h, w), dtype=torch.uint8, device="cuda")
    with pytest.raises(VstabNativeError):
        handle.dis_flow(g)


@pytest.mark.parametrize("case", cases.SMALL_STABILIZER_CASES, ids=[c["name"] for c in cases.SMALL_STABILIZER_CASES])
def test_small_clip_matches_reference_golden(case):
    from vstab_b200 import classic, flow, pipeline

    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    frames = cases.make_frames(case)
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
    driver = flow if case["node"] == "flow" else classic
    res = driver.stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                  case["keep_fov"], case["padding_rgb"], case["fps"])
    meta = res.meta
    assert meta["transform_mode_applied"] == gmeta["transform_mode_applied"]
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
        assert mine["mode"] == ref["mode"]
        parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}")
        assert abs(mine["confidence"] - ref["confidence"]) <= 1e-3
    for mine, ref in zip(meta["stabilization_warp"]["per_frame"], gmeta["stabilization_warp"]["per_frame"]):
        parity.assert_transform_close(mine["applied_matrix"], ref["applied_matrix"], f"frame {ref['index']}")
    parity.compare_nested(gmeta, json.loads(json.dumps(meta)), "meta", atol=2e-4, rtol=2e-4)
    assert tuple(res.frames.shape) == tuple(gold["shape"])
    assert float(np.abs(res.frames - gold["frames"]).max()) <= parity.TOL_PIXEL["bilinear"]
    assert int((np.asarray(res.masks) != gold["masks"]).sum()) <= 8  # identical rule, matrices equal to ~1e-6
