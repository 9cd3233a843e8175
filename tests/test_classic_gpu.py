"""Video Stabilizer Classic on the CUDA path: GFTT + LK vs the C oracle, node vs reference goldens."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import classic_ref as CR
from tests import cases, parity
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(960, 540), (832, 480), (640, 360), (540, 960), (333, 187), (121, 73)])
def test_gftt_lk_matches_oracle(handle, size):
    w, h = size
    import synth
    from oracle.gray_np import gray_u8

    n = 3
    base = synth.base_texture(w + 2, w, h).numpy()
    clip = synth.render_clip_numpy(base, synth.shake_matrices(n, w + 3, w, h, amount=1.5), w, h)
    gray = np.stack([gray_u8(f) for f in clip])
    prev, curr, det = handle.gftt_lk(torch.from_numpy(gray).cuda(), 400)
    torch.cuda.synchronize()
    prev, curr, det = prev.cpu().numpy(), curr.cpu().numpy(), det.cpu().numpy()
    for p in range(n - 1):
        feats = CR.good_features(gray[p])
        assert int(det[p]) == len(feats)
        assert np.array_equal(prev[p, : len(feats)], feats)       # identical corners, identical order
        assert np.isnan(prev[p, len(feats):]).all()
        want, st = CR.pyr_lk(gray[p], gray[p + 1], feats, exact=True)  # cv2's lane order: bit-exact against the wheel
        got = curr[p, : len(feats)]
        lost = np.isnan(got).any(axis=1)
        assert int((lost != (st == 0)).sum()) == 0
        assert np.array_equal(got[~lost], want[~lost]), float(np.abs(got[~lost] - want[~lost]).max())


CLASSIC_CASES = [c for c in cases.STABILIZER_CASES if c["node"] == "classic"]


@pytest.mark.parametrize("case", CLASSIC_CASES, ids=[c["name"] for c in CLASSIC_CASES])
def test_classic_matches_reference_golden(case):
    from vstab_b200 import classic, pipeline

    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    frames = cases.make_frames(case)
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
    res = classic.stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                   case["keep_fov"], case["padding_rgb"], case["fps"])
    meta = res.meta
    assert "flow_backend" not in meta and "residual" not in meta["estimated_motion"]["per_transition"][0]
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
        assert mine["mode"] == ref["mode"]
        parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}")
        assert abs(mine["confidence"] - ref["confidence"]) <= 0.01
    for mine, ref in zip(meta["stabilization_warp"]["per_frame"], gmeta["stabilization_warp"]["per_frame"]):
        parity.assert_transform_close(mine["applied_matrix"], ref["applied_matrix"], f"frame {ref['index']}")
    # corners and tracks carry cv2's bits; what remains is the float32 rounding of the fitted matrices (like Flow)
    parity.compare_nested(gmeta, json.loads(json.dumps(meta)), "meta", atol=2e-5, rtol=2e-5)
    f, y, x, hh, ww = gold["patch0_at"]
    assert float(np.abs(res.frames[f, y:y + hh, x:x + ww] - gold["patch0"]).max()) <= 1e-3
    assert tuple(res.frames.shape) == tuple(gold["shape"])


def test_long_clips_are_tracked_in_passes(handle, monkeypatch):
    """vstab_gftt_lk works through a clip in passes of 256 pairs (bounded workspace); the frame shared by two passes gets
    its pyramid twice.  With 4 pairs per pass on a 14-frame clip: same corners, tracks and counts as one pass."""
    import torch

    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (96 + 40, 160 + 40), dtype=np.uint8)
    base = np.asarray(torch.nn.functional.avg_pool2d(torch.from_numpy(base.astype(np.float32))[None, None], 3, 1, 1)[0, 0]).astype(np.uint8)
    frames = np.stack([base[20 + (i * 3) % 7: 116 + (i * 3) % 7, 20 + (i * 5) % 9: 180 + (i * 5) % 9] for i in range(14)])
    gray = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    monkeypatch.delenv("VSTAB_LK_PAIRS_PER_PASS", raising=False)
    want = [t.cpu().numpy() for t in handle.gftt_lk(gray, max_corners=200)]
    monkeypatch.setenv("VSTAB_LK_PAIRS_PER_PASS", "4")
    got = [t.cpu().numpy() for t in handle.gftt_lk(gray, max_corners=200)]
    assert want[2].min() >= 12  # a real test: corners were found in every frame
    for a, b in zip(want, got):
        assert np.array_equal(a, b, equal_nan=True)
