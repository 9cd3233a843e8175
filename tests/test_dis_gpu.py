"""CUDA DIS (vstab_dis_flow) vs the C oracle (bit-exact against cv2) and the reference goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import dis_ref
from tests import cases
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _gpu_flow(handle, frames_u8, want_flow=True):
    g = torch.from_numpy(np.ascontiguousarray(frames_u8)).cuda()
    flow, grid = handle.dis_flow(g, want_flow=want_flow, grid_step=8)
    torch.cuda.synchronize()
    return (flow.cpu().numpy() if flow is not None else None), grid.cpu().numpy()


@pytest.mark.parametrize("case", cases.DIS_CASES, ids=[c["name"] for c in cases.DIS_CASES])
def test_dis_matches_reference_golden(handle, case):
    gold = np.load(os.path.join(GOLDEN_DIR, f"dis_{case['name']}.npz"))
    prev, curr = cases.make_gray_pair(case)
    flow, grid = _gpu_flow(handle, np.stack([prev, curr]))
    err = float(np.abs(grid[0] - gold["grid"]).max())
    assert err <= 1e-3, err                     # well inside the 0.05 px transform tolerance
    assert np.array_equal(grid[0], gold["grid"]), err   # and in fact the same bits
    assert np.array_equal(flow[0][::8, ::8], grid[0])


# 832x480 / 640x360: the resident refinement kernel on clusters of 4 / 2 row bands; 160x960 and 960x160: thin levels
# (many rows per band, few cells per row and the other way round); 333x187: odd sizes, single-CTA levels only
@pytest.mark.parametrize("size", [(320, 180), (640, 360), (832, 480), (960, 540), (540, 960), (160, 960), (960, 160), (333, 187)])
def test_dis_bit_exact_vs_oracle_batch(handle, size):
    """Several pairs in one launch, each compared with the C oracle."""
    w, h = size
    n = 4
    rng = np.random.default_rng(w)
    import synth
    from oracle.gray_np import gray_u8

    base = synth.base_texture(w, w, h).numpy()
    mats = synth.shake_matrices(n, w + 1, w, h, amount=2.0)
    clip = synth.render_clip_numpy(base, mats, w, h)
    gray = np.stack([gray_u8(f) for f in clip])
    flow, grid = _gpu_flow(handle, gray)
    for p in range(n - 1):
        want = dis_ref.calc(gray[p], gray[p + 1])
        err = float(np.abs(flow[p] - want).max())
        assert err <= 1e-3, (p, err)
        assert np.array_equal(flow[p], want), (p, err)


def test_dis_white_noise(handle):
    rng = np.random.default_rng(0)
    i0 = rng.integers(0, 256, (270, 480), dtype=np.uint8)
    i1 = np.roll(i0, (2, 3), (0, 1))
    flow, _ = _gpu_flow(handle, np.stack([i0, i1]))
    want = dis_ref.calc(i0, i1)
    assert float(np.abs(flow[0] - want).max()) <= 1e-3
    assert np.array_equal(flow[0], want)
