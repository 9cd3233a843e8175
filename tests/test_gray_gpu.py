"""CUDA gray + INTER_AREA working image (vstab_gray_working) vs the numpy oracle: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import gray_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(832, 480), (1920, 1080), (1280, 720), (3840, 2160), (1000, 562), (121, 73)])
def test_gray_working_bit_exact(handle, size):
    w, h = size
    rng = np.random.default_rng(w)
    n = 2
    rgb = rng.random((n, h, w, 3), dtype=np.float32)
    rgb[0, :4, :4] = 1.0  # saturating values
    rgb[0, 4:8, :4] = 0.0
    got = handle.gray_working(torch.from_numpy(rgb).cuda()).cpu().numpy()
    ws = gray_np.working_size(w, h)
    for i in range(n):
        want = gray_np.gray_for_estimation(rgb[i], ws)
        assert got[i].shape == want.shape
        assert int((got[i] != want).sum()) == 0


def test_gray_matches_cv2_when_available(handle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    rgb = rng.random((1, 720, 1280, 3), dtype=np.float32)
    got = handle.gray_working(torch.from_numpy(rgb).cuda()).cpu().numpy()[0]
    g = np.clip(cv2.cvtColor(rgb[0], cv2.COLOR_RGB2GRAY) * 255.0, 0, 255).astype(np.uint8)
    want = cv2.resize(g, (960, 540), interpolation=cv2.INTER_AREA)
    assert int((got != want).sum()) == 0
