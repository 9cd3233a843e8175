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


def test_area_tables_are_recycled_when_the_cache_is_full(handle):
    """A long-lived process sees many frame sizes; every non-integer INTER_AREA ratio needs a coverage table per
    (source, destination) length and the handle keeps 32 of them.  More sizes than that must keep working
    (slots recycled round robin) and sizes seen before must give the same bytes when their table is rebuilt."""
    rng = np.random.default_rng(9)
    widths = list(range(962, 1004))  # 42 widths -> 42 x tables (+ the y tables) > 32 slots
    first = None
    for k, w in enumerate(widths + widths[:2]):
        h = 300 + (k % 3)
        rgb = rng.random((1, h, w, 3), dtype=np.float32) if k < len(widths) else first[k - len(widths)][0]
        got = handle.gray_working(torch.from_numpy(rgb).cuda()).cpu().numpy()[0]
        want = gray_np.gray_for_estimation(rgb[0], gray_np.working_size(w, rgb.shape[1]))
        assert got.shape == want.shape and int((got != want).sum()) == 0, (k, w)
        if k < 2:
            first = (first or []) + [(rgb, got)]
        if k >= len(widths):
            assert np.array_equal(got, first[k - len(widths)][1])
