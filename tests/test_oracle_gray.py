"""Pins oracle/gray_np.py against live cv2 (bit-exact)."""
import numpy as np
import pytest

from oracle import gray_np

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("size", [(832, 480), (1920, 1080), (1280, 720), (1000, 562), (121, 73)])
def test_gray_and_area_bit_exact(size):
    w, h = size
    rgb = np.random.default_rng(w).random((h, w, 3), dtype=np.float32)
    g_ref = np.clip(cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY) * 255.0, 0, 255).astype(np.uint8)
    assert np.array_equal(gray_np.gray_u8(rgb), g_ref)
    ws = gray_np.working_size(w, h)
    if ws is not None:
        assert np.array_equal(gray_np.resize_area_u8(g_ref, ws), cv2.resize(g_ref, ws, interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("shape", [((240, 135), (120, 67)), ((120, 67), (60, 33)), ((60, 33), (30, 16)), ((960, 540), (240, 135)), ((3840, 2160), (960, 540))])
def test_area_pyramid_shapes(shape):
    (sw, sh), (dw, dh) = shape
    src = np.random.default_rng(sw).integers(0, 256, (sh, sw), dtype=np.uint8)
    assert np.array_equal(gray_np.resize_area_u8(src, (dw, dh)), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA))


def test_working_size_rule():
    assert gray_np.working_size(1920, 1080) == (960, 540)
    assert gray_np.working_size(3840, 2160) == (960, 540)
    assert gray_np.working_size(1280, 720) == (960, 540)
    assert gray_np.working_size(832, 480) is None
    assert gray_np.working_size(1080, 1920) == (540, 960)
