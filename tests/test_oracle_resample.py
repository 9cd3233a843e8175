"""Pins oracle/resample_np.py + oracle/apply_np.py against live cv2 and the reference goldens."""
import json
import os

import numpy as np
import pytest

from oracle import apply_np, resample_np as R
from tests import cases
from tests.conftest import GOLDEN_DIR

cv2 = pytest.importorskip("cv2")


def _rand_matrix(rng, kind, shift=15.0):
    th, s = rng.normal(0, 0.01), 1 + rng.normal(0, 0.01)
    tx, ty = rng.normal(0, shift, 2)
    m = np.array([[s * np.cos(th), -s * np.sin(th), tx], [s * np.sin(th), s * np.cos(th), ty], [0, 0, 1]])
    if kind == "persp":
        m[2, :2] = rng.normal(0, 2e-5, 2)
    return m.astype(np.float32)


@pytest.mark.parametrize("size", [(121, 73), (320, 200), (832, 480)])
@pytest.mark.parametrize("kind", ["sim", "persp"])
def test_warp_matches_cv2(size, kind):
    rng = np.random.default_rng(hash((size, kind)) % 1000)
    w, h = size
    src = rng.random((h, w, 3), dtype=np.float32)
    border = [0.5, 0.25, 0.75]
    for out_size in [(w, h), (w + 10, h + 8)]:
        m = _rand_matrix(rng, kind)
        for interp, flag, tol in [("bilinear", cv2.INTER_LINEAR, 0.0), ("bicubic", cv2.INTER_CUBIC, 0.0)]:
            ref = cv2.warpPerspective(src, m, out_size, flags=flag, borderMode=cv2.BORDER_CONSTANT, borderValue=border)
            mine = R.warp_np(src, m, out_size, interp, border)
            assert float(np.abs(ref - mine).max()) <= tol, (interp, out_size)


@pytest.mark.parametrize("size", [(121, 73), (832, 480), (1920, 1080)])
def test_mask_rule_p_matches_cv2(size):
    rng = np.random.default_rng(5)
    w, h = size
    ones = np.ones((h, w), np.float32)
    cv2.setNumThreads(1)  # one stripe => the IPP path never falls back to Rule C
    try:
        for kind in ["sim", "persp"]:
            for out_size in [(w, h), (w + 10, h + 8)]:
                m = _rand_matrix(rng, kind)
                cov = cv2.warpPerspective(ones, m, out_size, flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0) > 0.5
                assert int((R.coverage_np(m, (w, h), out_size, R.RULE_P) != cov).sum()) == 0
        for t in [(3.0, -2.0), (0.5, 0.5), (-7.5, 4.0)]:  # integer / half-pixel translations: exact ties
            m = np.array([[1, 0, t[0]], [0, 1, t[1]], [0, 0, 1]], np.float32)
            cov = cv2.warpPerspective(ones, m, (w, h), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0) > 0.5
            assert int((R.coverage_np(m, (w, h), (w, h), R.RULE_P) != cov).sum()) == 0
    finally:
        cv2.setNumThreads(-1)


def test_mask_rule_c_stripe_fallback():
    """With 8 threads a fully-outside destination stripe flips cv2 to the classic rule (A.3)."""
    if cv2.getNumThreads() < 8:
        pytest.skip("needs the 8-stripe configuration")
    w, h = 832, 480
    ones = np.ones((h, w), np.float32)
    m = np.array([[1, 0, -0.3], [0, 1, h / 8 + 2.0], [0, 0, 1]], np.float32)
    cov = cv2.warpPerspective(ones, m, (w, h), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0) > 0.5
    assert int((R.coverage_np(m, (w, h), (w, h), R.RULE_C) != cov).sum()) == 0


SMALL = [c for c in cases.MOTION_APPLY_CASES if c["store"] == "full"]


@pytest.mark.parametrize("case", SMALL, ids=[c["name"] for c in SMALL])
def test_apply_oracle_matches_reference_golden(case):
    gold = np.load(os.path.join(GOLDEN_DIR, f"apply_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"apply_{case['name']}_meta.json")) as fh:
        meta = json.load(fh)
    frames = cases.make_frames(case)
    out_f, out_m, _, _ = apply_np.apply_motion_np(frames, meta, case["padding_rgb"], case["framing"], case["interp"],
                                                  case["blur"], case["samples"])
    assert out_f.shape == gold["frames"].shape
    assert np.array_equal(out_f, gold["frames"])  # bilinear, bicubic (cv2's row-sum order) and the blur accumulation alike
    assert np.array_equal(out_m, gold["masks"]) if case["blur"] == 0.0 else float(np.abs(out_m - gold["masks"]).max()) <= 1e-6


def test_bicubic_in_cv2_row_order_is_bit_exact():
    """warp_np (cubic_rows=True, the default and what the CUDA resampler does): the bicubic sum in cv2's own order (row
    sums first) carries cv2's bits on every pixel, interior and border alike; one running sum over the 16 taps
    (cubic_rows=False) only stays within 4.8e-7."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for k in range(10):
        w, h = int(rng.integers(40, 300)), int(rng.integers(40, 200))
        ow, oh = w + int(rng.integers(-5, 30)), h + int(rng.integers(-5, 30))
        src = rng.random((h, w, 3), dtype=np.float32)
        th, sc = rng.normal(0, 0.05), 1 + rng.normal(0, 0.05)
        M = np.array([[sc * np.cos(th), -sc * np.sin(th), rng.normal(0, 6)], [sc * np.sin(th), sc * np.cos(th), rng.normal(0, 6)],
                      [0, 0, 1]], np.float64)
        if k % 3 == 0:
            M[2, :2] = rng.normal(0, 2e-4, 2)
        M = M.astype(np.float32)
        border = tuple(float(x) for x in (rng.integers(0, 256, 3) / 255.0).astype(np.float32))
        ref = cv2.warpPerspective(src, M, (ow, oh), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT, borderValue=border)
        assert np.array_equal(ref, R.warp_np(src, M, (ow, oh), "bicubic", border, cubic_rows=True)), k
        assert np.array_equal(ref, R.warp_np(src, M, (ow, oh), "bicubic", border)), k
        assert float(np.abs(ref - R.warp_np(src, M, (ow, oh), "bicubic", border, cubic_rows=False)).max()) <= 1e-6
