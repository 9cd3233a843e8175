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


def _cv_rule(m, src, out):
    """Which rule the live wheel applied to this call (None when both rules give the same mask)."""
    ones = np.ones((src[1], src[0]), np.float32)
    cov = cv2.warpPerspective(ones, m, out, flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0) > 0.5
    dp = int((cov != R.coverage_np(m, src, out, R.RULE_P)).sum())
    dc = int((cov != R.coverage_np(m, src, out, R.RULE_C)).sum())
    assert dp == 0 or dc == 0, (dp, dc)  # one of the two rules always explains the wheel's mask exactly
    return None if dp == dc == 0 else (R.RULE_P if dp == 0 else R.RULE_C)


@pytest.mark.parametrize("threads", [8, 4, 3])
def test_mask_rule_auto_matches_what_cv2_picks(threads):
    """the wheel switches the WHOLE call to Rule C when one of its destination stripes misses
    the source.  resample_np.auto_rule restates the predicate (stripe count = min(threads, ceil(W'H'/2^14)), boundaries
    (s H' + S/2) // S, forward quad of the source's pixel centres vs the stripe's pixel-centre rectangle); here against
    the live wheel on random large off-canvas warps and on shifts one row either side of every flip."""
    if (os.cpu_count() or 1) < threads:
        pytest.skip("not enough CPUs for this cv2 thread count")
    cv2.setNumThreads(threads)
    try:
        if cv2.getNumThreads() != threads:
            pytest.skip("cv2 did not accept the thread count")
        rng = np.random.default_rng(threads)
        decided = 0
        for trial in range(150):
            w, h = [(832, 480), (320, 200), (640, 360), (400, 300), (200, 120), (256, 256)][trial % 6]
            out = (w, h) if trial % 3 else (w + int(rng.integers(-40, 80)), h + int(rng.integers(-40, 80)))
            th, sc = rng.normal(0, 0.15), 1 + rng.normal(0, 0.15)
            m = np.array([[sc * np.cos(th), -sc * np.sin(th), rng.normal(0, 0.6 * w) + 0.3],
                          [sc * np.sin(th), sc * np.cos(th), rng.normal(0, 0.6 * h)], [0, 0, 1]])
            if trial % 4 == 0:
                m[2, :2] = rng.normal(0, 3e-4, 2)
            m = m.astype(np.float32)
            want = _cv_rule(m, (w, h), out)
            if want is not None:
                decided += 1
                assert R.auto_rule(m, (w, h), out, threads) == want, (trial, (w, h), out)
        assert decided >= 100
        for w, h in [(832, 480), (400, 300), (640, 360)]:
            stripes = R.mask_stripes((w, h), threads)
            for edge, b in (("top", stripes[0][1]), ("bottom", h - stripes[-1][0])):
                for d in (-1.5, -1.0, -0.5, -0.1, 0.0, 0.1, 0.5, 1.0):
                    ty = (b + d) if edge == "top" else -(b + d)
                    m = np.array([[1, 0, -0.3], [0, 1, ty], [0, 0, 1]], np.float32)
                    want = _cv_rule(m, (w, h), (w, h))
                    assert want is not None and R.auto_rule(m, (w, h), (w, h), threads) == want, (w, h, edge, d)
    finally:
        cv2.setNumThreads(-1)
    assert R.auto_rule(np.array([[1, 0, -0.3], [0, 1, 300.0], [0, 0, 1]], np.float32), (832, 480), (832, 480), 1) == R.RULE_P
