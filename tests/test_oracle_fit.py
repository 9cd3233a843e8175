"""Pins oracle/fit_np.py (RANSAC replay, LS / DLT+LM, median) against live cv2."""
import numpy as np
import pytest

from oracle import dis_ref, fit_np
from tests import cases

cv2 = pytest.importorskip("cv2")


def _flows():
    case = cases.DIS_CASES[2]  # 320x180
    p, c = cases.make_gray_pair(case)
    flow = dis_ref.calc(p, c)
    rng = np.random.default_rng(1)
    out = {"clean": flow}
    blk = flow.copy()
    blk[40:120, 60:200] += np.array([9, -6], np.float32)
    out["block"] = blk
    noisy = flow.copy()
    m = rng.random(flow.shape[:2]) < 0.6
    noisy[m] += rng.normal(0, 20, (int(m.sum()), 2)).astype(np.float32)
    out["noise60"] = noisy
    H0 = np.array([[1.01, 0.01, 3], [-0.01, 0.99, -2], [9e-5, -6e-5, 1]])
    ys, xs = np.mgrid[0 : flow.shape[0], 0 : flow.shape[1]].astype(np.float64)
    w = H0[2, 0] * xs + H0[2, 1] * ys + 1
    persp = np.stack([(H0[0, 0] * xs + H0[0, 1] * ys + H0[0, 2]) / w - xs, (H0[1, 0] * xs + H0[1, 1] * ys + H0[1, 2]) / w - ys], -1)
    out["persp"] = persp.astype(np.float32) + rng.normal(0, 0.05, flow.shape).astype(np.float32)
    return out


FLOWS = _flows()


@pytest.mark.parametrize("name", list(FLOWS))
def test_similarity_matches_cv2(name):
    prev, curr, _ = fit_np.grid_correspondences(FLOWS[name])
    A, inl = cv2.estimateAffinePartial2D(prev, curr, method=cv2.RANSAC, ransacReprojThreshold=2.0, maxIters=2000, confidence=0.992)
    A2, inl2 = fit_np.estimate_affine_partial_2d(prev, curr)
    assert np.array_equal(inl.ravel(), inl2)          # same winning hypothesis => same consensus set
    assert np.abs(A - A2).max() <= 1e-12


@pytest.mark.parametrize("name", list(FLOWS))
def test_homography_matches_cv2(name):
    prev, curr, _ = fit_np.grid_correspondences(FLOWS[name])
    H, inl = cv2.findHomography(prev, curr, method=cv2.RANSAC, ransacReprojThreshold=2.5, maxIters=2000, confidence=0.992)
    H2, inl2 = fit_np.find_homography(prev, curr)
    assert int((inl.ravel() != inl2).sum()) <= 1      # float-threshold ties only
    assert np.abs(H - H2).max() <= 1e-8


@pytest.mark.parametrize("name", list(FLOWS))
def test_median_matches_numpy(name):
    prev, curr, _ = fit_np.grid_correspondences(FLOWS[name])
    assert np.array_equal(fit_np.median_shift(prev, curr), np.median(curr - prev, axis=0).astype(np.float32))
