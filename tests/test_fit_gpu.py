"""CUDA model fit (vstab_fit_batch) vs the numpy oracle (itself pinned to cv2)."""
import numpy as np
import pytest
import torch

from oracle import fit_np
from tests.test_oracle_fit import FLOWS

pytestmark = pytest.mark.gpu


def _gpu(handle, flows):
    from vstab_b200 import _native

    grid = np.stack([f[::8, ::8] for f in flows]).astype(np.float32)
    raw = handle.fit_grid(torch.from_numpy(np.ascontiguousarray(grid)).cuda(), 8, 7)
    torch.cuda.synchronize()
    return _native.decode_fit_results(raw)


def test_all_candidates_match_oracle(handle):
    names = list(FLOWS)
    d = _gpu(handle, [FLOWS[n] for n in names])
    for p, name in enumerate(names):
        prev, curr, total = fit_np.grid_correspondences(FLOWS[name])
        assert int(d["n_total"][p, 0]) == total and int(d["n_valid"][p, 1]) == len(prev)
        # translation: same bits
        t = fit_np.median_shift(prev, curr)
        assert float(d["matrix"][p, 0, 0, 2]) == float(t[0]) and float(d["matrix"][p, 0, 1, 2]) == float(t[1])
        res = float(np.abs((prev + t) - curr).mean())
        assert abs(d["residual"][p, 0] - res) <= 1e-5 * max(1.0, res)
        # similarity: same consensus size, parameters to 1e-9
        A, inl = fit_np.estimate_affine_partial_2d(prev, curr)
        assert int(d["ok"][p, 1]) == 1 and int(d["n_inliers"][p, 1]) == int(inl.sum()), name
        assert np.abs(d["matrix"][p, 1, :2] - A).max() <= 1e-9, name
        res = float(np.abs((prev @ A[:, :2].T + A[:, 2]) - curr).mean())
        assert abs(d["residual"][p, 1] - res) <= 1e-9 * max(1.0, res)
        # perspective: consensus within float-threshold ties, model well inside the parity tolerance
        H, inl = fit_np.find_homography(prev, curr)
        assert int(d["ok"][p, 2]) == 1, name
        assert abs(int(d["n_inliers"][p, 2]) - int(inl.sum())) <= 2, name
        assert np.abs(d["matrix"][p, 2] - H).max() <= 1e-6, name


def test_nan_points_are_filtered_and_few_points_reported(handle):
    from vstab_b200 import _native

    rng = np.random.default_rng(0)
    prev = rng.uniform(0, 500, (2, 40, 2)).astype(np.float32)
    curr = prev + np.array([1.5, -2.25], np.float32)
    curr[0, 5:] = np.nan  # pair 0: only 5 valid points
    curr[1, ::7] = np.nan
    raw = handle.fit_points(torch.from_numpy(prev).cuda(), torch.from_numpy(curr).cuda(), 7)
    d = _native.decode_fit_results(raw)
    assert int(d["n_valid"][0, 0]) == 5 and int(d["n_valid"][1, 0]) == 40 - len(range(0, 40, 7))
    ok = np.isfinite(curr[1]).all(axis=1)
    A, inl = fit_np.estimate_affine_partial_2d(prev[1][ok], curr[1][ok])
    assert np.abs(d["matrix"][1, 1, :2] - A).max() <= 1e-9
    assert float(d["matrix"][1, 0, 0, 2]) == 1.5 and float(d["matrix"][1, 0, 1, 2]) == -2.25


def test_few_points_noise_and_outliers_match_oracle(handle):
    """Small frames give the fit 60-160 grid points and Classic as few as a dozen tracks: random clouds of 12..200
    points with noise and up to 70 % outliers, every pair against the oracle (pinned to cv2 in tests/test_oracle_fit.py)."""
    from vstab_b200 import _native

    rng = np.random.default_rng(7)
    P, K = 48, 200
    prev = np.full((P, K, 2), np.nan, np.float32)
    curr = np.full((P, K, 2), np.nan, np.float32)
    for p in range(P):
        n = int(rng.integers(12, K + 1))
        W, H = int(rng.integers(40, 960)), int(rng.integers(24, 540))
        a = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1).astype(np.float32)
        th, sc, t = rng.normal(0, 0.01), 1 + rng.normal(0, 0.01), rng.normal(0, 3, 2)
        A = np.array([[sc * np.cos(th), -sc * np.sin(th), t[0]], [sc * np.sin(th), sc * np.cos(th), t[1]]])
        b = (a @ A[:, :2].T + A[:, 2]).astype(np.float32)
        b += rng.normal(0, rng.choice([0.0, 0.02, 0.3, 1.0]), b.shape).astype(np.float32)
        m = rng.random(n) < rng.choice([0, 0.1, 0.4, 0.7])
        b[m] += rng.normal(0, 15, (int(m.sum()), 2)).astype(np.float32)
        prev[p, :n], curr[p, :n] = a, b
    d = _native.decode_fit_results(handle.fit_points(torch.from_numpy(prev).cuda(), torch.from_numpy(curr).cuda(), 7))
    for p in range(P):
        ok = np.isfinite(curr[p]).all(axis=1)
        a, b = prev[p][ok], curr[p][ok]
        assert int(d["n_valid"][p, 1]) == len(a)
        t = fit_np.median_shift(a, b)
        assert float(d["matrix"][p, 0, 0, 2]) == float(t[0]) and float(d["matrix"][p, 0, 1, 2]) == float(t[1])
        A, inl = fit_np.estimate_affine_partial_2d(a, b)
        assert int(d["ok"][p, 1]) == int(A is not None), p
        if A is not None:
            assert int(d["n_inliers"][p, 1]) == int(inl.sum()), p
            assert np.abs(d["matrix"][p, 1, :2] - A).max() <= 1e-9, p
        Hm, hin = fit_np.find_homography(a, b)
        assert int(d["ok"][p, 2]) == int(Hm is not None), p
        if Hm is not None and hin.sum() >= 0.5 * len(a):
            assert abs(int(d["n_inliers"][p, 2]) - int(hin.sum())) <= 2, p
            pts = np.concatenate([a[hin.astype(bool)].astype(np.float64), np.ones((int(hin.sum()), 1))], 1)
            q0, q1 = pts @ Hm.T, pts @ d["matrix"][p, 2].T
            assert np.abs(q0[:, :2] / q0[:, 2:] - q1[:, :2] / q1[:, 2:]).max() <= 1e-3, p  # px; north_star allows 0.05
