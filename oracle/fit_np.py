"""numpy restatement of the reference's per-pair model fit.  TEST INFRASTRUCTURE ONLY.

Follows nodes/video_stabilizer_flow.py:141-210 (_estimate_motion_flow: 8-px grid sampling, finite
filter, perspective -> similarity -> translation ladder with the 0.15 / 0.10 acceptance
thresholds) and nodes/video_stabilizer_classic.py:112-158, with the cv2 4.13.0.92 estimators
restated from the published algorithm of modules/calib3d/src/ptsetreg.cpp / fundam.cpp:

  RANSAC driver   RNG(0xffffffffffffffff) multiply-with-carry stream, `modelPoints` distinct
                  indices per draw (re-draw the current index on a duplicate), strict `>`
                  acceptance, RANSACUpdateNumIters with confidence 0.992
  similarity      2-point closed form in double, squared error in double -> float, thr 2.0;
                  final model = least squares over the winner's inlier set (cv2 runs LM on a
                  linear problem: identical to ~1e-14)
  perspective     4-point subsets (collinearity + orientation checks), normalised DLT, float
                  reprojection error, thr 2.5; DLT on the inliers + <=10 Levenberg-Marquardt steps
  translation     per-axis median of float32 shifts

Pinned against live cv2 in tests/test_oracle_fit.py.
"""
from __future__ import annotations

import math

import numpy as np

MASK32 = 0xFFFFFFFF
RNG_COEFF = 4164903690


class CvRNG:
    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self) -> int:
        self.state = ((self.state & MASK32) * RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & MASK32

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else self.next() % (b - a) + a


def ransac_update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def grid_correspondences(flow: np.ndarray, step: int = 8):
    """flow.py:141-152: prev grid points (float32), curr = prev + flow, finite filter."""
    h, w = flow.shape[:2]
    ys = np.arange(0, h, step, dtype=np.int32)
    xs = np.arange(0, w, step, dtype=np.int32)
    gy, gx = np.meshgrid(ys, xs, indexing="ij")
    prev = np.stack([gx.ravel(), gy.ravel()], axis=1).astype(np.float32)
    curr = prev + flow[gy, gx].reshape(-1, 2)
    ok = np.isfinite(curr).all(axis=1)
    return prev[ok], curr[ok], int(len(prev))


# ------------------------------------------------------------------------------- similarity ----

def _similarity_from_2(f, t):
    x1, y1, x2, y2 = float(f[0, 0]), float(f[0, 1]), float(f[1, 0]), float(f[1, 1])
    X1, Y1, X2, Y2 = float(t[0, 0]), float(t[0, 1]), float(t[1, 0]), float(t[1, 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.float64(1.0) / np.float64((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2))
        s0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2))
        s1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2))
        s2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2))
        s3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2))
    return np.array([[s0, -s1, s2], [s1, s0, s3]], dtype=np.float64)


def _affine_err(model, prev, curr):
    px, py = prev[:, 0].astype(np.float64), prev[:, 1].astype(np.float64)
    a = model[0, 0] * px + model[0, 1] * py + model[0, 2] - curr[:, 0].astype(np.float64)
    b = model[1, 0] * px + model[1, 1] * py + model[1, 2] - curr[:, 1].astype(np.float64)
    return (a * a + b * b).astype(np.float32)


def _draw_subset(rng: CvRNG, count: int, k: int):
    idx = []
    for _ in range(k):
        v = rng.uniform(0, count)
        while v in idx:
            v = rng.uniform(0, count)
        idx.append(v)
    return idx


def similarity_ls(prev, curr):
    """Closed-form least-squares [a -b tx; b a ty] in double (what cv2's LM refine converges to)."""
    x, y = prev[:, 0].astype(np.float64), prev[:, 1].astype(np.float64)
    X, Y = curr[:, 0].astype(np.float64), curr[:, 1].astype(np.float64)
    xc, yc, Xc, Yc = x.mean(), y.mean(), X.mean(), Y.mean()
    xd, yd, Xd, Yd = x - xc, y - yc, X - Xc, Y - Yc
    den = (xd * xd + yd * yd).sum()
    a = (xd * Xd + yd * Yd).sum() / den
    b = (xd * Yd - yd * Xd).sum() / den
    return np.array([[a, -b, Xc - a * xc + b * yc], [b, a, Yc - b * xc - a * yc]], dtype=np.float64)


def estimate_affine_partial_2d(prev, curr, thresh=2.0, max_iters=2000, confidence=0.992):
    """cv2.estimateAffinePartial2D(prev, curr, RANSAC, thresh, max_iters, confidence) -> (2x3 | None, mask)."""
    count = len(prev)
    if count < 2:
        return None, np.zeros(count, np.uint8)
    rng = CvRNG()
    t = np.float32(thresh * thresh)
    best_mask, best_count, niters = None, 0, max(max_iters, 1)
    if count == 2:
        return _similarity_from_2(prev, curr), np.ones(count, np.uint8)
    it = 0
    while it < niters:
        idx = _draw_subset(rng, count, 2)
        model = _similarity_from_2(prev[idx], curr[idx])
        err = _affine_err(model, prev, curr)
        with np.errstate(invalid="ignore"):
            mask = err <= t
        good = int(mask.sum())
        if good > max(best_count, 1):
            best_mask, best_count = mask, good
            niters = ransac_update_num_iters(confidence, (count - good) / count, 2, niters)
        it += 1
    if best_count <= 0:
        return None, np.zeros(count, np.uint8)
    return similarity_ls(prev[best_mask], curr[best_mask]), best_mask.astype(np.uint8)


# ------------------------------------------------------------------------------ translation ----

def median_shift(prev, curr):
    """np.median(curr - prev, axis=0) in float32 (mean of the two middle values for even n)."""
    shifts = (curr - prev).astype(np.float32)
    return np.median(shifts, axis=0).reshape(-1).astype(np.float32)


# ------------------------------------------------------------------------------ perspective ----

def _collinear(pts, count):
    i = count - 1
    for j in range(i):
        dx1, dy1 = float(pts[j, 0]) - float(pts[i, 0]), float(pts[j, 1]) - float(pts[i, 1])
        for k in range(j):
            dx2, dy2 = float(pts[k, 0]) - float(pts[i, 0]), float(pts[k, 1]) - float(pts[i, 1])
            if abs(dx2 * dy1 - dy2 * dx1) <= np.finfo(np.float32).eps * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3(a):
    return (a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
            + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]))


def _homography_subset_ok(src, dst, count):
    if _collinear(src, count) or _collinear(dst, count):
        return False
    if count == 4:
        negative = 0
        for t in ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)):
            A = [[float(src[k, 0]), float(src[k, 1]), 1.0] for k in t]
            B = [[float(dst[k, 0]), float(dst[k, 1]), 1.0] for k in t]
            negative += _det3(A) * _det3(B) < 0
        if negative not in (0, 4):
            return False
    return True


def homography_dlt(src, dst):
    """HomographyEstimatorCallback::runKernel: normalised DLT, eigenvector of the smallest eigenvalue."""
    M = src.astype(np.float64)
    m = dst.astype(np.float64)
    n = len(M)
    cM, cm = M.mean(axis=0), m.mean(axis=0)
    sM, sm = np.abs(M - cM).sum(axis=0), np.abs(m - cm).sum(axis=0)
    eps = np.finfo(np.float64).eps
    if (np.abs(sM) < eps).any() or (np.abs(sm) < eps).any():
        return None
    sM, sm = n / sM, n / sm
    x, y = (m[:, 0] - cm[0]) * sm[0], (m[:, 1] - cm[1]) * sm[1]
    X, Y = (M[:, 0] - cM[0]) * sM[0], (M[:, 1] - cM[1]) * sM[1]
    one, zero = np.ones(n), np.zeros(n)
    Lx = np.stack([X, Y, one, zero, zero, zero, -x * X, -x * Y, -x], axis=1)
    Ly = np.stack([zero, zero, zero, X, Y, one, -y * X, -y * Y, -y], axis=1)
    LtL = Lx.T @ Lx + Ly.T @ Ly
    w, v = np.linalg.eigh(LtL)
    h0 = v[:, 0].reshape(3, 3)
    inv_hnorm = np.array([[1.0 / sm[0], 0, cm[0]], [0, 1.0 / sm[1], cm[1]], [0, 0, 1.0]])
    hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1.0]])
    H = inv_hnorm @ h0 @ hnorm2
    return H / H[2, 2]


def _homography_err(H, src, dst):
    hf = H.astype(np.float32).reshape(-1)
    Mx, My = src[:, 0].astype(np.float32), src[:, 1].astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        ww = np.float32(1.0) / (hf[6] * Mx + hf[7] * My + np.float32(1.0))
        dx = (hf[0] * Mx + hf[1] * My + hf[2]) * ww - dst[:, 0].astype(np.float32)
        dy = (hf[3] * Mx + hf[4] * My + hf[5]) * ww - dst[:, 1].astype(np.float32)
        return (dx * dx + dy * dy).astype(np.float32)


def _homography_lm(H, src, dst, iters=10):
    """cv::LMSolver on the 8-parameter forward reprojection error (HomographyRefineCallback)."""
    M = src.astype(np.float64)
    m = dst.astype(np.float64)

    def residual_jac(h):
        Mx, My = M[:, 0], M[:, 1]
        ww = h[6] * Mx + h[7] * My + 1.0
        ww = np.where(np.abs(ww) > np.finfo(np.float64).eps, 1.0 / ww, 0.0)
        xi = (h[0] * Mx + h[1] * My + h[2]) * ww
        yi = (h[3] * Mx + h[4] * My + h[5]) * ww
        err = np.empty(2 * len(M))
        err[0::2] = xi - m[:, 0]
        err[1::2] = yi - m[:, 1]
        J = np.zeros((2 * len(M), 8))
        J[0::2, 0], J[0::2, 1], J[0::2, 2] = Mx * ww, My * ww, ww
        J[0::2, 6], J[0::2, 7] = -Mx * ww * xi, -My * ww * xi
        J[1::2, 3], J[1::2, 4], J[1::2, 5] = Mx * ww, My * ww, ww
        J[1::2, 6], J[1::2, 7] = -Mx * ww * yi, -My * ww * yi
        return err, J

    # cv::LMSolverImpl::run: lambda starts at 1, halves on a good step (R > .75) and snaps to 0 once
    # below 0.75 (pure Gauss-Newton); D = diag(JtJ) of the initial point; eps = FLT_EPSILON.
    x = H.reshape(-1)[:8].copy()
    r, J = residual_jac(x)
    S = float(r @ r)
    A, v = J.T @ J, J.T @ r
    D = np.diag(A).copy()
    lam, lc = 1.0, 0.75
    eps_stop = float(np.finfo(np.float32).eps)
    deps = np.finfo(np.float64).eps
    it = 0
    while True:
        Ap = A + np.diag(lam * D)
        d = np.linalg.solve(Ap, v)
        xd = x - d
        rd, _ = residual_jac(xd)
        Sd = float(rd @ rd)
        temp_d = -(A @ d) + 2 * v
        dS = float(d @ temp_d)
        R = (S - Sd) / (dS if abs(dS) > deps else 1.0)
        if R > 0.75:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < 0.25:
            t = float(d @ v)
            nu = (Sd - S) / (t if abs(t) > deps else 1.0) + 2
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0:
                inv = np.linalg.inv(A)
                lam = lc = 1.0 / max(deps, float(np.abs(np.diag(inv)).max()))
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x = xd
            r, J = residual_jac(x)
            A, v = J.T @ J, J.T @ r
        it += 1
        if not (it < iters and np.abs(d).max() >= eps_stop and np.abs(r).max() >= eps_stop):
            break
    out = np.ones(9)
    out[:8] = x
    return out.reshape(3, 3)


def find_homography(prev, curr, thresh=2.5, max_iters=2000, confidence=0.992):
    """cv2.findHomography(prev, curr, RANSAC, thresh, maxIters, confidence) -> (3x3 | None, mask)."""
    count = len(prev)
    if count < 4:
        return None, np.zeros(count, np.uint8)
    rng = CvRNG()
    t = np.float32(thresh * thresh)
    best_mask, best_count, niters = None, 0, max(max_iters, 1)
    it = 0
    if count == 4:
        H = homography_dlt(prev, curr)
        return H, np.ones(count, np.uint8)
    while it < niters:
        found = False
        for _ in range(10000):
            idx = _draw_subset(rng, count, 4)
            if _homography_subset_ok(prev[idx], curr[idx], 4):
                found = True
                break
        if not found:
            if it == 0:
                return None, np.zeros(count, np.uint8)
            break
        H = homography_dlt(prev[idx], curr[idx])
        it += 1
        if H is None:
            continue
        err = _homography_err(H, prev, curr)
        with np.errstate(invalid="ignore"):
            mask = err <= t
        good = int(mask.sum())
        if good > max(best_count, 3):
            best_mask, best_count = mask, good
            niters = ransac_update_num_iters(confidence, (count - good) / count, 4, niters)
    if best_count <= 0:
        return None, np.zeros(count, np.uint8)
    src, dst = prev[best_mask], curr[best_mask]
    H = homography_dlt(src, dst)
    if H is None:
        return None, np.zeros(count, np.uint8)
    H = _homography_lm(H, src, dst, 10)
    # cv2 >= 4.5 re-evaluates the mask with the refined model (verified black-box: the returned mask
    # equals thresholding the returned H, not the winning minimal hypothesis)
    with np.errstate(invalid="ignore"):
        final_mask = _homography_err(H, prev, curr) <= t
    return H, final_mask.astype(np.uint8)


# ----------------------------------------------------------------------------------- ladder ----

LADDER = {
    "perspective": ("perspective", "similarity", "translation"),
    "similarity": ("similarity", "translation"),
    "translation": ("translation",),
}


def estimate_from_points(prev, curr, n_total, requested: str, *, min_points: int = 12, with_residual: bool = True):
    """flow.py:153-210 (min_points 12) / classic.py:101-158 (min_points 8, no residual).
    Returns (matrix f32 3x3, mode, confidence, residual)."""
    eye = np.eye(3, dtype=np.float32)
    if len(prev) < min_points:
        return eye, "translation", 0.0, 0.0
    for mode in LADDER[requested]:
        if mode == "perspective" and len(prev) >= 4:
            H, inl = find_homography(prev, curr)
            if H is not None:
                conf = float(inl.sum()) / float(len(prev))
                if conf >= 0.15:
                    res = float(np.abs((prev @ H[:2, :2].T + H[:2, 2]) - curr).mean()) if with_residual else 0.0
                    return H.astype(np.float32), "perspective", conf, res
        elif mode == "similarity" and len(prev) >= 3:
            A, inl = estimate_affine_partial_2d(prev, curr)
            if A is not None:
                conf = float(inl.sum()) / float(len(prev))
                if conf >= 0.1:
                    m = np.vstack([A, np.array([0.0, 0.0, 1.0], dtype=np.float32)])
                    res = float(np.abs((prev @ A[:, :2].T + A[:, 2]) - curr).mean()) if with_residual else 0.0
                    return m.astype(np.float32), "similarity", conf, res
        elif mode == "translation":
            delta = median_shift(prev, curr)
            tx, ty = float(delta[0]), float(delta[1])
            m = np.array([[1.0, 0.0, tx], [0.0, 1.0, ty], [0.0, 0.0, 1.0]], dtype=np.float32)
            conf = float(len(prev)) / float(n_total)
            res = float(np.abs((prev + np.array([tx, ty], dtype=np.float32)) - curr).mean()) if with_residual else 0.0
            return m, "translation", conf, res
    return eye, "translation", 0.0, 0.0
