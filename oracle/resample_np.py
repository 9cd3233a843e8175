"""numpy restatement of the reference's resampling path.  TEST INFRASTRUCTURE ONLY.

Follows (reference file:line, relative to /root/reference):
  * nodes/video_stabilizer_flow.py:560-588      warp + mask loop body
  * nodes/motion_apply.py:75-122                _warp_with_matrices
  * nodes/motion_apply.py:125-202               _blurred_matrix_samples / _warp_with_motion_blur
and the semantics of the cv2 4.13.0.92 calls they make, as probed in
SURVEY.md Appendix A.1-A.3 (classic fixed-point remap: double coordinates,
1/32-px half-to-even quantisation, f32 separable weight tables, per-tap
BORDER_CONSTANT; padding mask = closed-rectangle test on continuous coords).

Pinned in tests/test_oracle_resample.py against the live cv2 wheel.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS  # 32

RULE_P = 0  # closed rectangle on continuous source coordinates (IPP HAL path)
RULE_C = 1  # classic: round-half-even then range check


def invert3(m32: np.ndarray) -> np.ndarray:
    """cv::invert for a 3x3 in double: cofactors times 1/det (A.1)."""
    m = np.asarray(m32, dtype=np.float32).astype(np.float64)
    a, b, c = m[0]
    d, e, f = m[1]
    g, h, i = m[2]
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
    if det == 0.0:
        return np.zeros((3, 3), dtype=np.float64)
    r = 1.0 / det
    return np.array(
        [
            [(e * i - f * h) * r, (c * h - b * i) * r, (b * f - c * e) * r],
            [(f * g - d * i) * r, (a * i - c * g) * r, (c * d - a * f) * r],
            [(d * h - e * g) * r, (b * g - a * h) * r, (a * e - b * d) * r],
        ],
        dtype=np.float64,
    )


def _homog(mi: np.ndarray, out_w: int, out_h: int):
    xs = np.arange(out_w, dtype=np.float64)[None, :]
    ys = np.arange(out_h, dtype=np.float64)[:, None]
    X = mi[0, 0] * xs + mi[0, 1] * ys + mi[0, 2]
    Y = mi[1, 0] * xs + mi[1, 1] * ys + mi[1, 2]
    W = mi[2, 0] * xs + mi[2, 1] * ys + mi[2, 2]
    return X, Y, W


def fixed_point_coords(mi: np.ndarray, out_w: int, out_h: int):
    """Integer texel + 5-bit fraction per destination pixel (A.1)."""
    X, Y, W = _homog(mi, out_w, out_h)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(W != 0.0, INTER_TAB / W, 0.0)
    lim_lo, lim_hi = float(-(2**31)), float(2**31 - 1)
    fx = np.clip(X * s, lim_lo, lim_hi)
    fy = np.clip(Y * s, lim_lo, lim_hi)
    ix = np.rint(fx).astype(np.int64)  # half-to-even
    iy = np.rint(fy).astype(np.int64)
    sx = np.clip(ix >> INTER_BITS, -32768, 32767)
    sy = np.clip(iy >> INTER_BITS, -32768, 32767)
    ax = (ix & (INTER_TAB - 1)).astype(np.int64)
    ay = (iy & (INTER_TAB - 1)).astype(np.int64)
    return sx, sy, ax, ay


def _linear_tab() -> np.ndarray:
    a = np.arange(INTER_TAB, dtype=np.float32) * np.float32(1.0 / INTER_TAB)
    return np.stack([np.float32(1.0) - a, a], axis=1).astype(np.float32)  # [32][2]


def _cubic_tab() -> np.ndarray:
    A = np.float32(-0.75)
    x = np.arange(INTER_TAB, dtype=np.float32) * np.float32(1.0 / INTER_TAB)
    one = np.float32(1.0)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    xm = one - x
    c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.float32)  # [32][4]


LINEAR_TAB = _linear_tab()
CUBIC_TAB = _cubic_tab()


def _gather(src: np.ndarray, yy: np.ndarray, xx: np.ndarray, border: np.ndarray) -> np.ndarray:
    h, w = src.shape[:2]
    inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
    vals = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
    return np.where(inside[..., None], vals, border[None, None, :].astype(np.float32))


def warp_np(src: np.ndarray, m32: np.ndarray, out_size, interp: str, border, cubic_rows: bool = True) -> np.ndarray:
    """cv2.warpPerspective(src, m32, out_size, INTER_LINEAR|INTER_CUBIC, BORDER_CONSTANT, border).

    cubic_rows: order of the 16-term bicubic sum of pixels whose 4x4 footprint lies inside the source.  True (default,
    what the CUDA resampler does since round 2): cv2's own order, found black-box: the four taps of a row are summed
    first and the row sums are added to the running sum (remapBicubic's `sum += S[0]*w[4] + S[cn]*w[5] + ...`);
    bit-exact against the wheel.  False: one running sum over the 16 taps, row-major (<= 4.8e-7 from cv2; kept to show
    the difference)."""
    src = np.asarray(src, dtype=np.float32)
    if src.ndim == 2:
        src = src[..., None]
    c = src.shape[2]
    border = np.broadcast_to(np.asarray(border, dtype=np.float32).reshape(-1), (c,)) if np.ndim(border) else np.full((c,), border, np.float32)
    out_w, out_h = int(out_size[0]), int(out_size[1])
    mi = invert3(m32)
    sx, sy, ax, ay = fixed_point_coords(mi, out_w, out_h)
    h, w = src.shape[:2]
    out = np.zeros((out_h, out_w, c), dtype=np.float32)
    bvec = border[None, None, :].astype(np.float32)
    if interp == "bilinear":
        tx = LINEAR_TAB[ax]
        ty = LINEAR_TAB[ay]
        for k1 in range(2):
            for k2 in range(2):
                wgt = (ty[..., k1] * tx[..., k2]).astype(np.float32)
                out = out + _gather(src, sy + k1, sx + k2, border) * wgt[..., None]
        # remapBilinear: a footprint entirely outside the source is the border colour itself
        gone = (sx >= w) | (sx + 1 < 0) | (sy >= h) | (sy + 1 < 0)
        out = np.where(gone[..., None], bvec, out)
    elif interp == "bicubic":
        tx = CUBIC_TAB[ax]
        ty = CUBIC_TAB[ay]
        bx, by = sx - 1, sy - 1
        edge = np.zeros((out_h, out_w, c), dtype=np.float32) + bvec  # cv + sum (S - cv) * w over in-range taps
        for k1 in range(4):
            row = None
            for k2 in range(4):
                wgt = (ty[..., k1] * tx[..., k2]).astype(np.float32)
                yy, xx = by + k1, bx + k2
                tap = _gather(src, yy, xx, border)
                if cubic_rows:
                    row = tap * wgt[..., None] if row is None else row + tap * wgt[..., None]
                else:
                    out = out + tap * wgt[..., None]
                inside = ((yy >= 0) & (yy < h) & (xx >= 0) & (xx < w))[..., None]
                edge = np.where(inside, edge + (tap - bvec) * wgt[..., None], edge)
            if cubic_rows:
                out = row if k1 == 0 else out + row
        interior = (bx >= 0) & (bx < w - 3) & (by >= 0) & (by < h - 3)
        gone = (bx >= w) | (bx + 3 < 0) | (by >= h) | (by + 3 < 0)
        out = np.where(interior[..., None], out, edge)
        out = np.where(gone[..., None], bvec, out)
    else:
        raise ValueError(interp)
    return out.astype(np.float32)


def coverage_np(m32: np.ndarray, src_size, out_size, rule: int = RULE_P) -> np.ndarray:
    """(cv2.warpPerspective(ones, m32, out, INTER_NEAREST, CONSTANT 0) > 0.5) as bool (A.3)."""
    w, h = int(src_size[0]), int(src_size[1])
    out_w, out_h = int(out_size[0]), int(out_size[1])
    mi = invert3(m32)
    X, Y, W = _homog(mi, out_w, out_h)
    with np.errstate(divide="ignore", invalid="ignore"):
        sx = X / W
        sy = Y / W
    if rule == RULE_C:
        with np.errstate(invalid="ignore"):
            sx = np.rint(sx)
            sy = np.rint(sy)
    with np.errstate(invalid="ignore"):
        ok = (sx >= 0.0) & (sx <= w - 1.0) & (sy >= 0.0) & (sy <= h - 1.0)
    return ok


def _clip_halfplane(poly, a, b, c):
    """Sutherland-Hodgman: the part of the polygon with a x + b y + c >= 0."""
    out = []
    for i, p in enumerate(poly):
        q = poly[(i + 1) % len(poly)]
        fp, fq = a * p[0] + b * p[1] + c, a * q[0] + b * q[1] + c
        if fp >= 0:
            out.append(p)
        if (fp >= 0) != (fq >= 0):
            t = fp / (fp - fq)
            out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


def mask_stripes(out_size, threads: int):
    """Row ranges [r0, r1) of the destination stripes the wheel's IPP path of warpPerspective(INTER_NEAREST) works in:
    min(cv2.getNumThreads(), ceil(W' H' / 2^14)) stripes, boundaries (s H' + S // 2) // S.  Found black-box (threshold of
    the Rule P -> Rule C flip under vertical shifts, 11 sizes x 7 thread counts)."""
    ow, oh = int(out_size[0]), int(out_size[1])
    s = max(1, min(int(threads), -(-(ow * oh) // 16384)))
    return [((k * oh + s // 2) // s, ((k + 1) * oh + s // 2) // s) for k in range(s)]


def auto_rule(m32, src_size, out_size, threads: int) -> int:
    """The rule cv2 applies to ONE warpPerspective(ones, m32, out_size, INTER_NEAREST) call: Rule P, unless
    the forward-mapped quad of the source's pixel centres misses one destination stripe entirely -- then the IPP call is
    refused and OpenCV's own code (Rule C) handles the whole image."""
    w, h = int(src_size[0]), int(src_size[1])
    ow = int(out_size[0])
    m = np.asarray(m32, dtype=np.float32).astype(np.float64).reshape(3, 3)
    quad = []
    for x, y in ((0.0, 0.0), (w - 1.0, 0.0), (w - 1.0, h - 1.0), (0.0, h - 1.0)):
        X, Y, W = (m[0, 0] * x + m[0, 1] * y + m[0, 2], m[1, 0] * x + m[1, 1] * y + m[1, 2], m[2, 0] * x + m[2, 1] * y + m[2, 2])
        with np.errstate(divide="ignore", invalid="ignore"):
            quad.append((float(np.float64(X) / np.float64(W)), float(np.float64(Y) / np.float64(W))))
    for r0, r1 in mask_stripes(out_size, threads):
        if r1 <= r0:
            continue
        poly = quad
        for a, b, c in ((1.0, 0.0, 0.0), (-1.0, 0.0, ow - 1.0), (0.0, 1.0, -float(r0)), (0.0, -1.0, float(r1 - 1))):
            poly = _clip_halfplane(poly, a, b, c)
            if not poly:
                return RULE_C
    return RULE_P


def mask_np(m32, src_size, out_size, rule: int = RULE_P) -> np.ndarray:
    """Padding mask exactly as the reference post-processes it (flow.py:583-584)."""
    mask = np.float32(1.0) - coverage_np(m32, src_size, out_size, rule).astype(np.float32)
    mask[mask < 1e-3] = 0.0
    return mask


def blur_sample_matrices(matrices, idx: int, motion_blur: float, sample_count: int):
    """motion_apply.py:125-134 -- f64 linear interpolation towards the next matrix."""
    n = len(matrices)
    if n <= 1:
        return [np.asarray(matrices[idx], dtype=np.float64)]
    base = np.asarray(matrices[idx], dtype=np.float64)
    if idx < n - 1:
        delta = np.asarray(matrices[idx + 1], dtype=np.float64) - base
    else:
        delta = base - np.asarray(matrices[idx - 1], dtype=np.float64)
    ts = np.linspace(0.0, float(motion_blur), int(sample_count), dtype=np.float64)
    return [base + delta * t for t in ts]


def warp_blur_np(src, matrices, idx, out_size, interp, border, motion_blur, samples, rule=RULE_P):
    """motion_apply.py:137-202 for one frame: f32 accumulate in sample order, /S, soft mask."""
    s = int(np.clip(samples, 3, 33))
    src = np.asarray(src, dtype=np.float32)
    out_w, out_h = int(out_size[0]), int(out_size[1])
    acc = np.zeros((out_h, out_w, src.shape[2]), dtype=np.float32)
    cov = np.zeros((out_h, out_w), dtype=np.float32)
    for m in blur_sample_matrices(matrices, idx, motion_blur, s):
        m32 = np.asarray(m, dtype=np.float32)
        acc += warp_np(src, m32, out_size, interp, border)
        cov += coverage_np(m32, (src.shape[1], src.shape[0]), out_size, rule).astype(np.float32)
    frame = acc / np.float32(s)
    mask = np.float32(1.0) - cov / np.float32(s)
    mask[mask < 1e-3] = 0.0
    return frame.astype(np.float32), mask.astype(np.float32)
