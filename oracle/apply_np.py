"""numpy restatement of the reference's Motion Apply engine.  TEST INFRASTRUCTURE ONLY.

Follows nodes/motion_apply.py of the reference: apply_motion :297-429 (framing dispatch),
_warp_with_matrices :75-122, _warp_with_motion_blur :137-202, _common_valid_mask :205-228,
_center_crop_matrix_from_common :231-285, _expand_matrices :288-294, with the cv2 calls replaced
by oracle.resample_np.  Pinned against tests/golden/apply_*.npz (outputs of the real reference).
"""
from __future__ import annotations

import math

import numpy as np

from . import resample_np as R


def _bboxes(matrices, w, h):
    corners = np.array([[0.0, 0.0, 1.0], [w, 0.0, 1.0], [0.0, h, 1.0], [w, h, 1.0]], dtype=np.float64).T
    lo, hi = [], []
    for m in matrices:
        q = m @ corners
        q = q / q[2]
        lo.append([q[0].min(), q[1].min()])
        hi.append([q[0].max(), q[1].max()])
    return np.array(lo), np.array(hi)


def _expand(matrices, w, h):
    lo, hi = _bboxes(matrices, w, h)
    x0, y0 = float(lo[:, 0].min()), float(lo[:, 1].min())
    x1, y1 = float(hi[:, 0].max()), float(hi[:, 1].max())
    t = np.array([[1.0, 0.0, -x0], [0.0, 1.0, -y0], [0.0, 0.0, 1.0]], dtype=np.float32)
    size = (max(int(math.ceil(x1 - x0)), 1), max(int(math.ceil(y1 - y0)), 1))
    return [t @ m for m in matrices], size


def _center_crop(common, out_size):
    ow, oh = out_size
    cx, cy = (ow - 1) * 0.5, (oh - 1) * 0.5
    asp = ow / float(oh)

    def dims(s):
        cw = max(1.0, ow / s)
        ch = cw / asp
        if ch > oh:
            ch = oh / s
            cw = ch * asp
        return cw, ch

    def fits(s):
        cw, ch = dims(s)
        x0 = int(np.ceil(cx - cw * 0.5)); y0 = int(np.ceil(cy - ch * 0.5))
        x1 = int(np.floor(cx + cw * 0.5)); y1 = int(np.floor(cy + ch * 0.5))
        if x0 < 0 or y0 < 0 or x1 >= ow or y1 >= oh or x1 <= x0 or y1 <= y0:
            return False
        return bool(common[y0:y1 + 1, x0:x1 + 1].all())

    lo, hi = 0.0, 1.0
    if not fits(1.0):
        while hi <= 4.0 and not fits(hi):
            hi *= 1.25
        if hi > 4.0:
            return None
    for _ in range(32):
        mid = (lo + hi) * 0.5
        if mid < 1.0:
            mid = 1.0
        if fits(mid):
            hi = mid
        else:
            lo = mid
    s = float(hi)
    cw = ow / s
    ch = cw / asp
    if ch > oh:
        ch = oh / s
        cw = ch * asp
    return np.array([[s, 0.0, -s * (cx - cw * 0.5)], [0.0, s, -s * (cy - ch * 0.5)], [0.0, 0.0, 1.0]], dtype=np.float64)


def apply_motion_np(frames, meta, padding_rgb, framing="crop_and_pad", interp="bilinear", blur=0.0, samples=9,
                    rule=R.RULE_P):
    """frames [N,H,W,3] f32; meta = {"motion_meta": v2 block}.  Returns (frames', masks'[N,H',W',1], out_size, framing)."""
    block = meta["motion_meta"]
    mats = [np.asarray(e["matrix"], dtype=np.float64) for e in block["per_frame"]]
    w, h = block["input_size"]
    out_size = tuple(block["output_size"])
    n = len(mats)
    blur = float(np.clip(blur, 0.0, 1.0))
    samples = int(np.clip(samples, 3, 33))
    border = (np.array(padding_rgb, dtype=np.float32) / 255.0).tolist()
    masks_zero = False
    effective = "crop_and_pad" if framing == "pad" else framing
    if effective == "crop":
        common = np.ones((out_size[1], out_size[0]), dtype=bool)
        for m in mats:
            common &= R.coverage_np(np.asarray(m, dtype=np.float32), (w, h), out_size, rule)
        crop = _center_crop(common, out_size)
        if crop is None:
            effective = "crop_and_pad"
        else:
            mats = [crop @ m for m in mats]
            masks_zero = True
    elif effective == "expand":
        mats, out_size = _expand(mats, w, h)
    out_f = np.empty((n, out_size[1], out_size[0], 3), np.float32)
    out_m = np.zeros((n, out_size[1], out_size[0], 1), np.float32)
    for i in range(n):
        if blur <= 0.0:
            m32 = np.asarray(mats[i], dtype=np.float32)
            out_f[i] = R.warp_np(frames[i], m32, out_size, interp, border)
            if not masks_zero:
                out_m[i, ..., 0] = R.mask_np(m32, (w, h), out_size, rule)
        else:
            if n <= 1:  # one sample, still divided by the nominal count (reference behaviour)
                m32 = np.asarray(mats[i], dtype=np.float32)
                out_f[i] = R.warp_np(frames[i], m32, out_size, interp, border) / np.float32(samples)
                if not masks_zero:
                    cov = R.coverage_np(m32, (w, h), out_size, rule).astype(np.float32) / np.float32(samples)
                    mk = np.float32(1.0) - cov
                    mk[mk < 1e-3] = 0.0
                    out_m[i, ..., 0] = mk
            else:
                f, mk = R.warp_blur_np(frames[i], mats, i, out_size, interp, border, blur, samples, rule)
                out_f[i] = f
                if not masks_zero:
                    out_m[i, ..., 0] = mk
    return out_f, out_m, out_size, effective
