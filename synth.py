"""Seeded synthetic 'jittered video' used by tests and bench.py.

Not product code and not oracle code: a data generator.  A smooth random base texture is
rendered through per-frame camera-shake matrices; frame i is a pure function of (seed, i) so a
frame-range shard can regenerate exactly its own frames (plus the one-frame halo) on its GPU.
"""
from __future__ import annotations

import numpy as np
import torch

MARGIN = 64


def base_texture(seed: int, width: int, height: int) -> torch.Tensor:
    """[(H+128),(W+128),3] float32 in 0..1: low-res noise upsampled bicubically (CPU torch)."""
    rng = np.random.default_rng(seed)
    low = rng.random((height // 8 + 10, width // 8 + 10, 3), dtype=np.float32)
    t = torch.from_numpy(low).permute(2, 0, 1)[None]
    up = torch.nn.functional.interpolate(
        t, size=(height + 2 * MARGIN, width + 2 * MARGIN), mode="bicubic", align_corners=False
    )
    return up[0].permute(1, 2, 0).clamp_(0.0, 1.0).contiguous()


def shake_matrices(n: int, seed: int, width: int, height: int, *, perspective: bool = False,
                   amount: float = 1.0) -> np.ndarray:
    """[n,3,3] float64 camera-shake matrices (frame 0 = identity): a handheld-like sum of a few
    low-frequency sinusoids plus a little white jitter, about the frame centre."""
    rng = np.random.default_rng(seed + 7919)
    t = np.arange(n, dtype=np.float64) / 16.0
    scale_px = amount * max(width, height) / 1920.0

    def wobble(amp, k=3):
        out = np.zeros(n)
        for _ in range(k):
            f = rng.uniform(0.3, 2.5)
            out += amp / k * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
        out += rng.normal(0.0, amp * 0.15, n)
        return out - out[0]

    tx, ty = wobble(9.0 * scale_px), wobble(7.0 * scale_px)
    rot = wobble(np.deg2rad(0.35) * amount)
    zoom = wobble(0.004 * amount)
    cx, cy = width * 0.5, height * 0.5
    mats = np.zeros((n, 3, 3))
    for i in range(n):
        s = 1.0 + zoom[i]
        c, sn = s * np.cos(rot[i]), s * np.sin(rot[i])
        m = np.array([[c, -sn, cx - c * cx + sn * cy + tx[i]], [sn, c, cy - sn * cx - c * cy + ty[i]], [0, 0, 1.0]])
        if perspective:
            m[2, 0], m[2, 1] = rng.normal(0.0, 2e-5 * 1920.0 / max(width, height), 2) * (i > 0)
        mats[i] = m
    return mats


def render_matrices(mats: np.ndarray) -> np.ndarray:
    """Matrices that map the (margin-padded) base texture to frame i: M_i @ T(-margin,-margin)."""
    shift = np.array([[1.0, 0.0, -MARGIN], [0.0, 1.0, -MARGIN], [0.0, 0.0, 1.0]])
    return np.stack([m @ shift for m in mats], axis=0).astype(np.float32)


def render_clip_cuda(handle, base_dev: torch.Tensor, mats: np.ndarray, width: int, height: int,
                     start: int = 0, stop: int | None = None) -> torch.Tensor:
    """Frames [start, stop) rendered on the GPU with the fused resampler -> [n,H,W,3] float32."""
    stop = len(mats) if stop is None else stop
    fwd = torch.from_numpy(render_matrices(mats[start:stop]).reshape(-1, 1, 9)).to(base_dev.device)
    n = stop - start
    out = torch.empty((n, height, width, 3), dtype=torch.float32, device=base_dev.device)
    src = base_dev[None]
    for i in range(n):  # one source image, n matrices: launch per frame against the same base
        handle.warp_fused(src, fwd[i : i + 1], (width, height), "bilinear", (0.5, 0.5, 0.5),
                          want_mask=False, out=out[i : i + 1])
    return out


def render_clip_numpy(base: np.ndarray, mats: np.ndarray, width: int, height: int) -> np.ndarray:
    """CPU twin of render_clip_cuda built on the numpy oracle (small clips only)."""
    from oracle.resample_np import warp_np

    fwd = render_matrices(mats)
    return np.stack([warp_np(base, fwd[i], (width, height), "bilinear", (0.5, 0.5, 0.5)) for i in range(len(fwd))])
