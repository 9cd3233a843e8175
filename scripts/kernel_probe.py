"""Quick device-resident timing of the fused resampler (development aid, not the bench)."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader

vstab_loader.load()
from vstab_b200 import _native


def main():
    dev = torch.device("cuda", 0)
    h = _native.get_handle(dev)
    w, hh, n = 1920, 1080, 32
    src = torch.rand((n, hh, w, 3), device=dev)
    rng = np.random.default_rng(0)
    mats = []
    for i in range(n):
        th, s = rng.normal(0, 0.004), 1 + rng.normal(0, 0.003)
        mats.append([s * np.cos(th), -s * np.sin(th), rng.normal(0, 8), s * np.sin(th), s * np.cos(th), rng.normal(0, 6), 0, 0, 1])
    fwd = torch.tensor(mats, dtype=torch.float32, device=dev).reshape(n, 1, 9)
    dst = torch.empty((n, hh, w, 3), device=dev)
    mask = torch.empty((n, hh, w), device=dev)
    bytes_per_frame = 12 * hh * w + 12 * hh * w + 4 * hh * w
    out = {}
    for interp in ("bilinear", "bicubic"):
        for stage in (0, 1):
            for _ in range(3):
                h.warp_fused(src, fwd, (w, hh), interp, (0.5, 0.5, 0.5), stage_mode=stage, out=dst, mask_out=mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                h.warp_fused(src, fwd, (w, hh), interp, (0.5, 0.5, 0.5), stage_mode=stage, out=dst, mask_out=mask)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[f"{interp}_stage{stage}"] = {"ms_per_launch": ms, "fps": n / ms * 1e3, "GBps": bytes_per_frame * n / ms / 1e6}
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        g = h.gray_working(src)
    t0.record(); g = h.gray_working(src); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    out["gray_x2"] = {"ms": ms, "GBps": 12 * hh * w * n / ms / 1e6}
    print(json.dumps(out, indent=1))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def blur_probe():
    """BASELINE config 4 shape: bicubic, 33 shutter samples, 1080p."""
    import vstab_loader
    from vstab_b200.motion_apply import sample_matrices
    import synth

    dev = torch.device("cuda", 0)
    h = _native.get_handle(dev)
    w, hh, n = 1920, 1080, 16
    src = torch.rand((n, hh, w, 3), device=dev)
    mats = synth.shake_matrices(n, 0, w, hh)
    out = {}
    for interp in ("bilinear", "bicubic"):
        for s in (5, 33):
            fwd = torch.from_numpy(sample_matrices(list(mats), 0.5, s)).to(dev)
            dst = torch.empty((n, hh + 8, w + 10, 3), device=dev)
            mask = torch.empty((n, hh + 8, w + 10), device=dev)
            for _ in range(2):
                h.warp_fused(src, fwd, (w + 10, hh + 8), interp, (0.5, 0.5, 0.5), out=dst, mask_out=mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                h.warp_fused(src, fwd, (w + 10, hh + 8), interp, (0.5, 0.5, 0.5), out=dst, mask_out=mask)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            taps = (4 if interp == "bilinear" else 16) * s
            out[f"{interp}_S{s}"] = {"ms_per_frame": ms / n, "fps": n / ms * 1e3, "Gtaps_per_s": taps * (w + 10) * (hh + 8) * n / ms / 1e6}
    print(json.dumps(out, indent=1))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "blur":
    blur_probe()
