"""Resampler time vs projective strength at 3840x2160 (the limiter of BASELINE config 5 at scale):
forward matrices [[1,0,0],[0,1,0],[g,g/2,1]] -- w runs from 1 at the origin to 1 + g (3840 + 1080) at the far corner."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
from vstab_b200 import _native

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 3840, 2160, 48
src = torch.rand((N, H, W, 3), device=dev)
dst = torch.empty((N, H, W, 3), device=dev); msk = torch.empty((N, H, W), device=dev)
out = []
for g in (0.0, 1e-5, 2e-5, 4e-5, 8e-5, 1.6e-4, 3.2e-4, -2e-5, -4e-5, -8e-5):
    m = np.tile(np.array([1, 0, 0, 0, 1, 0, g, g / 2, 1], np.float32), (N, 1, 1))
    fwd = torch.from_numpy(m).to(dev)
    for _ in range(2):
        h.warp_fused(src, fwd, (W, H), "bilinear", (0.5, 0.5, 0.5), out=dst, mask_out=msk, want_pad_count=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        _, _, pad = h.warp_fused(src, fwd, (W, H), "bilinear", (0.5, 0.5, 0.5), out=dst, mask_out=msk, want_pad_count=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out.append({"g": g, "w_far_corner": round(1 + g * (W + H / 2), 3), "ms_per_frame": round(ms / N, 4),
                "GBs": round(N * 232243200 / (ms * 1e-3) / 1e9, 1), "padded_frac": round(float(pad.sum()) / (N * W * H), 3)})
    print(json.dumps(out[-1]), flush=True)
