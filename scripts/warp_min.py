"""Single launches of the fused resampler at bench shapes (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader
vstab_loader.load()
from vstab_b200 import _native

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
for (w, hh, n, ow, oh) in [(256, 128, 2, 256, 128), (1920, 1080, 121, 1920, 1080), (1920, 1080, 40, 1800, 1012)]:
    src = torch.rand((n, hh, w, 3), device=dev)
    fwd = torch.tensor([[1, 0, 1.5, 0, 1, -0.75, 0, 0, 1]] * n, dtype=torch.float32, device=dev).reshape(n, 1, 9)
    try:
        dst, mask, pad = h.warp_fused(src, fwd, (ow, oh), "bilinear", (0.5, 0.5, 0.5), want_pad_count=True)
        torch.cuda.synchronize()
        print("ok", (w, hh, n, ow, oh), float(dst.mean()), float(mask.mean()), int(pad.sum()))
    except Exception as e:
        print("FAIL", (w, hh, n, ow, oh), e)
    del src
