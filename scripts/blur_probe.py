"""BASELINE config 4 shape on a few frames (Motion Apply, 33 shutter samples, 1080p -> 1930x1088):
ms per frame of the multi-sample resampler with CUDA events, plus a checksum of the output so two
builds can be compared bit for bit.  `python scripts/blur_probe.py once` runs one launch per
interpolation (what an ncu capture wants)."""
import hashlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native
from vstab_b200.motion_apply import sample_matrices

once = len(sys.argv) > 1 and sys.argv[1] == "once"
dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
w, hh, n = 1920, 1080, (2 if once else 12)
ow, oh = w + 10, hh + 8
mats = synth.shake_matrices(n, 0, w, hh)
src = synth.render_clip_cuda(h, synth.base_texture(0, w, hh).to(dev), mats, w, hh)
shift = np.array([[1, 0, 5], [0, 1, 4], [0, 0, 1]], dtype=np.float64)
fwd_mats = [shift @ m for m in mats.astype(np.float64)]
out = {}
for interp in ("bicubic", "bilinear"):
    for s in ((33,) if once else (33, 9, 3)):
        fwd = torch.from_numpy(sample_matrices(fwd_mats, 0.5, s)).to(dev)
        dst = torch.empty((n, oh, ow, 3), device=dev)
        mask = torch.empty((n, oh, ow), device=dev)
        reps = 1 if once else 3
        if not once:
            h.warp_fused(src, fwd, (ow, oh), interp, (0.5, 0.5, 0.5), out=dst, mask_out=mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _, _, pad = h.warp_fused(src, fwd, (ow, oh), interp, (0.5, 0.5, 0.5), out=dst, mask_out=mask, want_pad_count=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        digest = hashlib.sha256(dst.cpu().numpy().tobytes() + mask.cpu().numpy().tobytes() + pad.cpu().numpy().tobytes()).hexdigest()[:16]
        taps = (4 if interp == "bilinear" else 16) * s
        out[f"{interp}_S{s}"] = {"ms_per_frame": round(ms / n, 4), "fps": round(n / ms * 1e3, 1),
                                 "Gtaps_per_s": round(taps * ow * oh * n / ms / 1e6, 1), "sha": digest}
print(json.dumps(out, indent=1))
