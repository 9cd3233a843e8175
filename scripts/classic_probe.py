"""One Classic (GFTT + LK) stabilize of 241 x 1280x720 frames, device-resident (profiling aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native, classic, pipeline, stabilizer_core as core

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
n, w, hh = 241, 1280, 720
mats = synth.shake_matrices(n, 0, w, hh)
clip = synth.render_clip_cuda(h, synth.base_texture(0, w, hh).to(dev), mats, w, hh)
ctx = pipeline.VideoContext(clip, pipeline.FrameAdapter(np.float32, False, "0_1", "torch", False), w, hh, 3, None, "sequence", {})
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    core.PHASE_LOG = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    classic.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device")
    torch.cuda.synchronize()
    print("ms", round((time.perf_counter() - t0) * 1e3, 2), [(a[:24], round(b * 1e3, 2)) for a, b in core.PHASE_LOG])
