"""BASELINE config 5 shape on one GPU's share: Flow DIS, perspective, camera_lock on N 3840x2160 frames,
device-resident (profiling aid: `ncu --metrics gpu__time_duration.sum` launch list of one call)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native, flow, pipeline, stabilizer_core as core

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
n, w, hh = int(sys.argv[2]) if len(sys.argv) > 2 else 48, 3840, 2160
mats = synth.shake_matrices(n, 0, w, hh, perspective=True)
clip = synth.render_clip_cuda(h, synth.base_texture(0, w, hh).to(dev), mats, w, hh)
ctx = pipeline.VideoContext(clip, pipeline.FrameAdapter(np.float32, False, "0_1", "torch", False), w, hh, 3, None, "sequence", {})
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    core.PHASE_LOG = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = flow.stabilize_frames(ctx, "crop_and_pad", "perspective", True, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device")
    torch.cuda.synchronize()
    print("ms", round((time.perf_counter() - t0) * 1e3, 2), "mode", r.meta["transform_mode_applied"], [(a[:24], round(b * 1e3, 2)) for a, b in core.PHASE_LOG])
