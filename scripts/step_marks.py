"""GPU-side view of one device-resident Flow step on the bench clip: CUDA events recorded on the launching stream at the
phase boundaries (stabilizer_core.GPU_MARKS) and the host-side phase log of the same step.  elapsed(start -> estimation
kernels) is gray + DIS + fit on the GPU, the next interval is what the GPU idles while the host copies the table and
solves the trajectory, the last one the resampler.  Development aid."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native, flow, pipeline, stabilizer_core as core

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, 121
mats = synth.shake_matrices(N, 0, W, H)
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), mats, W, H)
ctx = pipeline.VideoContext(clip, pipeline.FrameAdapter(np.float32, False, "0_1", "torch", False), W, H, 3, None, "sequence", {})
run = lambda: flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device")
for _ in range(5):
    run()
torch.cuda.synchronize()
rows = []
for _ in range(5):
    core.GPU_MARKS, core.PHASE_LOG = [], []
    t0 = time.perf_counter()
    run()
    end = torch.cuda.Event(enable_timing=True); end.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    marks = core.GPU_MARKS + [("end", end)]
    gpu = {f"{a[0]} -> {b[0]}": round(a[1].elapsed_time(b[1]), 3) for a, b in zip(marks[:-1], marks[1:])}
    rows.append({"wall_ms": round(wall, 3), "gpu_ms": gpu, "host_ms": {a[:24]: round(b * 1e3, 3) for a, b in core.PHASE_LOG}})
core.GPU_MARKS = core.PHASE_LOG = None
print(json.dumps(rows[-3:], indent=1))
