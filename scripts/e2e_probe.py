"""Host-to-host timing of the Flow path (resident vs streamed), development aid."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native, flow, pipeline

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, 121
mats = synth.shake_matrices(N, 0, W, H)
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), mats, W, H)
host = torch.empty(clip.shape, dtype=torch.float32, pin_memory=True); host.copy_(clip); torch.cuda.synchronize()
del clip
for limit in (None, "0", None, "0"):
    if limit is None: os.environ.pop("VSTAB_RESIDENT_LIMIT_MB", None)
    else: os.environ["VSTAB_RESIDENT_LIMIT_MB"] = limit
    free, total = torch.cuda.mem_get_info(dev)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx = pipeline.normalize_video_input(host, dev)
        res = flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="host")
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("limit", limit, "streamed", ctx.streamed, "free_GB", round(free / 2**30, 1), "ms", [round(t, 1) for t in ts])
