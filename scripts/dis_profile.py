"""DIS on the 120 pairs of the bench clip (121 x 1080p -> 960x540 working images), CUDA-event timed.
Profiling aid: the command `ncu --set full -k regex:'vr_fused_kernel|patch_search_kernel'` is run on."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, 121
mats = synth.shake_matrices(N, 0, W, H)
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), mats, W, H)
gray = h.gray_working(clip, (960, 540))
del clip
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
h.dis_flow(gray, want_flow=False, grid_step=8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    _, grid = h.dis_flow(gray, want_flow=False, grid_step=8)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"dis_120_pairs_ms": e0.elapsed_time(e1) / reps, "grid_sha": __import__("hashlib").sha1(grid.cpu().numpy().tobytes()).hexdigest()}))
