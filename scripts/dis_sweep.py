"""A/B sweep of the DIS scheduling knobs (env read per call by dis.cu) on the bench clip's 120 pairs.
Prints one JSON line per configuration: ms per call (CUDA events, 5 calls) and the SHA of the sampled flow grid
(every configuration must print the same SHA: the knobs only change scheduling)."""
import os, sys, json, hashlib, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, int(os.environ.get("SWEEP_FRAMES", "121"))
mats = synth.shake_matrices(N, 0, W, H)
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), mats, W, H)
gray = h.gray_working(clip, (960, 540))
del clip
KNOBS = ("VSTAB_PS_NOPACK", "VSTAB_PS_WPC", "VSTAB_DIS_GROUPS", "VSTAB_VR_CLUSTER", "VSTAB_DIS_STAGGER", "VSTAB_PS_SPW", "VSTAB_VR_RESIDENT",
         "VSTAB_VR_RESIDENT_THREADS", "VSTAB_DIS_PREP_ASYNC", "VSTAB_VR_ONCHIP")

def run(cfg, reps=5):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        os.environ[k] = str(v)
    for _ in range(2):
        h.dis_flow(gray, want_flow=False, grid_step=8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _, grid = h.dis_flow(gray, want_flow=False, grid_step=8)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"lib": os.path.basename(_native.LIB_PATH), **cfg, "ms": round(e0.elapsed_time(e1) / reps, 4),
                      "sha": hashlib.sha1(grid.cpu().numpy().tobytes()).hexdigest()[:12]}), flush=True)

configs = json.loads(os.environ["SWEEP_CONFIGS"]) if "SWEEP_CONFIGS" in os.environ else [{}]
for c in configs:
    run(c)
