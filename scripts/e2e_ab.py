"""A/B of the host-result path on ONE box: padding mask over the link as bytes (VSTAB_MASK_BYTES=1, default) or as float32 (=0),
through the driver (flow.stabilize_frames, pinned clip), the node (nodes.VideoStabilizerFlow.execute, pinned IMAGE) and the node
with a pageable IMAGE; alternating, wall-clock per call with a device synchronize on both sides.  Also times the host's
uint8 -> float32 widening alone and prints the host-side phase log of one call per setting.  Development aid."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from baseline import refload
refload.install_stubs()
from vstab_b200 import _native, flow, nodes, pipeline, stabilizer_core as core

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, int(os.environ.get("AB_FRAMES", "121"))
REPS = int(os.environ.get("AB_REPS", "3"))
mats = synth.shake_matrices(N, 0, W, H)
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), mats, W, H)
pinned = torch.empty(clip.shape, dtype=torch.float32, pin_memory=True); pinned.copy_(clip); torch.cuda.synchronize()
pageable = torch.empty(clip.shape, dtype=torch.float32); pageable.copy_(pinned)
del clip
out = {"threads": torch.get_num_threads(), "cpus": len(os.sched_getaffinity(0))}

u8 = torch.randint(0, 2, (N, H, W), dtype=torch.uint8).pin_memory()
f32 = torch.empty((N, H, W), dtype=torch.float32, pin_memory=True)
f32.copy_(u8)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); f32.copy_(u8); ts.append((time.perf_counter() - t0) * 1e3)
out["widen_250M_ms"] = [round(t, 2) for t in ts]
del u8, f32

def driver(image):
    ctx = pipeline.normalize_video_input(image, dev)
    res = flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="host")
    return pipeline.reconstruct_video(res.frames, ctx), pipeline.convert_masks_for_output(res.masks)

def node(image):
    return nodes.VideoStabilizerFlow.execute(image, 16.0, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, "#7F7F7F")

def wall(fn, image):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = fn(image)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3

import gc
for name, fn, image in (("driver_pinned", driver, pinned), ("node_pinned", node, pinned), ("node_pageable", node, pageable)):
    for flag in ("1", "0"):
        os.environ["VSTAB_MASK_BYTES"] = flag
        wall(fn, image)
    gc.collect(); gc.freeze()
    res = {"1": [], "0": []}
    for _ in range(REPS):
        for flag in ("1", "0"):
            os.environ["VSTAB_MASK_BYTES"] = flag
            res[flag].append(round(wall(fn, image), 1))
    out[name] = {"bytes": res["1"], "float32": res["0"]}
    for flag in ("1", "0"):
        os.environ["VSTAB_MASK_BYTES"] = flag
        core.PHASE_LOG = []
        t = wall(fn, image)
        out[name]["phases_" + ("bytes" if flag == "1" else "float32")] = [(a[:28], round(b * 1e3, 2)) for a, b in core.PHASE_LOG] + [("wall", round(t, 1))]
        core.PHASE_LOG = None
print(json.dumps(out, indent=1))
