"""Per-repetition host-side breakdown of the node call with a page-locked IMAGE (development aid): which phase takes the
extra time on the slow repetitions."""
import gc, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from baseline import refload
refload.install_stubs()
from vstab_b200 import _native, flow, nodes, pipeline, stabilizer_core as core

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, 121
mats = synth.shake_matrices(N, 0, W, H)
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), mats, W, H)
pinned = torch.empty(clip.shape, dtype=torch.float32, pin_memory=True); pinned.copy_(clip); torch.cuda.synchronize()
keep_clip = os.environ.get("REPS_KEEP_CLIP", "1") == "1"
if not keep_clip:
    del clip
REPS = int(os.environ.get("REPS", "10"))

def node(image):
    return nodes.VideoStabilizerFlow.execute(image, 16.0, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, "#7F7F7F")

def driver(image):
    ctx = pipeline.normalize_video_input(image, dev)
    res = flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="host")
    return pipeline.reconstruct_video(res.frames, ctx), pipeline.convert_masks_for_output(res.masks)

orig_norm = pipeline.normalize_video_input
marks = {}
def timed_norm(*a, **k):
    t0 = time.perf_counter(); r = orig_norm(*a, **k); marks["normalize"] = time.perf_counter() - t0; return r
pipeline.normalize_video_input = timed_norm
nodes.normalize_video_input = timed_norm
rows = []
for name, fn in (("node", node), ("driver", driver), ("node", node), ("driver", driver)):
    fn(pinned); gc.collect(); gc.freeze()
    for r in range(REPS):
        core.PHASE_LOG = []; marks.clear()
        free0 = torch.cuda.mem_get_info(dev)[0]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(pinned)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
        del out
        rows.append({"path": name, "rep": r, "wall": round(wall, 1), "normalize": round(marks.get("normalize", 0) * 1e3, 1),
                     "phases": [round(b * 1e3, 1) for a, b in core.PHASE_LOG], "reserved_GB": round(torch.cuda.memory_reserved(dev) / 2**30, 2),
                     "free_GB": round(free0 / 2**30, 1)})
        core.PHASE_LOG = None
for r in rows:
    print(json.dumps(r))
