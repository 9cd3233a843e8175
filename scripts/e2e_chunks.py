"""Does the size of the upload's cudaMemcpyAsync calls change how often a node call stalls while enqueueing them?
Interleaved A/B on one box: every repetition runs the node call once per setting (development aid)."""
import gc, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from baseline import refload
refload.install_stubs()
from vstab_b200 import _native, nodes, pipeline

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, 121
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), synth.shake_matrices(N, 0, W, H), W, H)
pinned = torch.empty(clip.shape, dtype=torch.float32, pin_memory=True); pinned.copy_(clip); torch.cuda.synchronize()
del clip
settings = {"64MB": 64 << 20, "256MB": 256 << 20, "768MB": 768 << 20, "whole clip": 4 << 30}
def node():
    return nodes.VideoStabilizerFlow.execute(pinned, 16.0, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, "#7F7F7F")
for v in settings.values():
    pipeline.UPLOAD_CHUNK_BYTES = v
    node()
gc.collect(); gc.freeze()
out = {k: [] for k in settings}
for rep in range(int(os.environ.get("REPS", "16"))):
    for k, v in settings.items():
        pipeline.UPLOAD_CHUNK_BYTES = v
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = node()
        torch.cuda.synchronize(); out[k].append(round((time.perf_counter() - t0) * 1e3, 1))
        del r
import statistics
print(json.dumps({k: {"median": statistics.median(v), "mean": round(statistics.mean(v), 1), "max": max(v), "all": v} for k, v in out.items()}, indent=1))
