"""Phase timing of one device-resident Flow step (development aid)."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native, flow, pipeline, hostmath as hm, stabilizer_core as core

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
W, H, N = 1920, 1080, 121
mats = synth.shake_matrices(N, 0, W, H)
clip = synth.render_clip_cuda(h, synth.base_texture(0, W, H).to(dev), mats, W, H)
torch.cuda.synchronize()
ctx = pipeline.VideoContext(clip, pipeline.FrameAdapter(np.float32, False, "0_1", "torch", False), W, H, 3, None, "sequence", {})

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r

out = {}
out["gray_ms"], gray = t(lambda: h.gray_working(clip, (960, 540)))
out["dis_ms"], (_, grid) = t(lambda: h.dis_flow(gray, want_flow=False, grid_step=8))
out["fit_sim+trans_ms"], raw = t(lambda: h.fit_grid(grid, 8, 3))
out["fit_sim_only_ms"], _ = t(lambda: h.fit_grid(grid, 8, 2))
out["fit_all_ms"], _ = t(lambda: h.fit_grid(grid, 8, 7))
out["decode_ms"], d = t(lambda: _native.decode_fit_results(raw))
out["estimate_total_ms"], cands = t(lambda: flow.estimate_candidates(ctx, 960, 540, "similarity").to_host())
t0 = time.perf_counter()
for _ in range(5):
    chosen, active, _ = core.replay_mode_ladder(cands, "similarity", with_residual=True)
out["ladder_ms"] = (time.perf_counter() - t0) / 5 * 1e3
out["full_step_ms"], res = t(lambda: flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device"))
fwd = torch.eye(3, device=dev).reshape(1, 1, 9).repeat(N, 1, 1).contiguous()
out["warp_only_ms"], _ = t(lambda: h.warp_fused(clip, fwd, (W, H), "bilinear", (0.5, 0.5, 0.5), want_pad_count=True))
print(json.dumps(out, indent=1))
core.PHASE_LOG = []
flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device")
print("PHASES", [(a, round(b * 1e3, 3)) for a, b in core.PHASE_LOG]); core.PHASE_LOG = None
import cProfile, pstats, io
dst = torch.empty((N, H, W, 3), device=dev); msk = torch.empty((N, H, W), device=dev)
out2 = {}
out2["warp_prealloc_identity_ms"], _ = t(lambda: h.warp_fused(clip, fwd, (W, H), "bilinear", (0.5, 0.5, 0.5), out=dst, mask_out=msk))
jit = torch.from_numpy(np.stack([np.array([[1.001, -0.002, 3.3], [0.002, 0.999, -2.1], [0, 0, 1]], np.float32).reshape(9)] * N)).to(dev).reshape(N, 1, 9).contiguous()
out2["warp_prealloc_jitter_ms"], _ = t(lambda: h.warp_fused(clip, jit, (W, H), "bilinear", (0.5, 0.5, 0.5), out=dst, mask_out=msk))
out2["warp_prealloc_jitter_padcount_ms"], _ = t(lambda: h.warp_fused(clip, jit, (W, H), "bilinear", (0.5, 0.5, 0.5), out=dst, mask_out=msk, want_pad_count=True))
out2["warp_32frames_ms"], _ = t(lambda: h.warp_fused(clip[:32], jit[:32], (W, H), "bilinear", (0.5, 0.5, 0.5), out=dst[:32], mask_out=msk[:32]))
print(json.dumps(out2, indent=1))
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    res = flow.stabilize_frames(ctx, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device")
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
