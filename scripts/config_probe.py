"""Device-resident frames/s of every BASELINE.json config on one B200, with the cv2 call sequence of the
reference (oracle/cv_path.py) timed on a bounded sample of the same clip on the host cores.

Measurement aid for profiles/README.md; bench.py stays the contract for config 2."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
import synth
from vstab_b200 import _native, classic, flow, motion_apply, pipeline
from oracle import cv_path

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)


def clip_of(n, w, hh, seed=0, perspective=False):
    mats = synth.shake_matrices(n, seed, w, hh, perspective=perspective)
    c = synth.render_clip_cuda(h, synth.base_texture(seed, w, hh).to(dev), mats, w, hh)
    torch.cuda.synchronize()
    return c, mats


def ctx_of(clip):
    n, hh, w, _ = clip.shape
    return pipeline.VideoContext(clip, pipeline.FrameAdapter(np.float32, False, "0_1", "torch", False), w, hh, 3, None, "sequence", {})


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def cpu_timed(fn):
    t0 = time.perf_counter(); fn(); return time.perf_counter() - t0


out = {}
# config 1: shake -> Motion Apply bilinear crop_and_pad, 81 x 832x480
n, w, hh = 81, 832, 480
clip, mats = clip_of(n, w, hh)
meta = {"motion_meta": {"version": 2, "source": "probe", "frame_count": n, "fps": 16.0, "input_size": [w, hh], "output_size": [w, hh],
                        "matrix_convention": "input_to_output", "per_frame": [{"index": i, "matrix": m.tolist()} for i, m in enumerate(mats.astype(np.float64))]}}
ctx = ctx_of(clip)
t = timed(lambda: motion_apply.apply_motion(ctx, meta, (127, 127, 127), framing_mode="crop_and_pad", interpolation="bilinear", output="device"))
host = clip.cpu().numpy()
tc = cpu_timed(lambda: cv_path.apply_motion(list(host), [m for m in mats.astype(np.float32)], (w, hh), (w, hh), (127, 127, 127), "crop_and_pad", "bilinear", 0.0, 3))
out["cfg1 motion_apply bilinear 81x832x480"] = {"gpu_fps": n / t, "cpu_fps": n / tc, "cpu_sample": f"{n} frames"}
del clip, host

# config 3: Classic, similarity, 241 x 1280x720
n, w, hh = 241, 1280, 720
clip, _ = clip_of(n, w, hh)
ctx = ctx_of(clip)
for mode in ("translation", "similarity"):
    t = timed(lambda: classic.stabilize_frames(ctx, "crop_and_pad", mode, False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device"))
    ns = 41
    host = clip[:ns].cpu().numpy()
    tc = cpu_timed(lambda: cv_path.stabilize(list(host), "classic", "crop_and_pad", mode, False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0))
    out[f"cfg3 classic {mode} 241x1280x720"] = {"gpu_fps": n / t, "cpu_fps": ns / tc, "cpu_sample": f"first {ns} frames"}
del clip, host

# config 4: Motion Apply bicubic expand, blur 0.5 Ultra (33 samples), 121 x 1920x1080
n, w, hh = 121, 1920, 1080
clip, mats = clip_of(n, w, hh)
meta = {"motion_meta": {"version": 2, "source": "probe", "frame_count": n, "fps": 16.0, "input_size": [w, hh], "output_size": [w, hh],
                        "matrix_convention": "input_to_output", "per_frame": [{"index": i, "matrix": m.tolist()} for i, m in enumerate(mats.astype(np.float64))]}}
ctx = ctx_of(clip)
t = timed(lambda: motion_apply.apply_motion(ctx, meta, (127, 127, 127), framing_mode="expand", interpolation="bicubic", motion_blur=0.5,
                                            motion_blur_samples=33, output="device"), reps=2)
ns = 4
host = clip[:ns].cpu().numpy()
tc = cpu_timed(lambda: cv_path.apply_motion(list(host), [m for m in mats[:ns].astype(np.float32)], (w, hh), (w, hh), (127, 127, 127), "expand", "bicubic", 0.5, 33))
out["cfg4 motion_apply bicubic expand blur33 121x1920x1080"] = {"gpu_fps": n / t, "cpu_fps": ns / tc, "cpu_sample": f"first {ns} frames"}
del clip, host

# config 5 shape: Flow DIS perspective camera_lock at 3840x2160 (48 frames on one GPU)
n, w, hh = 48, 3840, 2160
clip, _ = clip_of(n, w, hh, perspective=True)
ctx = ctx_of(clip)
t = timed(lambda: flow.stabilize_frames(ctx, "crop_and_pad", "perspective", True, 0.7, 0.5, 0.6, (127, 127, 127), 16.0, output="device"), reps=2)
ns = 6
host = clip[:ns].cpu().numpy()
tc = cpu_timed(lambda: cv_path.stabilize(list(host), "flow", "crop_and_pad", "perspective", True, 0.7, 0.5, 0.6, (127, 127, 127), 16.0))
out["cfg5 flow perspective camera_lock 48x3840x2160 (one GPU's shard)"] = {"gpu_fps": n / t, "cpu_fps": ns / tc, "cpu_sample": f"first {ns} frames"}
print(json.dumps(out, indent=1))
