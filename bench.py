#!/usr/bin/env python
"""bench.py -- frames/sec of 1080p Video Stabilizer Flow (DIS, similarity, crop_and_pad,
strength 0.7, smooth 0.5) on N B200s, with the fused-warp HBM roofline and the host-CPU reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 121 synthetic jittered 1920x1080 float32 frames per GPU.  With
N > 1 the clip is N x 121 frames sharded by contiguous frame range (one halo frame per rank, one
NCCL all-gather of the per-pair candidate table, one of the per-frame padded-pixel counts): weak
scaling.  One "step" = one full pass of the hot path over the (sharded) clip.

  value   frames/s with the clip already resident in HBM and the results left in HBM
  e2e     frames/s through the node-level API with HOST tensors: pinned-host -> HBM upload of
          the frames and HBM -> pinned-host download of frames + masks inside the timed region
  roofline  fused resampler (vstab_warp_fused): algorithmic bytes (12HW read + 12H'W' + 4H'W'
          written per frame) / CUDA-event duration of its launches inside the timed steps,
          against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's OpenCV path (oracle/cv_path.py: same cv2 calls per frame / pair as
          the reference's _stabilize_frames) timed once on the host cores over the same clip
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WIDTH, HEIGHT, FRAMES_PER_GPU = 1920, 1080, 121
PARAMS = dict(framing="crop_and_pad", mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6,
              padding_rgb=(127, 127, 127), fps=16.0)
WORKLOAD = "Video Stabilizer Flow DIS, similarity, crop_and_pad, strength=0.7 smooth=0.5, 121 synthetic jittered 1920x1080 f32 frames per GPU"
METRIC = "frames/sec 1080p Flow stabilize"
REFERENCE_SAMPLE_FRAMES = 41


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            }
            while not self._stop_evt.is_set():
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                self._stop_evt.wait(0.01)
        except Exception as exc:  # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def settle_gc():
    """After warm-up, both arms: collect once and move every surviving object (torch, numpy, cv2 module
    state) to the permanent generation, the way a long-running serving process does after start-up, so a
    full collection inside the timed steps only walks the objects of the steps themselves."""
    import gc

    gc.collect()
    gc.freeze()


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on the host cores (oracle/cv_path.py:
    the Python reference cannot travel to the GPU box; this restates its cv2 call sequence and is
    checked against the unmodified reference in tests/test_cv_path_vs_reference.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2

    import synth
    from oracle import cv_path

    n = REFERENCE_SAMPLE_FRAMES
    base = synth.base_texture(0, WIDTH, HEIGHT)
    mats = synth.shake_matrices(FRAMES_PER_GPU, 0, WIDTH, HEIGHT)[:n]
    fwd = synth.render_matrices(mats)
    frames = np.stack([cv2.warpPerspective(base.numpy(), fwd[i], (WIDTH, HEIGHT), flags=cv2.INTER_LINEAR) for i in range(n)])
    clip = torch.from_numpy(frames)

    def step():
        t0 = time.perf_counter()
        out_f, out_m, _ = cv_path.stabilize(clip, "flow", PARAMS["framing"], PARAMS["mode"], PARAMS["camera_lock"], PARAMS["strength"],
                                            PARAMS["smooth"], PARAMS["keep_fov"], PARAMS["padding_rgb"], PARAMS["fps"])
        torch.from_numpy(out_f), torch.from_numpy(out_m[..., 0])
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    settle_gc()
    times = [step() for _ in range(args.steps)]
    sec = float(np.mean(times))
    fps = n / sec
    sample = f"{n} of the 121 frames per step (same clip, same parameters), cv2 {cv2.__version__} with {cv2.getNumThreads()} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": host_cores(), "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist

    import synth
    import vstab_loader

    vstab_loader.load()
    from vstab_b200 import _native, flow, pipeline, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    total_frames = FRAMES_PER_GPU * world
    shard = None
    if world > 1:
        shard = sharding.init_from_env(total_frames)
    h = _native.get_handle(dev)

    # ---- synthetic clip: every rank renders exactly the frames it must hold (own range + halo) ----
    mats = synth.shake_matrices(total_frames, 0, WIDTH, HEIGHT)
    base_dev = synth.base_texture(0, WIDTH, HEIGHT).to(dev)
    lo, hi = shard.load_range if shard is not None else (0, total_frames)
    clip_dev = synth.render_clip_cuda(h, base_dev, mats, WIDTH, HEIGHT, lo, hi)
    del base_dev
    torch.cuda.synchronize()

    def make_context(frames):
        return pipeline.VideoContext(frames, pipeline.FrameAdapter(np.float32, False, "0_1", "torch", False), WIDTH, HEIGHT, 3, None, "sequence", {})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return flow.stabilize_frames(make_context(clip_dev), PARAMS["framing"], PARAMS["mode"], PARAMS["camera_lock"], PARAMS["strength"],
                                     PARAMS["smooth"], PARAMS["keep_fov"], PARAMS["padding_rgb"], PARAMS["fps"], output="device", shard=shard)

    def timed(step_fn, steps, warmup, collect_warp=False):
        for _ in range(warmup):
            step_fn()
        settle_gc()
        barrier()
        if collect_warp:
            pipeline.WARP_LAUNCH_LOG = []
        launches0 = h.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        log = pipeline.WARP_LAUNCH_LOG
        pipeline.WARP_LAUNCH_LOG = None
        return float(t.item()), h.launch_count - launches0, log

    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms, launches, warp_log = timed(device_step, args.steps, args.warmup, collect_warp=True)
    clocks = sampler.stop()
    if os.environ.get("VSTAB_BENCH_PHASES"):  # diagnostics: host-side phase times of three more steps, every rank, to stderr
        from vstab_b200 import stabilizer_core as core

        for _ in range(3):
            core.PHASE_LOG = []
            barrier()
            t0 = time.perf_counter()
            device_step()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            print(f"[phases rank {rank}] step {wall:.2f} ms " + ", ".join(f"{a[:18]} {b * 1e3:.2f}" for a, b in core.PHASE_LOG), file=sys.stderr, flush=True)
        core.PHASE_LOG = None
    ms_per_step = total_ms / args.steps
    value = total_frames / (ms_per_step * 1e-3)

    # fused-warp roofline from the launches inside the timed steps
    warp_ms = sum(a.elapsed_time(b) for a, b, _, _ in warp_log)
    warp_bytes = sum(x[3] for x in warp_log)
    peak, peak_src = measured_peak_gbs()
    achieved = warp_bytes / (warp_ms * 1e-3) / 1e9 if warp_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": "warp_stream_kernel<bilinear> (+ warp_plan_kernel, timed together)", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "launches": len(warp_log), "avg_launch_ms": warp_ms / max(len(warp_log), 1),
        "algorithmic_bytes_per_frame": 12 * HEIGHT * WIDTH + 16 * HEIGHT * WIDTH,
        "algorithmic_bytes_per_launch": warp_bytes / max(len(warp_log), 1),
        "share_of_step": warp_ms / total_ms,
    }
    traffic_file = os.path.join(ROOT, "profiles", "warp_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                t = json.load(fh)
                # ncu capture was a 32-frame launch; DRAM traffic scales per frame (no inter-frame reuse)
                roofline["traffic"] = t["dram_bytes_per_frame"] * (sum(x[2] for x in warp_log) / max(len(warp_log), 1))
                roofline["traffic_source"] = t.get("source")
        except Exception:
            pass

    # ---- end to end through the public API with host tensors ----
    e2e = None
    if not args.no_e2e:
        own_lo, own_hi = shard.frame_range if shard is not None else (0, total_frames)
        host_clip = torch.empty(clip_dev.shape, dtype=torch.float32, pin_memory=True)
        host_clip.copy_(clip_dev)
        torch.cuda.synchronize()
        h2d = host_clip.numel() * 4
        d2h_holder = {}

        def e2e_step():
            ctx = pipeline.normalize_video_input(host_clip, dev)
            res = flow.stabilize_frames(ctx, PARAMS["framing"], PARAMS["mode"], PARAMS["camera_lock"], PARAMS["strength"], PARAMS["smooth"],
                                        PARAMS["keep_fov"], PARAMS["padding_rgb"], PARAMS["fps"], output="host", shard=shard)
            frames_out = pipeline.reconstruct_video(res.frames, ctx)
            masks_out = pipeline.convert_masks_for_output(res.masks)
            d2h_holder["bytes"] = frames_out.numel() * 4 + masks_out.numel() * 4

        e2e_steps = max(2, min(args.steps, 3))
        e2e_ms, _, _ = timed(e2e_step, e2e_steps, 1)
        e2e_ms /= e2e_steps
        e2e = {"value": total_frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d * world),
               "d2h_bytes_per_step": int(d2h_holder["bytes"] * world), "ms_per_step": e2e_ms,
               "note": "pinned host clip -> HBM, results -> pinned host, both inside the timed region"}
        del host_clip

    # ---- CPU baseline on the host cores (rank 0, N == 1 only) ----
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        import cv2

        from oracle import cv_path

        host = clip_dev.cpu()
        t0 = time.perf_counter()
        out_f, out_m, _ = cv_path.stabilize(host, "flow", PARAMS["framing"], PARAMS["mode"], PARAMS["camera_lock"], PARAMS["strength"],
                                            PARAMS["smooth"], PARAMS["keep_fov"], PARAMS["padding_rgb"], PARAMS["fps"])
        sec = time.perf_counter() - t0
        cpu = {"value": FRAMES_PER_GPU / sec, "unit": "frames/s", "cores": host_cores(), "kind": "port",
               "sample": f"the full 121-frame clip once ({sec:.1f} s), cv2 {cv2.__version__} with {cv2.getNumThreads()} threads"}
        del out_f, out_m, host

    launches_t = torch.tensor([launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(launches_t)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_total": total_frames, "sharding": "frame-range, 1 halo frame, all-gather of per-pair candidates" if world > 1 else "none",
                       "l2": "inputs larger than L2 (3.0 GB clip per GPU vs 126 MB)", "parallelism": f"frame-shard x{world}"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_t.item()), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
