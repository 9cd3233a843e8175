"""BASELINE config 5 at scale: Video Stabilizer Flow (DIS), perspective, camera_lock, 3840x2160 frames, frame-range
sharded over N B200s with the NCCL all-gather of the per-pair candidate models.

    python scripts/cfg5_scale.py [--frames-per-gpu 250] [--steps 3] [--warmup 2] [--e2e-frames 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/cfg5_scale.py ...

The 2000-frame clip of BASELINE.json configs[4] is 199 GB of float32 input: it exists only sharded (250 frames = 24.9 GB
per GPU at N = 8), so the scaling is measured the way the clip is run -- every GPU holds 250 frames (own range + one halo
frame, rendered on the device from (seed, index)) and N = 1, 2, 4, 8 process 250 N frames: frames/s(N) / frames/s(1).
Rank 0 prints one JSON line: device-resident frames/s (CUDA events, max over ranks), host-side phase times of the slowest
rank, and an end-to-end leg (pinned host shard -> HBM -> pinned host results) on `--e2e-frames` frames per GPU.
`--cpu-frames K` (N = 1 only) times the unmodified reference (baseline/_ref) on the first K frames of the same clip.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

W, H = 3840, 2160
ARGS = ("crop_and_pad", "perspective", True, 0.7, 0.5, 0.6, (127, 127, 127), 16.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-gpu", type=int, default=250)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--e2e-frames", type=int, default=32)
    ap.add_argument("--cpu-frames", type=int, default=0)
    ap.add_argument("--probe", action="store_true", help="per-rank kernel timings + NVML clocks / power / throttle reasons")
    args = ap.parse_args()

    import torch.distributed as dist

    import synth
    import vstab_loader

    vstab_loader.load()
    from vstab_b200 import _native, flow, pipeline, sharding, stabilizer_core as core

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = pipeline.bind_host_to_gpu(local, world)
    h = _native.get_handle(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_case(frames_per_gpu, steps, warmup, host_io):
        total = frames_per_gpu * world
        shard = None
        if world > 1:
            if not dist.is_initialized():
                sharding.init_from_env(total)
            shard = sharding.FrameShard(rank, world, total, None, dev)
        mats = synth.shake_matrices(total, 0, W, H, perspective=True)
        lo, hi = shard.load_range if shard is not None else (0, total)
        base = synth.base_texture(0, W, H).to(dev)
        clip = synth.render_clip_cuda(h, base, mats, W, H, lo, hi)
        del base
        src = clip
        if host_io:
            src = torch.empty(clip.shape, dtype=torch.float32, pin_memory=True)
            src.copy_(clip)
            del clip
            torch.cuda.synchronize()

        def step():
            if host_io:
                ctx = pipeline.normalize_video_input(src, dev)
                res = flow.stabilize_frames(ctx, *ARGS, output="host", shard=shard)
                return pipeline.reconstruct_video(res.frames, ctx), pipeline.convert_masks_for_output(res.masks), res.meta
            ctx = pipeline.VideoContext(src, pipeline.FrameAdapter(np.float32, False, "0_1", "torch", False), W, H, 3, None, "sequence", {})
            res = flow.stabilize_frames(ctx, *ARGS, output="device", shard=shard)
            return res.frames, res.masks, res.meta

        out = None
        for _ in range(warmup):
            del out  # results of the previous step go back to torch's pinned-host / device caches before the next one allocates
            out = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            del out
            out = step()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # host-side phases of one more step, slowest rank's view
        core.PHASE_LOG = []
        core.GPU_MARKS = []
        pipeline.WARP_LAUNCH_LOG = []
        barrier()
        del out
        t0 = time.perf_counter()
        out = step()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        phases = {a[:40]: round(b * 1e3, 3) for a, b in core.PHASE_LOG}
        marks = core.GPU_MARKS
        warp_ms = round(sum(a.elapsed_time(b) for a, b, _, _ in pipeline.WARP_LAUNCH_LOG), 3)
        pipeline.WARP_LAUNCH_LOG = None
        gpu = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 3) for i in range(1, len(marks))}
        gpu["of which: resampler launches (CUDA events around vstab_warp_fused)"] = warp_ms
        core.PHASE_LOG = None
        core.GPU_MARKS = None
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, gpu)
            gpu = {k: [e.get(k) for e in every] for k in gpu}
        meta = out[2]
        info = {"frames_total": total, "ms_per_step": float(ms.item()), "frames_per_s": total / (float(ms.item()) * 1e-3),
                "mode_applied": meta["transform_mode_applied"], "rank0_step_wall_ms": round(wall, 3), "rank0_phases_ms": phases,
                "gpu_ms_between_marks_per_rank": gpu}
        if host_io:
            info["h2d_bytes_per_step"] = int(src.numel() * 4 * world)
            info["d2h_bytes_per_step"] = int((out[0].numel() + out[1].numel()) * 4 * world)
        del src, out
        torch.cuda.empty_cache()
        return info

    def kernel_probe(frames_per_gpu):
        """Every rank times its own kernels with CUDA events -- gray, DIS, fit, resampler, each alone, all ranks at the same
        moment (barrier first) -- while NVML is sampled: tells apart "the GPUs got slower" (clocks, power cap when 8 GPUs
        stream from HBM at once) from "the ranks wait for each other / for the host"."""
        import threading

        total = frames_per_gpu * world
        mats = synth.shake_matrices(total, 0, W, H, perspective=True)
        lo, hi = sharding.FrameShard(rank, world, total).load_range if world > 1 else (0, total)
        base = synth.base_texture(0, W, H).to(dev)
        clip = synth.render_clip_cuda(h, base, mats, W, H, lo, hi)
        del base
        n = clip.shape[0]
        samples = {"sm": [], "mem": [], "power": [], "reasons": set()}
        stop = threading.Event()

        def sample():
            try:
                import pynvml as nv

                nv.nvmlInit()
                hd = nv.nvmlDeviceGetHandleByIndex(local)
                while not stop.is_set():
                    samples["sm"].append(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM))
                    samples["mem"].append(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_MEM))
                    samples["power"].append(nv.nvmlDeviceGetPowerUsage(hd) / 1000.0)
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(hd)
                    for name, bit in (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal", 0x20), ("hw_thermal", 0x40), ("hw_power_brake", 0x80)):
                        if mask & bit:
                            samples["reasons"].add(name)
                    stop.wait(0.005)
            except Exception as exc:
                samples["reasons"].add(f"nvml:{type(exc).__name__}")

        th = threading.Thread(target=sample, daemon=True)
        th.start()

        def timed(fn, reps=3):
            fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                r = fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps, r

        out = {}
        out["gray_ms"], gray = timed(lambda: h.gray_working(clip, (960, 540)))
        out["dis_ms"], (_, grid) = timed(lambda: h.dis_flow(gray, want_flow=False, grid_step=8))
        out["fit_ms"], _ = timed(lambda: h.fit_grid(grid, 8, 7))
        fwd = torch.from_numpy(np.tile(np.array([1.001, -0.002, 3.3, 0.002, 0.999, -2.1, 1e-6, 0, 1], np.float32), (n, 1, 1))).to(dev)
        dst = torch.empty((n, H, W, 3), dtype=torch.float32, device=dev)
        msk = torch.empty((n, H, W), dtype=torch.float32, device=dev)
        out["warp_ms"], _ = timed(lambda: h.warp_fused(clip, fwd, (W, H), "bilinear", (0.5, 0.5, 0.5), out=dst, mask_out=msk, want_pad_count=True))
        stop.set()
        th.join(timeout=2)
        out["sm_mhz_min_median"] = [int(min(samples["sm"] or [0])), int(np.median(samples["sm"] or [0]))]
        out["mem_mhz_min"] = int(min(samples["mem"] or [0]))
        out["power_w_max"] = round(max(samples["power"] or [0]), 1)
        out["reasons"] = sorted(samples["reasons"])
        del clip, dst, msk, gray, grid
        torch.cuda.empty_cache()
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, out)
            return gathered
        return [out]

    line = {"config": "BASELINE configs[4]: Flow DIS, perspective, camera_lock, 3840x2160 f32, frame-range sharded", "n_gpus": world,
            "frames_per_gpu": args.frames_per_gpu, "steps": args.steps, "warmup": args.warmup, "placement_rank0": placement}
    line["device_resident"] = run_case(args.frames_per_gpu, args.steps, args.warmup, host_io=False)
    if args.probe:
        line["kernel_probe_per_rank"] = kernel_probe(args.frames_per_gpu)
    if args.e2e_frames > 0:
        line["e2e"] = run_case(args.e2e_frames, max(1, min(args.steps, 2)), 2, host_io=True)
        line["e2e"]["note"] = f"{args.e2e_frames} frames per GPU: pinned host shard -> HBM -> pinned host results inside the timed region"
    if args.cpu_frames > 0 and world == 1:
        import cv2

        from baseline import refload

        k = args.cpu_frames
        mats = synth.shake_matrices(args.frames_per_gpu, 0, W, H, perspective=True)[:k]
        fwd = synth.render_matrices(mats)
        base = synth.base_texture(0, W, H).numpy()
        clip = torch.from_numpy(np.stack([cv2.warpPerspective(base, fwd[i], (W, H), flags=cv2.INTER_LINEAR) for i in range(k)]))
        ref = refload.load()
        t0 = time.perf_counter()
        ref.video_stabilizer_flow.VideoStabilizerFlow.execute(clip, 16.0, "crop_and_pad", "perspective", True, 0.7, 0.5, 0.6, "#7F7F7F")
        sec = time.perf_counter() - t0
        line["cpu_reference"] = {"frames": k, "frames_per_s": k / sec, "kind": "reference", "cores": len(os.sched_getaffinity(0)),
                                 "note": f"unmodified reference node on the first {k} frames of the same clip, cv2 {cv2.__version__}, {cv2.getNumThreads()} threads"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
