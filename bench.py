#!/usr/bin/env python
"""bench.py -- frames/sec of 1080p Video Stabilizer Flow (DIS, similarity, crop_and_pad,
strength 0.7, smooth 0.5) on N B200s, with the fused-warp HBM roofline and the host-CPU reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 121 synthetic jittered 1920x1080 float32 frames per GPU.  With
N > 1 the clip is N x 121 frames sharded by contiguous frame range (one halo frame per rank, one
NCCL all-gather of the per-pair candidate table, one of the per-frame padded-pixel counts): weak
scaling.  One "step" = one full pass of the hot path over the (sharded) clip.

  value     frames/s with the clip already resident in HBM and the results left in HBM
  e2e       frames/s through the NODE (`nodes.VideoStabilizerFlow.execute`, the call ComfyUI makes) with a page-locked
            CPU IMAGE tensor in and CPU tensors out: upload and download inside the timed region.  Calls are timed one
            by one (they return host tensors); `value` is the MEDIAN call, the mean and every sample are listed
            (`ms_per_step_mean`, `ms_per_step_all`): on a box whose other GPUs serve other jobs the odd call is 20-90 ms slow.
            `e2e.pageable_node` is the same call with a pageable IMAGE (what a stock graph hands over),
            `e2e.pinned_driver` the driver below the node with a pinned input (what round 1 reported);
            `e2e.host_link` is the measured ceiling of this box's host link (plain pinned cudaMemcpyAsync up and
            down, one copy per call) and `e2e.frac_of_link` = time the bytes need at that rate / time taken
  roofline  fused resampler (vstab_warp_fused): algorithmic bytes (12HW read + 12H'W' + 4H'W'
            written per frame) / CUDA-event duration of its launches inside the timed steps,
            against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, a copy of /root/reference made by build()) through its own
            node `VideoStabilizerFlow.execute` on the host cores over the same clip; `oracle/cv_path.py` (a port
            of its cv2 call sequence) only when the copy is absent
  parity    the GPU result of the full 121-frame clip against that CPU result: per-pair transforms, applied
            matrices, pixels, masks, meta tree (north_star tolerances)
"""
from __future__ import annotations

import argparse
import json
import math
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
              padding_rgb=(127, 127, 127), padding_color="#7F7F7F", fps=16.0)
WORKLOAD = "Video Stabilizer Flow DIS, similarity, crop_and_pad, strength=0.7 smooth=0.5, 121 synthetic jittered 1920x1080 f32 frames per GPU"
METRIC = "frames/sec 1080p Flow stabilize"


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


# ------------------------------------------------------------------------------------ the CPU reference ----

def cpu_clip(n_frames: int) -> torch.Tensor:
    """The first `n_frames` of the bench clip rendered on the host with cv2 (bilinear, same matrices and base
    texture as the GPU render: the fused resampler is bit-exact against cv2 for bilinear, so the clips agree)."""
    import cv2

    import synth

    base = synth.base_texture(0, WIDTH, HEIGHT)
    mats = synth.shake_matrices(FRAMES_PER_GPU, 0, WIDTH, HEIGHT)[:n_frames]
    fwd = synth.render_matrices(mats)
    frames = np.stack([cv2.warpPerspective(base.numpy(), fwd[i], (WIDTH, HEIGHT), flags=cv2.INTER_LINEAR) for i in range(n_frames)])
    return torch.from_numpy(frames)


class CpuReference:
    """One callable for the CPU arm: the unmodified reference's node when baseline/_ref exists, else the port."""

    def __init__(self):
        import cv2

        from baseline import refload

        self.cv2 = cv2
        if refload.available():
            self.kind = "reference"
            self.ref = refload.load()
            self.what = "unmodified reference (baseline/_ref) VideoStabilizerFlow.execute"
        else:
            from oracle import cv_path  # the one other place bench.py may execute oracle/ (task statement, section 4)

            self.kind = "port"
            self.cv_path = cv_path
            self.what = "oracle/cv_path.py (port of the reference's cv2 call sequence; baseline/_ref absent)"

    def __call__(self, clip: torch.Tensor):
        """-> (frames [N,H,W,3] CPU tensor, masks [N,H,W] CPU tensor, meta)."""
        p = PARAMS
        if self.kind == "reference":
            out = self.ref.video_stabilizer_flow.VideoStabilizerFlow.execute(
                clip, p["fps"], p["framing"], p["mode"], p["camera_lock"], p["strength"], p["smooth"], p["keep_fov"], p["padding_color"])
            return out[0], out[1], out[2]
        f, m, meta = self.cv_path.stabilize(clip, "flow", p["framing"], p["mode"], p["camera_lock"], p["strength"], p["smooth"],
                                            p["keep_fov"], p["padding_rgb"], p["fps"])
        return torch.from_numpy(f), torch.from_numpy(m[..., 0]), meta

    def describe(self, n):
        return (f"{self.what}, all {n} frames of the clip per step, cv2 {self.cv2.__version__} with {self.cv2.getNumThreads()} threads")


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores, same clip, same
    parameters, all 121 frames of one GPU's share per step (at N > 1 the sharded clip is N x 121 frames; the CPU arm
    keeps timing 121 -- a bounded sample, and a rate).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = CpuReference()
    clip = cpu_clip(FRAMES_PER_GPU)
    n = FRAMES_PER_GPU

    def step():
        t0 = time.perf_counter()
        cpu(clip)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    settle_gc()
    times = [step() for _ in range(args.steps)]
    sec = float(np.mean(times))
    fps = n / sec
    sample = cpu.describe(n)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": host_cores(), "kind": cpu.kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- parity ----

def _decompose(m):
    m = np.asarray(m, dtype=np.float64)
    m = m / m[2, 2]
    return m[0, 2], m[1, 2], math.degrees(math.atan2(m[1, 0], m[0, 0])), math.hypot(m[0, 0], m[1, 0])


def _nested_diff(a, b, path, out):
    """Key-set equality + largest numeric difference of two JSON trees (reference: scripts/compare_refactor_behavior.py:196-217)."""
    if isinstance(a, dict):
        if not isinstance(b, dict) or set(a) != set(b):
            out["keys_equal"] = False
            out.setdefault("first_key_mismatch", path)
            return
        for k in a:
            _nested_diff(a[k], b[k], f"{path}.{k}", out)
    elif isinstance(a, (list, tuple)):
        if not isinstance(b, (list, tuple)) or len(a) != len(b):
            out["keys_equal"] = False
            out.setdefault("first_key_mismatch", path)
            return
        for i, (x, y) in enumerate(zip(a, b)):
            _nested_diff(x, y, f"{path}[{i}]", out)
    elif isinstance(a, bool) or a is None or isinstance(a, str):
        if a != b:
            out["values_equal"] = False
            out.setdefault("first_value_mismatch", path)
    elif isinstance(a, (int, float)):
        if not isinstance(b, (int, float)) or isinstance(b, bool):
            out["values_equal"] = False
            out.setdefault("first_value_mismatch", path)
            return
        d = abs(float(a) - float(b))
        rel = d / max(abs(float(a)), abs(float(b)), 1e-300)
        if d > out["max_abs"]:
            out["max_abs"], out["max_abs_at"] = d, path
        out["max_excess"] = max(out["max_excess"], min(d / 2e-5, rel / 2e-5) if d > 0 else 0.0)


def parity_block(gpu_res, ref_frames, ref_masks, ref_meta, dev):
    """GPU result (device tensors + meta) of the full bench clip against the CPU reference's."""
    meta = json.loads(json.dumps(gpu_res.meta))
    ref_meta = json.loads(json.dumps(ref_meta))
    worst = [0.0, 0.0, 0.0]
    modes_equal = True
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], ref_meta["estimated_motion"]["per_transition"]):
        a, b = _decompose(mine["matrix"]), _decompose(ref["matrix"])
        worst = [max(worst[0], abs(a[0] - b[0]), abs(a[1] - b[1])), max(worst[1], abs(a[2] - b[2])), max(worst[2], abs(a[3] - b[3]))]
        modes_equal &= mine["mode"] == ref["mode"]
    applied = max(float(np.abs(np.asarray(x["applied_matrix"]) - np.asarray(y["applied_matrix"])).max())
                  for x, y in zip(meta["stabilization_warp"]["per_frame"], ref_meta["stabilization_warp"]["per_frame"]))
    tree = {"keys_equal": True, "values_equal": True, "max_abs": 0.0, "max_abs_at": None, "max_excess": 0.0}
    _nested_diff(ref_meta, meta, "meta", tree)
    # pixels and masks, chunk by chunk on the device
    gf, gm = gpu_res.frames, gpu_res.masks[..., 0]
    n = gf.shape[0]
    px_max, px_bad, mask_bad, px_total = 0.0, 0, 0, 0
    same_shape = tuple(gf.shape) == tuple(ref_frames.shape) and tuple(gm.shape) == tuple(ref_masks.shape)
    if same_shape:
        for a in range(0, n, 16):
            b = min(a + 16, n)
            rf = ref_frames[a:b].to(dev, non_blocking=True)
            rm = ref_masks[a:b].to(dev, non_blocking=True)
            d = (gf[a:b] - rf).abs()
            px_max = max(px_max, float(d.max()))
            px_bad += int((d > 1e-3).sum())
            px_total += d.numel()
            mask_bad += int((gm[a:b] != rm).sum())
    within = (modes_equal and worst[0] <= 0.05 and worst[1] <= 0.01 and worst[2] <= 1e-4 and same_shape and px_max <= 1e-3
              and tree["keys_equal"] and tree["values_equal"])
    return {
        "against": "CPU reference result of the same 121-frame clip (cpu_baseline leg)", "pairs": len(meta["estimated_motion"]["per_transition"]),
        "pair_modes_equal": modes_equal,
        "pair_transform_max_delta": {"translation_px": worst[0], "rotation_deg": worst[1], "scale": worst[2]},
        "applied_matrix_max_abs_diff": applied,
        "pixel_max_abs_err": px_max, "pixels_over_1e-3": px_bad, "pixels": px_total,
        "mask_mismatch_px": mask_bad, "mask_px": int(gm.numel()),
        "meta": {"keys_equal": tree["keys_equal"], "non_numeric_equal": tree["values_equal"], "max_abs_diff": tree["max_abs"],
                 "max_abs_diff_at": tree["max_abs_at"], "within_2e-5_abs_or_rel": tree["max_excess"] <= 1.0},
        "tolerances": "north_star: 0.05 px / 0.01 deg / 1e-4 scale per pair, 1e-3 pixels (bilinear), mask bit-exact given identical matrices",
        "within_north_star": bool(within),
    }


# -------------------------------------------------------------------------------------- host link ----

def host_link_probe(dev, world, barrier):
    """Ceiling of the host link of THIS box for the e2e figure: plain pinned cudaMemcpyAsync (torch copy_) of 1 GiB,
    host->device and device->host, one copy per call, every rank at the same time (all N GPUs pull on the host
    memory system together, like the e2e step does).  GB/s per GPU, best of 3."""
    nbytes = 1 << 30
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    host.fill_(1)
    out = {}
    for name, (dst, src) in {"h2d": (devbuf, host), "d2h": (host, devbuf)}.items():
        best = 0.0
        for _ in range(3):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[name + "_gbs"] = best
    del host, devbuf
    return out


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
    placement = pipeline.bind_host_to_gpu(local_rank, world)  # NUMA node / core share of this rank (pinned buffers follow)
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

    def timed_each(step_fn, steps, warmup):
        """End-to-end calls, one by one: every call returns host tensors, so each is bracketed by a barrier and a device
        synchronize of its own; the per-call times are max-reduced over the ranks.  Returns (median, mean, all) in ms: the
        host side of the link (pinned allocations, the DMA engines of a box whose other GPUs serve other jobs) produces
        the odd call 20-90 ms slower than its neighbours (scripts/e2e_reps.py names the phase: enqueueing / running the
        upload), so the MEDIAN is what `e2e.value` quotes and every sample is listed next to it."""
        for _ in range(warmup):
            step_fn()
        settle_gc()
        times = []
        for _ in range(steps):
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step_fn()
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor(times, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times = [float(x) for x in t.tolist()]
        return float(np.median(times)), float(np.mean(times)), [round(x, 2) for x in times]

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
        from baseline import refload

        refload.install_stubs()  # ComfyUI is not in this image: comfy_api.latest / comfy.utils stand-ins (sockets, NodeOutput, ProgressBar)
        from vstab_b200 import nodes

        link = host_link_probe(dev, world, barrier)
        pinned_clip = torch.empty(clip_dev.shape, dtype=torch.float32, pin_memory=True)
        pinned_clip.copy_(clip_dev)
        torch.cuda.synchronize()
        h2d = pinned_clip.numel() * 4
        d2h_holder = {}

        def pinned_driver_step():
            ctx = pipeline.normalize_video_input(pinned_clip, dev)
            res = flow.stabilize_frames(ctx, PARAMS["framing"], PARAMS["mode"], PARAMS["camera_lock"], PARAMS["strength"], PARAMS["smooth"],
                                        PARAMS["keep_fov"], PARAMS["padding_rgb"], PARAMS["fps"], output="host", shard=shard)
            frames_out = pipeline.reconstruct_video(res.frames, ctx)
            masks_out = pipeline.convert_masks_for_output(res.masks)
            # what actually crossed the link: float32 frames + the 0 / 1 padding mask as one byte per pixel (widened to the
            # float32 MASK on the host inside the timed region); pipeline counts the bytes of the copies it issues
            d2h_holder["bytes"] = pipeline.LAST_D2H_BYTES
            d2h_holder["result_bytes"] = frames_out.numel() * 4 + masks_out.numel() * 4

        e2e_steps = max(2, min(args.steps, 7))
        pin_ms, pin_mean, pin_all = timed_each(pinned_driver_step, e2e_steps, 1)
        d2h = d2h_holder["bytes"]
        ideal_ms = (h2d / (link["h2d_gbs"] * 1e9) + d2h / (link["d2h_gbs"] * 1e9)) * 1e3
        e2e = {"unit": "frames/s", "h2d_bytes_per_step": int(h2d * world), "d2h_bytes_per_step": int(d2h * world),
               "result_bytes_per_step": int(d2h_holder["result_bytes"] * world),
               "mask_on_the_link": "uint8 (binary mask, widened to float32 on the host)" if pipeline.mask_bytes_enabled() else "float32",
               "host_link": {**link, "how": "plain pinned cudaMemcpyAsync of 1 GiB each way, one copy per call, all ranks at once, best of 3",
                             "ms_for_the_step_bytes": ideal_ms, "placement": placement},
               "statistic": f"median of {e2e_steps} calls timed one by one (host wall clock between device synchronizes, max over ranks); mean and every sample listed",
               "pinned_driver": {"value": total_frames / (pin_ms * 1e-3), "ms_per_step": pin_ms, "frac_of_link": ideal_ms / pin_ms,
                                 "ms_per_step_mean": pin_mean, "ms_per_step_all": pin_all,
                                 "note": "pinned host clip -> flow.stabilize_frames(output='host') -> pinned results (round 1's e2e)"}}
        if world == 1:
            # the call ComfyUI makes: nodes.VideoStabilizerFlow.execute(IMAGE, widgets...) -> (IMAGE, MASK, JSON), CPU tensors both ways
            def node_step(image):
                out = nodes.VideoStabilizerFlow.execute(image, PARAMS["fps"], PARAMS["framing"], PARAMS["mode"], PARAMS["camera_lock"],
                                                        PARAMS["strength"], PARAMS["smooth"], PARAMS["keep_fov"], PARAMS["padding_color"])
                assert out[0].shape[0] == total_frames and not out[0].is_cuda and not out[1].is_cuda

            node_ms, node_mean, node_all = timed_each(lambda: node_step(pinned_clip), e2e_steps, 1)
            e2e.update({"value": total_frames / (node_ms * 1e-3), "ms_per_step": node_ms, "frac_of_link": ideal_ms / node_ms,
                        "ms_per_step_mean": node_mean, "ms_per_step_all": node_all,
                        "note": "nodes.VideoStabilizerFlow.execute(page-locked CPU IMAGE) -> CPU IMAGE + MASK + meta; upload and download inside the timed region"})
            # ... and with what a stock ComfyUI graph hands over: a PAGEABLE tensor (staged through two pinned bounce buffers)
            pageable = torch.empty(clip_dev.shape, dtype=torch.float32)
            pageable.copy_(pinned_clip)
            page_ms, page_mean, page_all = timed_each(lambda: node_step(pageable), e2e_steps, 1)
            e2e["pageable_node"] = {"value": total_frames / (page_ms * 1e-3), "ms_per_step": page_ms, "frac_of_link": ideal_ms / page_ms,
                                    "ms_per_step_mean": page_mean, "ms_per_step_all": page_all,
                                    "note": "the same node call with a pageable CPU IMAGE"}
            del pageable
        else:
            # sharded runs have no single-process node call: every rank drives its frame range through the driver
            e2e.update({"value": e2e["pinned_driver"]["value"], "ms_per_step": pin_ms, "frac_of_link": ideal_ms / pin_ms,
                        "ms_per_step_mean": pin_mean, "ms_per_step_all": pin_all,
                        "note": "per-rank pinned host shard -> flow.stabilize_frames(shard, output='host') -> pinned results"})
        del pinned_clip

    # ---- CPU baseline on the host cores + parity of the GPU result against it (rank 0, N == 1 only) ----
    cpu = None
    parity = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        ref = CpuReference()
        host = clip_dev.cpu()
        t0 = time.perf_counter()
        ref_frames, ref_masks, ref_meta = ref(host)
        sec = time.perf_counter() - t0
        cpu = {"value": FRAMES_PER_GPU / sec, "unit": "frames/s", "cores": host_cores(), "kind": ref.kind,
               "sample": f"{ref.describe(FRAMES_PER_GPU)}, once ({sec:.1f} s)"}
        del host
        parity = parity_block(device_step(), ref_frames, ref_masks, ref_meta, dev)
        del ref_frames, ref_masks

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
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "gpu_launches": int(launches_t.item()), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
