"""Shared pipeline body of the Flow and Classic stabilizers on the CUDA hot path.

Behavioural mirror of ``_stabilize_frames`` in the reference
(nodes/video_stabilizer_flow.py:213-640 and nodes/video_stabilizer_classic.py:163-567, which are
near-verbatim copies of each other): fps resolution, early outs, per-pair estimation, sticky mode
ladder, path cumsum + box smoothing, framing (crop_and_pad recentring / expand / crop), warp +
padding mask, and the ``meta`` dictionary whose key set is part of the drop-in contract.

Where the reference loops over frames calling cv2, this module launches batched kernels:
  estimation   gray+area (K1/K2) -> DIS or GFTT+LK (K3 | K5/K6) -> all candidate models (K7-K9)
  ladder       replayed on the host over the candidate table (tiny; also what makes frame-range
               sharding exact, see sharding.py)
  warp         one fused launch per chunk (K10-K12) with per-frame padded-pixel counts
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple

import math

import numpy as np
import torch

from . import _native, hostmath as hm
from .motion_meta import applied_motion_meta_from_matrices, applied_motion_meta_from_stabilization_warp
from .pipeline import VideoContext, fused_warp

MODE_NAMES = _native.MODE_NAMES

# set to a list to collect (label, seconds) host-side phase timings of the last call (bench / debugging)
PHASE_LOG = None
# set to a list to collect (label, torch.cuda.Event) marks recorded on the launching stream at the phase boundaries of
# the last call: elapsed_time between consecutive marks is what the GPU (not the host) spent there, collectives included
GPU_MARKS = None


def _gpu_mark(label, device):
    if GPU_MARKS is not None and device is not None and torch.device(device).type == "cuda":
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(device))
        GPU_MARKS.append((label, ev))


def _mark(label, t0):
    import time

    if PHASE_LOG is not None:
        PHASE_LOG.append((label, time.perf_counter() - t0))
    return time.perf_counter()


class _Nvtx:
    """NVTX ranges around the phases of a call (estimate / table / ladder / solve / resample / meta / finish), visible in
    Nsight Systems and in ncu's NVTX filters.  On when a CUDA device is there; VSTAB_NVTX=0 switches them off."""

    def __init__(self):
        import os

        self.on = os.environ.get("VSTAB_NVTX", "1") != "0" and torch.cuda.is_available()
        self.open = False

    def phase(self, name: Optional[str]) -> None:
        if not self.on:
            return
        if self.open:
            torch.cuda.nvtx.range_pop()
            self.open = False
        if name is not None:
            torch.cuda.nvtx.range_push(f"vstab.{name}")
            self.open = True


@dataclass
class StabilizationResult:
    frames: Any  # [N,H',W',3] float32 (numpy view of a pinned CPU tensor, or CUDA tensor)
    masks: Any   # [N,H',W',1] float32
    meta: Dict[str, Any]


@dataclass
class PairCandidates:
    """Candidate models of every frame pair, indexed [pair, mode] with mode = VSTAB_MODE_*."""
    matrix: np.ndarray      # [P,3,3,3] float64 (pre-float32-cast)
    residual: np.ndarray    # [P,3] float64
    n_inliers: np.ndarray   # [P,3] int
    n_valid: np.ndarray     # [P,3] int
    n_total: np.ndarray     # [P,3] int
    ok: np.ndarray          # [P,3] int
    min_points: int = 12    # Flow: 12 finite grid points; Classic: 8 tracked features
    detected: Optional[np.ndarray] = None  # Classic: corners found per pair (<12 => identity)
    raw: Optional[np.ndarray] = None       # [P,3,12] float64 words the columns were decoded from (vstab_fit_result)

    def raw_words(self) -> np.ndarray:
        """The table as [P,3,12] vstab_fit_result words (what vstab_host_trajectory reads): the block it was decoded
        from, or the columns packed again (tables that were built or gathered as columns)."""
        if self.raw is not None:
            return self.raw
        p = self.matrix.shape[0]
        arr = np.zeros((p, 3, 12), dtype=np.float64)
        arr[..., :9] = np.asarray(self.matrix, dtype=np.float64).reshape(p, 3, 9)
        arr[..., 9] = self.residual
        ints = np.stack([self.n_inliers, self.n_valid, self.n_total, self.ok], axis=-1).astype(np.int32)
        arr[..., 10:12] = np.ascontiguousarray(ints).view(np.float64).reshape(p, 3, 2)
        return arr

    def to_array(self) -> np.ndarray:
        """Flat float64 [P, 3*14 + 1] table (what shards all-gather)."""
        p = self.matrix.shape[0]
        cols = [self.matrix.reshape(p, 27), self.residual, self.n_inliers, self.n_valid, self.n_total, self.ok]
        det = self.detected if self.detected is not None else np.full((p,), -1)
        return np.concatenate([np.asarray(c, dtype=np.float64).reshape(p, -1) for c in cols] + [det.reshape(p, 1).astype(np.float64)], axis=1)

    @staticmethod
    def from_array(arr: np.ndarray, min_points: int) -> "PairCandidates":
        p = arr.shape[0]
        ints = np.rint(arr[:, 30:43]).astype(np.int64)  # the five integer blocks in one pass
        det = ints[:, 12]
        return PairCandidates(
            matrix=arr[:, :27].reshape(p, 3, 3, 3).copy(), residual=arr[:, 27:30].copy(), n_inliers=ints[:, 0:3],
            n_valid=ints[:, 3:6], n_total=ints[:, 6:9], ok=ints[:, 9:12], min_points=min_points,
            detected=None if (det < 0).all() else det,
        )


    @staticmethod
    def from_raw(raw: np.ndarray, min_points: int, detected: Optional[np.ndarray] = None) -> "PairCandidates":
        """raw [P,3,12] float64 words of vstab_fit_result (include/vstab.h) -> columns."""
        arr = np.ascontiguousarray(raw, dtype=np.float64).reshape(-1, 3, 12)
        ints = arr[..., 10:12].copy().view(np.int32)  # [P,3,4]: n_inliers, n_valid, n_total, ok
        return PairCandidates(
            matrix=arr[..., :9].reshape(arr.shape[0], 3, 3, 3).copy(), residual=arr[..., 9].copy(), n_inliers=ints[..., 0].copy(),
            n_valid=ints[..., 1].copy(), n_total=ints[..., 2].copy(), ok=ints[..., 3].copy(), min_points=min_points,
            detected=None if detected is None else np.asarray(detected).astype(np.int64), raw=arr,
        )


@dataclass
class DeviceCandidates:
    """The candidate table as the fit kernels left it in HBM: nothing has waited for the GPU yet.  A frame-range
    shard all-gathers `raw` device to device (sharding.FrameShard.gather_candidates); `to_host()` is the one
    blocking copy of the estimation phase."""
    raw: torch.Tensor                 # [P,3,12] float64 words (vstab_fit_result), CUDA
    min_points: int
    detected: Optional[torch.Tensor] = None  # Classic: [P] int32, CUDA
    send: Optional[torch.Tensor] = None      # [cap,3,12] all-gather send buffer `raw` is a view of (sharded runs)

    def to_host(self) -> PairCandidates:
        return self.to_host_words().decode()

    def to_host_words(self) -> "HostWords":
        """The blocking copy alone; the columns are decoded when somebody asks for them (HostWords.decode)."""
        det = None if self.detected is None else self.detected.cpu().numpy()
        return HostWords(self.raw.cpu().numpy(), self.min_points, det)


@dataclass
class HostWords:
    """The candidate table on the host as the fit kernels wrote it ([P,3,12] vstab_fit_result words).  The trajectory
    helper of libvstab reads the words directly; the per-column view (PairCandidates) is only needed by the fallback
    ladder and by the meta builder, which runs while the resampler is in flight."""
    raw: np.ndarray
    min_points: int
    detected: Optional[np.ndarray] = None

    def raw_words(self) -> np.ndarray:
        return self.raw

    def decode(self) -> PairCandidates:
        return PairCandidates.from_raw(self.raw, self.min_points, self.detected)


Estimator = Callable[[VideoContext, int, int, str], Any]  # -> PairCandidates | DeviceCandidates


def _accepts_requested(cands: PairCandidates, mode: str) -> np.ndarray:
    """Per pair: would the ladder accept `mode` itself (no fallback)?  Vectorised."""
    k = _native.MODE_INDEX[mode]
    n_valid = cands.n_valid.max(axis=1)
    ok = n_valid >= cands.min_points
    if cands.detected is not None:
        ok &= cands.detected >= 12
    if mode == "translation":
        return ok
    conf = cands.n_inliers[:, k] / np.maximum(n_valid, 1).astype(np.float64)
    need, thr = (4, 0.15) if mode == "perspective" else (3, 0.1)
    return ok & (cands.ok[:, k] != 0) & (n_valid >= need) & (conf >= thr)


class LadderEntries:
    """Per-pair outcome of the ladder as parallel columns (mode, confidence, residual) next to the
    float32 matrix stack.  Indexing / iterating yields the reference's per-pair tuples
    (matrix, mode, confidence, residual); the columns are what the meta builder consumes, so long
    clips do not pay for a tuple and a matrix view per pair on every rank."""

    __slots__ = ("matrices", "modes", "confidences", "residuals")

    def __init__(self, matrices, modes, confidences, residuals):
        self.matrices, self.modes, self.confidences, self.residuals = matrices, modes, confidences, residuals

    def __len__(self):
        return len(self.modes)

    def __getitem__(self, i):
        return (self.matrices[i], self.modes[i], self.confidences[i], self.residuals[i])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def replay_mode_ladder(cands: PairCandidates, requested_mode: str, *, with_residual: bool):
    """The reference's per-pair fallback ladder with its clip-wide sticky downgrade
    (flow.py:156-210 + :324-339; classic.py:104-158 + :264-272), replayed over the table.
    Returns (LadderEntries of (matrix f32, mode, confidence, residual), final active mode, matrix
    stack).  The common case -- every pair accepts the requested model -- is answered from array
    operations; the per-pair loop only starts at the first pair that falls back."""
    active = requested_mode
    eye = np.eye(3, dtype=np.float32)
    total = cands.matrix.shape[0]
    accepted = _accepts_requested(cands, requested_mode)
    first_fallback = int(np.argmin(accepted)) if not accepted.all() else total
    modes: List[str] = []
    confs: List[float] = []
    resid: List[Optional[float]] = []
    mats32 = np.zeros((0, 3, 3), np.float32)
    if first_fallback > 0:
        k = _native.MODE_INDEX[requested_mode]
        n_valid = cands.n_valid.max(axis=1)[:first_fallback]
        mats32 = cands.matrix[:first_fallback, k].astype(np.float32)
        if requested_mode == "translation":
            denom = cands.detected[:first_fallback] if cands.detected is not None else cands.n_total[:first_fallback, k]
            confs = (n_valid / denom.astype(np.float64)).tolist()
        else:
            confs = (cands.n_inliers[:first_fallback, k] / n_valid.astype(np.float64)).tolist()
        resid = cands.residual[:first_fallback, k].tolist() if with_residual else [None] * first_fallback
        modes = [requested_mode] * first_fallback
        if first_fallback == total:
            return LadderEntries(mats32, modes, confs, resid), active, mats32
    tail = []
    for p in range(first_fallback, total):
        n_valid = int(cands.n_valid[p].max())
        chosen = None
        too_few = n_valid < cands.min_points
        if cands.detected is not None and int(cands.detected[p]) < 12:
            too_few = True
        if not too_few:
            for mode in hm.MODE_LADDER[active]:
                k = _native.MODE_INDEX[mode]
                ok = bool(cands.ok[p, k])
                if mode == "perspective" and n_valid >= 4 and ok:
                    conf = float(cands.n_inliers[p, k]) / float(n_valid)
                    if conf >= 0.15:
                        chosen = (cands.matrix[p, k].astype(np.float32), mode, conf, float(cands.residual[p, k]))
                        break
                elif mode == "similarity" and n_valid >= 3 and ok:
                    conf = float(cands.n_inliers[p, k]) / float(n_valid)
                    if conf >= 0.1:
                        chosen = (cands.matrix[p, k].astype(np.float32), mode, conf, float(cands.residual[p, k]))
                        break
                elif mode == "translation":
                    denom = int(cands.detected[p]) if cands.detected is not None else int(cands.n_total[p, k])
                    conf = float(n_valid) / float(denom)
                    chosen = (cands.matrix[p, k].astype(np.float32), mode, conf, float(cands.residual[p, k]))
                    break
        if chosen is None:
            chosen = (eye.copy(), "translation", 0.0, 0.0)
        matrix, used, conf, res = chosen
        if used != active:
            active = used
        tail.append(matrix)
        modes.append(used)
        confs.append(conf)
        resid.append(res if with_residual else None)
    stacked = np.concatenate([mats32, np.stack(tail, axis=0)], axis=0) if tail else mats32
    return LadderEntries(stacked, modes, confs, resid), active, stacked


def accepted_ladder_entries(cands: PairCandidates, mode: str, *, with_residual: bool) -> LadderEntries:
    """LadderEntries columns (mode, confidence, residual) of a clip whose every pair accepted `mode` itself: what
    replay_mode_ladder returns in that case, without the matrix stack (vstab_host_trajectory produced it already)."""
    k = _native.MODE_INDEX[mode]
    total = cands.matrix.shape[0]
    n_valid = cands.n_valid.max(axis=1)
    if mode == "translation":
        denom = cands.detected if cands.detected is not None else cands.n_total[:, k]
        confs = (n_valid / denom.astype(np.float64)).tolist()
    else:
        confs = (cands.n_inliers[:, k] / n_valid.astype(np.float64)).tolist()
    resid = cands.residual[:, k].tolist() if with_residual else [None] * total
    return LadderEntries(None, [mode] * total, confs, resid)


def _native_host_solve() -> bool:
    """The trajectory solve runs in libvstab's host helpers (csrc/hostsolve.cu) unless VSTAB_HOST_SOLVE=0 asks for the
    numpy formulation of hostmath.py -- same operations, same bits (tests/test_host_solve_cpu.py), ~0.4 ms slower."""
    import os

    return os.environ.get("VSTAB_HOST_SOLVE", "1") != "0"


class _Progress:
    """update_absolute every 10 items, like the reference's progress_stride bookkeeping."""

    def __init__(self, bar, total: int):
        self.bar, self.total, self.done = bar, total, 0

    def advance(self, count: int, stride: int = 10) -> None:
        if self.bar is None or count <= 0:
            self.done += max(count, 0)
            return
        left = count
        while left > 0:
            step = min(stride, left)
            self.done += step
            left -= step
            self.bar.update_absolute(self.done, self.total)

    def finish(self) -> None:
        if self.bar is not None:
            self.bar.update_absolute(self.total, self.total)


def stabilize_frames(
    context: VideoContext,
    framing_mode: str,
    transform_mode: str,
    camera_lock: bool,
    strength: float,
    smooth: float,
    keep_fov: float,
    padding_rgb: Tuple[int, int, int],
    frame_rate: float,
    *,
    estimator: Estimator,
    flavour: str,  # "flow" | "classic": selects meta keys and motion_meta source
    progress_bar: Any = None,
    interrupt_check: Optional[Callable[[], None]] = None,
    output: str = "host",
    shard=None,
) -> StabilizationResult:
    nvtx = _Nvtx()
    try:
        with hm.gc_paused() as pause:
            return _stabilize_frames(context, framing_mode, transform_mode, camera_lock, strength, smooth, keep_fov, padding_rgb,
                                     frame_rate, estimator=estimator, flavour=flavour, progress_bar=progress_bar,
                                     interrupt_check=interrupt_check, output=output, shard=shard, nvtx=nvtx, gc_pause=pause)
    finally:
        nvtx.phase(None)


def _stabilize_frames(context, framing_mode, transform_mode, camera_lock, strength, smooth, keep_fov, padding_rgb, frame_rate, *,
                      estimator, flavour, progress_bar, interrupt_check, output, shard, nvtx, gc_pause=None) -> StabilizationResult:
    is_flow = flavour == "flow"
    total_frames = len(context)
    width, height = context.width, context.height
    fps_effective, fps_requested = hm.resolve_fps_for_stabilizer(frame_rate, context.fps)
    flow_keys = {"flow_backend": "DIS", "flow_fallback_reason": None} if is_flow else {}
    source_tag = "estimated_flow" if is_flow else "estimated_classic"

    def attach(meta):
        try:
            meta["motion_meta"] = applied_motion_meta_from_stabilization_warp(
                meta["stabilization_warp"], fps=fps_effective, source=source_tag
            )
        except (KeyError, TypeError, ValueError, np.linalg.LinAlgError):
            pass
        return meta

    def check():
        if interrupt_check is not None:
            interrupt_check()

    if shard is not None:
        total_frames = shard.total_frames
    estimation_steps = max(0, total_frames - 1)
    progress = _Progress(progress_bar, estimation_steps + total_frames)

    if total_frames == 1:
        meta = {
            "frames": 1,
            "note": "Single-frame input; bypassed stabilization.",
            "transform_mode": transform_mode,
            "framing_mode": framing_mode,
            **({"keep_fov_applied": False} if is_flow else {}),
            **flow_keys,
            "stabilization_warp": hm.build_stabilization_warp_meta(
                source_size=(width, height), output_size=(width, height), framing_mode=framing_mode,
                applied_matrices=[np.eye(3, dtype=np.float32)],
            ),
            "fps_requested": fps_requested,
            "fps_effective": fps_effective,
        }
        progress.finish()
        frames, masks = (context if shard is None else shard.owned_context(context)).untouched(output)
        return StabilizationResult(frames, masks, attach(meta))

    # ---- estimation: all candidate models of every pair -----------------------------------------
    work = hm.working_estimation_size(width, height)
    work_w, work_h = work if work is not None else (width, height)
    import time

    t0 = time.perf_counter()
    nvtx.phase("estimate")
    mark_dev = getattr(context, "device", None)
    _gpu_mark("start", mark_dev)
    cands = estimator(context, work_w, work_h, transform_mode)
    _gpu_mark("estimation kernels", mark_dev)
    t0 = _mark("estimate (gray+flow/track+fit enqueued)", t0)
    nvtx.phase("candidate_table")
    if shard is not None:
        cands = shard.gather_candidates(cands)
    _gpu_mark("candidate all-gather", mark_dev)
    if isinstance(cands, DeviceCandidates):
        cands = cands.to_host_words()
    t0 = _mark("candidate table on the host (waits for the GPU; all-gather when sharded)", t0)
    nvtx.phase("ladder+solve")
    base_mode = transform_mode
    use_native = _native_host_solve()
    solved = hm.native_trajectory(cands, transform_mode, (width, height), work) if use_native else None
    if solved is not None:
        # every pair accepted the requested model: per-pair float32 transforms at full size and the cumulative path in
        # one C pass; the ladder's meta columns are put together after the resampler is launched
        matrices, path = solved
        chosen, active_mode = None, transform_mode
    else:
        if isinstance(cands, HostWords):
            cands = cands.decode()
        chosen, active_mode, stacked = replay_mode_ladder(cands, transform_mode, with_residual=is_flow)
        if work is not None:
            stacked = hm.rescale_transforms_to_full(stacked, (width, height), work)
        matrices = stacked  # [P,3,3] float32, full-resolution per-pair transforms
        delta_params = hm.matrices_to_params(matrices, base_mode)
        path = np.zeros((total_frames, delta_params.shape[1]), dtype=np.float64)
        np.cumsum(delta_params, axis=0, out=path[1:])  # sequential adds, same as path[i] = path[i-1] + delta
    t0 = _mark("ladder", t0)
    progress.advance(estimation_steps)
    check()

    strength = hm.clip01(strength)
    smooth = hm.clip01(smooth)
    if camera_lock:
        smooth = max(smooth, 0.85)
    solved = hm.native_target(path, strength, smooth, fps_effective, camera_lock) if use_native else None
    target_path, diffs = solved if solved is not None else hm.numpy_target(path, strength, smooth, fps_effective, camera_lock)
    delta_full = diffs

    keep_fov_clamped = hm.clip01(keep_fov)
    keep_fov_applied = framing_mode == "crop" and keep_fov_clamped > 1e-6
    stabilization_scale = 1.0

    if framing_mode == "crop":
        from .crop import solve_crop_framing  # deferred: the crop solvers are their own module

        crop = solve_crop_framing(
            context, base_mode, delta_full, path, target_path, keep_fov_clamped, transform_mode, camera_lock,
            strength, smooth, fps_requested, fps_effective, padding_rgb, flow_keys, is_flow, attach, progress,
            check, output, shard=shard,
        )
        if isinstance(crop, StabilizationResult):
            return crop
        final_matrices, apply_matrices, crop_meta, stabilization_scale = crop
        output_size = (width, height)
    else:
        output_size = (width, height)
        crop_meta = None

    # bounding boxes of the frame corners under the applied matrices: box = (inner rectangle, union, all-affine flag)
    framed = hm.native_framing(delta_full, base_mode, width, height) if use_native and framing_mode != "crop" else None
    if framed is not None:
        apply_matrices, mins, maxs, box = framed
        final_matrices = apply_matrices
        inner = (box[0], box[1], box[2], box[3])
    else:
        box = None
        if framing_mode != "crop":
            apply_matrices = hm.params_to_matrices(delta_full, base_mode)  # [N,3,3] float32
            final_matrices = apply_matrices
        mins, maxs = hm.compute_bounding_boxes(apply_matrices, width, height)
        inner = hm.inner_rectangle(mins, maxs)
    framing_meta: Dict[str, Any] = {
        "mode": framing_mode,
        "input_size": [width, height],
        "padding_color_rgb": [int(c) for c in padding_rgb],
        "min_content_ratio": hm.min_content_ratio(mins, maxs, width, height, inner),
    }
    if framing_mode == "crop":
        framing_meta.update(crop_meta)
        if keep_fov_applied:
            framing_meta["keep_fov_requested"] = keep_fov_clamped
        note = framing_meta.pop("_keep_fov_note", None)
        if note:
            framing_meta["keep_fov_note"] = note
    elif framing_mode == "crop_and_pad":
        x0, y0, x1, y1 = (float(v) for v in inner)
        iw, ih = max(1.0, x1 - x0), max(1.0, y1 - y0)
        off_x = width * 0.5 - (x0 + x1) * 0.5
        off_y = height * 0.5 - (y0 + y1) * 0.5
        final_matrices = hm.translate_matrices(apply_matrices, off_x, off_y, affine=box is not None and box[8] == 1.0)
        framing_meta.update(
            {
                "safe_region_origin": [x0, y0],
                "safe_region_size": [iw, ih],
                "actual_content_ratio": min(iw / width, ih / height),
                "center_offset": [off_x, off_y],
            }
        )
    else:
        if box is not None:
            x_lo, y_lo, x_hi, y_hi = (float(v) for v in box[4:8])
            output_size = (max(int(math.ceil(x_hi - x_lo)), 1), max(int(math.ceil(y_hi - y_lo)), 1))
            final_matrices = hm.translate_matrices(apply_matrices, -x_lo, -y_lo, affine=box[8] == 1.0)
        else:
            shift, output_size = hm.prepare_expand_transform(mins, maxs)
            final_matrices = hm.left_multiply(shift, apply_matrices)
        framing_meta["expanded_size"] = list(output_size)

    # ---- warp + mask: one fused launch per chunk -------------------------------------------------
    lo, hi = (0, total_frames) if shard is None else shard.frame_range
    final_matrices = np.asarray(final_matrices, dtype=np.float32)
    fwd = final_matrices[lo:hi].reshape(-1, 1, 9)
    t0 = _mark("host path/framing solve", t0)
    nvtx.phase("resample")
    # sharded runs on GPUs: the per-frame padded-pixel counts are all-gathered device to device right behind the
    # resampler (they only feed the meta), and come back with the one copy that waits for it
    pad_gather = shard.device_pad_gather() if shard is not None else None
    _gpu_mark("idle while the host solves the trajectory", mark_dev)
    pending = fused_warp(
        context if shard is None else shard.owned_context(context), fwd, output_size, "bilinear",
        hm.border_value(padding_rgb), want_mask=True, want_pad_count=True, output=output, defer=True,
        **({"pad_transform": pad_gather} if pad_gather is not None else {}),
    )

    _gpu_mark("resampler + pad-count all-gather", mark_dev)
    # the kernels above are in flight: build the meta tree on the host meanwhile.  Sharded runs materialise
    # the per-frame lists only for shard.meta_frame_range (by default the rank's own frames: the per-frame
    # meta is sharded like the frames it describes, sharding.merge_sharded_meta() reassembles it); entries
    # keep their clip-wide indices.  Transitions go with the frame they end in.
    nvtx.phase("meta")
    if chosen is None:
        if isinstance(cands, HostWords):
            cands = cands.decode()
        chosen = accepted_ladder_entries(cands, transform_mode, with_residual=is_flow)
    effective_diffs = hm.matrices_to_params(np.asarray(apply_matrices), base_mode) if framing_mode == "crop" else np.array(delta_full)
    stabilization_scale = float(np.clip(stabilization_scale, 0.0, 1.0))
    strength_effective = strength * stabilization_scale
    effective_target_path = path + effective_diffs
    m_lo, m_hi = (0, total_frames) if shard is None else shard.meta_frame_range
    t_lo, t_hi = max(m_lo, 1) - 1, max(m_hi - 1, max(m_lo, 1) - 1)
    per_transition = []
    if t_hi > t_lo:
        mat_lists = matrices[t_lo:t_hi].tolist()
        columns = zip(chosen.modes[t_lo:t_hi], chosen.confidences[t_lo:t_hi], chosen.residuals[t_lo:t_hi], mat_lists)
        if is_flow:
            per_transition = [{"index": i, "mode": mode, "confidence": conf, "residual": res, "matrix": m}
                              for i, (mode, conf, res, m) in enumerate(columns, t_lo)]
        else:
            per_transition = [{"index": i, "mode": mode, "confidence": conf, "matrix": m}
                              for i, (mode, conf, _, m) in enumerate(columns, t_lo)]

    warp_meta = hm.build_stabilization_warp_meta(
        source_size=(width, height), output_size=output_size, framing_mode=framing_mode,
        applied_matrices=final_matrices[m_lo:m_hi], first_index=m_lo,
    )
    motion_block = None
    if m_hi > m_lo:
        try:
            motion_block = applied_motion_meta_from_matrices(
                final_matrices[m_lo:m_hi], source_size=(width, height), output_size=output_size, fps=fps_effective,
                source=source_tag, first_index=m_lo,
            )
        except (KeyError, TypeError, ValueError, np.linalg.LinAlgError):
            pass
    path_list, target_list, effective_list = (
        path[m_lo:m_hi].tolist(), target_path[m_lo:m_hi].tolist(), effective_target_path[m_lo:m_hi].tolist()
    )

    if gc_pause is not None:
        gc_pause.collect_young()  # the meta tree's generation-0 pass, while the resampler is still running
    t0 = _mark("meta build (overlaps the warp)", t0)
    nvtx.phase("finish")
    frames_out, masks_out, pad_counts = pending()
    t0 = _mark("wait for warp + pad counts", t0)
    if shard is not None:
        pad_counts = shard.unpack_pad_counts(pad_counts) if pad_gather is not None else shard.gather_pad_counts(pad_counts)
        t0 = _mark("all-gather pad counts", t0)
    pixels = int(output_size[0]) * int(output_size[1])
    pad_counts = np.asarray(pad_counts)
    padded_ratios = (pad_counts.astype(np.float32) / np.float32(pixels)).astype(np.float64)  # hm.padded_fractions as an array
    framing_meta["padding_detected"] = bool(pad_counts.max(initial=0) > 0)
    progress.advance(total_frames)
    check()

    meta = {
        "frames": total_frames,
        "transform_mode_requested": transform_mode,
        "transform_mode_applied": active_mode,
        "camera_lock": camera_lock,
        "strength": strength,
        "strength_effective": strength_effective,
        "smooth": smooth,
        "fps_requested": fps_requested,
        "fps_effective": fps_effective,
        "framing": framing_meta,
        "keep_fov_applied": keep_fov_applied,
        "padding_color_rgb": [int(c) for c in padding_rgb],
        **flow_keys,
        "stabilization_warp": warp_meta,
        "estimated_motion": {
            "per_transition": per_transition,
            "path": path_list,
            "target_path": target_list,
            "target_path_effective": effective_list,
        },
        "padding_fraction_mean": float(np.mean(padded_ratios)),
        "padding_fraction_max": float(np.max(padded_ratios)),
    }
    if motion_block is not None:
        meta["motion_meta"] = motion_block
    if shard is not None and (m_lo, m_hi) != (0, total_frames):
        meta["shard"] = {"rank": shard.rank, "world": shard.world, "frame_range": list(shard.frame_range),
                         "meta_frame_range": [m_lo, m_hi]}
    if output == "host":
        return StabilizationResult(frames_out.numpy(), masks_out.numpy()[..., None], meta)
    return StabilizationResult(frames_out, masks_out[..., None], meta)
