"""Frame-range sharding of one clip across the GPUs of a node (one process per GPU).

The reference is a single process; its only cross-frame couplings on this path are
  (1) the clip-wide sticky mode downgrade of the fit ladder (flow.py:324-339),
  (2) the O(N) trajectory smoothing / framing solve over ALL per-pair matrices (flow.py:356-533),
  (3) padding_fraction_mean/max over all frames (flow.py:636-637).
So the path shards by contiguous frame range with exactly one exchange: every rank estimates ALL
candidate models for its own pairs (it loads one halo frame before its range; no pixels cross
ranks), the per-pair candidate table (344 B per pair) is all-gathered over NCCL/NVLink, every rank
replays the ladder and the smoothing/framing solve redundantly and bit-identically, and warps only
its own frames.  A second tiny all-gather collects the per-frame padded-pixel counts for the meta.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .stabilizer_core import DeviceCandidates, HostWords, PairCandidates

TABLE_COLS = 43
META_OWN_RANGE, META_EVERY_RANK = -2, -1


def split_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous split of [0, total): the first `total % world` ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass
class FrameShard:
    rank: int
    world: int
    total_frames: int
    group: Optional[dist.ProcessGroup] = None
    device: Optional[torch.device] = None  # device of the communication buffers (cuda for NCCL)
    # Who materialises the per-frame lists of `meta` (tens of thousands of Python objects on long clips):
    #   META_OWN_RANGE  every rank the entries of its own frames -- the per-frame meta is sharded like the
    #                   frames it describes; merge_sharded_meta() concatenates the ranks' metas when the
    #                   whole tree is wanted (default: no rank carries O(total frames) of list building)
    #   META_EVERY_RANK every rank the whole clip
    #   r >= 0          rank r the whole clip, the others only the scalar part
    meta_rank: int = -2

    @property
    def builds_meta(self) -> bool:
        return self.meta_rank == META_EVERY_RANK or self.meta_rank == self.rank

    @property
    def meta_frame_range(self) -> Tuple[int, int]:
        """Frames whose per-frame meta entries this rank materialises (empty range: scalars only)."""
        if self.builds_meta:
            return 0, self.total_frames
        if self.meta_rank == META_OWN_RANGE:
            return self.frame_range
        return 0, 0

    @property
    def frame_range(self) -> Tuple[int, int]:
        return split_range(self.total_frames, self.world, self.rank)

    @property
    def load_range(self) -> Tuple[int, int]:
        """Frames this rank must hold: its own range plus one halo frame before it."""
        lo, hi = self.frame_range
        return max(lo - 1, 0), hi

    @property
    def pair_range(self) -> Tuple[int, int]:
        """Global pair indices (pair p = frames p, p+1) estimated by this rank: those ENDING in its range."""
        lo, hi = self.frame_range
        return max(lo, 1) - 1, max(hi - 1, max(lo, 1) - 1)

    def pair_counts(self) -> List[int]:
        out = []
        for r in range(self.world):
            lo, hi = split_range(self.total_frames, self.world, r)
            a = max(lo, 1) - 1
            out.append(max(hi - 1, a) - a)
        return out

    # -- hooks used by stabilizer_core.stabilize_frames ------------------------------------------
    @property
    def on_gpu(self) -> bool:
        return self.device is not None and self.device.type == "cuda"

    # Send buffer of the candidate all-gather, float64 words: cap x 36 table words (the fit kernels write them there
    # themselves), then min_points, a "has detected counts" flag and cap per-pair detected counts (Classic).
    def _send_words(self, cap: int) -> int:
        return cap * 36 + 2 + cap

    def wrap_estimator(self, estimate: Callable) -> Callable:
        """The local context holds frames load_range; every consecutive local pair is ours.  On GPUs the fit
        kernels write their result table straight into this rank's all-gather send buffer (no staging copy)."""

        def run(context, work_w, work_h, requested_mode):
            n_local = max(len(context) - 1, 0)
            if self.on_gpu:
                cap = max(max(self.pair_counts()), 1)
                send = torch.zeros((self._send_words(cap),), dtype=torch.float64, device=self.device)
                table = send[: cap * 36].view(cap, 3, 12)
                if n_local == 0:
                    return DeviceCandidates(table[:0], min_points=0, send=send)
                cands = estimate(context, work_w, work_h, requested_mode, out_raw=table[:n_local])
                if isinstance(cands, DeviceCandidates):
                    cands.send = send
                return cands
            if n_local == 0:
                z = np.zeros
                return PairCandidates(z((0, 3, 3, 3)), z((0, 3)), z((0, 3), int), z((0, 3), int), z((0, 3), int), z((0, 3), int))
            return estimate(context, work_w, work_h, requested_mode)

        return run

    def _all_gather_rows(self, local: np.ndarray, counts: List[int]) -> np.ndarray:
        width = local.shape[1] if local.ndim == 2 else 1
        cap = max(max(counts), 1)
        dev = self.device if self.device is not None else torch.device("cpu")
        send = torch.zeros((cap, width), dtype=torch.float64, device=dev)
        if local.shape[0]:
            send[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local.reshape(local.shape[0], width))).to(dev)
        recv = torch.empty((self.world, cap, width), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(recv.view(-1, width), send, group=self.group)
        host = recv.cpu().numpy()
        return np.concatenate([host[r, : counts[r]] for r in range(self.world)], axis=0)

    def _gather_device_table(self, local: DeviceCandidates) -> "HostWords":
        """ONE all-gather (NCCL, device to device) of the fit kernels' output plus the few words every rank must agree
        on, then ONE blocking copy of the gathered buffer (~300 B per pair) to the host."""
        counts = self.pair_counts()
        n_local = int(local.raw.shape[0])
        if n_local != counts[self.rank]:
            raise RuntimeError(f"rank {self.rank}: estimated {n_local} pairs, expected {counts[self.rank]}")
        cap = max(max(counts), 1)
        words = self._send_words(cap)
        send = local.send
        if send is None or send.numel() != words:
            send = torch.zeros((words,), dtype=torch.float64, device=self.device)
            send[: cap * 36].view(cap, 3, 12)[:n_local] = local.raw
        # min_points and the presence of `detected` agree on every rank that has pairs (same node); a rank without
        # pairs cannot know them, so the host takes the maximum over ranks
        tail = [float(local.min_points), 1.0 if local.detected is not None else 0.0]
        send[cap * 36 : cap * 36 + 2] = torch.tensor(tail, dtype=torch.float64).to(self.device, non_blocking=True)
        if local.detected is not None and n_local:
            send[cap * 36 + 2 : cap * 36 + 2 + n_local] = local.detected.to(torch.float64)
        recv = torch.empty((self.world * words,), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        host = recv.cpu().numpy().reshape(self.world, words)
        table = np.concatenate([host[r, : cap * 36].reshape(cap, 3, 12)[: counts[r]] for r in range(self.world)], axis=0)
        min_points = int(host[:, cap * 36].max())
        detected = None
        if host[:, cap * 36 + 1].max() > 0:
            detected = np.rint(np.concatenate([host[r, cap * 36 + 2 : cap * 36 + 2 + counts[r]] for r in range(self.world)])).astype(np.int64)
        # the words as they are: the trajectory helper reads them directly, the columns are decoded behind the resampler launch
        return HostWords(table, min_points, detected)

    def device_pad_gather(self):
        """fused_warp hook (pad_transform): all-gather the per-frame padded-pixel counts device to device on the
        stream that just launched the resampler.  None on the CPU (gloo tests use gather_pad_counts)."""
        import os

        if not self.on_gpu or os.environ.get("VSTAB_PAD_GATHER") == "host":  # "host": A/B switch, gather after the download
            return None
        sizes = [b - a for a, b in (split_range(self.total_frames, self.world, r) for r in range(self.world))]
        cap = max(max(sizes), 1)

        def gather(pad: torch.Tensor) -> torch.Tensor:
            send = torch.zeros((cap,), dtype=torch.int32, device=self.device)
            send[: pad.shape[0]] = pad
            recv = torch.empty((self.world * cap,), dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            return recv

        return gather

    def unpack_pad_counts(self, gathered: np.ndarray) -> np.ndarray:
        """Host view of device_pad_gather's result: [world * cap] -> clip-wide [total_frames] int64."""
        sizes = [b - a for a, b in (split_range(self.total_frames, self.world, r) for r in range(self.world))]
        cap = max(max(sizes), 1)
        g = np.asarray(gathered).reshape(self.world, cap)
        return np.concatenate([g[r, : sizes[r]] for r in range(self.world)]).astype(np.int64)

    def gather_candidates(self, local):
        if isinstance(local, DeviceCandidates):
            return self._gather_device_table(local)
        counts = self.pair_counts()
        if local.matrix.shape[0] != counts[self.rank]:
            raise RuntimeError(f"rank {self.rank}: estimated {local.matrix.shape[0]} pairs, expected {counts[self.rank]}")
        table = self._all_gather_rows(local.to_array() if counts[self.rank] else np.zeros((0, TABLE_COLS)), counts)
        return PairCandidates.from_array(table, local.min_points)

    def owned_context(self, context):
        """Drop the halo frame so frame k of the result is global frame frame_range[0] + k."""
        lo, _ = self.frame_range
        halo = lo - self.load_range[0]
        if halo == 0:
            return context
        return context.sliced(halo)

    def gather_pad_counts(self, local_counts) -> np.ndarray:
        counts = [split_range(self.total_frames, self.world, r) for r in range(self.world)]
        sizes = [hi - lo for lo, hi in counts]
        rows = np.asarray(local_counts, dtype=np.float64).reshape(-1, 1)
        return np.rint(self._all_gather_rows(rows, sizes)[:, 0]).astype(np.int64)


def merge_sharded_meta(metas: List[dict]) -> dict:
    """The single-process `meta` from the metas of all ranks (rank order) of a META_OWN_RANGE run:
    scalars from rank 0, per-frame / per-transition lists concatenated, motion_meta.frame_count restored."""
    import copy

    def covers_frames(m):
        lo, hi = m.get("shard", {}).get("meta_frame_range", (0, 1))
        return hi > lo

    metas = [m for m in metas if covers_frames(m)] or metas[:1]  # ranks without frames contribute nothing
    out = copy.deepcopy(metas[0])
    out.pop("shard", None)
    for m in metas[1:]:
        out["stabilization_warp"]["per_frame"].extend(copy.deepcopy(m["stabilization_warp"]["per_frame"]))
        est, src = out["estimated_motion"], m["estimated_motion"]
        for key in ("per_transition", "path", "target_path", "target_path_effective"):
            est[key].extend(copy.deepcopy(src[key]))
        if "motion_meta" in out and "motion_meta" in m:
            out["motion_meta"]["per_frame"].extend(copy.deepcopy(m["motion_meta"]["per_frame"]))
        elif "motion_meta" in out:
            del out["motion_meta"]  # a rank could not build its block: the merged tree has none either
    if "motion_meta" in out:
        out["motion_meta"]["frame_count"] = len(out["motion_meta"]["per_frame"])
    return out


def init_from_env(total_frames: int, backend: Optional[str] = None) -> FrameShard:
    """torchrun-style bootstrap (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    return FrameShard(rank, world, total_frames, None, torch.device("cuda", local) if use_cuda else torch.device("cpu"))
