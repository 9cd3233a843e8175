"""Clip container and host<->HBM plumbing for the node layer.

The reference keeps a ``VideoContext`` holding a list of HWC float32 numpy frames
(nodes/stabilizer_utils.py:62-72) filled by ``_normalize_video_input`` (:150-197) and packs the
results back with ``_reconstruct_video`` / ``_convert_masks_for_output`` (:200-221, :1055-1077).
Here the clip lives in HBM as one ``[N,H,W,3]`` float32 tensor; the adapter rules (uint8 or
0..255 floats are divided by 255 per frame, CHW frames are transposed, 1 channel is repeated,
alpha is dropped, dict inputs carry fps) are kept.  torch is used for device memory, streams and
the copies only -- every pixel operation of the path runs in libvstab kernels.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Literal, Optional, Tuple

import numpy as np
import torch

from . import _native

# frames per host->device / device->host copy chunk (keeps staging buffers ~0.8 GB at 1080p)
CHUNK_BYTES = 768 << 20
# bytes per cudaMemcpyAsync of a page-locked clip on its way up (scripts/e2e_chunks.py: A/B of the enqueue behaviour)
UPLOAD_CHUNK_BYTES = CHUNK_BYTES

# bench.py sets this to a list to collect (start_event, end_event, frames, algorithmic_bytes) of
# every vstab_warp_fused launch, recorded on the launching stream.
WARP_LAUNCH_LOG = None


def _timed_warp(h, src, fwd, out_size, interp, border, **kw):
    if WARP_LAUNCH_LOG is None:
        return h.warp_fused(src, fwd, out_size, interp, border, **kw)
    st = torch.cuda.current_stream(src.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    res = h.warp_fused(src, fwd, out_size, interp, border, **kw)
    e1.record(st)
    n, sh, sw, _ = src.shape
    ow, oh = int(out_size[0]), int(out_size[1])
    WARP_LAUNCH_LOG.append((e0, e1, int(n), int(n) * (12 * sh * sw + 12 * oh * ow + (4 * oh * ow if kw.get("want_mask", True) else 0))))
    return res


@dataclass
class FrameAdapter:
    dtype: Any
    channel_first: bool
    value_range: Literal["0_1", "0_255"]
    origin: Literal["numpy", "torch"]
    squeeze_last_dim: bool


@dataclass
class VideoContext:
    """Clip resident in HBM (``frames``), or -- when it is too large for that -- a host tensor that is
    streamed through the device twice (``host``: first pass for the gray working images, second pass
    for the resampler).  Both kinds hand out normalised device chunks."""

    frames: Optional[torch.Tensor]  # [N,H,W,3] float32, CUDA, values 0..1 (None when streamed)
    adapter: FrameAdapter
    width: int
    height: int
    channels: int
    fps: Optional[float]
    template_kind: Literal["dict", "sequence"]
    template_meta: Dict[str, Any] = field(default_factory=dict)
    host: Optional[torch.Tensor] = None  # [N,H,W,C] float32 / uint8 CPU tensor (streamed clips only)
    stream_device: Optional[torch.device] = None
    # float32 clip whose per-frame range rule (max > 1.5 => / 255) has not run yet: the estimation's luma kernel applies
    # it in the same read of the source (gray_working below); paths without an estimation pass call settle_range()
    range_pending: bool = False

    def __len__(self) -> int:
        return int((self.frames if self.frames is not None else self.host).shape[0])

    @property
    def streamed(self) -> bool:
        return self.frames is None

    @property
    def device(self) -> torch.device:
        return self.frames.device if self.frames is not None else self.stream_device

    def sliced(self, start: int) -> "VideoContext":
        """The same clip without its first `start` frames (frame-range shards drop their halo frame)."""
        import dataclasses

        if self.frames is not None:
            return dataclasses.replace(self, frames=self.frames[start:])
        return dataclasses.replace(self, host=self.host[start:])

    def settle_range(self) -> None:
        """Apply a pending range rule now (own kernels, in place; no host round trip)."""
        if self.range_pending and self.frames is not None:
            _native.get_handle(self.frames.device).range_normalize(self.frames)
        self.range_pending = False

    def chunk(self, a: int, b: int) -> torch.Tensor:
        """Frames [a, b) as a normalised [b-a,H,W,3] float32 device tensor."""
        if self.frames is not None:
            self.settle_range()
            return self.frames[a:b]
        dev = self.host[a:b].to(self.stream_device, non_blocking=True)
        return _normalize_device(dev)[0]

    def chunk_frames(self) -> int:
        return frames_per_chunk(self.height, self.width, 3)

    def untouched(self, output: str) -> Tuple[Any, Any]:
        """(frames, all-zero masks) for the paths that hand the input back unchanged
        (keep_fov bypass, single frame): device tensors, or numpy arrays for output="host"."""
        n = len(self)
        self.settle_range()
        if self.streamed:
            if output != "host":
                raise _native.VstabNativeError("a streamed clip can only be returned with output='host'")
            out = torch.empty((n, self.height, self.width, 3), dtype=torch.float32)
            step = self.chunk_frames()
            for a in range(0, n, step):
                out[a : a + step].copy_(self.chunk(a, min(a + step, n)))
            return out.numpy(), np.zeros((n, self.height, self.width, 1), dtype=np.float32)
        masks = torch.zeros((n, self.height, self.width, 1), dtype=torch.float32, device=self.frames.device)
        if output == "host":
            return self.frames.cpu().numpy(), masks.cpu().numpy()
        return self.frames, masks


def gray_working(context: VideoContext, size: Tuple[int, int], first: int = 0, last: Optional[int] = None) -> torch.Tensor:
    """K1+K2 over frames [first, last) of the clip: [n,h,w] uint8 on the device.  A streamed clip is
    uploaded chunk by chunk; only the working images (<= 960 px) stay resident."""
    h = _native.get_handle(context.device)
    last = len(context) if last is None else last
    if not context.streamed:
        if context.range_pending:
            if (first, last) == (0, len(context)):
                # the adapter's `max > 1.5 => / 255` rides on the luma kernel's read of the source
                gray, _flags = h.gray_working_adapt(context.frames, size)
                context.range_pending = False
                return gray
            context.settle_range()
        return h.gray_working(context.frames[first:last], size)
    step = context.chunk_frames()
    parts = [h.gray_working(context.chunk(a, min(a + step, last)), size) for a in range(first, last, step)]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


def _require_device(device) -> torch.device:
    if not torch.cuda.is_available():
        raise _native.VstabNativeError(
            "no CUDA device visible: the stabilizer path runs on a B200 only; there is no CPU fallback"
        )
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def frames_per_chunk(height: int, width: int, channels: int = 3) -> int:
    return max(1, CHUNK_BYTES // (height * width * channels * 4))


_BOUNCE: Dict[Any, List[torch.Tensor]] = {}
BOUNCE_BYTES = 256 << 20


def _bounce_buffers(nbytes: int) -> List[torch.Tensor]:
    """Two pinned staging buffers per process (grow-only), for host tensors that are not page-locked."""
    bufs = _BOUNCE.get("up")
    if bufs is None or bufs[0].numel() < nbytes:
        bufs = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        _BOUNCE["up"] = bufs
    return bufs


def _upload_batched(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    """CPU [N,H,W,C] tensor -> device, in chunks.

    Pinned sources go up with one asynchronous copy per chunk.  A PAGEABLE source -- what ComfyUI hands a node --
    would make every cudaMemcpyAsync a synchronous, single-threaded staging copy inside the driver; instead each chunk
    is copied into one of two pinned bounce buffers by torch's multi-threaded host copy while the DMA of the previous
    chunk is still in flight, so the upload runs at min(host memcpy rate, link rate)."""
    if t.is_cuda:
        return t.to(device)
    n = t.shape[0]
    out = torch.empty(t.shape, dtype=t.dtype, device=device)
    if torch.device(device).type != "cuda":  # host-side tests drive the adapters without a GPU
        out.copy_(t)
        return out
    frame_bytes = max(1, t[0].numel() * t.element_size())
    if t.is_pinned():
        step = max(1, UPLOAD_CHUNK_BYTES // frame_bytes)
        for a in range(0, n, step):
            out[a : a + step].copy_(t[a : a + step], non_blocking=True)
        return out
    step = max(1, BOUNCE_BYTES // frame_bytes)
    bufs = _bounce_buffers(step * frame_bytes)
    stream = torch.cuda.current_stream(device)
    busy: List[Optional[torch.cuda.Event]] = [None, None]
    for k, a in enumerate(range(0, n, step)):
        b = min(a + step, n)
        slot = k & 1
        if busy[slot] is not None:
            busy[slot].synchronize()  # the DMA that last read this buffer is done
        stage = bufs[slot][: (b - a) * frame_bytes].view(t.dtype).view((b - a,) + tuple(t.shape[1:]))
        stage.copy_(t[a:b])  # host -> pinned, torch's parallel copy
        out[a:b].copy_(stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        busy[slot] = ev
    return out


def bind_host_to_gpu(local_rank: int, world: int) -> Dict[str, Any]:
    """One-process-per-GPU deployments: run this rank's host threads on the cores of the GPU's NUMA node (a fair share
    of them when several ranks share a node) and prefer that node's memory for the pinned staging buffers, so that
    uploads and downloads do not cross the socket interconnect.  Best effort -- containers often pin the cpuset or hide
    sysfs; the returned dict says what was done.  Never called by the nodes themselves (a ComfyUI host owns its own
    affinity); bench.py and torchrun-style launchers call it once per rank."""
    import os

    info: Dict[str, Any] = {"numa_node": None, "cpus": None, "mempolicy": None}
    try:
        allowed = sorted(os.sched_getaffinity(0))
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = -1
        try:
            with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
                node = int(fh.read().strip())
        except OSError:
            pass
        info["pci"] = bdf
        cpus = allowed
        if node >= 0:
            info["numa_node"] = node
            try:
                with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
                    local = set()
                    for part in fh.read().strip().split(","):
                        lo, _, hi = part.partition("-")
                        local.update(range(int(lo), int(hi or lo) + 1))
                near = [c for c in allowed if c in local]
                if near:
                    cpus = near
            except OSError:
                pass
        if world > 1:
            # ranks whose GPUs sit on the same node share its cores evenly; without NUMA information all ranks
            # share the allowed set the same way (keeps N ranks x intra-op threads from oversubscribing the box)
            share = max(1, len(cpus) // world) if cpus is allowed or node < 0 else max(1, len(cpus) * 2 // world)
            start = (local_rank * share) % max(len(cpus) - share + 1, 1)
            cpus = cpus[start : start + share]
        if cpus and set(cpus) != set(allowed):
            os.sched_setaffinity(0, cpus)
        torch.set_num_threads(max(1, min(len(cpus), 16)))
        info["cpus"] = f"{cpus[0]}-{cpus[-1]} ({len(cpus)})" if cpus else None
        if node >= 0:
            import ctypes

            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            rc = libc.syscall(238, 1, ctypes.byref(mask), 16 * 64 + 1)  # set_mempolicy(MPOL_PREFERRED, node)
            info["mempolicy"] = "preferred" if rc == 0 else f"unchanged (errno {ctypes.get_errno()})"
    except Exception as exc:  # pragma: no cover - depends on the box
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info


def _frame_to_hwc(arr: np.ndarray):
    """Per-frame layout rules of _to_numpy_frame (stabilizer_utils.py:96-147), host slow path."""
    channel_first = False
    squeeze = False
    if arr.ndim == 3 and arr.shape[0] in (1, 3, 4) and arr.shape[0] < arr.shape[-1]:
        channel_first = True
        arr = np.moveaxis(arr, 0, -1)
    elif arr.ndim == 4 and arr.shape[0] == 1:
        arr = arr[0]
    if arr.ndim == 2:
        arr = arr[..., None]
        squeeze = True
    elif arr.ndim == 3 and arr.shape[2] == 1:
        squeeze = True
    return arr, channel_first, squeeze


def _ensure_rgb_host(arr: np.ndarray) -> np.ndarray:
    c = arr.shape[2]
    if c == 1:
        return np.repeat(arr, 3, axis=2)
    if c > 3:
        return arr[..., :3]
    return arr


def _normalize_device(dev: torch.Tensor, source_ptr: Optional[int] = None, defer_range: bool = False) -> Tuple[torch.Tensor, str]:
    """Adapter rules of _to_numpy_frame / _ensure_rgb (stabilizer_utils.py:96-147) on a device chunk
    [n,H,W,C] (float32 or uint8): per-frame `max > 1.5 => /255`, 1 channel repeated, alpha dropped.
    defer_range: leave the range rule of a float32 CUDA clip to the estimation's luma kernel (value_range "deferred")."""
    device = dev.device
    # IEEE division by a TENSOR 255 (torch turns division by a Python scalar into a multiplication
    # by the reciprocal on CUDA, which is 1 ulp off numpy's `arr /= 255.0` for some values)
    div255 = torch.full((), 255.0, dtype=torch.float32, device=device)
    if dev.dtype == torch.uint8:
        frames = torch.div(dev.to(torch.float32), div255)
        value_range = "0_255"
    elif dev.is_cuda:
        # float32 on the device: the range rule runs in libvstab (vstab_range_normalize), in place on the uploaded copy
        frames = dev.contiguous()
        if source_ptr is not None and frames.data_ptr() == source_ptr:
            frames = frames.clone()  # the caller handed us its own CUDA tensor: never write into it
        if defer_range and frames.shape[3] == 3:  # the reference's max() runs over ALL channels, alpha included
            value_range = "deferred"
        else:
            flags = _native.get_handle(frames.device).range_normalize(frames)
            value_range = "0_255" if int(flags[0]) == 1 else "0_1"
    else:
        frames = dev  # host tensors (CPU-side tests of the adapter rules)
        peaks = frames.reshape(frames.shape[0], -1).amax(dim=1)
        big = peaks > 1.5
        value_range = "0_255" if bool(big[0]) else "0_1"
        if bool(big.any()):
            if source_ptr is not None and frames.data_ptr() == source_ptr:
                frames = frames.clone()
            frames[big] = torch.div(frames[big], div255)
    if frames.shape[3] == 1:
        frames = frames.expand(-1, -1, -1, 3)
    elif frames.shape[3] > 3:
        frames = frames[..., :3]
    elif frames.shape[3] == 2:
        raise ValueError("2-channel frames are not supported.")
    return frames.contiguous(), value_range


def _must_stream(n: int, height: int, width: int, device: torch.device) -> bool:
    """True when the float32 RGB clip should not be made resident: it would not leave room for the
    staging buffers of a host-output run.  VSTAB_RESIDENT_LIMIT_MB overrides the budget (tests)."""
    import os

    need = n * height * width * 3 * 4
    limit = os.environ.get("VSTAB_RESIDENT_LIMIT_MB")
    if limit is not None:
        return need > int(limit) << 20
    free, _total = torch.cuda.mem_get_info(device)
    return need > 0.7 * free - 4 * CHUNK_BYTES


def normalize_video_input(value: Any, device=None, defer_range: bool = False) -> VideoContext:
    """ComfyUI IMAGE (or list / dict of frames) -> clip resident in HBM.

    Fast path: a 4-D torch tensor [B,H,W,C] (float32 / uint8) is uploaded as a whole and the
    per-frame range test (max > 1.5 => /255) runs on the device.  Everything else (lists of
    numpy / CHW / mixed frames) is normalised frame by frame on the host first.
    """
    device = _require_device(device)
    if isinstance(value, dict):
        seq = None
        for key in ("frames", "images", "video"):
            if key in value:
                seq = value[key]
                break
        if seq is None:
            raise ValueError("Video input dictionary must contain 'frames'.")
        kind: Literal["dict", "sequence"] = "dict"
        extra = {k: v for k, v in value.items() if k not in ("frames", "images", "video")}
        fps = extra.get("fps")
    else:
        seq, kind, extra, fps = value, "sequence", {}, None

    fast = (
        isinstance(seq, torch.Tensor)
        and seq.dim() == 4
        and seq.shape[0] > 0
        and seq.dtype in (torch.float32, torch.uint8)
        # a [B,H,W,C] tensor whose frames would not be mistaken for CHW by the reference
        and not (seq.shape[1] in (1, 3, 4) and seq.shape[1] < seq.shape[3])
    )
    if fast:
        origin_dtype = np.uint8 if seq.dtype == torch.uint8 else np.float32
        squeeze = seq.shape[3] == 1
        if seq.shape[3] == 2:
            raise ValueError("2-channel frames are not supported.")
        n, hh, ww = int(seq.shape[0]), int(seq.shape[1]), int(seq.shape[2])
        if not seq.is_cuda and _must_stream(n, hh, ww, device):
            # value range of the adapter = what the reference decides on the first frame
            first = seq[0]
            value_range = "0_255" if (first.dtype == torch.uint8 or float(first.max()) > 1.5) else "0_1"
            adapter = FrameAdapter(origin_dtype, False, value_range, "torch", bool(squeeze))
            return VideoContext(None, adapter, ww, hh, 3, fps, kind, extra, host=seq.contiguous(), stream_device=device)
        dev = _upload_batched(seq.contiguous(), device)
        frames, value_range = _normalize_device(dev, source_ptr=seq.data_ptr(), defer_range=defer_range)
        pending = value_range == "deferred"
        adapter = FrameAdapter(origin_dtype, False, "0_1" if pending else value_range, "torch", bool(squeeze))
    else:
        host: List[np.ndarray] = []
        adapter = None
        for frame in seq:
            origin = "torch" if isinstance(frame, torch.Tensor) else "numpy"
            arr = frame.detach().cpu().numpy() if origin == "torch" else np.asarray(frame)
            arr, channel_first, squeeze = _frame_to_hwc(arr)
            dtype = arr.dtype
            if dtype == np.uint8 or (arr.size and float(arr.max()) > 1.5):
                arr = arr.astype(np.float32)
                arr /= 255.0
                value_range = "0_255"
            else:
                arr = np.ascontiguousarray(arr, dtype=np.float32)
                value_range = "0_1"
            this = FrameAdapter(dtype, channel_first, value_range, origin, squeeze)
            if adapter is None:
                adapter = this
            elif this.channel_first != adapter.channel_first or this.origin != adapter.origin:
                raise ValueError("Mixed tensor layouts within the same video sequence are not supported.")
            host.append(_ensure_rgb_host(arr))
        if not host:
            raise ValueError("The input video sequence is empty.")
        if host[0].shape[2] != 3:
            raise ValueError("2-channel frames are not supported.")
        stacked = torch.from_numpy(np.ascontiguousarray(np.stack(host, axis=0), dtype=np.float32))
        frames = _upload_batched(stacked, device)

    n, h, w, c = frames.shape
    ctx = VideoContext(frames, adapter, int(w), int(h), int(c), fps, kind, extra)
    ctx.range_pending = bool(fast and pending)
    return ctx


def download(t: torch.Tensor, pin: bool = True) -> torch.Tensor:
    """Device tensor -> CPU tensor (pinned when possible), chunked along dim 0."""
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin and torch.cuda.is_available())
    n = t.shape[0]
    if n == 0:
        return out
    step = max(1, CHUNK_BYTES // max(1, t[0].numel() * t.element_size()))
    for a in range(0, n, step):
        out[a : a + step].copy_(t[a : a + step], non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out


def reconstruct_video(frames: Any, context: VideoContext) -> Any:
    """[N,H',W',3] float32 CPU tensor (or dict payload) following ComfyUI conventions."""
    if isinstance(frames, np.ndarray):
        frames = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.float32))
    if frames.shape[0] == 0:
        frames = torch.zeros((1, context.height, context.width, 3), dtype=torch.float32)
    if frames.is_cuda:
        frames = download(frames)
    if context.template_kind == "dict":
        payload = dict(context.template_meta)
        payload["frames"] = frames
        return payload
    return frames


def convert_masks_for_output(masks: Any) -> torch.Tensor:
    """[N,H',W'] float32 CPU tensor, 1 = padding; empty -> zeros [1,1,1]."""
    if isinstance(masks, np.ndarray):
        masks = torch.from_numpy(np.ascontiguousarray(masks, dtype=np.float32))
    if masks.shape[0] == 0:
        return torch.zeros((1, 1, 1), dtype=torch.float32)
    if masks.dim() == 4:
        masks = masks[..., 0]
    if masks.is_cuda:
        masks = download(masks)
    return masks.contiguous()


PINNED_RESULT_LIMIT = 64 << 30  # VSTAB_PINNED_RESULT_LIMIT_MB overrides

# bytes the last fused_warp(output="host") call copied device -> host (frames + mask as it crossed the link); bench.py reads it
LAST_D2H_BYTES = 0


def mask_bytes_enabled() -> bool:
    """Binary masks travel device -> host as uint8 unless VSTAB_MASK_BYTES=0 (A/B switch; the results are identical)."""
    import os

    return os.environ.get("VSTAB_MASK_BYTES", "1") != "0"


def _host_result(shape) -> torch.Tensor:
    """Host tensor a result is downloaded into: page-locked (the download then runs at link rate) unless the clip is
    larger than the pinned budget or the allocation fails -- then plain pageable memory, like the reference's numpy
    result (the copies still work, the driver stages them)."""
    import os

    nbytes = 4
    for d in shape:
        nbytes *= int(d)
    limit = os.environ.get("VSTAB_PINNED_RESULT_LIMIT_MB")
    limit = int(limit) << 20 if limit is not None else PINNED_RESULT_LIMIT
    if nbytes <= limit:
        try:
            return torch.empty(shape, dtype=torch.float32, pin_memory=True)
        except RuntimeError:
            pass
    return torch.empty(shape, dtype=torch.float32)


_POOL: Dict[Any, torch.Tensor] = {}


def release_pool() -> None:
    """Give the pooled device result buffers (output="device") back to torch's allocator."""
    _POOL.clear()


def _pooled(shape, device, tag: str) -> torch.Tensor:
    """Device-resident result buffers are reused from call to call (output="device" hands out views
    of them: consume a result before asking for the next one).  Avoids re-allocating gigabytes per
    clip; host results always get fresh pinned tensors."""
    key = (tag, tuple(shape), str(device))
    buf = _POOL.get(key)
    if buf is None:
        for k in [k for k in _POOL if k[0] == tag and k[2] == str(device)]:
            del _POOL[k]
        buf = torch.empty(shape, dtype=torch.float32, device=device)
        _POOL[key] = buf
    return buf


def fused_warp(
    context: VideoContext,
    fwd: np.ndarray,
    out_size: Tuple[int, int],
    interpolation: str,
    border: Tuple[float, float, float],
    *,
    want_mask: bool = True,
    want_pad_count: bool = False,
    mask_rule: Optional[int] = None,
    output: Literal["host", "device"] = "host",
    defer: bool = False,
    pad_transform=None,
):
    """Run the fused resampler over the whole clip.

    pad_transform: device -> device callable applied to the [N] int32 padded-pixel counts right after the last launch
    (frame-range shards all-gather them there); the returned counts are then whatever it produced.

    With defer=True everything is enqueued and a callable is returned; calling it waits for the
    GPU and yields (frames, masks, pad_counts).  The caller can build its host-side results while
    the kernels and copies are in flight.


    fwd: [N,S,9] float32 forward matrices (host).  Returns (frames, masks, pad_counts) where
    frames is [N,H',W',3] and masks [N,H',W'] (None when want_mask is False).  With
    output="host" the clip is processed in chunks: while chunk k is copied device->host on a
    side stream, chunk k+1 is being resampled; the results land in pinned CPU tensors.
    """
    h = _native.get_handle(context.device)
    dev = context.device
    n = len(context)
    ow, oh = int(out_size[0]), int(out_size[1])
    if mask_rule is None:
        mask_rule = _native.default_mask_rule()  # Rule P unless VSTAB_MASK_RULE says otherwise
    fwd_t = torch.from_numpy(np.ascontiguousarray(fwd, dtype=np.float32)).to(dev, non_blocking=True)
    if fwd_t.dim() == 2:
        fwd_t = fwd_t.view(n, 1, 9)
    if output == "device" and context.streamed:
        raise _native.VstabNativeError(
            "this clip is streamed through the device (it does not fit in HBM next to its results): ask for output='host'"
        )
    if output == "device":
        dst_buf = _pooled((n, oh, ow, 3), dev, "warp_dst")
        if context.frames.untyped_storage().data_ptr() == dst_buf.untyped_storage().data_ptr():
            # the clip IS the previous result (e.g. Flow with output="device" fed into Motion Apply at the same
            # size): resampling it into the pooled buffer would read and write the same memory
            dst_buf = torch.empty((n, oh, ow, 3), dtype=torch.float32, device=dev)
        mask_buf = _pooled((n, oh, ow), dev, "warp_mask") if want_mask else None
        dst, mask, pad = _timed_warp(
            h, context.frames, fwd_t, (ow, oh), interpolation, border,
            mask_rule=mask_rule, want_mask=want_mask, want_pad_count=want_pad_count, out=dst_buf, mask_out=mask_buf,
        )

        if pad is not None and pad_transform is not None:
            pad = pad_transform(pad)

        def finish_device():
            return dst, mask, (pad.cpu().numpy().astype(np.int64) if pad is not None else None)

        return finish_device if defer else finish_device()

    frames_cpu = _host_result((n, oh, ow, 3))
    masks_cpu = _host_result((n, oh, ow)) if want_mask else None
    # a single-sample mask is exactly 0 / 1: it crosses the link as bytes (a quarter of the float32 MASK, ~10 % of the
    # whole transfer) and is widened into masks_cpu by the host while the frames are still being copied
    pack_mask = want_mask and fwd_t.shape[1] == 1 and mask_bytes_enabled()
    pads: List[torch.Tensor] = []
    step = frames_per_chunk(oh, ow, 4)
    if context.streamed:
        step = min(step, context.chunk_frames())
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    mask_bytes_cpu = None
    if pack_mask:
        try:  # page-locked like the results; when the host cannot pin it, the mask travels as float32 like before
            mask_bytes_cpu = torch.empty((n, oh, ow), dtype=torch.uint8, pin_memory=True)
        except RuntimeError:
            pack_mask = False
    bufs = [
        (
            torch.empty((min(step, n), oh, ow, 3), dtype=torch.float32, device=dev),
            torch.empty((min(step, n), oh, ow), dtype=torch.float32, device=dev) if want_mask else None,
            torch.empty((min(step, n), oh, ow), dtype=torch.uint8, device=dev) if pack_mask else None,
        )
        for _ in range(2 if n > step else 1)
    ]
    odd = torch.zeros((1,), dtype=torch.int32, device=dev) if pack_mask else None
    mask_ready: List[Tuple[int, int, torch.cuda.Event]] = []
    copied = [None] * len(bufs)
    d2h_bytes = 0
    for k, a in enumerate(range(0, n, step)):
        b = min(a + step, n)
        dbuf, mbuf, ubuf = bufs[k % len(bufs)]
        if copied[k % len(bufs)] is not None:
            main.wait_event(copied[k % len(bufs)])  # buffer free again
        # streamed clips: the upload of chunk k+1 (this stream) overlaps the download of chunk k (copy stream)
        dst, mask, pad = _timed_warp(
            h, context.chunk(a, b), fwd_t[a:b], (ow, oh), interpolation, border,
            mask_rule=mask_rule, want_mask=want_mask, want_pad_count=want_pad_count,
            out=dbuf[: b - a], mask_out=(mbuf[: b - a] if mbuf is not None else None),
        )
        if pack_mask:
            h.mask_pack_u8(mask, ubuf[: b - a], odd)
        if pad is not None:
            pads.append(pad)
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            if pack_mask:  # the mask first: the host widens it while the frames of the chunk are on the link
                mask_bytes_cpu[a:b].copy_(ubuf[: b - a], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
                mask_ready.append((a, b, ready))
                d2h_bytes += (b - a) * oh * ow
            frames_cpu[a:b].copy_(dst, non_blocking=True)
            d2h_bytes += (b - a) * oh * ow * 12
            if masks_cpu is not None and not pack_mask:
                masks_cpu[a:b].copy_(mask, non_blocking=True)
                d2h_bytes += (b - a) * oh * ow * 4
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            copied[k % len(bufs)] = ev
    global LAST_D2H_BYTES
    LAST_D2H_BYTES = d2h_bytes
    pad_all = None
    if pads:
        pad_all = torch.cat(pads) if len(pads) > 1 else pads[0]
        if pad_transform is not None:
            pad_all = pad_transform(pad_all)
    keep_alive = (bufs, fwd_t)  # the copy stream still reads the staging buffers after this function returns

    def finish_host(_keep=keep_alive):
        for a, b, ready in mask_ready:
            ready.synchronize()
            masks_cpu[a:b].copy_(mask_bytes_cpu[a:b])  # uint8 -> float32 on the host cores (torch's parallel copy)
        copy_stream.synchronize()
        main.synchronize()
        pad_np = pad_all.cpu().numpy().astype(np.int64) if pad_all is not None else None
        if odd is not None and int(odd.item()) != 0:
            raise _native.VstabNativeError("a single-sample padding mask held a value other than 0 / 1: refusing to return it as bytes")
        return frames_cpu, masks_cpu, pad_np

    return finish_host if defer else finish_host()
