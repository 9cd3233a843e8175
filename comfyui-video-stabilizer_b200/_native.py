"""ctypes binding of libvstab.so (include/vstab.h).

There is deliberately NO fallback: if the CUDA library is missing or no B200 is visible the
import of the compute entry points fails loudly (``VstabNativeError``).  Host code above this
module only ever hands CUDA torch tensors to it; torch is plumbing (device memory + streams).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSTAB_LIB") or os.path.join(_PKG_DIR, "libvstab.so")  # VSTAB_LIB: A/B builds of the same ABI

INTERP = {"bilinear": 0, "bicubic": 1}
MASK_RULE_P = 0
MASK_RULE_C = 1
MASK_RULE_AUTO = 2


def mask_rule_auto(threads: int) -> int:
    """VSTAB_MASK_RULE_AUTO_THREADS(threads) of include/vstab.h: the rule cv2 would pick per call on a machine whose
    cv2.getNumThreads() is `threads` (Rule P unless one of its destination stripes misses the source)."""
    return MASK_RULE_AUTO | (max(1, int(threads)) << 8)


def default_mask_rule() -> int:
    """The nodes' mask rule.  Rule P (= cv2 with one thread, and cv2 at any thread count for ordinary stabilisation
    jitter) unless the environment asks for the reference machine's behaviour: VSTAB_MASK_RULE=auto (thread count =
    this host's CPU count, like cv2's default), auto:<threads>, P or C.  The node schema is untouched."""
    v = os.environ.get("VSTAB_MASK_RULE", "P").strip().lower()
    if v in ("p", "0", ""):
        return MASK_RULE_P
    if v in ("c", "1"):
        return MASK_RULE_C
    if v.startswith("auto"):
        _, _, t = v.partition(":")
        return mask_rule_auto(int(t) if t else (os.cpu_count() or 1))
    raise VstabNativeError(f"VSTAB_MASK_RULE={v!r}: expected P, C, auto or auto:<threads>")
STAGE_AUTO = 0
STAGE_GLOBAL = 1
MODE_INDEX = {"translation": 0, "similarity": 1, "perspective": 2}
MODE_NAMES = ("translation", "similarity", "perspective")


class VstabNativeError(RuntimeError):
    pass


class FitResult(C.Structure):
    _fields_ = [
        ("matrix", C.c_double * 9),
        ("residual", C.c_double),
        ("n_inliers", C.c_int32),
        ("n_valid", C.c_int32),
        ("n_total", C.c_int32),
        ("ok", C.c_int32),
    ]


FIT_RESULT_DOUBLES = C.sizeof(FitResult) // 8  # 12 x 8-byte words

_lib = None
_lib_lock = threading.Lock()


def load_library() -> C.CDLL:
    """dlopen libvstab.so and declare every prototype of include/vstab.h."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise VstabNativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        vp, i32, fp = C.c_void_p, C.c_int, C.POINTER(C.c_float)
        lib.vstab_abi_version.restype = i32
        lib.vstab_create.argtypes = [i32, C.POINTER(vp)]
        lib.vstab_create.restype = i32
        lib.vstab_destroy.argtypes = [vp]
        lib.vstab_destroy.restype = None
        lib.vstab_last_error.argtypes = [vp]
        lib.vstab_last_error.restype = C.c_char_p
        lib.vstab_launch_count.argtypes = [vp]
        lib.vstab_launch_count.restype = C.c_uint64
        lib.vstab_host_libm.argtypes = [i32, vp, vp, vp, i32]
        lib.vstab_host_libm.restype = i32
        lib.vstab_host_trajectory.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, C.POINTER(i32)]
        lib.vstab_host_trajectory.restype = i32
        lib.vstab_host_framing.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp]
        lib.vstab_host_framing.restype = i32
        lib.vstab_host_shift.argtypes = [vp, i32, C.c_float, C.c_float, vp]
        lib.vstab_host_shift.restype = i32
        lib.vstab_host_target.argtypes = [vp, i32, i32, i32, C.c_double, i32, vp, vp]
        lib.vstab_host_target.restype = i32
        lib.vstab_working_size.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32)]
        lib.vstab_working_size.restype = i32
        lib.vstab_gray_working.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp]
        lib.vstab_gray_working.restype = i32
        lib.vstab_gray_working_adapt.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, vp]
        lib.vstab_gray_working_adapt.restype = i32
        lib.vstab_range_normalize.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
        lib.vstab_range_normalize.restype = i32
        lib.vstab_warp_fused.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, i32, i32, fp, i32, i32, vp, vp, vp, vp]
        lib.vstab_warp_fused.restype = i32
        lib.vstab_common_coverage.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
        lib.vstab_common_coverage.restype = i32
        lib.vstab_coverage_bbox.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
        lib.vstab_coverage_bbox.restype = i32
        lib.vstab_mask_pack_u8.argtypes = [vp, vp, C.c_size_t, vp, vp, vp]
        lib.vstab_mask_pack_u8.restype = i32
        if hasattr(lib, "vstab_dis_flow"):
            lib.vstab_dis_flow.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, vp]
            lib.vstab_dis_flow.restype = i32
            lib.vstab_dis_flow_at.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp]
            lib.vstab_dis_flow_at.restype = i32
        if hasattr(lib, "vstab_gftt_lk"):
            lib.vstab_gftt_lk.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
            lib.vstab_gftt_lk.restype = i32
        if hasattr(lib, "vstab_fit_batch"):
            lib.vstab_fit_batch.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
            lib.vstab_fit_batch.restype = i32
        if lib.vstab_abi_version() != 1:
            raise VstabNativeError("libvstab.so ABI version mismatch; rebuild it")
        _lib = lib
        return lib


_LIBM_OPS = {"atan2": 0, "log": 1, "exp": 2, "cos": 3, "sin": 4}


def host_libm(op: str, a, b=None):
    """glibc libm over a float64 array (vstab_host_libm): bit-identical to map(math.<op>, ...), without a Python call per
    element.  Host-only: needs the library, not a GPU."""
    import numpy as np

    lib = load_library()
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty_like(a)
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.float64)
        bp = b.ctypes.data_as(C.c_void_p)
    rc = lib.vstab_host_libm(_LIBM_OPS[op], a.ctypes.data_as(C.c_void_p), bp, out.ctypes.data_as(C.c_void_p), int(a.size))
    if rc != 0:
        raise VstabNativeError(f"vstab_host_libm({op}) failed: {rc}")
    return out


def host_trajectory(raw, detected, min_points: int, mode: str, source_size, working_size):
    """vstab_host_trajectory: raw [P,3,12] float64 fit words -> (matrices [P,3,3] float32 at full size, path [P+1,K]
    float64), or None when some pair does not accept `mode` itself (the caller replays the fallback ladder)."""
    import numpy as np

    lib = load_library()
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    p = raw.shape[0]
    det = None
    if detected is not None:
        det = np.ascontiguousarray(detected, dtype=np.int32)
    k = (2, 4, 8)[MODE_INDEX[mode]]
    matrices = np.empty((p, 3, 3), dtype=np.float32)
    path = np.empty((p + 1, k), dtype=np.float64)
    first = C.c_int(0)
    ww, wh = (0, 0) if working_size is None else (int(working_size[0]), int(working_size[1]))
    rc = lib.vstab_host_trajectory(raw.ctypes.data, None if det is None else det.ctypes.data, p, int(min_points), MODE_INDEX[mode],
                                   int(source_size[0]), int(source_size[1]), ww, wh, matrices.ctypes.data, path.ctypes.data,
                                   C.byref(first))
    if rc != 0:
        raise VstabNativeError(f"vstab_host_trajectory failed: {rc}")
    if first.value != p:
        return None
    return matrices, path


def host_framing(diffs, mode: str, width: int, height: int):
    """vstab_host_framing: diffs [N,K] float64 -> (apply [N,3,3] float32, mins [N,2], maxs [N,2], box [9])."""
    import numpy as np

    lib = load_library()
    diffs = np.ascontiguousarray(diffs, dtype=np.float64)
    n = diffs.shape[0]
    apply = np.empty((n, 3, 3), dtype=np.float32)
    mins = np.empty((n, 2), dtype=np.float64)
    maxs = np.empty((n, 2), dtype=np.float64)
    box = np.empty((9,), dtype=np.float64)
    rc = lib.vstab_host_framing(diffs.ctypes.data, n, MODE_INDEX[mode], int(width), int(height), apply.ctypes.data,
                                mins.ctypes.data, maxs.ctypes.data, box.ctypes.data)
    if rc != 0:
        raise VstabNativeError(f"vstab_host_framing failed: {rc}")
    return apply, mins, maxs, box


def host_target(path, window: int, strength: float, camera_lock: bool):
    """vstab_host_target: (target, diffs) float64 arrays shaped like path, or None for windows numpy sums through BLAS."""
    import numpy as np

    lib = load_library()
    path = np.ascontiguousarray(path, dtype=np.float64)
    target = np.empty_like(path)
    diffs = np.empty_like(path)
    rc = lib.vstab_host_target(path.ctypes.data, int(path.shape[0]), int(path.shape[1]), int(window), float(strength), int(bool(camera_lock)),
                               target.ctypes.data, diffs.ctypes.data)
    return (target, diffs) if rc == 0 else None


def host_shift(apply, off_x, off_y):
    """vstab_host_shift: translate affine float32 matrices; None when a matrix is not affine (use the float32 matmul)."""
    import numpy as np

    lib = load_library()
    out = np.empty_like(apply)
    rc = lib.vstab_host_shift(apply.ctypes.data, int(apply.shape[0]), float(np.float32(off_x)), float(np.float32(off_y)), out.ctypes.data)
    return out if rc == 0 else None


def working_size(width: int, height: int) -> tuple[int, int]:
    """Host helper (no GPU needed): nodes/stabilizer_utils.py:248-268."""
    lib = load_library()
    w, h = C.c_int(), C.c_int()
    rc = lib.vstab_working_size(int(width), int(height), C.byref(w), C.byref(h))
    if rc != 0:
        raise ValueError("invalid frame size")
    return w.value, h.value


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _check_cuda(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise VstabNativeError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise VstabNativeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise VstabNativeError(f"{name} must be contiguous")


class Handle:
    """One libvstab handle per (device, thread); kernels go to torch's current stream."""

    def __init__(self, device: int | torch.device | None = None):
        if not torch.cuda.is_available():
            raise VstabNativeError("no CUDA device visible: libvstab needs a B200 (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise VstabNativeError("libvstab needs a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        hp = C.c_void_p()
        rc = self.lib.vstab_create(self.device.index, C.byref(hp))
        if rc != 0:
            raise VstabNativeError(self.lib.vstab_last_error(None).decode())
        self._h = hp

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.vstab_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise VstabNativeError(f"libvstab error {rc}: {self.lib.vstab_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self.lib.vstab_launch_count(self._h))

    # -- K1 + K2 ---------------------------------------------------------------------------------
    def gray_working(self, rgb: torch.Tensor, work_size: tuple[int, int] | None = None) -> torch.Tensor:
        """rgb [N,H,W,3] f32 cuda -> gray [N,h,w] u8 at the working (estimation) size."""
        _check_cuda(rgb, torch.float32, "rgb")
        n, h, w, c = rgb.shape
        if c != 3:
            raise VstabNativeError("rgb must have 3 channels")
        ww, wh = work_size if work_size is not None else working_size(w, h)
        out = torch.empty((n, wh, ww), dtype=torch.uint8, device=rgb.device)
        self._check(self.lib.vstab_gray_working(self._h, rgb.data_ptr(), n, h, w, out.data_ptr(), wh, ww, _stream_ptr(rgb.device)))
        return out

    def gray_working_adapt(self, rgb: torch.Tensor, work_size: tuple[int, int] | None = None):
        """K1 + K2 with the input adapter's range rule fused in (include/vstab.h, vstab_gray_working_adapt):
        rgb [N,H,W,3] f32 cuda, READ AND WRITTEN (frames whose max exceeds 1.5 are divided by 255 in place)
        -> (gray [N,h,w] u8, flags [N] int32 on the device: 1 = that frame was 0..255 content and has been rescaled)."""
        _check_cuda(rgb, torch.float32, "rgb")
        n, h, w, c = rgb.shape
        if c != 3:
            raise VstabNativeError("rgb must have 3 channels")
        ww, wh = work_size if work_size is not None else working_size(w, h)
        out = torch.empty((n, wh, ww), dtype=torch.uint8, device=rgb.device)
        flags = torch.empty((n,), dtype=torch.int32, device=rgb.device)
        self._check(self.lib.vstab_gray_working_adapt(self._h, rgb.data_ptr(), n, h, w, out.data_ptr(), wh, ww, flags.data_ptr(),
                                                      _stream_ptr(rgb.device)))
        return out, flags

    def range_normalize(self, frames: torch.Tensor) -> torch.Tensor:
        """The adapter's range rule alone, in place: frames [N,H,W,C] f32 cuda -> flags [N] int32 (1 = rescaled by 1/255)."""
        _check_cuda(frames, torch.float32, "frames")
        n, h, w, c = frames.shape
        flags = torch.empty((n,), dtype=torch.int32, device=frames.device)
        self._check(self.lib.vstab_range_normalize(self._h, frames.data_ptr(), n, h, w, c, flags.data_ptr(), _stream_ptr(frames.device)))
        return flags

    # -- K10 + K11 + K12 -------------------------------------------------------------------------
    def warp_fused(
        self,
        src: torch.Tensor,
        fwd: torch.Tensor,
        out_size: tuple[int, int],
        interp: str,
        border: tuple[float, float, float],
        *,
        mask_rule: int = MASK_RULE_P,
        stage_mode: int = STAGE_AUTO,
        want_mask: bool = True,
        want_pad_count: bool = False,
        out: torch.Tensor | None = None,
        mask_out: torch.Tensor | None = None,
    ):
        """src [N,H,W,3] f32, fwd [N,S,9] f32 forward matrices -> (dst [N,H',W',3], mask [N,H',W'] | None, pad_count [N] | None)."""
        _check_cuda(src, torch.float32, "src")
        _check_cuda(fwd, torch.float32, "fwd")
        n, sh, sw, c = src.shape
        if c != 3:
            raise VstabNativeError("src must have 3 channels")
        if fwd.dim() == 2:
            fwd = fwd.view(n, 1, 9)
        if fwd.shape[0] != n or fwd.shape[2] != 9:
            raise VstabNativeError("fwd must be [N,S,9]")
        samples = int(fwd.shape[1])
        ow, oh = int(out_size[0]), int(out_size[1])
        dst = out if out is not None else torch.empty((n, oh, ow, 3), dtype=torch.float32, device=src.device)
        _check_cuda(dst, torch.float32, "out")
        mask = None
        if want_mask:
            mask = mask_out if mask_out is not None else torch.empty((n, oh, ow), dtype=torch.float32, device=src.device)
            _check_cuda(mask, torch.float32, "mask_out")
        pad = torch.empty((n,), dtype=torch.int32, device=src.device) if want_pad_count else None
        b = (C.c_float * 3)(*[float(x) for x in border])
        self._check(
            self.lib.vstab_warp_fused(
                self._h, src.data_ptr(), n, sh, sw, fwd.data_ptr(), samples, INTERP[interp], oh, ow, b,
                int(mask_rule), int(stage_mode), dst.data_ptr(), _ptr(mask), _ptr(pad), _stream_ptr(src.device),
            )
        )
        return dst, mask, pad

    def mask_pack_u8(self, mask: torch.Tensor, out: torch.Tensor, odd: torch.Tensor) -> None:
        """Binary float32 mask -> bytes (vstab_mask_pack_u8); `odd` (uint32 [1], zeroed by the caller) collects bit 0 when a
        value other than 0 / 1 shows up."""
        _check_cuda(mask, torch.float32, "mask")
        _check_cuda(out, torch.uint8, "out")
        if not (mask.is_contiguous() and out.is_contiguous()) or out.numel() != mask.numel():
            raise VstabNativeError("mask_pack_u8: contiguous buffers of equal length expected")
        self._check(self.lib.vstab_mask_pack_u8(self._h, mask.data_ptr(), mask.numel(), out.data_ptr(), odd.data_ptr(), _stream_ptr(mask.device)))

    def common_coverage(self, fwd: torch.Tensor, src_size, out_size, mask_rule: Optional[int] = None) -> torch.Tensor:
        _check_cuda(fwd, torch.float32, "fwd")
        mask_rule = default_mask_rule() if mask_rule is None else mask_rule
        n = int(fwd.shape[0])
        sw, sh = int(src_size[0]), int(src_size[1])
        ow, oh = int(out_size[0]), int(out_size[1])
        out = torch.empty((oh, ow), dtype=torch.uint8, device=fwd.device)
        self._check(self.lib.vstab_common_coverage(self._h, fwd.data_ptr(), n, sh, sw, oh, ow, int(mask_rule), out.data_ptr(), _stream_ptr(fwd.device)))
        return out

    def coverage_bbox(self, fwd: torch.Tensor, src_size, out_size, mask_rule: Optional[int] = None) -> torch.Tensor:
        """fwd [N,9] f32 -> int32 [N,4] (xmin, ymin, xmax, ymax) of the 3x3-closed coverage; xmax < 0 = empty."""
        _check_cuda(fwd, torch.float32, "fwd")
        mask_rule = default_mask_rule() if mask_rule is None else mask_rule
        n = int(fwd.shape[0])
        sw, sh = int(src_size[0]), int(src_size[1])
        ow, oh = int(out_size[0]), int(out_size[1])
        out = torch.empty((n, 4), dtype=torch.int32, device=fwd.device)
        self._check(self.lib.vstab_coverage_bbox(self._h, fwd.data_ptr(), n, sh, sw, oh, ow, int(mask_rule), out.data_ptr(), _stream_ptr(fwd.device)))
        return out

    # -- K3 + K4 ---------------------------------------------------------------------------------
    def dis_flow(self, gray: torch.Tensor, *, want_flow: bool = False, grid_step: int = 8, first_pair: int = 0):
        """gray [N,h,w] u8 -> (flow [N-1,h,w,2] | None, grid [N-1,gh,gw,2] | None).
        first_pair: clip-wide index of the pair (gray[0], gray[1]); only pair 0 of a clip meets the backend
        object in its configured state (include/vstab.h, vstab_dis_flow_at)."""
        _check_cuda(gray, torch.uint8, "gray")
        if not hasattr(self.lib, "vstab_dis_flow"):
            raise VstabNativeError("libvstab.so was built without vstab_dis_flow")
        n, h, w = gray.shape
        npairs = max(n - 1, 0)
        flow = torch.empty((npairs, h, w, 2), dtype=torch.float32, device=gray.device) if want_flow else None
        grid = None
        if grid_step > 0:
            gh, gw = (h + grid_step - 1) // grid_step, (w + grid_step - 1) // grid_step
            grid = torch.empty((npairs, gh, gw, 2), dtype=torch.float32, device=gray.device)
        self._check(self.lib.vstab_dis_flow_at(self._h, gray.data_ptr(), n, h, w, int(first_pair), _ptr(flow), _ptr(grid),
                                               int(max(grid_step, 0)), _stream_ptr(gray.device)))
        return flow, grid

    # -- K5 + K6 ---------------------------------------------------------------------------------
    def gftt_lk(self, gray: torch.Tensor, max_corners: int = 400):
        """gray [N,h,w] u8 -> (prev [N-1,K,2], curr [N-1,K,2] with NaN rows for missing / lost, detected [N-1] int32)."""
        _check_cuda(gray, torch.uint8, "gray")
        if not hasattr(self.lib, "vstab_gftt_lk"):
            raise VstabNativeError("libvstab.so was built without vstab_gftt_lk")
        n, h, w = gray.shape
        p = max(n - 1, 0)
        prev = torch.empty((p, max_corners, 2), dtype=torch.float32, device=gray.device)
        curr = torch.empty((p, max_corners, 2), dtype=torch.float32, device=gray.device)
        det = torch.zeros((p,), dtype=torch.int32, device=gray.device)
        self._check(self.lib.vstab_gftt_lk(self._h, gray.data_ptr(), n, h, w, int(max_corners), prev.data_ptr(), curr.data_ptr(),
                                           det.data_ptr(), _stream_ptr(gray.device)))
        return prev, curr, det

    # -- K4 + K7..K9 -----------------------------------------------------------------------------
    def _fit_out(self, p: int, device, out: Optional[torch.Tensor]) -> torch.Tensor:
        if out is None:
            return torch.zeros((p, 3, FIT_RESULT_DOUBLES), dtype=torch.float64, device=device)
        _check_cuda(out, torch.float64, "out")
        if tuple(out.shape) != (p, 3, FIT_RESULT_DOUBLES):
            raise VstabNativeError(f"out must be [{p},3,{FIT_RESULT_DOUBLES}] float64")
        return out

    def fit_grid(self, grid_flow: torch.Tensor, grid_step: int, mode_mask: int = 7, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """grid_flow [P,gh,gw,2] sampled flow -> raw result words [P,3,12] float64 (see FitResult).
        out: zero-initialised destination (e.g. a slice of an all-gather send buffer)."""
        _check_cuda(grid_flow, torch.float32, "grid_flow")
        if not hasattr(self.lib, "vstab_fit_batch"):
            raise VstabNativeError("libvstab.so was built without vstab_fit_batch")
        p, gh, gw, _ = grid_flow.shape
        out = self._fit_out(p, grid_flow.device, out)
        self._check(self.lib.vstab_fit_batch(self._h, None, grid_flow.data_ptr(), p, gh * gw, gw, gh, int(grid_step), int(mode_mask), out.data_ptr(), _stream_ptr(grid_flow.device)))
        return out

    def fit_points(self, prev: torch.Tensor, curr: torch.Tensor, mode_mask: int = 7, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """prev/curr [P,K,2] correspondences (NaN rows = invalid) -> raw result words [P,3,12] float64."""
        _check_cuda(prev, torch.float32, "prev")
        _check_cuda(curr, torch.float32, "curr")
        p, k, _ = prev.shape
        out = self._fit_out(p, prev.device, out)
        self._check(self.lib.vstab_fit_batch(self._h, prev.data_ptr(), curr.data_ptr(), p, k, 0, 0, 0, int(mode_mask), out.data_ptr(), _stream_ptr(prev.device)))
        return out


_handles: dict[tuple[int, int], Handle] = {}


def get_handle(device: int | torch.device | None = None) -> Handle:
    if not torch.cuda.is_available():
        raise VstabNativeError("no CUDA device visible: libvstab needs a B200 (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, threading.get_ident())
    h = _handles.get(key)
    if h is None:
        h = Handle(torch.device("cuda", idx))
        _handles[key] = h
    return h


def decode_fit_results(raw: torch.Tensor):
    """raw [P,3,12] float64 words -> dict of numpy arrays indexed [pair, mode]."""
    import numpy as np

    arr = np.ascontiguousarray(raw.detach().cpu().numpy())
    ints = arr[..., 10:12].copy().view(np.int32)  # [P,3,4]
    return {
        "matrix": arr[..., :9].reshape(arr.shape[0], 3, 3, 3).copy(),
        "residual": arr[..., 9].copy(),
        "n_inliers": ints[..., 0].copy(),
        "n_valid": ints[..., 1].copy(),
        "n_total": ints[..., 2].copy(),
        "ok": ints[..., 3].copy(),
    }
