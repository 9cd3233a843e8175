"""CPU-only checks: the C-ABI library loads and exports every symbol include/vstab.h declares, the
host helpers match the unmodified reference (when present), the node schema is the reference's."""
import ctypes
import json
import os
import re
import sys
import types

import numpy as np
import pytest

import vstab_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vstab_loader.load()
from vstab_b200 import _native, hostmath as hm, motion_meta as mm  # noqa: E402


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vstab.h")).read()
    declared = sorted(set(re.findall(r"VSTAB_API\s+[\w\s\*]+?\b(vstab_\w+)\s*\(", header)))
    assert len(declared) >= 10, declared
    lib = ctypes.CDLL(_native.LIB_PATH)
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, missing
    assert lib.vstab_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.VstabNativeError):
        _native.get_handle()
    from vstab_b200 import pipeline

    with pytest.raises(_native.VstabNativeError):
        pipeline.normalize_video_input(torch.zeros((2, 8, 8, 3)))
    lib = _native.load_library()
    h = ctypes.c_void_p()
    assert lib.vstab_create(0, ctypes.byref(h)) != 0  # no device => error code + message, no handle
    assert lib.vstab_last_error(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "comfyui-video-stabilizer_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            text = open(os.path.join(pkg, name)).read()
            assert "oracle" not in text.replace("oracle/dis_ref.c", ""), name
            assert "import cv2" not in text, name


def test_working_size_matches_host_rule():
    for w, h in [(1920, 1080), (3840, 2160), (1280, 720), (832, 480), (1080, 1920), (961, 541), (2000, 300)]:
        ws = hm.working_estimation_size(w, h)
        assert _native.working_size(w, h) == (ws if ws is not None else (w, h))


def test_padding_color_parser():
    assert hm.parse_padding_color("#7F7F7F") == (127, 127, 127)
    assert hm.parse_padding_color("#abc") == (0xAA, 0xBB, 0xCC)
    assert hm.parse_padding_color("10, 300,-4") == (10, 255, 0)
    assert hm.parse_padding_color("12") == (127, 127, 127)  # not 3 or 6 hex digits
    assert hm.parse_padding_color("zzz") == (127, 127, 127)
    assert hm.parse_padding_color(0x102030) == (0x10, 0x20, 0x30)


def test_motion_meta_contract_messages():
    block = mm.build_motion_meta_v2(source="x", frame_count=2, fps=16.0, input_size=(8, 6), output_size=(8, 6),
                                    matrices=[np.eye(3), np.eye(3)])
    assert mm.resolve_motion_meta({"motion_meta": block}).frame_count == 2
    bad = json.loads(json.dumps(block))
    bad["per_frame"][1]["matrix"][2] = [0, 0, 0]
    with pytest.raises(ValueError, match=r"motion_meta.per_frame\[1\].matrix is not invertible."):
        mm.validate_motion_meta(bad)
    bad = json.loads(json.dumps(block))
    bad["per_frame"][1]["index"] = 5
    with pytest.raises(ValueError, match=r"per_frame\[1\].index must be 1, got 5"):
        mm.validate_motion_meta(bad)
    with pytest.raises(ValueError, match="version must be 2"):
        mm.validate_motion_meta({"version": 1})
    with pytest.raises(ValueError, match="meta must contain motion_meta or stabilization_warp"):
        mm.resolve_motion_meta({})


@pytest.mark.reference
def test_host_helpers_equal_reference(reference_nodes):
    ref = reference_nodes.stabilizer_utils
    rng = np.random.default_rng(0)
    for mode in ("translation", "similarity", "perspective"):
        for _ in range(50):
            th, s = rng.normal(0, 0.02), 1 + rng.normal(0, 0.02)
            m = np.array([[s * np.cos(th), -s * np.sin(th), rng.normal(0, 9)], [s * np.sin(th), s * np.cos(th), rng.normal(0, 9)],
                          [rng.normal(0, 1e-5), rng.normal(0, 1e-5), 1]], dtype=np.float32)
            full = hm.rescale_transform_to_full(m, (1920, 1080), (960, 540))
            assert np.array_equal(full, ref._rescale_transform_to_full(m, (1920, 1080), (960, 540)))
            p = hm.matrix_to_params(full, mode)
            assert np.array_equal(p, ref._matrix_to_params(full, mode))
            assert np.array_equal(hm.params_to_matrix(p * 0.3, mode), ref._params_to_matrix(p * 0.3, mode))
    path = np.cumsum(rng.normal(0, 2, (121, 4)), axis=0)
    for smooth, fps in [(0.5, 16.0), (1.0, 24.0), (0.05, 1.0), (0.0, 30.0)]:
        assert np.array_equal(hm.smooth_path(path, smooth, fps), ref._smooth_path(path, smooth, fps))
    mats = [hm.params_to_matrix(r, "similarity") for r in rng.normal(0, 0.01, (30, 4)) * np.array([300, 300, 1, 1])]
    for a, b in zip(hm.compute_bounding_boxes(mats, 1920, 1080), ref._compute_bounding_boxes(mats, 1920, 1080)):
        assert np.array_equal(a, b)
    lo, hi = ref._compute_bounding_boxes(mats, 1920, 1080)
    assert hm.min_content_ratio(lo, hi, 1920, 1080) == ref._min_content_ratio(lo, hi, 1920, 1080)
    t1, s1 = hm.prepare_expand_transform(lo, hi)
    t2, s2 = ref._prepare_expand_transform(lo, hi)
    assert np.array_equal(t1, t2) and s1 == s2
    for text in ["#404040", "1,2,3", "#12", 99999999, "0x10"]:
        assert hm.parse_padding_color(text) == ref._parse_padding_color(text)


@pytest.mark.reference
def test_motion_meta_equals_reference(reference_nodes):
    ref = reference_nodes.motion_meta
    mats = [np.array([[1.01, 0.02, 3.0], [-0.02, 0.99, -2.0], [1e-5, 0, 1.0]]) * (1 + 0.01 * i) for i in range(5)]
    warp = hm.build_stabilization_warp_meta(source_size=(64, 48), output_size=(70, 50), framing_mode="expand", applied_matrices=mats)
    assert warp == reference_nodes.stabilizer_utils._build_stabilization_warp_meta(
        source_size=(64, 48), output_size=(70, 50), framing_mode="expand", applied_matrices=mats)
    assert mm.applied_motion_meta_from_stabilization_warp(warp, 16.0, "s") == ref.applied_motion_meta_from_stabilization_warp(warp, 16.0, "s")
    assert mm.motion_meta_from_stabilization_warp(warp, 24.0, "s") == ref.motion_meta_from_stabilization_warp(warp, 24.0, "s")


def _install_comfy_stubs():
    class _Socket:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k

    class _Kind:
        Input = Output = _Socket

    io = types.SimpleNamespace(
        ComfyNode=type("ComfyNode", (), {}), Custom=lambda name: _Kind, Image=_Kind, Mask=_Kind, Float=_Kind, Combo=_Kind,
        Boolean=_Kind, Color=_Kind, Int=_Kind, String=_Kind, NumberDisplay=types.SimpleNamespace(slider="slider"),
        Schema=type("Schema", (), {"__init__": lambda self, **k: self.__dict__.update(k)}),
        NodeOutput=lambda *a: a,
    )
    latest = types.ModuleType("comfy_api.latest")
    latest.ComfyExtension = type("ComfyExtension", (), {})
    latest.io = io
    sys.modules.setdefault("comfy_api", types.ModuleType("comfy_api"))
    sys.modules["comfy_api.latest"] = latest
    comfy = types.ModuleType("comfy")
    comfy.__path__ = []
    utils = types.ModuleType("comfy.utils")
    utils.ProgressBar = type("ProgressBar", (), {"__init__": lambda self, total: None, "update_absolute": lambda self, *a, **k: None})
    sys.modules["comfy"] = comfy
    sys.modules["comfy.utils"] = utils


def test_node_schema_matches_reference_contract():
    """ids / display names / socket names and order pinned by the reference's scripts/check_node_schema.py:11-64."""
    _install_comfy_stubs()
    sys.modules.pop("vstab_b200.nodes", None)
    from vstab_b200 import nodes

    stab_inputs = ["frames", "frame_rate", "framing_mode", "transform_mode", "camera_lock", "strength", "smooth", "keep_fov", "padding_color"]
    expected = {
        nodes.VideoStabilizerClassic: ("video_stabilizer_classic", "Video Stabilizer Classic", stab_inputs, ["frames_stabilized", "padding_mask", "meta"]),
        nodes.VideoStabilizerFlow: ("video_stabilizer_flow", "Video Stabilizer Flow", stab_inputs, ["frames_stabilized", "padding_mask", "meta"]),
        nodes.VideoStabilizerMotionApply: ("video_stabilizer_motion_apply", "Video Stabilizer Motion Apply",
                                           ["frames", "motion_meta", "framing_mode", "interpolation", "padding_color", "motion_blur", "motion_blur_quality"],
                                           ["frames", "padding_mask", "meta"]),
    }
    for cls, (node_id, display, ins, outs) in expected.items():
        schema = cls.define_schema()
        assert schema.node_id == node_id and schema.display_name == display
        assert [s.args[0] for s in schema.inputs] == ins
        assert [s.args[0] for s in schema.outputs] == outs
    defaults = {s.args[0]: s.kwargs.get("default") for s in nodes.VideoStabilizerFlow.define_schema().inputs}
    assert defaults["frame_rate"] == 16.0 and defaults["framing_mode"] == "crop_and_pad" and defaults["transform_mode"] == "similarity"
    assert defaults["strength"] == 0.7 and defaults["smooth"] == 0.5 and defaults["keep_fov"] == 0.6 and defaults["padding_color"] == "#7F7F7F"


def test_meta_helpers_for_long_clips_keep_the_reference_results(monkeypatch):
    """The O(frames) host pieces were vectorised for long / sharded clips: same values, same errors."""
    from vstab_b200 import hostmath as hm, motion_meta as mm, pipeline

    counts = [0, 1, 10, 12345, 2073599, 2073600]
    assert hm.padded_fractions(counts, 1920 * 1080) == [hm.padded_fraction(c, 1920 * 1080) for c in counts]

    rng = np.random.default_rng(3)
    mats = np.tile(np.eye(3, dtype=np.float32), (50, 1, 1))
    mats[:, :2, 2] = rng.normal(0, 20, (50, 2))
    mats[7] *= np.float32(1e-6)  # tiny but invertible: must pass like np.linalg.inv does
    block = mm.applied_motion_meta_from_matrices(mats, source_size=(64, 48), output_size=(64, 48), fps=16.0, source="t")
    assert block["frame_count"] == 50 and block["per_frame"][7]["matrix"] == mats[7].astype(np.float64).tolist()
    for bad in (np.zeros((3, 3), np.float32), np.array([[1, 2, 3], [2, 4, 6], [0, 0, 1]], np.float32)):
        broken = mats.copy()
        broken[11] = bad
        with pytest.raises(ValueError, match=r"per_frame\[11\]\.applied_matrix is not invertible"):
            mm.applied_motion_meta_from_matrices(broken, source_size=(64, 48), output_size=(64, 48), fps=16.0, source="t")
    broken = mats.copy()
    broken[3, 0, 0] = np.nan
    with pytest.raises(ValueError, match=r"per_frame\[3\]\.applied_matrix must contain finite numbers"):
        mm.applied_motion_meta_from_matrices(broken, source_size=(64, 48), output_size=(64, 48), fps=16.0, source="t")

    # streaming decision: explicit budget in MiB of the float32 RGB clip
    monkeypatch.setenv("VSTAB_RESIDENT_LIMIT_MB", "10")
    assert pipeline._must_stream(10, 480, 832, None)       # 47.9 MB
    assert not pipeline._must_stream(2, 480, 832, None)    # 9.6 MB


def test_stacked_conversions_ladder_columns_and_gc_pause():
    """Stacked matrix<->params equal the per-matrix helpers (themselves pinned to the reference above)
    on both dtypes incl. degenerate scales; ladder entries behave like the reference's tuples with and
    without a fallback; gc_paused restores the collector state."""
    import gc

    from vstab_b200 import hostmath as hm
    from vstab_b200.stabilizer_core import PairCandidates, replay_mode_ladder

    rng = np.random.default_rng(11)
    for dt in (np.float32, np.float64):
        m = np.tile(np.eye(3), (400, 1, 1)) + rng.normal(size=(400, 3, 3)) * 0.05
        m[:5, 0, 0] = 0
        m[:5, 1, 0] = 0            # a^2 + c^2 below the 1e-10 floor
        m[5:9, 0, 0] = 1e-5
        m[5:9, 1, 0] = 0           # exactly at the floor in float32
        m = m.astype(dt)
        for mode in hm.TRANSFORM_MODES:
            stacked = hm.matrices_to_params(m, mode)
            assert np.array_equal(stacked, np.stack([hm.matrix_to_params(x, mode) for x in m]))
            back = hm.params_to_matrices(stacked * 0.7, mode)
            assert np.array_equal(back, np.stack([hm.params_to_matrix(x, mode) for x in stacked * 0.7]))

    pairs = 9
    mats = np.tile(np.eye(3), (pairs, 3, 1, 1)) + rng.normal(size=(pairs, 3, 3, 3)) * 1e-3
    full = dict(residual=rng.random((pairs, 3)), n_inliers=np.full((pairs, 3), 900), n_valid=np.full((pairs, 3), 1000),
                n_total=np.full((pairs, 3), 1000), ok=np.ones((pairs, 3), int))
    entries, active, stack = replay_mode_ladder(PairCandidates(mats, **full), "similarity", with_residual=True)
    assert active == "similarity" and len(entries) == pairs and stack.dtype == np.float32
    matrix, mode, conf, resid = entries[4]
    assert mode == "similarity" and conf == 0.9 and resid == full["residual"][4, 1]
    assert np.array_equal(matrix, mats[4, 1].astype(np.float32)) and [e[1] for e in entries] == ["similarity"] * pairs
    # pair 3 loses the similarity model: it and every later pair fall back to translation (sticky)
    full["ok"][3, 1] = 0
    entries, active, stack = replay_mode_ladder(PairCandidates(mats, **full), "similarity", with_residual=False)
    assert active == "translation" and entries.modes == ["similarity"] * 3 + ["translation"] * 6
    assert entries.residuals == [None] * pairs and entries.confidences[3:] == [1.0] * 6
    assert np.array_equal(stack[2], mats[2, 1].astype(np.float32)) and np.array_equal(stack[5], mats[5, 0].astype(np.float32))
    table = PairCandidates(mats, **full).to_array()
    again = PairCandidates.from_array(table, 12)
    assert np.array_equal(again.ok, full["ok"]) and np.array_equal(again.n_valid, full["n_valid"]) and again.detected is None

    for state in (True, False):
        (gc.enable if state else gc.disable)()
        try:
            with hm.gc_paused():
                assert not gc.isenabled()
            assert gc.isenabled() == state
            with pytest.raises(KeyError):
                with hm.gc_paused():
                    raise KeyError("x")
            assert gc.isenabled() == state
        finally:
            gc.enable()


def test_node_execute_glue(monkeypatch):
    """execute() of the three nodes (reference: video_stabilizer_motion_apply.py:86-129, video_stabilizer_flow.py:734-763):
    padding colour parsing, blur quality -> sample count with the silent "Standard" fallback, progress totals, the
    interrupt hook, argument order into the drivers, (IMAGE, MASK, JSON) out.  Drivers and tensor adapters are stand-ins."""
    import numpy as np

    _install_comfy_stubs()
    sys.modules.pop("vstab_b200.nodes", None)
    from vstab_b200 import nodes

    bars = []

    class Bar:
        def __init__(self, total):
            self.total, self.calls = total, []
            bars.append(self)

        def update_absolute(self, done, total):
            self.calls.append((done, total))

    class Ctx:
        def __len__(self):
            return 5

    class Result:
        def __init__(self, meta):
            self.frames, self.masks, self.meta = "F", "M", meta

    seen = {}

    def fake_apply(context, meta, rgb, **kw):
        seen["apply"] = (meta, rgb, kw)
        for _ in range(5 + 5 * 9):
            kw["progress_callback"]()
        return Result({"k": 1})

    def fake_stab(context, *a, **kw):
        seen["stab"] = (a, kw)
        kw["interrupt_check"]()
        return Result({"frames": 5})

    monkeypatch.setattr(nodes, "ProgressBar", Bar)
    monkeypatch.setattr(nodes, "normalize_video_input", lambda frames, **kw: Ctx())
    monkeypatch.setattr(nodes, "reconstruct_video", lambda frames, ctx: ("video", frames))
    monkeypatch.setattr(nodes, "convert_masks_for_output", lambda masks: ("mask", masks))
    monkeypatch.setattr(nodes.motion_apply, "apply_motion", fake_apply)
    monkeypatch.setattr(nodes.flow, "stabilize_frames", fake_stab)
    monkeypatch.setattr(nodes.classic, "stabilize_frames", fake_stab)

    out = nodes.VideoStabilizerMotionApply.execute("frames", {"motion_meta": {}}, "crop", "bicubic", "#0AC85A", 0.5, "NoSuchQuality")
    meta, rgb, kw = seen["apply"]
    assert rgb == (10, 200, 90) and kw["framing_mode"] == "crop" and kw["interpolation"] == "bicubic"
    assert kw["motion_blur"] == 0.5 and kw["motion_blur_samples"] == 9  # unknown quality -> "Standard"
    assert out == (("video", "F"), ("mask", "M"), {"k": 1, "motion_apply": {"motion_blur_quality": "Standard"}})
    total = 5 * 9 + 5
    assert bars[-1].total == total and bars[-1].calls[-1] == (total, total) and len(bars[-1].calls) == total + 1
    assert max(d for d, _ in bars[-1].calls) == total

    nodes.VideoStabilizerMotionApply.execute("frames", {}, "expand", "bilinear", "#000000", 0.0, "Ultra")
    assert seen["apply"][2]["motion_blur_samples"] == 33 and bars[-1].total == 5  # blur off: one tick per frame

    class Interrupted(Exception):
        pass

    def boom():
        raise Interrupted

    monkeypatch.setattr(nodes, "model_management", types.SimpleNamespace(throw_exception_if_processing_interrupted=boom))
    for cls in (nodes.VideoStabilizerFlow, nodes.VideoStabilizerClassic):
        try:
            cls.execute("frames", 24.0, "expand", "perspective", True, 0.3, 0.9, 0.2, "#FF0000")
            raise AssertionError("the interrupt must propagate")
        except Interrupted:
            pass
        a, kw = seen["stab"]
        assert a == ("expand", "perspective", True, 0.3, 0.9, 0.2, (255, 0, 0), 24.0)
        assert bars[-1].total == 4 + 5 and kw["progress_bar"] is bars[-1]
    monkeypatch.setattr(nodes, "model_management", None)
    out = nodes.VideoStabilizerFlow.execute("frames", 16.0, "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, "#7F7F7F")
    assert out == (("video", "F"), ("mask", "M"), {"frames": 5}) and seen["stab"][0][6] == (127, 127, 127)


@pytest.mark.reference
def test_input_adapters_equal_the_reference_on_its_own_cases(monkeypatch, reference_nodes):
    """scripts/compare_refactor_behavior.py:289-324 compare_helpers: every container the reference accepts (list,
    batch, dict, frames wrapped in a leading 1, float64, uint8, a non-contiguous view, torch float32 / uint8) must reach
    the same normalised clip and the same reconstructed payload.  The adapter rules are torch tensor operations, so
    they run here with the device pinned to the CPU; "base" is the unmodified reference."""
    import numpy as np
    import torch

    from vstab_b200 import pipeline

    monkeypatch.setattr(pipeline, "_require_device", lambda device: torch.device("cpu"))
    monkeypatch.setattr(pipeline, "_must_stream", lambda *a: False)
    U = reference_nodes.stabilizer_utils
    rng = np.random.default_rng(3)
    frames = [rng.random((45, 73, 3), dtype=np.float32) for _ in range(8)]
    batch = np.stack(frames, axis=0)
    u8 = (batch * 255.0).round().clip(0, 255).astype(np.uint8)
    cases_ = {
        "list": frames,
        "batch": batch,
        "dict": {"frames": batch, "fps": 24.0},
        "wrapped_frames": [f[np.newaxis, ...] for f in frames],
        "float64": batch.astype(np.float64),
        "uint8": u8,
        "noncontiguous": np.ascontiguousarray(np.stack([f[:, ::-1, :] for f in frames], axis=0))[:, :, ::-1, :],
        "torch_f32": torch.from_numpy(batch.copy()),
        "torch_uint8": torch.from_numpy(u8.copy()),
        "float_0_255": torch.from_numpy(batch * np.float32(255.0)),
        "gray": torch.from_numpy(batch[..., :1].copy()),
        "rgba": torch.from_numpy(np.concatenate([batch, batch[..., :1]], axis=-1)),
        "chw_list": [np.moveaxis(f, -1, 0) for f in frames],
    }
    for name, value in cases_.items():
        base = U._normalize_video_input(value)
        head = pipeline.normalize_video_input(value)
        for attr in ("width", "height", "channels", "fps", "template_kind", "template_meta"):
            assert getattr(base, attr) == getattr(head, attr), (name, attr)
        assert len(head) == len(base.frames), name
        got = head.frames.numpy()
        for i, want in enumerate(base.frames):
            assert np.array_equal(got[i], want), (name, i, float(np.abs(got[i] - want).max()))
        b = U._reconstruct_video(base.frames, base)
        h = pipeline.reconstruct_video(head.frames, head)
        if isinstance(b, dict):
            assert set(b) == set(h) and b["fps"] == h["fps"], name
            b, h = b["frames"], h["frames"]
        assert b.dtype == h.dtype and tuple(b.shape) == tuple(h.shape) and np.array_equal(b.numpy(), h.numpy()), name
    for bad in ([], {"fps": 24.0}):
        for fn in (U._normalize_video_input, pipeline.normalize_video_input):
            with pytest.raises(ValueError):
                fn(bad)
    masks = [rng.random((45, 73, 1)).astype(np.float32) for _ in range(3)]
    assert np.array_equal(U._convert_masks_for_output(masks).numpy(), pipeline.convert_masks_for_output(np.stack(masks)).numpy())
    assert tuple(pipeline.convert_masks_for_output(np.zeros((0, 4, 4, 1), np.float32)).shape) == tuple(U._convert_masks_for_output([]).shape) == (1, 1, 1)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line with the contract's keys,
    measured on the unmodified reference's own node (baseline/_ref, copied from /root/reference by build()) or, where that copy
    is absent, on the oracle's port of its cv2 call sequence; all 121 frames per step, no GPU involved."""
    import json
    import subprocess

    pytest.importorskip("cv2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["metric"] == "frames/sec 1080p Flow stabilize" and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"] and "workload" in line["config"]
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None and line["dtype"] == "f32" and line["data"] == "synthetic"
    from baseline import refload

    assert line["cpu_baseline"]["kind"] == ("reference" if refload.available() else "port")
    assert "all 121 frames" in line["cpu_baseline"]["sample"]
