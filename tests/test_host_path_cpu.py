"""The product's host path (stabilizer_core: candidate table -> sticky ladder -> rescale -> params -> cumsum ->
box smoothing -> framing -> meta) end to end on the CPU, with the oracle standing in for the two GPU stages
(estimation: dis_ref + fit_np; resampler: resample_np), against outputs of the UNMODIFIED reference
(tests/golden, scripts/make_golden.py).  What the `-m gpu` node tests check with the CUDA kernels in place,
checked here without a GPU: if this passes and the kernel parity tests pass, the nodes are the reference's."""
import functools
import json
import os

import numpy as np
import pytest

from oracle import dis_ref, fit_np, gray_np, resample_np
from tests import cases, parity
from tests.conftest import GOLDEN_DIR


class _Clip:
    """What stabilizer_core needs from a VideoContext."""

    def __init__(self, frames, fps=None):
        self.frames, self.fps = frames, fps
        self.height, self.width = frames.shape[1:3]
        self.device = None

    def __len__(self):
        return len(self.frames)

    def sliced(self, start):
        return type(self)(self.frames[start:], self.fps)

    def untouched(self, output):
        return self.frames, np.zeros(self.frames.shape[:3] + (1,), np.float32)


def _oracle_estimator(context, work_w, work_h, requested, clip_pair_offset=0):
    """flow.estimate_candidates with the oracle in place of K1-K4 / K7-K9: all candidate models of every pair."""
    from vstab_b200.stabilizer_core import PairCandidates

    work = None if (work_w, work_h) == (context.width, context.height) else (work_w, work_h)
    gray = [gray_np.gray_for_estimation(f, work) for f in context.frames]
    backend = dis_ref.Backend()  # one object per clip, like the reference (flow.py:312)
    if clip_pair_offset > 0:     # a shard in the middle of the clip meets the object after its first calc()
        backend.params.finest_scale = dis_ref.select_scales(work_h, work_w)[0]
    P = len(gray) - 1
    m = np.tile(np.eye(3), (P, 3, 1, 1))
    res, inl, valid, total, ok = (np.zeros((P, 3)) for _ in range(5))
    for p in range(P):
        prev, curr, n_total = fit_np.grid_correspondences(backend.calc(gray[p], gray[p + 1]))
        valid[p], total[p] = len(prev), n_total
        t = fit_np.median_shift(prev, curr)
        m[p, 0, :2, 2] = t
        res[p, 0] = float(np.abs((prev + t) - curr).mean())
        inl[p, 0], ok[p, 0] = len(prev), 1
        if requested in ("similarity", "perspective") and len(prev) >= 3:
            A, mask = fit_np.estimate_affine_partial_2d(prev, curr)
            if A is not None:
                m[p, 1, :2] = A
                res[p, 1] = float(np.abs((prev @ A[:, :2].T + A[:, 2]) - curr).mean())
                inl[p, 1], ok[p, 1] = int(mask.sum()), 1
        if requested == "perspective" and len(prev) >= 4:
            H, mask = fit_np.find_homography(prev, curr)
            if H is not None:
                m[p, 2] = H
                res[p, 2] = float(np.abs((prev @ H[:2, :2].T + H[:2, 2]) - curr).mean())
                inl[p, 2], ok[p, 2] = int(mask.sum()), 1
    return PairCandidates(m, res, inl.astype(int), valid.astype(int), total.astype(int), ok.astype(int), min_points=12)


def _oracle_warp(context, fwd, out_size, interpolation, border, **kw):
    """pipeline.fused_warp(defer=True) with the numpy resampler: frames, masks, padded-pixel counts."""
    def run():
        frames, masks = [], []
        for f, m in zip(context.frames, np.asarray(fwd, np.float32).reshape(-1, 3, 3)):
            frames.append(resample_np.warp_np(f, m, out_size, interpolation, border))
            masks.append(resample_np.mask_np(m, (context.width, context.height), out_size))
        masks = np.stack(masks)
        return np.stack(frames), masks, (masks > 0).reshape(len(masks), -1).sum(axis=1)

    return run


def _run(monkeypatch, frames, framing, mode, camera_lock, strength, smooth, keep_fov, padding_rgb, fps):
    from vstab_b200 import stabilizer_core as core

    monkeypatch.setattr(core, "fused_warp", _oracle_warp)
    return core.stabilize_frames(_Clip(frames), framing, mode, camera_lock, strength, smooth, keep_fov, padding_rgb, fps,
                                 estimator=_oracle_estimator, flavour="flow", output="device")


def _check(res, gmeta, want_frames=None, want_masks=None, gold=None):
    meta = json.loads(json.dumps(res.meta))
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
        assert mine["mode"] == ref["mode"]
        parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}")
    parity.compare_nested(gmeta, meta, "meta", atol=2e-5, rtol=2e-5)  # the reference's own A/B tolerance
    if want_frames is not None:
        err = np.abs(np.asarray(res.frames) - want_frames)
        assert float((err > 2e-5).mean()) <= 1e-3 and float(err.max()) <= 0.04
        assert float((np.asarray(res.masks) != want_masks).mean()) <= 1e-3
    if gold is not None:
        assert tuple(np.asarray(res.frames).shape) == tuple(gold["shape"])
        f, y, x, hh, ww = gold["patch0_at"]
        assert float(np.abs(np.asarray(res.frames)[f, y:y + hh, x:x + ww] - gold["patch0"]).max()) <= parity.TOL_PIXEL["bilinear"]
        assert float(np.abs(np.asarray(res.masks)[f, y:y + hh, x:x + ww, 0] - gold["mpatch0"]).max()) == 0.0


SMALL_FLOW = [c for c in cases.SMALL_STABILIZER_CASES if c["node"] == "flow" and c["framing"] != "crop"]


@pytest.mark.parametrize("case", SMALL_FLOW, ids=[c["name"] for c in SMALL_FLOW])
def test_host_path_on_small_clips(monkeypatch, case):
    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    res = _run(monkeypatch, cases.make_frames(case), case["framing"], case["mode"], case["camera_lock"], case["strength"],
               case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"])
    _check(res, gmeta, gold["frames"], gold["masks"])


@pytest.mark.parametrize("scenario", cases.AB_SCENARIOS[:2], ids=[s[0] for s in cases.AB_SCENARIOS[:2]])
def test_host_path_on_the_reference_ab_clip(monkeypatch, scenario):
    name, framing, mode, keep_fov = scenario
    gold = np.load(os.path.join(GOLDEN_DIR, "ab_73x45.npz"))
    with open(os.path.join(GOLDEN_DIR, "ab_73x45_meta.json")) as fh:
        gmeta = json.load(fh)[f"flow.{name}"]
    a = cases.AB_ARGS
    res = _run(monkeypatch, gold["input"], framing, mode, a["camera_lock"], a["strength"], a["smooth"], keep_fov, a["padding_rgb"], a["fps"])
    _check(res, gmeta, gold[f"flow.{name}.frames"], gold[f"flow.{name}.masks"])


def test_host_path_with_a_working_size(monkeypatch):
    """1080p clip: estimation at 960x540, transforms rescaled to full resolution (stabilizer_utils.py:279-297)."""
    case = next(c for c in cases.STABILIZER_CASES if c["name"] == "flow_sim_pad_1080p")
    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    res = _run(monkeypatch, cases.make_frames(case), case["framing"], case["mode"], case["camera_lock"], case["strength"],
               case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"])
    _check(res, gmeta, gold=gold)


# ---- Motion Apply: the product's engine (meta resolution, framing, shutter samples) with the numpy resampler -------

def _oracle_sample_warp(context, fwd, out_size, interpolation, border, *, want_mask=True, **kw):
    """pipeline.fused_warp for [N,S,9] shutter samples: f32 accumulate in sample order, / S, soft mask."""
    import torch

    fwd = np.asarray(fwd, np.float32)
    n, s = fwd.shape[0], fwd.shape[1]
    ow, oh = int(out_size[0]), int(out_size[1])
    frames = np.zeros((n, oh, ow, 3), np.float32)
    masks = np.zeros((n, oh, ow), np.float32)
    for i in range(n):
        acc = np.zeros((oh, ow, 3), np.float32)
        cov = np.zeros((oh, ow), np.float32)
        for k in range(s):
            m = fwd[i, k].reshape(3, 3)
            acc += resample_np.warp_np(context.frames[i], m, out_size, interpolation, border)
            cov += resample_np.coverage_np(m, (context.width, context.height), out_size).astype(np.float32)
        frames[i] = acc / np.float32(s) if s > 1 else acc
        mk = np.float32(1.0) - cov / np.float32(s)
        mk[mk < 1e-3] = 0.0
        masks[i] = mk
    return torch.from_numpy(frames), (torch.from_numpy(masks) if want_mask else None), None


def _oracle_common_valid_mask(context, input_size, output_size, matrices, progress_callback=None):
    """motion_apply.common_valid_mask (vstab_common_coverage) in numpy."""
    from vstab_b200 import motion_apply as ma

    out = np.ones((output_size[1], output_size[0]), dtype=bool)
    for m in matrices:
        out &= resample_np.coverage_np(np.asarray(m, np.float32), input_size, output_size)
    ma._tick(progress_callback, len(matrices))
    return out


SMALL_APPLY = [c for c in cases.MOTION_APPLY_CASES if c["store"] == "full"]


@pytest.mark.parametrize("case", SMALL_APPLY, ids=[c["name"] for c in SMALL_APPLY])
def test_motion_apply_engine_on_the_cpu(monkeypatch, case):
    from vstab_b200 import motion_apply as ma

    monkeypatch.setattr(ma, "fused_warp", _oracle_sample_warp)
    monkeypatch.setattr(ma, "common_valid_mask", _oracle_common_valid_mask)
    gold = np.load(os.path.join(GOLDEN_DIR, f"apply_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"apply_{case['name']}_meta.json")) as fh:
        meta = json.load(fh)
    ticks = [0]
    res = ma.apply_motion(_Clip(cases.make_frames(case)), meta, case["padding_rgb"], framing_mode=case["framing"],
                          interpolation=case["interp"], motion_blur=case["blur"], motion_blur_samples=case["samples"],
                          progress_callback=lambda: ticks.__setitem__(0, ticks[0] + 1))
    assert res.frames.shape == gold["frames"].shape and res.masks.shape == gold["masks"].shape
    assert np.array_equal(res.frames, gold["frames"])  # bilinear and bicubic (cv2's row-sum order), with and without blur
    assert float(np.abs(res.masks - gold["masks"]).max()) <= (0.0 if case["blur"] == 0.0 else 1e-6)
    n, s = case["n"], (int(np.clip(case["samples"], 3, 33)) if case["blur"] > 0 else 1)
    assert ticks[0] == n * s + (n if case["framing"] == "crop" else 0)  # scripts/check_motion_meta.py:366-394
    assert res.meta["motion_apply"]["framing_mode"] == ("crop_and_pad" if case["framing"] == "pad" else case["framing"])


# ---- Classic: corners + tracks from the C oracle, same host path ------------------------------------------------

def _oracle_classic_estimator(context, work_w, work_h, requested, exact_lk=False):
    """classic.estimate_candidates with oracle/classic_ref.c in place of K5/K6 and fit_np in place of K7-K9."""
    from oracle import classic_ref as CR
    from vstab_b200.stabilizer_core import PairCandidates

    work = None if (work_w, work_h) == (context.width, context.height) else (work_w, work_h)
    gray = [gray_np.gray_for_estimation(f, work) for f in context.frames]
    P = len(gray) - 1
    m = np.tile(np.eye(3), (P, 3, 1, 1))
    res, inl, valid, total, ok = (np.zeros((P, 3)) for _ in range(5))
    detected = np.zeros(P, np.int64)
    for p in range(P):
        feats = CR.good_features(gray[p])
        detected[p] = len(feats)
        if len(feats) == 0:
            continue
        nxt, st = CR.pyr_lk(gray[p], gray[p + 1], feats, exact=exact_lk)
        prev, curr = feats[st == 1], nxt[st == 1].astype(np.float32)
        valid[p], total[p] = len(prev), 400
        if len(prev) == 0:
            continue
        m[p, 0, :2, 2] = fit_np.median_shift(prev, curr)
        inl[p, 0], ok[p, 0] = len(prev), 1
        if requested in ("similarity", "perspective") and len(prev) >= 3:
            A, mask = fit_np.estimate_affine_partial_2d(prev, curr)
            if A is not None:
                m[p, 1, :2] = A
                inl[p, 1], ok[p, 1] = int(mask.sum()), 1
        if requested == "perspective" and len(prev) >= 4:
            H, mask = fit_np.find_homography(prev, curr)
            if H is not None:
                m[p, 2] = H
                inl[p, 2], ok[p, 2] = int(mask.sum()), 1
    return PairCandidates(m, res, inl.astype(int), valid.astype(int), total.astype(int), ok.astype(int), min_points=8, detected=detected)


CLASSIC_AB = [s for s in cases.AB_SCENARIOS[:2]]


@pytest.mark.parametrize("scenario", CLASSIC_AB, ids=[s[0] for s in CLASSIC_AB])
def test_classic_host_path_on_the_reference_ab_clip(monkeypatch, scenario):
    """Tracks from the C oracle in cv2's lane order carry cv2's bits, so Classic meets the 2e-5 of the Flow chain."""
    from vstab_b200 import stabilizer_core as core

    name, framing, mode, keep_fov = scenario
    gold = np.load(os.path.join(GOLDEN_DIR, "ab_73x45.npz"))
    with open(os.path.join(GOLDEN_DIR, "ab_73x45_meta.json")) as fh:
        gmeta = json.load(fh)[f"classic.{name}"]
    a = cases.AB_ARGS
    monkeypatch.setattr(core, "fused_warp", _oracle_warp)
    res = core.stabilize_frames(_Clip(gold["input"]), framing, mode, a["camera_lock"], a["strength"], a["smooth"], keep_fov,
                                a["padding_rgb"], a["fps"], estimator=functools.partial(_oracle_classic_estimator, exact_lk=True),
                                flavour="classic", output="device")
    meta = json.loads(json.dumps(res.meta))
    assert meta["transform_mode_applied"] == gmeta["transform_mode_applied"]
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
        assert mine["mode"] == ref["mode"] and "residual" not in mine
        parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}")
    parity.compare_nested(gmeta, meta, "meta", atol=2e-5, rtol=2e-5)
    err = np.abs(np.asarray(res.frames) - gold[f"classic.{name}.frames"])
    assert float(err.mean()) <= 1e-3 and float(err.max()) <= 0.05


# ---- crop framing: keep_fov search + no-padding refinement with numpy coverage in place of the two coverage kernels ----

class _CoverageOracle:
    """Stand-in for the handle methods crop.py calls (vstab_coverage_bbox, vstab_common_coverage)."""

    @staticmethod
    def _cov(fwd, src_size, out_size):
        return [resample_np.coverage_np(np.asarray(m, np.float32).reshape(3, 3), src_size, out_size) for m in fwd.numpy()]

    def coverage_bbox(self, fwd, src_size, out_size, mask_rule=0):
        import torch

        out = []
        for c in self._cov(fwd, src_size, out_size):
            p = np.pad(c, 1, constant_values=False)   # cv2.dilate ignores what lies outside the image
            d = np.zeros_like(c)
            for dy in range(3):
                for dx in range(3):
                    d |= p[dy:dy + c.shape[0], dx:dx + c.shape[1]]
            p = np.pad(d, 1, constant_values=True)    # and so does cv2.erode
            e = np.ones_like(c)
            for dy in range(3):
                for dx in range(3):
                    e &= p[dy:dy + c.shape[0], dx:dx + c.shape[1]]
            ys, xs = np.nonzero(e)
            out.append([xs.min(), ys.min(), xs.max(), ys.max()] if len(xs) else [0, 0, -1, -1])
        return torch.tensor(out, dtype=torch.int32)

    def common_coverage(self, fwd, src_size, out_size, mask_rule=0):
        import torch

        common = np.ones((out_size[1], out_size[0]), bool)
        for c in self._cov(fwd, src_size, out_size):
            common &= c
        return torch.from_numpy(common.astype(np.uint8))


CROP_CPU = [c for c in cases.SMALL_STABILIZER_CASES + cases.CROP_CASES
            if c["node"] == "flow" and c["framing"] == "crop" and c["name"] in ("flow_trans_crop_121x73", "flow_sim_crop06_480p")]


@pytest.mark.parametrize("case", CROP_CPU, ids=[c["name"] for c in CROP_CPU])
def test_crop_solver_host_path(monkeypatch, case):
    import torch

    from vstab_b200 import crop

    monkeypatch.setattr(crop._native, "get_handle", lambda device: _CoverageOracle())
    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    from vstab_b200 import stabilizer_core as core

    monkeypatch.setattr(core, "fused_warp", _oracle_warp)
    clip = _Clip(cases.make_frames(case))
    clip.device = torch.device("cpu")
    res = core.stabilize_frames(clip, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                case["keep_fov"], case["padding_rgb"], case["fps"], estimator=_oracle_estimator, flavour="flow",
                                output="device")
    _check(res, gmeta, gold=gold)
    assert float(np.asarray(res.masks).max()) == 0.0
    cw, ch = res.meta["framing"]["crop_size"]
    assert abs(cw / ch - case["w"] / case["h"]) <= 1e-6


# ---- legacy inverse stabilization (helper + deprecated node) --------------------------------------------------------

def _inverse_golden():
    gold = np.load(os.path.join(GOLDEN_DIR, "inverse_73x45.npz"))
    with open(os.path.join(GOLDEN_DIR, "inverse_73x45_meta.json")) as fh:
        return gold, json.load(fh)


class _RgbClip(_Clip):
    channels = 3


@pytest.mark.parametrize("name", ["expand", "crop"])
def test_inverse_stabilization_helper_on_the_cpu(monkeypatch, name):
    """scripts/check_inverse_stabilization.py scenario; golden = the reference's _apply_inverse_stabilization."""
    from vstab_b200 import inverse

    monkeypatch.setattr(inverse, "fused_warp", _oracle_sample_warp)
    gold, metas = _inverse_golden()
    res = inverse.apply_inverse_stabilization(_RgbClip(gold[f"{name}.input"]), metas[f"{name}.in"], (127, 127, 127))
    assert np.array_equal(res.frames, gold[f"{name}.frames"]) and np.array_equal(res.masks, gold[f"{name}.masks"])
    parity.compare_nested(metas[f"{name}.helper_out"], json.loads(json.dumps(res.meta)), "meta", atol=0.0, rtol=0.0)
    if name == "expand":  # the script's own acceptance numbers (:169-173)
        err = np.abs(res.frames - gold["source"])
        assert float(np.quantile(err, 0.99)) <= 0.3 and float(err.mean()) <= 0.035
    else:
        assert float(res.masks.max()) > 0.0  # crop framing: some source pixels cannot be recovered (:178-179)
    bad = json.loads(json.dumps(metas[f"{name}.in"]))
    bad["stabilization_warp"]["per_frame"][2]["index"] = 7
    with pytest.raises(ValueError, match=r"per_frame\[2\]\.index must be 2"):
        inverse.apply_inverse_stabilization(_RgbClip(gold[f"{name}.input"]), bad, (0, 0, 0))
    with pytest.raises(ValueError, match="Frame count mismatch"):
        inverse.apply_inverse_stabilization(_RgbClip(gold[f"{name}.input"][:3]), metas[f"{name}.in"], (0, 0, 0))
    with pytest.raises(ValueError, match="must match stabilization_warp.output_size"):
        inverse.apply_inverse_stabilization(_RgbClip(gold["source"][:, :40]), metas[f"{name}.in"], (0, 0, 0))
    with pytest.raises(ValueError, match="stabilization_warp is required"):
        inverse.apply_inverse_stabilization(_RgbClip(gold["source"]), {}, (0, 0, 0))


def test_inverse_node_on_the_cpu(monkeypatch):
    """The deprecated Video Stabilizer Inverse node (nodes/video_stabilizer_inverse.py:60-100) against its own output."""
    import sys

    import torch

    from tests.test_abi_and_host import _install_comfy_stubs

    _install_comfy_stubs()
    sys.modules.pop("vstab_b200.nodes", None)
    from vstab_b200 import motion_apply as ma, nodes

    gold, metas = _inverse_golden()
    monkeypatch.setattr(ma, "fused_warp", _oracle_sample_warp)
    monkeypatch.setattr(nodes, "normalize_video_input", lambda frames: _RgbClip(frames.numpy()))
    monkeypatch.setattr(nodes, "reconstruct_video", lambda frames, ctx: torch.from_numpy(np.ascontiguousarray(frames)))
    monkeypatch.setattr(nodes, "convert_masks_for_output", lambda masks: torch.from_numpy(np.ascontiguousarray(masks[..., 0])))
    video, mask, meta = nodes.VideoStabilizerInverse.execute(torch.from_numpy(gold["expand.input"]), metas["expand.in"], "#0AC85A")
    assert np.array_equal(video.numpy(), gold["node.frames"]) and np.array_equal(mask.numpy(), gold["node.mask"])
    parity.compare_nested(metas["node.out"], json.loads(json.dumps(meta)), "meta", atol=0.0, rtol=0.0)
    schema = nodes.VideoStabilizerInverse.define_schema()
    assert schema.node_id == "video_stabilizer_inverse" and schema.is_deprecated is True
    assert [s.args[0] for s in schema.inputs] == ["frames", "meta", "padding_color"]
    assert [s.args[0] for s in schema.outputs] == ["frames_restored", "padding_mask", "meta"]


def test_check_motion_meta_script_head(monkeypatch):
    """scripts/check_motion_meta.py:144-205 with the product's modules: v2 block round trip, legacy stabilization_warp
    inversion, applied block, and the block Motion Apply picks by the size of the frames it is given (a full stabilizer
    meta replays forwards on source-sized frames and backwards on stabilized-sized ones)."""
    from vstab_b200 import hostmath as hm, motion_apply as ma, motion_meta as mm

    monkeypatch.setattr(ma, "fused_warp", _oracle_sample_warp)
    matrices = [np.eye(3), np.array([[1.0, 0.0, 2.5], [0.0, 1.0, -1.25], [0.0, 0.0, 1.0]])]
    block = mm.build_motion_meta_v2(source="generated_shake", frame_count=2, fps=16.0, input_size=(64, 48), output_size=(64, 48),
                                    matrices=matrices, generator={"node": "t"})
    r = mm.resolve_motion_meta({"motion_meta": block})
    assert r.frame_count == 2 and r.input_size == (64, 48) and r.output_size == (64, 48)
    warp = hm.build_stabilization_warp_meta(source_size=(80, 50), output_size=(96, 60), framing_mode="expand", applied_matrices=matrices)
    assert mm.motion_meta_from_stabilization_warp(warp, fps=24.0, source="legacy_stabilization") is not None
    legacy = mm.resolve_motion_meta({"stabilization_warp": warp})
    assert legacy.input_size == (96, 60) and legacy.output_size == (80, 50)
    assert np.allclose(legacy.per_frame[1].matrix, np.linalg.inv(matrices[1]))
    applied = mm.applied_motion_meta_from_stabilization_warp(warp, fps=24.0, source="estimated_flow")
    ra = mm.resolve_motion_meta({"motion_meta": applied})
    assert ra.input_size == (80, 50) and ra.output_size == (96, 60) and np.allclose(ra.per_frame[1].matrix, matrices[1])
    frames = np.random.default_rng(0).random((2, 50, 80, 3), dtype=np.float32)
    combined = {"stabilization_warp": warp, "motion_meta": applied}
    direct = ma.apply_motion(_RgbClip(frames), combined, (127, 127, 127))
    assert direct.frames.shape[1:3] == (60, 96) and direct.meta["motion_apply"]["source"] == "estimated_flow"
    back = ma.apply_motion(_RgbClip(direct.frames), combined, (127, 127, 127))
    assert back.frames.shape[1:3] == (50, 80) and back.meta["motion_apply"]["source"] == "legacy_stabilization"
    # frame 1 went 2.5 px right / 1.25 px up and back: the interior returns to the source up to two bilinear passes
    assert float(np.abs(back.frames[0] - frames[0]).max()) <= 1e-6
    with pytest.raises(ValueError, match="motion_meta or stabilization_warp"):
        mm.resolve_motion_meta({})


def test_check_motion_meta_script_tail_on_the_cpu(monkeypatch):
    """scripts/check_motion_meta.py:289-415 (identity apply, blur == 0 baseline path, expand canvas, blur determinism,
    tick counts, crop fallback) and the error messages: the bodies of the GPU tests, run with the numpy resampler."""
    import tests.test_reference_checks_gpu as g
    from vstab_b200 import motion_apply as ma

    monkeypatch.setattr(ma, "fused_warp", _oracle_sample_warp)
    monkeypatch.setattr(ma, "common_valid_mask", _oracle_common_valid_mask)
    monkeypatch.setattr(g, "_ctx", lambda frames: _RgbClip(np.asarray(frames)))
    g.test_identity_blur_zero_expand_determinism_and_ticks()
    g.test_errors_match_the_reference_messages()



def test_host_path_perspective_camera_lock(monkeypatch):
    """Perspective ladder + camera_lock (target path 0, smooth >= 0.85) on the 1080p golden clip."""
    case = next(c for c in cases.STABILIZER_CASES if c["name"] == "flow_persp_lock_1080p")
    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{case['name']}_meta.json")) as fh:
        gmeta = json.load(fh)
    res = _run(monkeypatch, cases.make_frames(case), case["framing"], case["mode"], case["camera_lock"], case["strength"],
               case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"])
    meta = json.loads(json.dumps(res.meta))
    assert meta["transform_mode_applied"] == "perspective" and meta["camera_lock"] is True
    for mine, ref in zip(meta["estimated_motion"]["per_transition"], gmeta["estimated_motion"]["per_transition"]):
        assert mine["mode"] == ref["mode"] == "perspective"
        parity.assert_transform_close(mine["matrix"], ref["matrix"], f"pair {ref['index']}")
    # homography refinement (LM) agrees with cv2 to ~1e-8 per entry, which the 8 perspective parameters carry through
    parity.compare_nested(gmeta, meta, "meta", atol=2e-4, rtol=2e-4)
    f, y, x, hh, ww = gold["patch0_at"]
    assert float(np.abs(np.asarray(res.frames)[f, y:y + hh, x:x + ww] - gold["patch0"]).max()) <= parity.TOL_PIXEL["bilinear"]


@pytest.mark.reference
def test_crop_solvers_equal_the_reference_on_random_paths(monkeypatch, reference_nodes):
    """keep_fov search + no-padding refinement + aspect rectangle against the unmodified reference's functions
    (stabilizer_utils.py:448-837) on random stabilisation paths, every keep_fov regime (disabled / met / clamped /
    failed / no overlap); the two coverage kernels are replaced by their numpy statement."""
    import torch

    from vstab_b200 import crop, hostmath as hm

    U = reference_nodes.stabilizer_utils
    monkeypatch.setattr(crop._native, "get_handle", lambda device: _CoverageOracle())
    rng = np.random.default_rng(17)
    dev = torch.device("cpu")
    seen = set()
    for trial in range(24):
        w, h = int(rng.integers(48, 200)), int(rng.integers(32, 120))
        n = int(rng.integers(2, 9))
        mode = ("translation", "similarity")[trial % 2]
        amp = float(rng.choice([0.5, 3.0, 12.0, 60.0, 400.0]))
        deltas = rng.normal(0, 1, (n, 2 if mode == "translation" else 4)) * (np.array([amp, amp]) if mode == "translation" else
                                                                               np.array([amp, amp, 0.01, 0.01]))
        keep = float(rng.choice([0.0, 0.3, 0.6, 0.9, 0.97]))
        margin = max(0.5, 0.02 * max(w, h))
        want = U._compute_crop_with_keep_fov_parametric(U._params_to_matrix, mode, list(deltas), w, h, keep, margin, return_masks=False)
        w_final, w_pre, _masks, w_ratio, w_status, w_note, w_scale, w_origin, w_size = want
        got = crop.compute_crop_with_keep_fov(dev, mode, deltas, w, h, keep, margin)
        g_final, g_pre, g_ratio, g_status, g_note, g_scale, g_origin, g_size = got
        assert (g_status, g_note, g_scale) == (w_status, w_note, w_scale), trial
        assert g_ratio == w_ratio and list(g_origin) == list(w_origin) and list(g_size) == list(w_size), trial
        assert np.array_equal(np.asarray(g_final), np.stack(w_final)) and np.array_equal(np.asarray(g_pre), np.stack(w_pre)), trial
        seen.add((w_status, w_note is None))
        r_final, _m, r_origin, r_size, r_eff = U._refine_no_padding_crop(w_final, w, h, 1)
        q_final, q_origin, q_size, q_eff = crop.refine_no_padding_crop(dev, np.stack(w_final), w, h, 1)
        assert q_eff == r_eff and list(q_origin) == list(r_origin) and list(q_size) == list(r_size), trial
        assert np.array_equal(np.asarray(q_final), np.stack(r_final)), trial
        mask = rng.random((h, w)) > 0.02
        mask[: int(rng.integers(0, 6))] = False
        assert crop.largest_aspect_ratio_rectangle(mask.astype(np.uint8), w, h) == U._largest_aspect_ratio_rectangle(mask.astype(np.uint8), w, h)
    assert len({s for s, _ in seen}) >= 3, seen  # the draw reached several regimes


@pytest.mark.reference
def test_flow_host_path_equals_the_live_reference_on_random_settings(monkeypatch, reference_nodes):
    """Random small clips x random node settings (framing, model, camera_lock, strength, smooth, keep_fov, fps, padding
    colour): the unmodified reference's _stabilize_frames, run here, against the product's host path with the oracle in
    place of the GPU stages.  Sweeps what the goldens sample: sticky ladder fall-backs on frames with too few grid
    points, perspective, crop regimes, expand canvases, the keep_fov bypass."""
    import torch

    import synth
    from vstab_b200 import crop, stabilizer_core as core

    ref = reference_nodes.video_stabilizer_flow
    U = reference_nodes.stabilizer_utils
    monkeypatch.setattr(core, "fused_warp", _oracle_warp)
    monkeypatch.setattr(crop._native, "get_handle", lambda device: _CoverageOracle())
    rng = np.random.default_rng(23)
    modes_seen, framings_seen = set(), set()
    ran = 0
    for trial in range(48):
        w, h = int(rng.integers(24, 140)), int(rng.integers(20, 100))
        if dis_ref.select_scales(h, w) is None:
            continue
        f0, c0 = dis_ref.select_scales(h, w)
        if min(h >> c0, w >> c0) < 8:      # cv2 itself reads out of bounds on these sizes
            continue
        n = int(rng.integers(2, 8))
        seed = 500 + trial
        base = synth.base_texture(seed, w, h).numpy()
        mats = synth.shake_matrices(n, seed, w, h, perspective=bool(trial % 3 == 0), amount=float(rng.choice([0.5, 2.0, 5.0])))
        frames = synth.render_clip_numpy(base, mats, w, h)
        framing = str(rng.choice(["crop_and_pad", "expand", "crop"]))
        mode = str(rng.choice(["translation", "similarity", "perspective"]))
        args = (framing, mode, bool(rng.random() < 0.3), float(rng.choice([0.0, 0.4, 0.7, 1.0])), float(rng.choice([0.0, 0.5, 1.0])),
                float(rng.choice([0.0, 0.6, 0.95, 1.0])), tuple(int(v) for v in rng.integers(0, 256, 3)), float(rng.choice([8.0, 16.0, 30.0])))
        want = ref._stabilize_frames(U._normalize_video_input([f for f in frames]), *args)
        clip = _Clip(frames)
        clip.device = torch.device("cpu")
        clip.untouched = lambda output, fr=frames: (fr, np.zeros(fr.shape[:3] + (1,), np.float32))
        got = core.stabilize_frames(clip, *args, estimator=_oracle_estimator, flavour="flow", output="device")
        gmeta, meta = json.loads(json.dumps(want.meta)), json.loads(json.dumps(got.meta))
        parity.compare_nested(gmeta, meta, f"trial {trial} {args} {w}x{h}x{n}: meta", atol=2e-5, rtol=2e-5)
        wf, wm = np.asarray(want.frames, np.float32), np.asarray(want.masks, np.float32)
        gf, gm = np.asarray(got.frames), np.asarray(got.masks)
        assert gf.shape == wf.shape and gm.shape == wm.shape, trial
        err = np.abs(gf - wf)
        assert float((err > 2e-5).mean()) <= 2e-3 and float(err.max()) <= 0.05, (trial, float((err > 2e-5).mean()), float(err.max()))
        assert float((gm != wm).mean()) <= 2e-3, trial
        modes_seen.add(meta.get("transform_mode_applied"))
        framings_seen.add(framing)
        ran += 1
    assert ran >= 30 and len(framings_seen) == 3 and len(modes_seen) >= 4, (ran, framings_seen, modes_seen)


@pytest.mark.reference
def test_classic_host_path_equals_the_live_reference_with_exact_lk(monkeypatch, reference_nodes):
    """The same sweep for Classic, with the tracks summed in cv2's lane order (classic_ref.pyr_lk(exact=True)): corners,
    tracks, fits, ladder and framing then agree with the unmodified reference to its own 2e-5 -- what the CUDA tracker
    gains by adopting that order (parity items left for a next round)."""
    import functools

    import torch

    import synth
    from vstab_b200 import crop, stabilizer_core as core

    ref = reference_nodes.video_stabilizer_classic
    U = reference_nodes.stabilizer_utils
    monkeypatch.setattr(core, "fused_warp", _oracle_warp)
    monkeypatch.setattr(crop._native, "get_handle", lambda device: _CoverageOracle())
    est = functools.partial(_oracle_classic_estimator, exact_lk=True)
    rng = np.random.default_rng(29)
    modes_seen, ran = set(), 0
    for trial in range(30):
        w, h = int(rng.integers(60, 260)), int(rng.integers(48, 160))
        n = int(rng.integers(2, 7))
        seed = 700 + trial
        base = synth.base_texture(seed, w, h).numpy()
        mats = synth.shake_matrices(n, seed, w, h, perspective=bool(trial % 3 == 0), amount=float(rng.choice([0.5, 2.0, 5.0])))
        frames = synth.render_clip_numpy(base, mats, w, h)
        framing = str(rng.choice(["crop_and_pad", "expand", "crop"]))
        mode = str(rng.choice(["translation", "similarity", "perspective"]))
        args = (framing, mode, bool(rng.random() < 0.3), float(rng.choice([0.0, 0.4, 0.7, 1.0])), float(rng.choice([0.0, 0.5, 1.0])),
                float(rng.choice([0.0, 0.6, 0.95, 1.0])), tuple(int(v) for v in rng.integers(0, 256, 3)), float(rng.choice([8.0, 16.0, 30.0])))
        want = ref._stabilize_frames(U._normalize_video_input([f for f in frames]), *args)
        clip = _Clip(frames)
        clip.device = torch.device("cpu")
        clip.untouched = lambda output, fr=frames: (fr, np.zeros(fr.shape[:3] + (1,), np.float32))
        got = core.stabilize_frames(clip, *args, estimator=est, flavour="classic", output="device")
        gmeta, meta = json.loads(json.dumps(want.meta)), json.loads(json.dumps(got.meta))
        parity.compare_nested(gmeta, meta, f"trial {trial} {args} {w}x{h}x{n}: meta", atol=2e-5, rtol=2e-5)
        err = np.abs(np.asarray(got.frames) - np.asarray(want.frames, np.float32))
        assert float((err > 2e-5).mean()) <= 2e-3 and float(err.max()) <= 0.05, (trial, float((err > 2e-5).mean()), float(err.max()))
        modes_seen.add(meta.get("transform_mode_applied"))
        ran += 1
    assert ran == 30 and len(modes_seen) >= 3, (ran, modes_seen)



@pytest.mark.reference
def test_motion_apply_engine_equals_the_live_reference_on_random_settings(monkeypatch, reference_nodes):
    """Shake Generator metas (every style of the reference's shake_noise, random seeds / amounts) x random Motion Apply
    settings through the unmodified apply_motion and through the product's engine with the numpy resampler."""
    from vstab_b200 import motion_apply as ma

    monkeypatch.setattr(ma, "fused_warp", _oracle_sample_warp)
    monkeypatch.setattr(ma, "common_valid_mask", _oracle_common_valid_mask)
    R, SN, U = reference_nodes.motion_apply, reference_nodes.shake_noise, reference_nodes.stabilizer_utils
    rng = np.random.default_rng(31)
    styles = sorted(SN.STYLES)
    effective = set()
    for trial in range(24):
        w, h, n = int(rng.integers(16, 120)), int(rng.integers(12, 90)), int(rng.integers(1, 7))
        style = styles[trial % len(styles)]
        block = SN.generate_shake_motion_meta(recipe=SN.STYLES[style], frame_count=n, width=w, height=h, fps=float(rng.choice([8.0, 16.0, 24.0])),
                                              amount=float(rng.choice([0.5, 2.0, 8.0, 40.0])), speed=float(rng.choice([0.5, 1.0, 3.0])),
                                              seed=int(rng.integers(0, 1000)), node="shake_generator", style=style)
        meta = {"motion_meta": block}
        frames = rng.random((n, h, w, 3), dtype=np.float32)
        kw = dict(framing_mode=str(rng.choice(["crop_and_pad", "pad", "crop", "expand"])), interpolation=str(rng.choice(["bilinear", "bicubic"])),
                  motion_blur=float(rng.choice([0.0, 0.0, 0.5, 1.0])), motion_blur_samples=int(rng.choice([1, 5, 9, 33, 50])))
        rgb = tuple(int(v) for v in rng.integers(0, 256, 3))
        ticks = [0, 0]
        want = R.apply_motion(U._normalize_video_input([f for f in frames]), meta, rgb, progress_callback=lambda: ticks.__setitem__(0, ticks[0] + 1), **kw)
        got = ma.apply_motion(_RgbClip(frames), json.loads(json.dumps(meta)), rgb, progress_callback=lambda: ticks.__setitem__(1, ticks[1] + 1), **kw)
        assert ticks[0] == ticks[1], (trial, kw, ticks)
        assert got.frames.shape == want.frames.shape and got.masks.shape == want.masks.shape, (trial, kw)
        assert np.array_equal(got.frames, want.frames), (trial, kw)
        assert float(np.abs(got.masks - want.masks).max()) <= (0.0 if kw["motion_blur"] == 0.0 else 1e-6), (trial, kw)
        assert json.loads(json.dumps(got.meta)) == json.loads(json.dumps(want.meta)), (trial, kw)
        effective.add(got.meta["motion_apply"]["framing_mode"] + ("+fallback" if "framing_fallback" in got.meta else ""))
    assert {"crop_and_pad", "crop", "expand"} <= {e.split("+")[0] for e in effective}, effective
