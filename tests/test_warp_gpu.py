"""CUDA fused resampler (vstab_warp_fused) vs the numpy oracle and the reference goldens.

Tolerances are north_star's: pixels <= 1e-3 (bilinear) / 2e-3 (bicubic) max-abs on float32 [0,1],
padding mask bit-exact.  The asserted bounds are much tighter because the kernel reproduces cv2's
arithmetic in cv2's order of operations: bilinear AND bicubic (row sums first, see oracle/resample_np.py) are
bit-exact against the oracle and against the goldens of the unmodified reference.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import apply_np, resample_np as R
from tests import cases
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

TOL = {"bilinear": 1e-3, "bicubic": 2e-3}      # north_star


def _rand_matrix(rng, kind, shift=15.0):
    th, s = rng.normal(0, 0.01), 1 + rng.normal(0, 0.01)
    tx, ty = rng.normal(0, shift, 2)
    m = np.array([[s * np.cos(th), -s * np.sin(th), tx], [s * np.sin(th), s * np.cos(th), ty], [0, 0, 1]])
    if kind == "persp":
        m[2, :2] = rng.normal(0, 2e-5, 2)
    return m.astype(np.float32)


def _run(handle, src, mats, out_size, interp, border, **kw):
    from vstab_b200 import _native

    dev = torch.device("cuda", 0)
    s = torch.from_numpy(np.ascontiguousarray(src)).to(dev)
    f = torch.from_numpy(np.ascontiguousarray(mats, dtype=np.float32)).to(dev)
    dst, mask, pad = handle.warp_fused(s, f, out_size, interp, border, want_pad_count=True, **kw)
    torch.cuda.synchronize()
    return dst.cpu().numpy(), mask.cpu().numpy(), pad.cpu().numpy()


@pytest.mark.parametrize("size", [(121, 73), (320, 200), (832, 480)])
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("stage", [0, 1])
def test_single_sample_matches_oracle(handle, size, interp, stage):
    rng = np.random.default_rng(7)
    w, h = size
    n = 4
    src = rng.random((n, h, w, 3), dtype=np.float32)
    border = (0.5, 0.25, 0.75)
    for out_size in [(w, h), (w + 10, h + 8)]:
        mats = np.stack([_rand_matrix(rng, "sim" if i % 2 == 0 else "persp") for i in range(n)])
        got, mask, pad = _run(handle, src, mats.reshape(n, 1, 9), out_size, interp, border, stage_mode=stage)
        for i in range(n):
            want = R.warp_np(src[i], mats[i], out_size, interp, border)
            err = float(np.abs(got[i] - want).max())
            assert err <= TOL[interp]
            assert np.array_equal(got[i], want), (interp, out_size, i, err)
            want_mask = R.mask_np(mats[i], (w, h), out_size, R.RULE_P)
            assert np.array_equal(mask[i], want_mask)
            assert int(pad[i]) == int(want_mask.sum())


@pytest.mark.parametrize("rule", [R.RULE_P, R.RULE_C])
def test_mask_rules_and_ties(handle, rule):
    """Integer and half-pixel translations put whole rows/columns exactly on the rule boundary."""
    w, h = 200, 120
    src = np.random.default_rng(1).random((1, h, w, 3), dtype=np.float32)
    for t in [(3.0, -2.0), (0.5, 0.5), (-7.5, 4.0), (0.0, 0.0), (w - 1.0, 0.0), (-(w - 1.0) - 0.5, 1.5)]:
        m = np.array([[1, 0, t[0]], [0, 1, t[1]], [0, 0, 1]], np.float32)
        _, mask, _ = _run(handle, src, m.reshape(1, 1, 9), (w, h), "bilinear", (0, 0, 0), mask_rule=rule)
        assert np.array_equal(mask[0], R.mask_np(m, (w, h), (w, h), rule)), t


def test_degenerate_matrices_do_not_crash(handle):
    """Singular / horizon-crossing maps: the staged-tile path must fall back, results = oracle."""
    w, h = 160, 96
    src = np.random.default_rng(2).random((1, h, w, 3), dtype=np.float32)
    mats = [
        np.array([[1, 0, 0], [0, 1, 0], [0.02, 0.0, 1]], np.float32),      # horizon inside the frame
        np.array([[3.0, 0, -100], [0, 3.0, -50], [0, 0, 1]], np.float32),  # 3x zoom in
        np.array([[0.2, 0, 30], [0, 0.2, 20], [0, 0, 1]], np.float32),     # 5x zoom out: huge footprint
        np.array([[0, -1, 100], [1, 0, 0], [0, 0, 1]], np.float32),        # 90 degree rotation
    ]
    for m in mats:
        got, mask, _ = _run(handle, src, m.reshape(1, 1, 9), (w, h), "bilinear", (0.1, 0.2, 0.3))
        want = R.warp_np(src[0], m, (w, h), "bilinear", (0.1, 0.2, 0.3))
        assert float(np.abs(got[0] - want).max()) <= 1e-6
        assert np.array_equal(mask[0], R.mask_np(m, (w, h), (w, h)))


@pytest.mark.parametrize("interp,samples", [("bilinear", 5), ("bicubic", 9), ("bilinear", 33)])
def test_motion_blur_matches_oracle(handle, interp, samples):
    from vstab_b200.motion_apply import sample_matrices

    rng = np.random.default_rng(9)
    w, h, n = 160, 96, 3
    src = rng.random((n, h, w, 3), dtype=np.float32)
    mats = [_rand_matrix(rng, "sim", 6.0).astype(np.float64) for _ in range(n)]
    fwd = sample_matrices(mats, 0.5, samples)
    got, mask, _ = _run(handle, src, fwd, (w + 4, h + 2), interp, (0.5, 0.5, 0.5))
    for i in range(n):
        want, want_mask = R.warp_blur_np(src[i], mats, i, (w + 4, h + 2), interp, (0.5, 0.5, 0.5), 0.5, samples)
        assert np.array_equal(got[i], want), float(np.abs(got[i] - want).max())
        assert np.array_equal(mask[i], want_mask)


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("kind", ["sim", "persp"])
def test_motion_blur_register_cached_tiles_equal_the_general_path(handle, interp, kind):
    """Interior blur tiles keep the texel footprint in registers across shutter samples
    (blur_interior_tile); VSTAB_STAGE_GLOBAL sends every tile through general_tile_body with global
    loads.  Same operations in the same order: every output bit, mask value and padded count agrees,
    for 16-byte and scalar output stores, small and fast motion (footprint moves every sample)."""
    from vstab_b200.motion_apply import sample_matrices

    rng = np.random.default_rng(21)
    w, h, n = 448, 260, 4
    src = rng.random((n, h, w, 3), dtype=np.float32)
    for motion, samples, out_size in [(2.0, 33, (w, h)), (25.0, 9, (w + 6, h + 5)), (0.3, 3, (w, h))]:
        mats = [_rand_matrix(rng, kind, motion).astype(np.float64) for _ in range(n)]
        fwd = sample_matrices(mats, 0.5, samples)
        staged = _run(handle, src, fwd, out_size, interp, (0.2, 0.4, 0.6), stage_mode=0)
        plain = _run(handle, src, fwd, out_size, interp, (0.2, 0.4, 0.6), stage_mode=1)
        assert np.array_equal(staged[0], plain[0]) and np.array_equal(staged[1], plain[1]) and np.array_equal(staged[2], plain[2])
        assert float(staged[1][:, 100:160, 200:260].max()) == 0.0  # the middle of the frame is covered by every sample


FULL = [c for c in cases.MOTION_APPLY_CASES]


@pytest.mark.parametrize("case", FULL, ids=[c["name"] for c in FULL])
def test_apply_motion_matches_reference_golden(case):
    """apply_motion (public driver) vs outputs of the real reference (tests/golden)."""
    from vstab_b200 import motion_apply, pipeline

    gold = np.load(os.path.join(GOLDEN_DIR, f"apply_{case['name']}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"apply_{case['name']}_meta.json")) as fh:
        meta = json.load(fh)
    frames = cases.make_frames(case)
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
    res = motion_apply.apply_motion(ctx, meta, case["padding_rgb"], framing_mode=case["framing"],
                                    interpolation=case["interp"], motion_blur=case["blur"],
                                    motion_blur_samples=case["samples"])
    tol = TOL[case["interp"]]
    tight = 0.0  # cv2's own order of additions in both interpolations; the blur accumulation is numpy's f32 sum / S
    if case["store"] == "full":
        assert res.frames.shape == gold["frames"].shape
        err = float(np.abs(res.frames - gold["frames"]).max())
        assert err <= tol and err <= tight, err
        if case["blur"] == 0.0:
            assert np.array_equal(res.masks, gold["masks"])
        else:
            assert float(np.abs(res.masks - gold["masks"]).max()) <= 1e-6
    else:
        assert tuple(res.frames.shape) == tuple(gold["shape"])
        fs = res.frames.reshape(res.frames.shape[0], -1).astype(np.float64).sum(axis=1)
        ms = res.masks.reshape(res.masks.shape[0], -1).astype(np.float64).sum(axis=1)
        assert np.array_equal(fs, gold["frame_sum"])
        assert np.allclose(ms, gold["mask_sum"], rtol=0, atol=1e-3)
        for k in range(3):
            f, y, x, hh, ww = gold[f"patch{k}_at"]
            err = float(np.abs(res.frames[f, y:y + hh, x:x + ww] - gold[f"patch{k}"]).max())
            assert err <= tol and err <= tight, (k, err)
            assert float(np.abs(res.masks[f, y:y + hh, x:x + ww, 0] - gold[f"mpatch{k}"]).max()) <= 1e-6


def test_full_size_properties_1080p(handle):
    """BASELINE-size checks that need no oracle: identity is exact, translation by whole pixels is a
    shifted copy, mask count equals the uncovered area."""
    dev = torch.device("cuda", 0)
    w, h, n = 1920, 1080, 4
    g = torch.Generator(device="cpu").manual_seed(3)
    src = torch.rand((n, h, w, 3), generator=g).to(dev)
    eye = torch.eye(3, dtype=torch.float32).reshape(1, 1, 9).repeat(n, 1, 1).to(dev)
    dst, mask, pad = handle.warp_fused(src, eye, (w, h), "bilinear", (0.5, 0.5, 0.5), want_pad_count=True)
    assert torch.equal(dst, src) and int(mask.sum()) == 0 and int(pad.sum()) == 0
    dst, mask, pad = handle.warp_fused(src, eye, (w, h), "bicubic", (0.5, 0.5, 0.5), want_pad_count=True)
    assert float((dst - src).abs().max()) <= 1e-6
    shift = torch.tensor([[1, 0, 7.0], [0, 1, -5.0], [0, 0, 1]], dtype=torch.float32).reshape(1, 1, 9).repeat(n, 1, 1).to(dev)
    dst, mask, pad = handle.warp_fused(src, shift, (w, h), "bilinear", (0.25, 0.5, 0.75), want_pad_count=True)
    assert torch.equal(dst[:, : h - 5, 7:], src[:, 5:, : w - 7])
    assert int(pad[0]) == w * h - (w - 7) * (h - 5)
    assert torch.equal(mask[:, : h - 5, 7:], torch.zeros_like(mask[:, : h - 5, 7:]))


@pytest.mark.parametrize("threads", [8, 4])
def test_mask_rule_auto_follows_cv2_per_frame_and_sample(handle, threads):
    """VSTAB_MASK_RULE_AUTO_THREADS(t): every (frame, shutter sample) gets the rule cv2 would pick for that call on a
    machine with t cv2 threads -- Rule P, or Rule C when one of the wheel's destination stripes misses the source
    (oracle: resample_np.auto_rule, pinned against the live wheel in tests/test_oracle_resample.py).  Shifts on both
    sides of the flip, single sample and motion blur whose samples straddle it; pixels must not change at all."""
    from vstab_b200 import _native
    from vstab_b200.motion_apply import sample_matrices

    rng = np.random.default_rng(17)
    w, h = 832, 480
    first = R.mask_stripes((w, h), threads)[0][1]
    shifts = [first - 2.0, first - 0.5, first + 0.5, first + 30.0, -(h - R.mask_stripes((w, h), threads)[-1][0]) - 1.0, 5.0]
    src = rng.random((len(shifts), h, w, 3), dtype=np.float32)
    mats = np.stack([np.array([[1, 0, -0.3], [0, 1, ty], [0, 0, 1]], np.float32) for ty in shifts])
    rules = [R.auto_rule(m, (w, h), (w, h), threads) for m in mats]
    assert R.RULE_P in rules and R.RULE_C in rules
    got, mask, pad = _run(handle, src, mats.reshape(-1, 1, 9), (w, h), "bilinear", (0.5, 0.5, 0.5), mask_rule=_native.mask_rule_auto(threads))
    ref, _, _ = _run(handle, src, mats.reshape(-1, 1, 9), (w, h), "bilinear", (0.5, 0.5, 0.5))
    assert np.array_equal(got, ref)  # the rule only concerns the mask
    for i, m in enumerate(mats):
        want = R.mask_np(m, (w, h), (w, h), rules[i])
        assert np.array_equal(mask[i], want), (i, shifts[i], rules[i])
        assert int(pad[i]) == int(want.sum())
    # one thread: AUTO is Rule P
    _, mask1, _ = _run(handle, src, mats.reshape(-1, 1, 9), (w, h), "bilinear", (0.5, 0.5, 0.5), mask_rule=_native.mask_rule_auto(1))
    for i, m in enumerate(mats):
        assert np.array_equal(mask1[i], R.mask_np(m, (w, h), (w, h), R.RULE_P))
    # motion blur: the shutter samples of frame 0 move across the flip
    seq = [np.array([[1, 0, -0.3], [0, 1, first - 6.0], [0, 0, 1]], np.float64), np.array([[1, 0, -0.3], [0, 1, first + 10.0], [0, 0, 1]], np.float64)]
    fwd = sample_matrices(seq, 1.0, 9)
    gotb, maskb, _ = _run(handle, src[:2], fwd, (w, h), "bilinear", (0.5, 0.5, 0.5), mask_rule=_native.mask_rule_auto(threads))
    for i in range(2):
        cov = np.zeros((h, w), np.float32)
        seen = set()
        for k in range(9):
            m32 = np.asarray(fwd[i, k], np.float32).reshape(3, 3)
            rule = R.auto_rule(m32, (w, h), (w, h), threads)
            seen.add(rule)
            cov += R.coverage_np(m32, (w, h), (w, h), rule).astype(np.float32)
        want = np.float32(1.0) - cov / np.float32(9)
        want[want < 1e-3] = 0.0
        assert np.array_equal(maskb[i], want), i
        if i == 0:
            assert seen == {R.RULE_P, R.RULE_C}


def test_mask_rule_auto_in_the_crop_coverage_kernels(handle):
    from vstab_b200 import _native

    w, h, threads = 640, 360, 8
    first = R.mask_stripes((w, h), threads)[0][1]
    mats = np.stack([np.array([[1, 0, -0.3], [0, 1, ty], [0, 0, 1]], np.float32) for ty in (3.0, first + 4.0, first - 3.0)])
    fwd = torch.from_numpy(mats.reshape(-1, 9)).cuda()
    rules = [R.auto_rule(m, (w, h), (w, h), threads) for m in mats]
    assert rules == [R.RULE_P, R.RULE_C, R.RULE_P]
    common = handle.common_coverage(fwd, (w, h), (w, h), _native.mask_rule_auto(threads)).cpu().numpy() > 0
    want = np.ones((h, w), bool)
    for m, r in zip(mats, rules):
        want &= R.coverage_np(m, (w, h), (w, h), r)
    assert np.array_equal(common, want)
    box = handle.coverage_bbox(fwd, (w, h), (w, h), _native.mask_rule_auto(threads)).cpu().numpy()
    fixed = {r: handle.coverage_bbox(fwd, (w, h), (w, h), r).cpu().numpy() for r in (R.RULE_P, R.RULE_C)}
    for i, r in enumerate(rules):  # bounding box of the 3x3-closed coverage: what the same kernel returns under that frame's rule
        assert np.array_equal(box[i], fixed[r][i]), (i, box[i])


def test_mask_rule_env_reaches_the_nodes_engine(monkeypatch):
    """VSTAB_MASK_RULE=auto:<threads> (no schema change): Motion Apply with an off-canvas shift of more than H'/8 rows gets
    cv2's Rule C mask for that frame and Rule P for the others, as the reference does on a machine with 8 cv2 threads."""
    from vstab_b200 import motion_apply, pipeline
    from vstab_b200.motion_meta import build_motion_meta_v2

    w, h = 640, 360
    rng = np.random.default_rng(4)
    frames = rng.random((3, h, w, 3), dtype=np.float32)
    mats = [np.array([[1, 0, -0.3], [0, 1, ty], [0, 0, 1]], np.float64) for ty in (2.0, 60.0, -7.0)]
    meta = {"motion_meta": build_motion_meta_v2(source="test", frame_count=3, fps=16.0, input_size=(w, h), output_size=(w, h), matrices=mats)}
    want_rules = [R.auto_rule(np.asarray(m, np.float32), (w, h), (w, h), 8) for m in mats]
    assert want_rules == [R.RULE_P, R.RULE_C, R.RULE_P]
    out = {}
    for env in ("P", "auto:8"):
        monkeypatch.setenv("VSTAB_MASK_RULE", env)
        ctx = pipeline.normalize_video_input(torch.from_numpy(frames))
        out[env] = motion_apply.apply_motion(ctx, meta, (127, 127, 127), framing_mode="crop_and_pad", interpolation="bilinear")
    for i, m in enumerate(mats):
        m32 = np.asarray(m, np.float32)
        assert np.array_equal(out["P"].masks[i, ..., 0], R.mask_np(m32, (w, h), (w, h), R.RULE_P))
        assert np.array_equal(out["auto:8"].masks[i, ..., 0], R.mask_np(m32, (w, h), (w, h), want_rules[i]))
    assert np.array_equal(out["P"].frames, out["auto:8"].frames)
    assert not np.array_equal(out["P"].masks[1], out["auto:8"].masks[1])
