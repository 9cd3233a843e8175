"""Seeded test cases shared by scripts/make_golden.py (reference side) and the parity tests.

Inputs are pure functions of the case dict, so the golden generator (build container, real
reference) and the GPU tests (no reference) see identical bytes.
"""
from __future__ import annotations

import numpy as np

PAD = (127, 127, 127)


def make_frames(case) -> np.ndarray:
    """[N,H,W,3] float32 in 0..1."""
    n, w, h = case["n"], case["w"], case["h"]
    kind = case.get("frames", "noise")
    if kind == "noise":  # white noise: worst case for resampling parity (gradient ~1/px)
        rng = np.random.default_rng(case["seed"])
        return rng.random((n, h, w, 3), dtype=np.float32)
    if kind == "texture_fast":
        # the same clip as "texture" for the BASELINE-length cases: rendered by the fused resampler when a GPU is there
        # (the -m gpu tests) and by cv2 where the goldens are made -- bilinear is bit-exact across oracle, cv2 and kernel
        # (tests/test_oracle_resample.py, tests/test_warp_gpu.py), so both sides see the same bytes
        import synth
        import torch

        mats = synth.shake_matrices(n, case["seed"], w, h, perspective=case.get("perspective", False), amount=case.get("amount", 1.0))
        base = synth.base_texture(case["seed"], w, h)
        if torch.cuda.is_available():
            import vstab_loader

            vstab_loader.load()
            from vstab_b200 import _native

            dev = torch.device("cuda", torch.cuda.current_device())
            return synth.render_clip_cuda(_native.get_handle(dev), base.to(dev), mats, w, h).cpu().numpy()
        import cv2

        fwd = synth.render_matrices(mats)
        return np.stack([cv2.warpPerspective(base.numpy(), fwd[i], (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                                             borderValue=(0.5, 0.5, 0.5)) for i in range(n)])
    if kind == "texture":  # smooth trackable texture rendered through shake matrices
        import synth

        base = synth.base_texture(case["seed"], w, h).numpy()
        mats = synth.shake_matrices(n, case["seed"], w, h, perspective=case.get("perspective", False),
                                    amount=case.get("amount", 1.0))
        return synth.render_clip_numpy(base, mats, w, h)
    raise ValueError(kind)


def make_motion_meta(case, ref=None):
    """motion_meta for Motion Apply cases.  'shake' needs the reference (generator); the
    resulting JSON is committed beside the golden outputs, tests read it from there."""
    n, w, h = case["n"], case["w"], case["h"]
    if case["meta"] == "shake":
        block = ref.shake_noise.generate_shake_motion_meta(
            recipe=ref.shake_noise.STYLES[case.get("style", "handheld")], frame_count=n, width=w, height=h,
            fps=16.0, amount=case.get("amount", 1.0), speed=1.0, seed=case["seed"], node="shake_generator",
            style=case.get("style", "handheld"),
        )
        return {"motion_meta": block}
    raise ValueError(case["meta"])


MOTION_APPLY_CASES = [
    dict(name="small_bilinear_pad", n=6, w=121, h=73, seed=11, meta="shake", amount=2.0, framing="crop_and_pad",
         interp="bilinear", blur=0.0, samples=9, padding_rgb=PAD, store="full", patches=[]),
    dict(name="small_bicubic_expand_blur", n=6, w=121, h=73, seed=12, meta="shake", amount=2.0, framing="expand",
         interp="bicubic", blur=0.5, samples=5, padding_rgb=(10, 200, 90), store="full", patches=[]),
    dict(name="small_bilinear_crop", n=6, w=120, h=72, seed=13, meta="shake", amount=1.0, framing="crop",
         interp="bilinear", blur=0.0, samples=9, padding_rgb=PAD, store="full", patches=[]),
    dict(name="small_bilinear_pad_blur", n=5, w=64, h=48, seed=14, meta="shake", amount=3.0, framing="crop_and_pad",
         interp="bilinear", blur=1.0, samples=33, padding_rgb=PAD, store="full", patches=[]),
    # BASELINE config 1: Shake Generator (handheld, seed 0) -> Motion Apply bilinear crop_and_pad, 81x832x480
    dict(name="cfg1_832x480", n=81, w=832, h=480, seed=0, meta="shake", amount=1.0, framing="crop_and_pad",
         interp="bilinear", blur=0.0, samples=9, padding_rgb=PAD, store="summary",
         patches=[(0, 0, 0, 48, 64), (40, 200, 400, 48, 64), (80, 432, 768, 48, 64)]),
    # BASELINE config 4 shape, shortened to 3 frames: bicubic expand, blur 0.5, Ultra (33 samples), 1080p
    dict(name="cfg4_1080p_x3", n=3, w=1920, h=1080, seed=4, meta="shake", amount=1.0, framing="expand",
         interp="bicubic", blur=0.5, samples=33, padding_rgb=PAD, store="summary",
         patches=[(0, 0, 0, 48, 64), (1, 500, 900, 48, 64), (2, 1030, 1850, 48, 64)]),
]

STABILIZER_CASES = [
    dict(name="flow_sim_pad_480p", node="flow", n=9, w=832, h=480, seed=21, frames="texture", framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         store="summary", patches=[(4, 100, 300, 48, 64)]),
    dict(name="flow_sim_pad_1080p", node="flow", n=7, w=1920, h=1080, seed=22, frames="texture", framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         store="summary", patches=[(3, 500, 900, 48, 64)]),
    dict(name="flow_trans_expand_480p", node="flow", n=7, w=832, h=480, seed=23, frames="texture", framing="expand",
         mode="translation", camera_lock=False, strength=1.0, smooth=1.0, keep_fov=0.6, padding_rgb=(0, 0, 0), fps=24.0,
         store="summary", patches=[(3, 100, 300, 48, 64)]),
    dict(name="flow_persp_lock_1080p", node="flow", n=6, w=1920, h=1080, seed=24, frames="texture", perspective=True,
         framing="crop_and_pad", mode="perspective", camera_lock=True, strength=0.7, smooth=0.5, keep_fov=0.6,
         padding_rgb=PAD, fps=16.0, store="summary", patches=[(3, 500, 900, 48, 64)]),
    dict(name="classic_sim_pad_720p", node="classic", n=7, w=1280, h=720, seed=25, frames="texture", framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         store="summary", patches=[(3, 300, 600, 48, 64)]),
    dict(name="classic_trans_pad_720p", node="classic", n=7, w=1280, h=720, seed=26, frames="texture", framing="crop_and_pad",
         mode="translation", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         store="summary", patches=[(3, 300, 600, 48, 64)]),
]

# Frame sizes of the reference's own script-level checks (scripts/compare_refactor_behavior.py:220: 73x45,
# scripts/check_crop_aspect_ratio.py:165: 121x73) and one size where cv2's DIS object changes state after the
# first pair (90x50).  Below ~91 px cv2 selects the pyramid levels itself and computes down to full resolution.
SMALL_STABILIZER_CASES = [
    dict(name="flow_sim_pad_73x45", node="flow", n=8, w=73, h=45, seed=71, frames="texture", framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         store="full", patches=[(4, 8, 8, 24, 32)]),
    dict(name="flow_sim_pad_90x50", node="flow", n=6, w=90, h=50, seed=72, frames="texture", amount=2.0, framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=1.0, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=24.0,
         store="full", patches=[(3, 8, 8, 24, 32)]),
    dict(name="flow_trans_crop_121x73", node="flow", n=6, w=121, h=73, seed=73, frames="texture", framing="crop",
         mode="translation", camera_lock=False, strength=1.0, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=24.0,
         store="full", patches=[(3, 8, 8, 24, 32)]),
    dict(name="classic_sim_pad_73x45", node="classic", n=8, w=73, h=45, seed=74, frames="texture", framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         store="full", patches=[(4, 8, 8, 24, 32)]),
]


CROP_CASES = [
    dict(name="flow_sim_crop06_480p", node="flow", n=8, w=832, h=480, seed=51, frames="texture", framing="crop",
         mode="similarity", camera_lock=False, strength=1.0, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=24.0,
         store="summary", patches=[(4, 100, 300, 48, 64)]),
    dict(name="flow_trans_crop00_480p", node="flow", n=8, w=832, h=480, seed=52, frames="texture", amount=2.0, framing="crop",
         mode="translation", camera_lock=False, strength=1.0, smooth=1.0, keep_fov=0.0, padding_rgb=PAD, fps=24.0,
         store="summary", patches=[(4, 100, 300, 48, 64)]),
    dict(name="flow_sim_crop095_480p", node="flow", n=8, w=832, h=480, seed=53, frames="texture", amount=3.0, framing="crop",
         mode="similarity", camera_lock=True, strength=1.0, smooth=0.5, keep_fov=0.95, padding_rgb=PAD, fps=24.0,
         store="summary", patches=[(4, 100, 300, 48, 64)]),
    dict(name="flow_sim_crop1_480p", node="flow", n=5, w=832, h=480, seed=54, frames="texture", framing="crop",
         mode="similarity", camera_lock=False, strength=1.0, smooth=0.5, keep_fov=1.0, padding_rgb=PAD, fps=24.0,
         store="summary", patches=[(2, 100, 300, 48, 64)]),
    dict(name="classic_sim_crop06_720p", node="classic", n=7, w=1280, h=720, seed=55, frames="texture", framing="crop",
         mode="similarity", camera_lock=False, strength=1.0, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=24.0,
         store="summary", patches=[(3, 300, 600, 48, 64)]),
]


# The BASELINE.json configurations AS QUOTED (full clip lengths): summaries (per-frame sums, three
# patches, the whole meta tree) of the unmodified reference's output, scripts/make_golden.py --only full.
FULL_CASES = [
    dict(name="cfg2_flow_1080p_121", kind="stab", node="flow", n=121, w=1920, h=1080, seed=0, frames="texture_fast", framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         patches=[(0, 0, 0, 48, 64), (60, 500, 900, 48, 64), (120, 1030, 1850, 48, 64)]),
    dict(name="cfg3_classic_sim_720p_241", kind="stab", node="classic", n=241, w=1280, h=720, seed=3, frames="texture_fast", framing="crop_and_pad",
         mode="similarity", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         patches=[(0, 0, 0, 48, 64), (120, 300, 600, 48, 64), (240, 670, 1210, 48, 64)]),
    dict(name="cfg3_classic_trans_720p_241", kind="stab", node="classic", n=241, w=1280, h=720, seed=3, frames="texture_fast", framing="crop_and_pad",
         mode="translation", camera_lock=False, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         patches=[(0, 0, 0, 48, 64), (120, 300, 600, 48, 64), (240, 670, 1210, 48, 64)]),
    dict(name="cfg4_apply_1080p_121", kind="apply", n=121, w=1920, h=1080, seed=4, frames="texture_fast", meta="shake", amount=1.0, framing="expand",
         interp="bicubic", blur=0.5, samples=33, padding_rgb=PAD,
         patches=[(0, 0, 0, 48, 64), (60, 500, 900, 48, 64), (120, 1030, 1850, 48, 64)]),
    dict(name="cfg5_flow_persp_lock_4k_6", kind="stab", node="flow", n=6, w=3840, h=2160, seed=5, frames="texture_fast", perspective=True,
         framing="crop_and_pad", mode="perspective", camera_lock=True, strength=0.7, smooth=0.5, keep_fov=0.6, padding_rgb=PAD, fps=16.0,
         patches=[(0, 0, 0, 48, 64), (3, 1000, 1800, 48, 64), (5, 2100, 3770, 48, 64)]),
]


def make_gray_pair(case):
    """Working-size uint8 pair for raw DIS parity: smooth texture + known similarity jitter."""
    import synth
    from oracle.gray_np import gray_u8

    w, h = case["w"], case["h"]
    base = synth.base_texture(case["seed"], w, h).numpy()
    mats = synth.shake_matrices(2, case["seed"], w, h, amount=case.get("amount", 1.0))
    clip = synth.render_clip_numpy(base, mats, w, h)
    return gray_u8(clip[0]), gray_u8(clip[1])


DIS_CASES = [
    dict(name="pair_960x540", w=960, h=540, seed=31, store="grid"),
    dict(name="pair_832x480", w=832, h=480, seed=32, store="grid"),
    dict(name="pair_320x180", w=320, h=180, seed=33, store="full"),
]


# scripts/compare_refactor_behavior.py:352-366 compare_stabilizers: the reference's own A/B gate between two
# revisions of itself ("base" = the unmodified reference, "head" = here the CUDA path), 8 frames of 73x45.
AB_SCENARIOS = [
    ("crop_and_pad_similarity", "crop_and_pad", "similarity", 0.6),
    ("expand_translation", "expand", "translation", 0.6),
    ("crop_keep_fov_bypass", "crop", "translation", 1.0),
]
AB_ARGS = dict(camera_lock=False, strength=0.7, smooth=0.5, padding_rgb=PAD, fps=24.0)
