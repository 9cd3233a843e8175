"""Generates tests/golden/* by running the UNMODIFIED reference from /root/reference.

Run in the build container only:  python scripts/make_golden.py [--only NAME]
The fixtures are small (selected patches / per-frame sums for the big shapes) and committed; the
GPU box has no /root/reference, so `-m gpu` tests check the CUDA path against these files.
Inputs are regenerated from seeds by tests/cases.py on both sides.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from tests import cases, ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _summaries(frames: np.ndarray, masks: np.ndarray, patches):
    out = {
        "frame_sum": frames.reshape(frames.shape[0], -1).astype(np.float64).sum(axis=1),
        "mask_sum": masks.reshape(masks.shape[0], -1).astype(np.float64).sum(axis=1),
        "shape": np.array(frames.shape),
    }
    for k, (f, y, x, hh, ww) in enumerate(patches):
        out[f"patch{k}"] = frames[f, y : y + hh, x : x + ww].copy()
        out[f"mpatch{k}"] = masks[f, y : y + hh, x : x + ww, 0].copy()
        out[f"patch{k}_at"] = np.array([f, y, x, hh, ww])
    return out


def gen_motion_apply(ref):
    """Motion Apply on small clips (full outputs) and on BASELINE config 1 / 4 shapes (summaries)."""
    for case in cases.MOTION_APPLY_CASES:
        frames = cases.make_frames(case)
        meta = cases.make_motion_meta(case, ref)
        ctx = ref.stabilizer_utils._normalize_video_input([f for f in frames])
        res = ref.motion_apply.apply_motion(
            ctx, meta, case["padding_rgb"], framing_mode=case["framing"], interpolation=case["interp"],
            motion_blur=case["blur"], motion_blur_samples=case["samples"],
        )
        name = case["name"]
        with open(os.path.join(GOLDEN, f"apply_{name}_meta.json"), "w") as fh:
            json.dump(meta, fh)
        if case["store"] == "full":
            np.savez_compressed(os.path.join(GOLDEN, f"apply_{name}.npz"), frames=res.frames, masks=res.masks)
        else:
            np.savez_compressed(os.path.join(GOLDEN, f"apply_{name}.npz"), **_summaries(res.frames, res.masks, case["patches"]))
        print("apply", name, res.frames.shape, res.meta["motion_apply"])


def gen_estimators(ref, only_crop=False, only_small=False):
    """Flow / Classic stabilizers: per-pair matrices, paths and final matrices (+ small outputs)."""
    import cv2

    todo = cases.SMALL_STABILIZER_CASES if only_small else cases.STABILIZER_CASES + cases.CROP_CASES
    for case in todo:
        if only_crop and case not in cases.CROP_CASES:
            continue
        frames = cases.make_frames(case)
        mod = ref.video_stabilizer_flow if case["node"] == "flow" else ref.video_stabilizer_classic
        ctx = ref.stabilizer_utils._normalize_video_input([f for f in frames])
        res = mod._stabilize_frames(
            ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
            case["keep_fov"], case["padding_rgb"], case["fps"],
        )
        name = case["name"]
        meta = res.meta
        with open(os.path.join(GOLDEN, f"stab_{name}_meta.json"), "w") as fh:
            json.dump(meta, fh)
        out_frames = np.asarray(res.frames, dtype=np.float32)
        out_masks = np.asarray(res.masks, dtype=np.float32)
        payload = _summaries(out_frames, out_masks, case["patches"])
        if case["store"] == "full":
            payload["frames"] = out_frames
            payload["masks"] = out_masks
        payload["cv2_threads"] = np.array([cv2.getNumThreads()])
        np.savez_compressed(os.path.join(GOLDEN, f"stab_{name}.npz"), **payload)
        print("stab", name, out_frames.shape, meta.get("transform_mode_applied"))


def gen_dis(ref):
    """Raw cv2.DISOpticalFlow (reference configuration) on working-size gray pairs."""
    be = ref.video_stabilizer_flow._create_flow_backend("DIS")
    for case in cases.DIS_CASES:
        prev, curr = cases.make_gray_pair(case)
        flow = be.calc(prev, curr, None)
        grid = flow[::8, ::8].copy()
        np.savez_compressed(os.path.join(GOLDEN, f"dis_{case['name']}.npz"), grid=grid,
                            flow_mean=flow.reshape(-1, 2).mean(axis=0), flow=flow if case["store"] == "full" else np.zeros(0))
        print("dis", case["name"], flow.shape, flow.reshape(-1, 2).mean(axis=0))


def gen_ab(ref):
    """The reference's A/B gate (scripts/compare_refactor_behavior.py:220-243, :352-366): its synthetic
    73x45 clip through both stabilizers in its three scenarios; the input frames are stored too."""
    import cv2

    width, height, count = 73, 45, 8
    yy, xx = np.mgrid[0:height, 0:width]
    base = np.zeros((height, width, 3), dtype=np.float32)
    base[..., 0] = xx / max(width - 1, 1)
    base[..., 1] = yy / max(height - 1, 1)
    base[..., 2] = (((xx // 5) + (yy // 7)) % 2).astype(np.float32)
    cv2.rectangle(base, (9, 8), (31, 24), (1.0, 0.2, 0.1), -1)
    cv2.circle(base, (width - 20, height - 14), 7, (0.1, 1.0, 0.3), -1)
    frames = []
    center = (width * 0.5, height * 0.5)
    for idx in range(count):
        matrix = cv2.getRotationMatrix2D(center, idx * 0.6, 1.0 + idx * 0.002)
        matrix[0, 2] += idx * 0.7
        matrix[1, 2] += idx * 0.35
        frames.append(cv2.warpAffine(base, matrix, (width, height), flags=cv2.INTER_LINEAR,
                                     borderMode=cv2.BORDER_REFLECT).astype(np.float32))
    payload = {"input": np.stack(frames)}
    metas = {}
    a = cases.AB_ARGS
    for node in ("classic", "flow"):
        mod = ref.video_stabilizer_flow if node == "flow" else ref.video_stabilizer_classic
        for name, framing, mode, keep_fov in cases.AB_SCENARIOS:
            ctx = ref.stabilizer_utils._normalize_video_input(frames)
            res = mod._stabilize_frames(ctx, framing, mode, a["camera_lock"], a["strength"], a["smooth"], keep_fov,
                                        a["padding_rgb"], a["fps"])
            payload[f"{node}.{name}.frames"] = np.asarray(res.frames, dtype=np.float32)
            payload[f"{node}.{name}.masks"] = np.asarray(res.masks, dtype=np.float32)
            metas[f"{node}.{name}"] = res.meta
            print("ab", node, name, payload[f"{node}.{name}.frames"].shape, res.meta.get("transform_mode_applied"))
    np.savez_compressed(os.path.join(GOLDEN, "ab_73x45.npz"), **payload)
    with open(os.path.join(GOLDEN, "ab_73x45_meta.json"), "w") as fh:
        json.dump(metas, fh)


def gen_inverse(ref):
    """Legacy inverse stabilization on the scenario of the reference's scripts/check_inverse_stabilization.py
    (:24-131: 7 source frames of 73x45, an expand and a crop stabilization made with cv2): the helper
    _apply_inverse_stabilization and the deprecated node's execute()."""
    import importlib
    import sys

    import cv2
    import torch

    U = ref.stabilizer_utils
    width, height = 73, 45
    yy, xx = np.mgrid[0:height, 0:width]
    base = np.zeros((height, width, 3), dtype=np.float32)
    base[..., 0] = xx / max(width - 1, 1)
    base[..., 1] = yy / max(height - 1, 1)
    base[..., 2] = (((xx // 6) + (yy // 5)) % 2).astype(np.float32)
    cv2.rectangle(base, (8, 7), (30, 24), (1.0, 0.2, 0.1), -1)
    cv2.circle(base, (width - 19, height - 13), 7, (0.1, 0.9, 0.3), -1)
    center = (width * 0.5, height * 0.5)
    source = []
    for idx in range(7):
        m = cv2.getRotationMatrix2D(center, idx * 0.45, 1.0 + idx * 0.0015)
        m[0, 2] += idx * 0.65
        m[1, 2] += idx * -0.35
        source.append(cv2.warpAffine(base, m, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT).astype(np.float32))

    def stabilize(matrices, out_size, framing):
        frames = [cv2.warpPerspective(f, m, out_size, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                                      borderValue=(0.5, 0.5, 0.5)).astype(np.float32) for f, m in zip(source, matrices)]
        meta = {"frames": len(frames), "framing": {"mode": framing},
                "stabilization_warp": U._build_stabilization_warp_meta(source_size=(width, height), output_size=out_size,
                                                                       framing_mode=framing, applied_matrices=matrices)}
        return np.stack(frames), meta

    shifts = [np.array([[1.0, 0.0, -i * 0.7], [0.0, 1.0, i * 0.4], [0.0, 0.0, 1.0]], dtype=np.float32) for i in range(7)]
    mins, maxs = U._compute_bounding_boxes(shifts, width, height)
    translate, out_size = U._prepare_expand_transform(mins, maxs)
    cases_ = {"expand": stabilize([translate @ m for m in shifts], out_size, "expand")}
    crop = np.array([[1.12, 0.0, -0.06 * width], [0.0, 1.12, -0.06 * height], [0.0, 0.0, 1.0]], dtype=np.float32)
    cases_["crop"] = stabilize([crop.copy() for _ in range(7)], (width, height), "crop")

    payload, metas = {"source": np.stack(source)}, {}
    for name, (frames, meta) in cases_.items():
        res = U._apply_inverse_stabilization(U._normalize_video_input([f for f in frames]), meta, (127, 127, 127))
        payload[f"{name}.input"] = frames
        payload[f"{name}.frames"] = np.stack(res.frames)
        payload[f"{name}.masks"] = np.stack(res.masks)
        metas[f"{name}.in"] = meta
        metas[f"{name}.helper_out"] = res.meta
    sys.modules["comfy_api.latest"].io.NodeOutput = lambda *a: a
    node = importlib.import_module("nodes.video_stabilizer_inverse").VideoStabilizerInverse
    frames, meta = cases_["expand"]
    video, mask, out_meta = node.execute(torch.from_numpy(frames), meta, "#0AC85A")
    payload["node.frames"] = video.numpy()
    payload["node.mask"] = mask.numpy()
    metas["node.out"] = out_meta
    np.savez_compressed(os.path.join(GOLDEN, "inverse_73x45.npz"), **payload)
    with open(os.path.join(GOLDEN, "inverse_73x45_meta.json"), "w") as fh:
        json.dump(metas, fh)
    print("inverse", {k: v.shape for k, v in payload.items()})


def gen_full(ref, only_name=None):
    """The BASELINE.json configurations as quoted (121 / 241 / 121 frames, and 4K): summaries + the whole meta."""
    for case in cases.FULL_CASES:
        if only_name and case["name"] != only_name:
            continue
        name = case["name"]
        frames = cases.make_frames(case)
        ctx = ref.stabilizer_utils._normalize_video_input([f for f in frames])
        if case["kind"] == "apply":
            meta_in = cases.make_motion_meta(case, ref)
            with open(os.path.join(GOLDEN, f"full_{name}_motion_meta.json"), "w") as fh:
                json.dump(meta_in, fh)
            res = ref.motion_apply.apply_motion(ctx, meta_in, case["padding_rgb"], framing_mode=case["framing"], interpolation=case["interp"],
                                                motion_blur=case["blur"], motion_blur_samples=case["samples"])
        else:
            mod = ref.video_stabilizer_flow if case["node"] == "flow" else ref.video_stabilizer_classic
            res = mod._stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                        case["keep_fov"], case["padding_rgb"], case["fps"])
        with open(os.path.join(GOLDEN, f"full_{name}_meta.json"), "w") as fh:
            json.dump(res.meta, fh)
        out_frames, out_masks = np.asarray(res.frames, dtype=np.float32), np.asarray(res.masks, dtype=np.float32)
        np.savez_compressed(os.path.join(GOLDEN, f"full_{name}.npz"), **_summaries(out_frames, out_masks, case["patches"]))
        print("full", name, out_frames.shape, flush=True)
        del frames, ctx, res, out_frames, out_masks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--name", default=None, help="with --only full: one case")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_import.load_reference()
    gens = {"apply": gen_motion_apply, "stab": gen_estimators, "dis": gen_dis, "crop": lambda r: gen_estimators(r, True),
            "small": lambda r: gen_estimators(r, only_small=True), "ab": gen_ab, "inverse": gen_inverse,
            "full": lambda r: gen_full(r, args.name)}
    for key, fn in gens.items():
        if args.only == key or (args.only is None and key not in ("crop", "small", "ab", "inverse", "full")):
            fn(ref)


if __name__ == "__main__":
    main()
