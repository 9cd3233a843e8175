"""CUDA gray + INTER_AREA working image (vstab_gray_working) vs the numpy oracle: bit-exact."""
import json

import numpy as np
import pytest
import torch

from oracle import gray_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(832, 480), (1920, 1080), (1280, 720), (3840, 2160), (1000, 562), (121, 73)])
def test_gray_working_bit_exact(handle, size):
    w, h = size
    rng = np.random.default_rng(w)
    n = 2
    rgb = rng.random((n, h, w, 3), dtype=np.float32)
    rgb[0, :4, :4] = 1.0  # saturating values
    rgb[0, 4:8, :4] = 0.0
    got = handle.gray_working(torch.from_numpy(rgb).cuda()).cpu().numpy()
    ws = gray_np.working_size(w, h)
    for i in range(n):
        want = gray_np.gray_for_estimation(rgb[i], ws)
        assert got[i].shape == want.shape
        assert int((got[i] != want).sum()) == 0


def test_gray_matches_cv2_when_available(handle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    rgb = rng.random((1, 720, 1280, 3), dtype=np.float32)
    got = handle.gray_working(torch.from_numpy(rgb).cuda()).cpu().numpy()[0]
    g = np.clip(cv2.cvtColor(rgb[0], cv2.COLOR_RGB2GRAY) * 255.0, 0, 255).astype(np.uint8)
    want = cv2.resize(g, (960, 540), interpolation=cv2.INTER_AREA)
    assert int((got != want).sum()) == 0


def test_area_tables_are_recycled_when_the_cache_is_full(handle):
    """A long-lived process sees many frame sizes; every non-integer INTER_AREA ratio needs a coverage table per
    (source, destination) length and the handle keeps 32 of them.  More sizes than that must keep working
    (slots recycled round robin) and sizes seen before must give the same bytes when their table is rebuilt."""
    rng = np.random.default_rng(9)
    widths = list(range(962, 1004))  # 42 widths -> 42 x tables (+ the y tables) > 32 slots
    first = None
    for k, w in enumerate(widths + widths[:2]):
        h = 300 + (k % 3)
        rgb = rng.random((1, h, w, 3), dtype=np.float32) if k < len(widths) else first[k - len(widths)][0]
        got = handle.gray_working(torch.from_numpy(rgb).cuda()).cpu().numpy()[0]
        want = gray_np.gray_for_estimation(rgb[0], gray_np.working_size(w, rgb.shape[1]))
        assert got.shape == want.shape and int((got != want).sum()) == 0, (k, w)
        if k < 2:
            first = (first or []) + [(rgb, got)]
        if k >= len(widths):
            assert np.array_equal(got, first[k - len(widths)][1])


@pytest.mark.parametrize("size", [(320, 180), (1280, 720), (1920, 1080)])
def test_fused_range_adapter_matches_the_reference_rule(handle, size):
    """vstab_gray_working_adapt: a float frame whose max() exceeds 1.5 is a 0..255 frame -- the reference
    divides it by 255 before everything else (stabilizer_utils.py:96-147).  Here the luma kernel raises the flag in its
    one read of the source, the working image of such frames is recomputed from value / 255 and the frame is divided in
    place, all on the device.  Mixed clip: 0..1 frames, 0..255 frames, a frame with a NaN (numpy's max() is NaN: not
    rescaled), a frame with one value above 1.5."""
    from oracle import gray_np
    from vstab_b200 import _native

    w, h = size
    rng = np.random.default_rng(w)
    clip = rng.random((6, h, w, 3), dtype=np.float32)
    clip[1] *= 255.0
    clip[3] *= 255.0
    clip[4, h // 2, w // 3, 1] = np.nan
    clip[4, 0, 0, 0] = 200.0            # would be "big", but the NaN wins
    clip[5, h - 1, w - 1, 2] = 1.75     # one element above 1.5 rescales the whole frame
    want_frames = clip.copy()
    want_scaled = []
    for i in range(len(clip)):
        big = bool(clip[i].max() > 1.5)  # numpy: NaN > 1.5 is False
        want_scaled.append(int(big))
        if big:
            want_frames[i] /= np.float32(255.0)
    work = _native.working_size(w, h)
    dev = torch.from_numpy(clip).cuda()
    gray, flags = handle.gray_working_adapt(dev, work)
    torch.cuda.synchronize()
    assert [int(f == 1) for f in flags.cpu().numpy()] == want_scaled == [0, 1, 0, 1, 0, 1]
    got_frames = dev.cpu().numpy()
    assert np.array_equal(got_frames, want_frames, equal_nan=True)
    size_arg = None if work == (w, h) else work
    for i in range(len(clip)):
        if i == 4:
            continue  # NaN luma: cv2's cast is undefined there
        assert np.array_equal(gray[i].cpu().numpy(), gray_np.gray_for_estimation(want_frames[i], size_arg)), i
    # the adapter alone (Motion Apply has no estimation pass)
    dev2 = torch.from_numpy(clip).cuda()
    flags2 = handle.range_normalize(dev2)
    assert [int(f == 1) for f in flags2.cpu().numpy()] == want_scaled
    assert np.array_equal(dev2.cpu().numpy(), want_frames, equal_nan=True)


def test_node_input_with_0_255_floats_goes_through_the_fused_adapter():
    """normalize_video_input(defer_range=True) + the Flow driver: a 0..255 float clip gives the result of the same clip / 255."""
    from tests import cases
    from vstab_b200 import flow, pipeline

    case = next(c for c in cases.STABILIZER_CASES if c["name"] == "flow_sim_pad_480p")
    frames = cases.make_frames(case)[:5]
    args = (case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"])
    ctx = pipeline.normalize_video_input(torch.from_numpy(frames), defer_range=True)
    assert ctx.range_pending
    want = flow.stabilize_frames(ctx, *args)
    assert not ctx.range_pending
    big = (frames * np.float32(255.0)).astype(np.float32)
    exact = (big / np.float32(255.0)) == frames  # compare only where the round trip is exact... the clip as the reference would see it:
    ctx2 = pipeline.normalize_video_input(torch.from_numpy(big), defer_range=True)
    got = flow.stabilize_frames(ctx2, *args)
    ref_ctx = pipeline.normalize_video_input(torch.from_numpy(big / np.float32(255.0)))
    ref = flow.stabilize_frames(ref_ctx, *args)
    assert np.array_equal(got.frames, ref.frames) and np.array_equal(got.masks, ref.masks)
    assert json.dumps(got.meta, sort_keys=True) == json.dumps(ref.meta, sort_keys=True)
    assert exact.mean() > 0.5 and want.frames.shape == got.frames.shape
