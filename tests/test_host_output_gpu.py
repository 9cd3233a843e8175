"""Host results of the fused resampler (pipeline.fused_warp, output="host"): the padding mask of a single-sample call
crosses the link as one byte per pixel (vstab_mask_pack_u8) and is widened on the host.  Frames, masks and padded-pixel
counts must equal the device-resident results bit for bit -- with the byte route on and off, with one chunk and with
several, on frame sizes whose pixel count is not a multiple of the kernel's 16-value vectors -- and a soft (motion-blur)
mask must keep travelling as float32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _context(frames):
    from vstab_b200 import pipeline

    return pipeline.normalize_video_input(torch.from_numpy(frames))


def _matrices(rng, n, w, h, samples=1):
    out = np.zeros((n, samples, 9), np.float32)
    for i in range(n):
        a, s = rng.normal(0, 0.03), 1 + rng.normal(0, 0.03)
        tx, ty = rng.normal(0, 0.08 * w), rng.normal(0, 0.08 * h)
        for k in range(samples):
            t = k / max(samples - 1, 1)
            out[i, k] = np.array([[s * np.cos(a), -s * np.sin(a), tx * (1 + 0.1 * t)], [s * np.sin(a), s * np.cos(a), ty * (1 - 0.1 * t)],
                                  [0, 0, 1]], np.float32).ravel()
    return out


@pytest.mark.parametrize("size", [(96, 64), (131, 77), (333, 201)])
@pytest.mark.parametrize("chunks", [1, 3])
def test_host_result_equals_device_result(monkeypatch, size, chunks):
    from vstab_b200 import pipeline

    w, h = size
    n = 7
    rng = np.random.default_rng(w * 1000 + h + chunks)
    ctx = _context(rng.random((n, h, w, 3), dtype=np.float32))
    fwd = _matrices(rng, n, w, h)
    out_size = (w + 9, h + 5)
    dev = pipeline.fused_warp(ctx, fwd, out_size, "bilinear", (0.1, 0.5, 0.9), want_mask=True, want_pad_count=True, output="device")
    want = [t.cpu().numpy() if torch.is_tensor(t) else np.asarray(t) for t in dev]
    assert set(np.unique(want[1])) <= {0.0, 1.0} and 0.0 < want[1].mean() < 1.0  # a real mask: padding and content
    if chunks > 1:
        monkeypatch.setattr(pipeline, "CHUNK_BYTES", (n // chunks) * out_size[0] * out_size[1] * 16)
    for flag in ("1", "0"):
        monkeypatch.setenv("VSTAB_MASK_BYTES", flag)
        frames, masks, pads = pipeline.fused_warp(ctx, fwd, out_size, "bilinear", (0.1, 0.5, 0.9), want_mask=True, want_pad_count=True,
                                                  output="host")
        assert masks.dtype == torch.float32 and not masks.is_cuda
        assert np.array_equal(frames.numpy(), want[0])
        assert np.array_equal(masks.numpy(), want[1])
        assert np.array_equal(np.asarray(pads), want[2])
        expect = n * out_size[0] * out_size[1] * (13 if flag == "1" else 16)
        assert pipeline.LAST_D2H_BYTES == expect


def test_soft_mask_of_a_motion_blur_call_stays_float32():
    from vstab_b200 import pipeline

    w, h, n, s = 160, 96, 4, 5
    rng = np.random.default_rng(3)
    ctx = _context(rng.random((n, h, w, 3), dtype=np.float32))
    fwd = _matrices(rng, n, w, h, samples=s)
    dev = pipeline.fused_warp(ctx, fwd, (w, h), "bilinear", (0.5, 0.5, 0.5), want_mask=True, output="device")
    frames, masks, _ = pipeline.fused_warp(ctx, fwd, (w, h), "bilinear", (0.5, 0.5, 0.5), want_mask=True, output="host")
    soft = dev[1].cpu().numpy()
    assert len(np.unique(soft)) > 2  # fractional coverage
    assert np.array_equal(masks.numpy(), soft) and np.array_equal(frames.numpy(), dev[0].cpu().numpy())
    assert pipeline.LAST_D2H_BYTES == n * w * h * 16


def test_mask_pack_flags_values_other_than_zero_and_one(handle):
    dev = torch.device("cuda", 0)
    for count in (5, 16, 4099):
        mask = (torch.arange(count, device=dev) % 3 == 0).to(torch.float32)
        out = torch.full((count,), 7, dtype=torch.uint8, device=dev)
        odd = torch.zeros((1,), dtype=torch.int32, device=dev)
        handle.mask_pack_u8(mask, out, odd)
        assert torch.equal(out.to(torch.float32), mask) and int(odd.item()) == 0
        mask[count - 1] = 0.25
        handle.mask_pack_u8(mask, out, odd)
        assert int(odd.item()) == 1
