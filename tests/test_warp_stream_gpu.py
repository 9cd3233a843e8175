"""The streaming resampler (persistent CTAs, TMA box loads, ticket scheduler, interior / edge / general
tiles) against the one-tile kernel run with global gathers (stage_mode=1): two independent code paths
over the same arithmetic must agree bit for bit on pixels, mask and padded counts -- for random
clips that push tiles through every branch (footprints leaving the frame, tiles fully outside,
boxes too large to stage, perspective, canvases larger and smaller than the source)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(2024)
    out = []
    for k in range(36):
        w = int(rng.choice([76, 96, 160, 320, 640])) + 4 * int(rng.integers(0, 5))
        h = int(rng.integers(24, 260))
        n = int(rng.integers(1, 8))
        kind = ["similarity", "perspective", "wild", "identity"][k % 4]
        ow = max(4, w + 4 * int(rng.integers(-8, 12))) if k % 3 else w
        if k % 5 == 4:
            ow += int(rng.choice([1, 2, 3]))  # output rows that cannot take 16-byte stores
        oh = max(1, h + int(rng.integers(-20, 40))) if k % 3 else h
        out.append(dict(w=w, h=h, n=n, kind=kind, ow=ow, oh=oh, seed=1000 + k, interp="bicubic" if k % 7 == 3 else "bilinear"))
    return out


def _matrices(case):
    rng = np.random.default_rng(case["seed"])
    mats = []
    for _ in range(case["n"]):
        if case["kind"] == "identity":
            m = np.eye(3)
        else:
            wild = case["kind"] == "wild"
            th = rng.normal(0, 0.08 if wild else 0.01)
            s = 1 + rng.normal(0, 0.2 if wild else 0.01)
            tx, ty = rng.normal(0, 120.0 if wild else 12.0, 2)
            m = np.array([[s * np.cos(th), -s * np.sin(th), tx], [s * np.sin(th), s * np.cos(th), ty], [0, 0, 1.0]])
            if case["kind"] == "perspective" or (wild and rng.random() < 0.5):
                m[2, :2] = rng.normal(0, 2e-4 if wild else 2e-5, 2)
        mats.append(m)
    return np.asarray(mats, dtype=np.float32)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c['w']}x{c['h']}to{c['ow']}x{c['oh']}_{c['kind']}_{c['interp']}_n{c['n']}")
def test_stream_kernel_equals_global_gather_kernel(handle, case):
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(case["seed"] + 7)
    src = torch.from_numpy(rng.random((case["n"], case["h"], case["w"], 3), dtype=np.float32)).to(dev)
    fwd = torch.from_numpy(_matrices(case).reshape(case["n"], 1, 9)).to(dev)
    border = tuple(float(v) for v in rng.random(3))
    res = []
    for stage in (0, 1):
        dst, mask, pad = handle.warp_fused(src, fwd, (case["ow"], case["oh"]), case["interp"], border, want_pad_count=True, stage_mode=stage)
        torch.cuda.synchronize()
        res.append((dst.cpu().numpy(), mask.cpu().numpy(), pad.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2], res[1][2])


def test_stream_kernel_is_deterministic_on_a_long_clip(handle):
    """Ticket order changes from run to run; results must not."""
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    n, h, w = 40, 272, 480
    src = torch.from_numpy(rng.random((n, h, w, 3), dtype=np.float32)).to(dev)
    mats = np.tile(np.eye(3, dtype=np.float32), (n, 1, 1))
    mats[:, 0, 2] = rng.normal(0, 9, n)
    mats[:, 1, 2] = rng.normal(0, 9, n)
    mats[:, 0, 1] = rng.normal(0, 0.01, n)
    fwd = torch.from_numpy(mats.reshape(n, 1, 9)).to(dev)
    first = None
    for _ in range(6):
        dst, mask, pad = handle.warp_fused(src, fwd, (w, h), "bilinear", (0.1, 0.2, 0.3), want_pad_count=True)
        torch.cuda.synchronize()
        got = (dst.cpu().numpy(), mask.cpu().numpy(), pad.cpu().numpy())
        if first is None:
            first = got
        else:
            assert all(np.array_equal(a, b) for a, b in zip(first, got))


@pytest.mark.parametrize("kind", ["zoom_out_0.9", "zoom_out_0.6", "zoom_out_0.35", "projective_far_side", "projective_strong", "rotate_zoom_out"])
def test_minifying_maps_take_the_mixed_tiles_and_keep_every_bit(handle, kind):
    """Maps that MINIFY (zoom-out, the far side of a projective correction: the late frames of a camera-locked perspective
    clip, BASELINE config 5) have tile footprints larger than the staged TMA box.  Round 2 serves them as "mixed" tiles
    (box where it reaches, L1 gathers elsewhere) instead of the general path: same bits as the global-gather kernel and
    as the numpy oracle, mask and padded counts included."""
    from oracle import resample_np as R

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(99)
    n, h, w = 3, 288, 512
    src_np = rng.random((n, h, w, 3), dtype=np.float32)
    cx, cy = w / 2, h / 2
    mats = []
    for i in range(n):
        if kind.startswith("zoom_out"):
            s = float(kind.split("_")[-1]) + 0.01 * i
            m = np.array([[s, 0, cx - s * cx + 3.3 * i], [0, s, cy - s * cy - 2.1 * i], [0, 0, 1.0]])
        elif kind == "rotate_zoom_out":
            s, th = 0.7, 0.2 + 0.05 * i
            c, sn = s * np.cos(th), s * np.sin(th)
            m = np.array([[c, -sn, cx - c * cx + sn * cy], [sn, c, cy - sn * cx - c * cy], [0, 0, 1.0]])
        else:
            g = (4e-4 if kind == "projective_far_side" else 2.5e-3) * (1 + 0.3 * i)
            m = np.array([[1, 0, 2.0 * i], [0, 1, -1.0 * i], [g, g / 2, 1.0]])
        mats.append(m)
    mats = np.asarray(mats, dtype=np.float32)
    src = torch.from_numpy(src_np).to(dev)
    fwd = torch.from_numpy(mats.reshape(n, 1, 9)).to(dev)
    border = (0.25, 0.5, 0.75)
    res = []
    for stage in (0, 1):
        dst, mask, pad = handle.warp_fused(src, fwd, (w, h), "bilinear", border, want_pad_count=True, stage_mode=stage)
        torch.cuda.synchronize()
        res.append((dst.cpu().numpy(), mask.cpu().numpy(), pad.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    for i in range(n):
        want = R.warp_np(src_np[i], mats[i], (w, h), "bilinear", border)
        assert np.array_equal(res[0][0][i], want), (kind, i, float(np.abs(res[0][0][i] - want).max()))
        assert np.array_equal(res[0][1][i], R.mask_np(mats[i], (w, h), (w, h)))
