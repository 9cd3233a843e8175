"""libvstab's host helpers for the trajectory solve (csrc/hostsolve.cu: vstab_host_trajectory / _framing / _shift)
against the numpy formulation of hostmath.py, which the other CPU tests pin to the unmodified reference: every output
byte for byte, on random candidate tables in all three models, with and without a working size, camera lock and
fallback pairs -- and the whole driver (stabilizer_core.stabilize_frames with stand-ins for the two GPU stages) with
the helpers on and off.  No GPU: the helpers are host code inside the library."""
import json

import numpy as np
import pytest

MODES = ("translation", "similarity", "perspective")


def _table(rng, pairs, bad=None, detected=False):
    """Random [P,3,12] vstab_fit_result words: plausible inter-frame motion in every model."""
    raw = np.zeros((pairs, 3, 12))
    for p in range(pairs):
        for k in range(3):
            a, s = rng.normal(0, 0.01), 1 + rng.normal(0, 0.01)
            m = np.array([[s * np.cos(a), -s * np.sin(a), rng.normal(0, 5)], [s * np.sin(a), s * np.cos(a), rng.normal(0, 5)], [0, 0, 1]])
            if k == 0:
                m[:2, :2] = np.eye(2)
            if k == 2:
                m[2, :2] = rng.normal(0, 1e-5, 2)
                m[:2, :2] += rng.normal(0, 1e-3, (2, 2))
            raw[p, k, :9] = m.ravel()
            raw[p, k, 9] = rng.random()
            ints = np.array([rng.integers(3000, 8000), 8160, 8160, 1], np.int32)
            if bad is not None and p == bad and k > 0:
                ints[0] = 10  # confidence far below the threshold: this pair falls back
            raw[p, k, 10:12] = ints.view(np.float64)
    det = rng.integers(12, 400, pairs).astype(np.int64) if detected else None
    return raw, det


def _numpy_route(core, hm, cands, mode, size, work):
    chosen, active, stacked = core.replay_mode_ladder(cands, mode, with_residual=True)
    if work is not None:
        stacked = hm.rescale_transforms_to_full(stacked, size, work)
    delta = hm.matrices_to_params(stacked, mode)
    path = np.zeros((len(delta) + 1, delta.shape[1]))
    np.cumsum(delta, axis=0, out=path[1:])
    return stacked, path, chosen, active


@pytest.mark.parametrize("mode", MODES)
def test_trajectory_and_framing_helpers_equal_numpy(mode):
    from vstab_b200 import hostmath as hm, stabilizer_core as core

    rng = np.random.default_rng(MODES.index(mode))
    for trial in range(120):
        pairs = int(rng.integers(1, 160))
        size = (int(rng.integers(64, 4000)), int(rng.integers(64, 2200)))
        work = hm.working_estimation_size(*size)
        raw, det = _table(rng, pairs, detected=trial % 4 == 3)
        cands = core.PairCandidates.from_raw(raw, 12 if det is None else 8, det)
        got = hm.native_trajectory(cands, mode, size, work)
        assert got is not None
        stacked, path, chosen, active = _numpy_route(core, hm, cands, mode, size, work)
        assert active == mode
        assert got[0].tobytes() == stacked.tobytes() and got[1].tobytes() == path.tobytes()
        entries = core.accepted_ladder_entries(cands, mode, with_residual=True)
        assert (entries.modes, entries.confidences, entries.residuals) == (chosen.modes, chosen.confidences, chosen.residuals)
        # a table that only exists as columns packs to the same words
        columns = core.PairCandidates(cands.matrix, cands.residual, cands.n_inliers, cands.n_valid, cands.n_total, cands.ok, 12)
        assert columns.raw_words().tobytes() == raw.tobytes()

        target = np.zeros_like(path) if trial % 5 == 0 else path + rng.random() * (hm.smooth_path(path, rng.random(), 16.0) - path)
        diffs = target - path
        apply, mins, maxs, box = hm.native_framing(diffs, mode, *size)
        want = hm.params_to_matrices(diffs, mode)
        want_mins, want_maxs = hm.compute_bounding_boxes(want, *size)
        assert apply.tobytes() == want.tobytes()
        assert mins.tobytes() == want_mins.tobytes() and maxs.tobytes() == want_maxs.tobytes()
        assert tuple(box[:4]) == tuple(hm.inner_rectangle(want_mins, want_maxs))
        assert tuple(box[4:8]) == (want_mins[:, 0].min(), want_mins[:, 1].min(), want_maxs[:, 0].max(), want_maxs[:, 1].max())
        affine = bool(np.all(want[:, 2, :2] == 0) and np.all(want[:, 2, 2] == 1))
        assert bool(box[8]) == affine
        ox, oy = rng.normal(0, 30, 2)
        shift = np.array([[1, 0, ox], [0, 1, oy], [0, 0, 1]], np.float32)
        assert hm.translate_matrices(apply, ox, oy, affine=affine).tobytes() == hm.left_multiply(shift, want).tobytes()


def test_target_path_helper_equals_numpy():
    """vstab_host_target (box filter in numpy's tap order for windows of up to 11 taps, strength blend, camera lock) against
    hostmath.numpy_target = np.convolve; longer windows are left to numpy (their summation order is the BLAS's)."""
    from vstab_b200 import hostmath as hm

    rng = np.random.default_rng(3)
    through_helper = 0
    for trial in range(1500):
        n, k = int(rng.integers(1, 300)), (2, 4, 8)[trial % 3]
        path = np.cumsum(rng.normal(0, 5, (n, k)), axis=0)
        path[0] = 0
        strength, smooth = float(rng.random()), float(np.clip(rng.random() * 1.2 - 0.1, 0, 1))
        fps, lock = float(rng.choice([1, 8, 12, 16, 24, 30, 60])), trial % 7 == 0
        if lock:
            smooth = max(smooth, 0.85)
        want = hm.numpy_target(path, strength, smooth, fps, lock)
        got = hm.native_target(path, strength, smooth, fps, lock)
        window = hm.smoothing_window(smooth, fps)
        if got is None:
            assert window > 11 and not lock and smooth > 0 and n > 2
            continue
        through_helper += 1
        assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes(), (trial, n, k, smooth, fps)
    assert through_helper > 500
    assert all(hm._TARGET_WINDOW_OK.values())  # this numpy sums short kernels in tap order; if not, the helper steps aside


def test_helpers_follow_numpy_on_non_finite_input():
    """NaN / inf / float32-overflowing / denormal entries in the candidate table: the helpers and the numpy route still leave
    the same bytes in every output (IEEE operations in the same order propagate them the same way)."""
    import warnings

    from vstab_b200 import hostmath as hm, stabilizer_core as core

    rng = np.random.default_rng(5)
    poison = [np.nan, np.inf, -np.inf, 1e39, -1e39, 1e-46]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for trial in range(120):
            mode = MODES[trial % 3]
            pairs = int(rng.integers(2, 40))
            raw, _ = _table(rng, pairs)
            raw[int(rng.integers(0, pairs)), int(rng.integers(0, 3)), int(rng.integers(0, 9))] = poison[trial % 6]
            cands = core.PairCandidates.from_raw(raw, 12)
            work = (960, 540) if trial % 2 else None
            got = hm.native_trajectory(cands, mode, (1920, 1080), work)
            stacked, path, _, _ = _numpy_route(core, hm, cands, mode, (1920, 1080), work)
            assert got is not None and got[0].tobytes() == stacked.tobytes() and got[1].tobytes() == path.tobytes()
            want = hm.numpy_target(path, 0.7, 0.5, 16.0, False)
            target = hm.native_target(path, 0.7, 0.5, 16.0, False)
            assert target is not None and target[0].tobytes() == want[0].tobytes() and target[1].tobytes() == want[1].tobytes()
            apply, mins, maxs, _ = hm.native_framing(want[1], mode, 1920, 1080)
            ref = hm.params_to_matrices(want[1], mode)
            ref_mins, ref_maxs = hm.compute_bounding_boxes(ref, 1920, 1080)
            assert apply.tobytes() == ref.tobytes() and mins.tobytes() == ref_mins.tobytes() and maxs.tobytes() == ref_maxs.tobytes()


def test_fallback_pairs_are_left_to_the_ladder():
    from vstab_b200 import hostmath as hm, stabilizer_core as core

    rng = np.random.default_rng(7)
    for trial in range(30):
        pairs = int(rng.integers(2, 60))
        raw, _ = _table(rng, pairs, bad=int(rng.integers(0, pairs)))
        cands = core.PairCandidates.from_raw(raw, 12)
        assert hm.native_trajectory(cands, "translation", (640, 360), None) is not None  # translation accepts everything
        assert hm.native_trajectory(cands, "similarity", (640, 360), None) is None
        assert hm.native_trajectory(cands, "perspective", (640, 360), None) is None
        few = raw.copy()
        ints = few[:, :, 10:12].copy().view(np.int32)
        ints[3 % pairs, :, 1] = 5  # fewer valid points than min_points
        few[:, :, 10:12] = ints.view(np.float64)
        assert hm.native_trajectory(core.PairCandidates.from_raw(few, 12), "translation", (640, 360), None) is None
    raw, det = _table(rng, 20, detected=True)
    det[4] = 3  # Classic: fewer than 12 corners => identity pair
    assert hm.native_trajectory(core.PairCandidates.from_raw(raw, 8, det), "similarity", (640, 360), None) is None


@pytest.mark.parametrize("framing", ("crop_and_pad", "expand"))
@pytest.mark.parametrize("mode", MODES)
def test_driver_with_and_without_the_helpers(monkeypatch, framing, mode):
    """stabilize_frames end to end (estimation and resampler replaced by cheap stand-ins): identical results and meta
    whether the solve runs in libvstab's host helpers or in numpy (VSTAB_HOST_SOLVE=0).  `crop` framing shares the
    trajectory helper and is covered by tests/test_host_path_cpu.py::test_crop_solver_host_path."""
    from vstab_b200 import stabilizer_core as core

    rng = np.random.default_rng(11)
    n, h, w = 40, 270, 480
    frames = rng.random((n, h, w, 3), dtype=np.float32)
    raw, _ = _table(rng, n - 1)

    class Clip:
        def __init__(self, f):
            self.frames, self.fps, self.device = f, None, None
            self.height, self.width = f.shape[1:3]

        def __len__(self):
            return len(self.frames)

    def estimator(context, work_w, work_h, requested):
        return core.PairCandidates.from_raw(raw.copy(), 12)

    seen = []

    def warp(context, fwd, out_size, interpolation, border, **kw):
        seen.append((np.asarray(fwd, np.float32).copy(), tuple(out_size)))
        count = len(context)
        return lambda: (np.zeros((count, out_size[1], out_size[0], 3), np.float32), np.zeros((count, out_size[1], out_size[0]), np.float32),
                        np.arange(count))

    monkeypatch.setattr(core, "fused_warp", warp)
    def run(flag):
        monkeypatch.setenv("VSTAB_HOST_SOLVE", flag)
        seen.clear()
        res = core.stabilize_frames(Clip(frames), framing, mode, False, 0.8, 0.6, 0.6, (1, 2, 3), 24.0, estimator=estimator,
                                    flavour="flow", output="device")
        return json.dumps(res.meta, sort_keys=True), [(f.tobytes(), s) for f, s in seen]

    on = run("1")
    off = run("0")
    assert on[1] == off[1]  # forward matrices handed to the resampler, byte for byte
    assert on[0] == off[0]  # the whole meta tree
