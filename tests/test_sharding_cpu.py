"""World-size-2 gloo test of the frame-range sharding host logic (no GPU, no pixels):
the gathered candidate table and the ladder replayed on it must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vstab_loader

vstab_loader.load()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_candidates(total_frames, seed=0):
    """A table with a sticky downgrade in the middle: pair 7 has a failing similarity fit."""
    from vstab_b200.stabilizer_core import PairCandidates

    rng = np.random.default_rng(seed)
    p = total_frames - 1
    m = np.tile(np.eye(3), (p, 3, 1, 1)).astype(np.float64)
    m[:, :, :2, 2] = rng.normal(0, 3, (p, 3, 2))
    n_valid = np.full((p, 3), 8160)
    n_inl = np.full((p, 3), 8000)
    n_inl[7, 1] = 300  # confidence 0.037 < 0.10 => similarity rejected => sticky translation afterwards
    return PairCandidates(m, rng.random((p, 3)), n_inl, n_valid, np.full((p, 3), 8160), np.ones((p, 3), int))


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vstab_loader

    vstab_loader.load()
    from vstab_b200.sharding import FrameShard
    from vstab_b200.stabilizer_core import PairCandidates, replay_mode_ladder

    shard = FrameShard(rank, world, total, None, torch.device("cpu"))
    full = _fake_candidates(total)
    a, b = shard.pair_range
    local = PairCandidates.from_array(full.to_array()[a:b], 12)
    gathered = shard.gather_candidates(local)
    chosen, active, _ = replay_mode_ladder(gathered, "similarity", with_residual=True)
    pads = shard.gather_pad_counts(np.arange(*shard.frame_range) * 10)
    out[rank] = (gathered.to_array(), [c[1] for c in chosen], active, pads, shard.frame_range, shard.load_range)
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    from vstab_b200.stabilizer_core import replay_mode_ladder

    total, world = 23, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    full = _fake_candidates(total)
    chosen, active, _ = replay_mode_ladder(full, "similarity", with_residual=True)
    modes = [c[1] for c in chosen]
    assert modes[6] == "similarity" and modes[7] == "translation" and modes[-1] == "translation" and active == "translation"
    for r in range(world):
        table, rmodes, ractive, pads, frange, lrange = out[r]
        assert np.array_equal(table, full.to_array())
        assert rmodes == modes and ractive == active
        assert np.array_equal(pads, np.arange(total) * 10)
    assert out[0][4] == (0, 12) and out[1][4] == (12, 23) and out[1][5] == (11, 23)


def test_split_covers_everything():
    from vstab_b200.sharding import FrameShard, split_range

    for total in (1, 2, 7, 121, 2000):
        for world in (1, 2, 3, 8):
            ranges = [split_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            pairs = []
            for r in range(world):
                a, b = FrameShard(r, world, total).pair_range
                pairs += list(range(a, b))
            assert pairs == list(range(max(total - 1, 0)))


def test_own_range_meta_merges_to_the_single_process_meta(monkeypatch):
    """META_OWN_RANGE: every rank materialises only the per-frame meta of its own frames (clip-wide indices);
    merge_sharded_meta() of the ranks' metas is the single-process meta, key for key and value for value.
    Host logic only: the estimator, the collectives and the resampler are stand-ins."""
    from vstab_b200 import stabilizer_core as core
    from vstab_b200.sharding import META_EVERY_RANK, FrameShard, merge_sharded_meta

    total, world, W, H = 23, 3, 640, 360
    full = _fake_candidates(total)
    pads_all = (np.arange(total) * 7) % 50

    class Ctx:
        def __init__(self, n):
            self.n, self.width, self.height, self.fps = n, W, H, None

        def __len__(self):
            return self.n

        def sliced(self, k):
            return self

    class Shard(FrameShard):
        def gather_candidates(self, local):
            return core.PairCandidates.from_array(full.to_array(), 12)

        def gather_pad_counts(self, local_counts):
            return pads_all.astype(np.int64)

    def fake_warp(context, fwd, *a, **k):
        return lambda: (np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(len(fwd), np.int64))

    monkeypatch.setattr(core, "fused_warp", fake_warp)
    monkeypatch.setattr(core, "StabilizationResult", lambda frames, masks, meta: meta)

    def run(shard):
        n = total if shard is None else shard.load_range[1] - shard.load_range[0]

        def est(ctx, w, h, mode):
            return full if shard is None else full  # the shard's gather stand-in supplies the table anyway

        meta = core.stabilize_frames(Ctx(n), "crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0,
                                     estimator=est, flavour="flow", output="device", shard=shard)
        return meta

    # single process: pad counts come from the warp stand-in (zeros); give it the same counts through a 1-rank shard
    single = run(Shard(0, 1, total, None, torch.device("cpu"), meta_rank=META_EVERY_RANK))
    assert "shard" not in single and len(single["stabilization_warp"]["per_frame"]) == total
    metas = [run(Shard(r, world, total, None, torch.device("cpu"))) for r in range(world)]
    for r, m in enumerate(metas):
        lo, hi = FrameShard(r, world, total).frame_range
        assert m["shard"]["frame_range"] == [lo, hi] and m["frames"] == total
        assert [e["index"] for e in m["stabilization_warp"]["per_frame"]] == list(range(lo, hi))
        assert [e["index"] for e in m["motion_meta"]["per_frame"]] == list(range(lo, hi))
        assert [e["index"] for e in m["estimated_motion"]["per_transition"]] == list(range(max(lo, 1) - 1, hi - 1))
        assert len(m["estimated_motion"]["path"]) == hi - lo and m["motion_meta"]["frame_count"] == hi - lo
    assert merge_sharded_meta(metas) == single
    # a rank that owns no frames (more ranks than frames) has nothing to contribute and must not break the merge
    empty = {k: v for k, v in metas[0].items() if k != "motion_meta"}
    empty["shard"] = dict(metas[0]["shard"], meta_frame_range=[5, 5])
    assert merge_sharded_meta(metas + [empty]) == single
    # a designated meta rank still builds the whole tree, the others only the scalar part
    whole = run(Shard(1, world, total, None, torch.device("cpu"), meta_rank=1))
    bare = run(Shard(0, world, total, None, torch.device("cpu"), meta_rank=1))
    assert whole == single and bare["estimated_motion"]["per_transition"] == [] and "motion_meta" not in bare
    assert bare["padding_fraction_max"] == single["padding_fraction_max"]


def test_shards_tell_dis_where_their_pairs_sit_in_the_clip(monkeypatch):
    """On small frames cv2's DIS treats the first pair of a CLIP differently (vstab_dis_flow_at): a frame-range shard
    must pass the clip-wide index of its first pair, a single-process run passes 0."""
    from vstab_b200 import _native, flow, pipeline
    from vstab_b200.sharding import FrameShard
    from vstab_b200.stabilizer_core import DeviceCandidates

    seen = []

    class FakeHandle:
        def dis_flow(self, gray, *, want_flow, grid_step, first_pair=0):
            seen.append(first_pair)
            return None, "grid"

        def fit_grid(self, grid, step, mask, out=None):
            return "raw"

    monkeypatch.setattr(_native, "get_handle", lambda device: FakeHandle())
    monkeypatch.setattr(pipeline, "gray_working", lambda context, size, a, b: "gray")

    class Ctx:
        device = torch.device("cpu")

        def __len__(self):
            return 9

    captured = {}
    monkeypatch.setattr(flow, "_core", lambda *a, estimator, **k: captured.setdefault("est", estimator))
    args = ("crop_and_pad", "similarity", False, 0.7, 0.5, 0.6, (127, 127, 127), 16.0)
    for shard, want in ((None, 0), (FrameShard(0, 3, 24), 0), (FrameShard(1, 3, 24), 7), (FrameShard(2, 3, 24), 15)):
        captured.clear()
        flow.stabilize_frames(Ctx(), *args, shard=shard)
        got = captured["est"](Ctx(), 73, 45, "similarity")
        assert isinstance(got, DeviceCandidates) and got.raw == "raw"  # the table stays where the fit kernels wrote it
        assert seen[-1] == want, (shard, seen[-1])
        if shard is not None:
            assert shard.pair_range[0] == want


def _real_worker(rank, world, port, case_name, out):
    """One rank of a frame-range sharded Flow run on the CPU: real FrameShard + gloo collectives + the product's host
    path, with the oracle in place of the GPU stages (tests/test_host_path_cpu.py)."""
    import functools

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vstab_loader

    vstab_loader.load()
    from tests import cases
    from tests.test_host_path_cpu import _Clip, _oracle_estimator, _oracle_warp
    from vstab_b200 import stabilizer_core as core
    from vstab_b200.sharding import FrameShard

    case = dict(next(c for c in cases.SMALL_STABILIZER_CASES if c["name"] == case_name.split("@")[0]))
    if "@" in case_name:  # "name@keep_fov": the same clip with another keep_fov
        case["keep_fov"] = float(case_name.split("@")[1])
    frames = cases.make_frames(case)
    shard = FrameShard(rank, world, len(frames), None, torch.device("cpu"))
    lo, hi = shard.load_range
    core.fused_warp = _oracle_warp
    if case["framing"] == "crop":  # numpy coverage in place of the two coverage kernels
        from tests.test_host_path_cpu import _CoverageOracle
        from vstab_b200 import crop

        crop._native.get_handle = lambda device: _CoverageOracle()
    est = shard.wrap_estimator(functools.partial(_oracle_estimator, clip_pair_offset=shard.pair_range[0]))  # as flow.stabilize_frames does
    clip = _Clip(frames[lo:hi])
    clip.device = torch.device("cpu")
    res = core.stabilize_frames(clip, case["framing"], case["mode"], case["camera_lock"], case["strength"],
                                case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"], estimator=est, flavour="flow",
                                output="device", shard=shard)
    out[rank] = (np.asarray(res.frames), np.asarray(res.masks), res.meta)
    dist.destroy_process_group()


def test_two_rank_flow_run_equals_the_reference(monkeypatch):
    """90x50 is the size where cv2's DIS object changes state after the first pair of a clip: rank 1 starts in the
    middle of the clip and must estimate its pairs in the later state.  Concatenated frames / merged meta of the two
    ranks == the single-process run, and both == the unmodified reference's output for the clip."""
    import json

    from tests import cases, parity
    from tests.conftest import GOLDEN_DIR
    from tests.test_host_path_cpu import _run
    from vstab_b200.sharding import merge_sharded_meta

    name = "flow_sim_pad_90x50"
    case = next(c for c in cases.SMALL_STABILIZER_CASES if c["name"] == name)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_real_worker, args=(2, _free_port(), name, out), nprocs=2, join=True)
    frames = np.concatenate([out[0][0], out[1][0]])
    masks = np.concatenate([out[0][1], out[1][1]])
    meta = merge_sharded_meta([out[0][2], out[1][2]])
    single = _run(monkeypatch, cases.make_frames(case), case["framing"], case["mode"], case["camera_lock"], case["strength"],
                  case["smooth"], case["keep_fov"], case["padding_rgb"], case["fps"])
    assert np.array_equal(frames, np.asarray(single.frames)) and np.array_equal(masks, np.asarray(single.masks))
    assert json.loads(json.dumps(meta)) == json.loads(json.dumps(single.meta))
    gold = np.load(os.path.join(GOLDEN_DIR, f"stab_{name}.npz"))
    with open(os.path.join(GOLDEN_DIR, f"stab_{name}_meta.json")) as fh:
        gmeta = json.load(fh)
    parity.compare_nested(gmeta, json.loads(json.dumps(meta)), "meta", atol=2e-5, rtol=2e-5)
    err = np.abs(frames - gold["frames"])
    assert float((err > 2e-5).mean()) <= 1e-3 and float(err.max()) <= 0.04


@pytest.mark.parametrize("keep_fov", [1.0, 0.6])
def test_two_rank_crop_framing_equals_the_single_process_run(monkeypatch, keep_fov):
    """`crop` framing under frame-range shards: the keep_fov ~= 1 bypass must hand back each rank's OWN
    frames (not its halo frame) with clip-wide frame counts and indices, and the keep_fov search / no-padding refinement
    must solve for the whole clip on every rank.  Concatenated frames / masks and the merged meta == single process."""
    import json

    import torch as _torch

    from tests import cases
    from tests.test_host_path_cpu import _Clip, _CoverageOracle, _oracle_estimator, _oracle_warp
    from vstab_b200 import crop, stabilizer_core as core
    from vstab_b200.sharding import merge_sharded_meta

    name = "flow_trans_crop_121x73"
    case = next(c for c in cases.SMALL_STABILIZER_CASES if c["name"] == name)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_real_worker, args=(2, _free_port(), f"{name}@{keep_fov}", out), nprocs=2, join=True)
    frames = np.concatenate([out[0][0], out[1][0]])
    masks = np.concatenate([out[0][1], out[1][1]])
    meta = merge_sharded_meta([out[0][2], out[1][2]])
    monkeypatch.setattr(core, "fused_warp", _oracle_warp)
    monkeypatch.setattr(crop._native, "get_handle", lambda device: _CoverageOracle())
    clip = _Clip(cases.make_frames(case))
    clip.device = _torch.device("cpu")
    single = core.stabilize_frames(clip, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                   keep_fov, case["padding_rgb"], case["fps"], estimator=_oracle_estimator, flavour="flow", output="device")
    assert frames.shape == np.asarray(single.frames).shape
    assert np.array_equal(frames, np.asarray(single.frames)) and np.array_equal(masks, np.asarray(single.masks))
    assert json.loads(json.dumps(meta)) == json.loads(json.dumps(single.meta))
    assert meta["frames"] == case["n"] and [e["index"] for e in meta["stabilization_warp"]["per_frame"]] == list(range(case["n"]))
    if keep_fov >= 0.9999:
        assert np.array_equal(frames, cases.make_frames(case)) and meta["transform_mode_applied"] == "identity"
