"""Frame-range sharded runs on real kernels vs the single-GPU run, bit for bit.

Two ranks, one process each (torch.multiprocessing spawn): with two or more B200s visible every rank has its own
device and the exchange is NCCL over NVLink, as in production; on a one-GPU box both ranks share cuda:0 and the
exchange is gloo (NCCL refuses two ranks on one device) -- the kernels, the halo frame, the candidate table, the sticky
ladder replay, the framing solve and the per-rank meta are the same code either way.  Concatenated frames / masks and
`merge_sharded_meta` of the ranks' metas must EQUAL the single-process result: np.array_equal and == on the JSON tree.
"""
import json
import os
import socket
import tempfile

import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _case(name):
    base, _, kf = name.partition("@")
    case = dict(next(c for c in cases.STABILIZER_CASES + cases.CROP_CASES + cases.SMALL_STABILIZER_CASES if c["name"] == base))
    if kf:
        case["keep_fov"] = float(kf)
    return case


def _drive(case, frames, shard=None):
    import vstab_loader

    vstab_loader.load()
    from vstab_b200 import classic, flow, pipeline

    driver = flow if case["node"] == "flow" else classic
    ctx = pipeline.normalize_video_input(torch.from_numpy(np.ascontiguousarray(frames)))
    return driver.stabilize_frames(ctx, case["framing"], case["mode"], case["camera_lock"], case["strength"], case["smooth"],
                                   case["keep_fov"], case["padding_rgb"], case["fps"], shard=shard)


def _worker(rank, world, port, name, clip_path, out_dir, use_nccl):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    device = torch.device("cuda", rank if use_nccl else 0)
    torch.cuda.set_device(device)
    dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
    import vstab_loader

    vstab_loader.load()
    from vstab_b200.sharding import FrameShard

    case = _case(name)
    frames = np.load(clip_path)
    shard = FrameShard(rank, world, len(frames), None, device if use_nccl else torch.device("cpu"))
    lo, hi = shard.load_range
    res = _drive(case, frames[lo:hi], shard)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), frames=np.asarray(res.frames), masks=np.asarray(res.masks))
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as fh:
        json.dump(res.meta, fh)
    dist.barrier()
    dist.destroy_process_group()


SHARDED = [
    "flow_sim_pad_480p",            # Flow, similarity, crop_and_pad (the bench configuration, smaller)
    "flow_trans_expand_480p",       # expand: the output size is a clip-wide quantity
    "flow_persp_lock_1080p",        # perspective ladder + camera_lock at a working size (BASELINE config 5 shape)
    "flow_sim_crop06_480p",         # crop: keep_fov search + padding-free refinement over the whole clip on every rank
    "flow_sim_crop06_480p@1.0",     # crop with keep_fov ~= 1: the bypass hands back each rank's own frames
    "flow_sim_pad_90x50",           # small frames: cv2's DIS object changes state after the first pair of the CLIP
    "classic_sim_pad_720p",         # Classic (GFTT + LK)
]


@pytest.mark.parametrize("name", SHARDED)
def test_two_rank_run_equals_single_gpu(name):
    import torch.multiprocessing as mp

    import vstab_loader

    vstab_loader.load()
    from vstab_b200.sharding import merge_sharded_meta

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    case = _case(name)
    frames = cases.make_frames(case)
    use_nccl = torch.cuda.device_count() >= WORLD
    with tempfile.TemporaryDirectory() as tmp:
        clip_path = os.path.join(tmp, "clip.npy")
        np.save(clip_path, frames)
        mp.spawn(_worker, args=(WORLD, _free_port(), name, clip_path, tmp, use_nccl), nprocs=WORLD, join=True)
        parts = [np.load(os.path.join(tmp, f"rank{r}.npz")) for r in range(WORLD)]
        got_frames = np.concatenate([p["frames"] for p in parts])
        got_masks = np.concatenate([p["masks"] for p in parts])
        metas = []
        for r in range(WORLD):
            with open(os.path.join(tmp, f"rank{r}.json")) as fh:
                metas.append(json.load(fh))
    single = _drive(case, frames)
    want_meta = json.loads(json.dumps(single.meta))
    assert got_frames.shape == np.asarray(single.frames).shape and got_masks.shape == np.asarray(single.masks).shape
    assert np.array_equal(got_frames, np.asarray(single.frames)), float(np.abs(got_frames - np.asarray(single.frames)).max())
    assert np.array_equal(got_masks, np.asarray(single.masks))
    for r, m in enumerate(metas):
        assert m["frames"] == len(frames) and m.get("shard", {}).get("rank", r) == r
    assert merge_sharded_meta(metas) == want_meta
