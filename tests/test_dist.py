"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding of
videos and the gather of per-track records."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from articulation3d_b200 import dist as a3d_dist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_matches_reference_split():
    # tools/opt_arti.py:116-123: chunk = ceil(n / array_cnt); ids[rank*chunk:(rank+1)*chunk]
    for n in (0, 1, 7, 8, 9, 256):
        for world in (1, 2, 3, 8):
            chunk = int(np.ceil(n / world)) if n else 0
            got = [list(a3d_dist.shard_range(n, r, world)) for r in range(world)]
            want = [list(range(n))[r * chunk:(r + 1) * chunk] for r in range(world)]
            assert got == want
            assert sorted(sum(got, [])) == list(range(n))


def _fake_optimize(videos, seeds, cfg=None, device=None):
    """Stands in for the GPU optimiser: fills plane['fit'] deterministically from the seed."""
    outs = []
    for (preds, planes), seed in zip(videos, seeds):
        rng = np.random.RandomState(seed)
        for cat in ("trans", "rot"):
            for plane in planes[cat]:
                frames = list(plane["ids"].keys())
                plane["has_rot"] = bool(rng.rand() < 0.7)
                plane["fit"] = {"rsq": rng.rand(3)}
                if plane["has_rot"]:
                    plane["std_axis"] = (torch.tensor(rng.randint(0, 640, 4)) if cat == "rot"
                                         else torch.tensor(rng.rand(2), dtype=torch.float32))
                    plane["fit"].update(frames=frames, center_frame=frames[0],
                                        angle_id=rng.randint(0, 45, len(frames)).astype(np.int32),
                                        inter=rng.randint(0, 1000, len(frames)).astype(np.int32),
                                        union=rng.randint(1000, 2000, len(frames)).astype(np.int32))
        outs.append(preds)
    return outs


def _make_videos(n):
    vids = []
    for v in range(n):
        planes = {"trans": [{"ids": {f: 0 for f in range(10 + v)}}],
                  "rot": [{"ids": {f: 1 for f in range(12)}}, {"ids": {f: 2 for f in range(3, 15)}}]}
        vids.append(([], planes))
    return vids


def _fake_optimize_lazy(videos, seeds, cfg=None, device=None):
    """As ``optimize_videos`` does for entries ``(preds, None)``: the planes appear in place (here: the ones
    _make_videos would have given the video whose seed this is)."""
    for i, ((preds, planes), seed) in enumerate(zip(videos, seeds)):
        if planes is None:
            videos[i] = (preds, _make_videos(seed - 100 + 1)[seed - 100][1])
    return _fake_optimize(videos, seeds, cfg=cfg, device=device)


def _worker(rank, world, port, n_videos, q, lazy=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vids = _make_videos(n_videos)
        if lazy:                                 # untracked videos; remote ones are never touched
            vids = [(p, None) for p, _ in vids]
        outs, mine, fr, tr = a3d_dist.optimize_videos_sharded(vids, list(range(100, 100 + n_videos)),
                                                              optimize_fn=_fake_optimize_lazy if lazy else _fake_optimize)
        q.put((rank, mine, fr.numpy(), tr.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_videos,lazy", [(5, False), (1, False), (5, True)])
def test_sharded_gather_world2_gloo(n_videos, lazy):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_videos, q, lazy)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process ground truth
    vids = _make_videos(n_videos)
    _fake_optimize(vids, list(range(100, 100 + n_videos)))
    fr, tr = a3d_dist.pack_records(list(range(n_videos)), [v[1] for v in vids])
    assert sorted(got[0][1] + got[1][1]) == list(range(n_videos))
    for _, _, f, t in got:                       # every rank holds the complete tables, in video order
        assert np.array_equal(f, fr.numpy()) and np.array_equal(t, tr.numpy())
    assert tr.shape == (3 * n_videos, a3d_dist.TRACK_COLS)
