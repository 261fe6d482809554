"""2+ GPU check of the sharded driver (run under torchrun): every rank optimises its block of
videos on its own GPU, records are gathered with NCCL, rank 0 compares with a single-GPU run.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import dist as a3d_dist, opt_utils, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_videos = 5
seeds = [300 + v for v in range(n_videos)]


def make():
    vids = []
    for s in seeds:
        preds, _ = synth.make_video(s, 3, 14, kinds=[0, 1, 0])
        vids.append((preds, opt_utils.track_planes(preds)))
    return vids


vids = make()
outs, mine, fr, tr = a3d_dist.optimize_videos_sharded(vids, seeds, device=dev)
print(f"rank {rank}: videos {mine}, gathered frame records {tuple(fr.shape)}, track records {tuple(tr.shape)}", flush=True)
if rank == 0:
    ref = make()
    opt_utils.optimize_videos(ref, seeds, device=dev)
    f0, t0 = a3d_dist.pack_records(list(range(n_videos)), [v[1] for v in ref])
    assert np.array_equal(fr.cpu().numpy(), f0.numpy()) and np.array_equal(tr.cpu().numpy(), t0.numpy())
    print("dist_check ok: sharded + NCCL gather == single GPU", flush=True)
dist.barrier()
dist.destroy_process_group()
