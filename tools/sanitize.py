"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel once, all three
scoring kernels, both projection kernels, odd resolution, RLE ingest, override_depth.
    compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import OptConfig, adapter, engine, opt_utils, rle, synth  # noqa: E402

cfg = OptConfig.scaled(200, 150)
preds, _ = synth.make_video(5, 2, 12, cfg, kinds=[0, 1])
for kernel, proj, sched, out in (("ldg", "exact", "cta", "rows"), ("tma", "filter", "cta", "full"),
                                 ("mma", "filter", "persistent", "rows")):
    os.environ["A3D_SCORE_KERNEL"] = kernel
    os.environ["A3D_PROJ_OUT"] = out                 # both output modes of the projection
    os.environ["A3D_PROJECT_KERNEL"] = proj          # both projection kernels, both CTA schedulers
    os.environ["A3D_PROJECT_SCHED"] = sched
    engine._tile_cache.clear()                       # the largest tile depends on the scheduler
    p = synth.clone_preds(preds)
    random.seed(1)
    planes = opt_utils.track_planes(p, cfg)
    out = opt_utils.optimize_planes(p, planes, '3dc', cfg=cfg, device="cuda:0")
    for cat in planes:
        for pl in planes[cat]:
            if pl.get('has_rot'):
                _ = pl['reg_masks'][next(iter(pl['reg_masks']))]
    print(kernel, [pl.get('has_rot') for cat in planes for pl in planes[cat]])
# the pipelined batch API (upload and preparation threads, shared workspace), both schedules
for sched in ("table", "chain"):
    os.environ["A3D_SCHEDULE"] = sched
    vids = []
    for v in range(3):
        pv, _ = synth.make_video(20 + v, 2, 11, cfg, kinds=[0, 1])
        vids.append((pv, opt_utils.track_planes(pv, cfg)))
    opt_utils.optimize_videos(vids, [1, 2, 3], cfg=cfg, device="cuda:0")
    print("optimize_videos", sched, [pl.get('has_rot') for _, planes in vids for cat in planes for pl in planes[cat]])
os.environ.pop("A3D_SCHEDULE")
# pinned host masks, videos tracked inside the pipeline (a3d_upload_masks with asynchronous copies, frames with
# untracked boxes as views, a3d_fetch_host_block for the descriptors)
vids = []
for v in range(3):
    pv, _ = synth.make_video(30 + v, 3, 12, cfg, kinds=[0, 1, 0])
    for q in pv:
        q.pred_masks = q.pred_masks.pin_memory()
    vids.append((pv, None))
opt_utils.optimize_videos(vids, [4, 5, 6], cfg=cfg, device="cuda:0")
print("optimize_videos lazy/pinned", [pl.get('has_rot') for _, planes in vids for cat in planes for pl in planes[cat]])
H, W = 150, 200
rles = [rle.encode(m.numpy() > 0.5) for m in preds[3].pred_masks]
pool = engine.rle_to_pool(rles, H, W, "cuda:0")
inst = [{"instances": [{"segmentation": r} for r in rles], "pred_plane": preds[3].pred_planes.clone()}]
adapter.override_depth_batch(inst, depths=torch.rand(1, H, W, device="cuda:0") + 1,
                             rays=torch.FloatTensor(adapter.get_K_inv_dot_xy_1(H, W)).cuda())
torch.cuda.synchronize()
print("sanitize run complete", pool.popc.tolist())
