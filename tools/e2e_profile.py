"""cProfile of the public API on workload-shaped clips (host-side overheads).
    python tools/e2e_profile.py [workload] [n_videos]"""
import cProfile
import os
import pstats
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import opt_utils, synth, workloads  # noqa: E402

wl = workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
n_videos = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = wl.cfg()
clips = []
for v in range(n_videos):
    preds, _ = synth.make_video(2020 + v, wl.tracks, wl.frames, cfg, kinds=[synth.KIND_ROT] * wl.tracks, device="cuda:0")
    for p in preds:
        p.pred_masks = p.pred_masks.cpu().pin_memory()
    clips.append(preds)
saved = [[(p.pred_tran_axis.clone(), p.pred_rot_axis.clone(), p.pred_planes.clone()) for p in c] for c in clips]


def run():
    for c, sv in zip(clips, saved):
        for p, (ta, ra, pl) in zip(c, sv):
            p.pred_tran_axis, p.pred_rot_axis, p.pred_planes = ta.clone(), ra.clone(), pl.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = opt_utils.Stats()
    if n_videos == 1:
        random.seed(2020)
        planes = opt_utils.track_planes(clips[0], cfg)
        opt_utils.optimize_planes(clips[0], planes, "3dc", cfg=cfg, device="cuda:0", stats=st)
    else:
        vids = [(c, opt_utils.track_planes(c, cfg)) for c in clips]
        opt_utils.optimize_videos(vids, [2020 + v for v in range(n_videos)], cfg=cfg, device="cuda:0", stats=st)
    torch.cuda.synchronize()
    return st, time.perf_counter() - t0


run()
for _ in range(4):
    st, dt = run()
    print("wall ms %.2f" % (1e3 * dt), st.schedule, st.passes, st.units_visited, st.units_computed,
          "-> %.3g units/s" % (st.units_visited / dt))
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(60)
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
