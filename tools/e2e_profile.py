"""cProfile of the public API on a workload-shaped clip (host-side overheads)."""
import cProfile
import os
import pstats
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import opt_utils, workloads  # noqa: E402

wl = workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
preds, cfg = workloads.make_clip(wl, 2020)
for p in preds:
    p.pred_masks = p.pred_masks.pin_memory()


def run():
    random.seed(2020)
    planes = opt_utils.track_planes(preds, cfg)
    st = opt_utils.Stats()
    opt_utils.optimize_planes(preds, planes, "3dc", cfg=cfg, device="cuda:0", stats=st)
    torch.cuda.synchronize()
    return st


run()
for _ in range(4):
    t0 = time.perf_counter()
    st = run()
    print("wall ms %.2f" % (1e3 * (time.perf_counter() - t0)), st.schedule, st.passes, st.units_visited, st.units_computed)
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

# phase split: session construction (H2D + pack) vs. the passes
videos = [(preds, [opt_utils.track_planes(preds, cfg)["rot"]])]
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = opt_utils._Session(videos, cfg, torch.device("cuda:0"))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("session (H2D %.1f MB + pack) ms %.2f -> %.1f GB/s" % (s.h2d_bytes / 1e6, 1e3 * (t1 - t0), s.h2d_bytes / (t1 - t0) / 1e9))
big = torch.cat([p.pred_masks for p in preds]).pin_memory()
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = big.to("cuda:0", non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("one pinned H2D of %.1f MB: %.2f ms -> %.1f GB/s" % (big.numel() * 4 / 1e6, 1e3 * (t1 - t0), big.numel() * 4 / (t1 - t0) / 1e9))
