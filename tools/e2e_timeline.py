"""Per-thread timeline of the public batch API on workload-shaped clips (who waits for whom).
    python tools/e2e_timeline.py [workload] [n_videos]
Wraps the stages of ``optimize_videos`` with wall-clock stamps (thread, label, start, end) and the device passes
with CUDA events; prints the merged timeline of the last of three runs."""
import os
import random
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import engine, opt_utils, synth, workloads  # noqa: E402

wl = workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
n_videos = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = wl.cfg()
clips = []
for v in range(n_videos):
    preds, _ = synth.make_video(2020 + v, wl.tracks, wl.frames, cfg, kinds=[synth.KIND_ROT] * wl.tracks, device="cuda:0")
    for p in preds:
        p.pred_masks = p.pred_masks.cpu().pin_memory()
    clips.append(preds)
saved = [[(p.pred_tran_axis.clone(), p.pred_rot_axis.clone(), p.pred_planes.clone()) for p in c] for c in clips]

LOG, DEV = [], []
T0 = [0.0]


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            LOG.append((threading.current_thread().name, label, t - T0[0], time.perf_counter() - T0[0]))
    if isinstance(obj.__dict__.get(name), staticmethod):
        w = staticmethod(w)
    setattr(obj, name, w)


S = opt_utils._Session
for n in ("_upload_worker", "_prepare_groups", "launch_tables", "finish_tables", "run", "run_rows"):
    wrap(S, n)
for n in ("_run_lists", "_drive", "_answer_chain", "_write_back", "track_planes"):
    wrap(opt_utils, n)
wrap(engine.DeviceBatch, "__init__", "DeviceBatch")
wrap(S, "__init__", "Session")
if os.environ.get("TL_FINE"):
    for n in ("host", "device_block", "host_done"):
        wrap(engine.Staging, n, "Staging." + n)
    for n in ("build_batch_rows", "plan_tiles_native", "build_batch"):
        wrap(engine, n)
    wrap(opt_utils, "_xforms_rows")
    wrap(opt_utils._VideoRows, "geometry")
if os.environ.get("TL_SWITCH"):
    sys.setswitchinterval(float(os.environ["TL_SWITCH"]))
_run_pass = engine.run_pass


def run_pass(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t = time.perf_counter()
    r = _run_pass(*a, **k)
    e1.record()
    DEV.append((e0, e1, a[2].n_jobs))
    LOG.append((threading.current_thread().name, "run_pass(host)", t - T0[0], time.perf_counter() - T0[0]))
    return r


engine.run_pass = run_pass


def run():
    if os.environ.get("TL_GC"):
        import gc
        gc.collect()
        gc.disable()
    for c, sv in zip(clips, saved):
        for p, (ta, ra, pl) in zip(c, sv):
            p.pred_tran_axis, p.pred_rot_axis, p.pred_planes = ta.clone(), ra.clone(), pl.clone()
    torch.cuda.synchronize()
    del LOG[:], DEV[:]
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    T0[0] = t0 = time.perf_counter()
    if os.environ.get("TL_PROBE"):
        x = clips[0][0].pred_boxes.tensor
        y = x.detach().cpu()
        LOG.append(("MainThread", "probe .cpu() of %s" % x.device, 0.0, time.perf_counter() - t0))
    st = opt_utils.Stats()
    vids = [(c, opt_utils.track_planes(c, cfg)) for c in clips]
    opt_utils.optimize_videos(vids, [2020 + v for v in range(n_videos)], cfg=cfg, device="cuda:0", stats=st)
    torch.cuda.synchronize()
    return st, time.perf_counter() - t0, base


for i in range(3):
    st, dt, base = run()
    print("wall ms %.2f passes %d -> %.3g units/s" % (1e3 * dt, st.passes, st.units_visited / dt))
rows = [(a, b, th, lb) for th, lb, a, b in LOG]
for e0, e1, nj in DEV:
    rows.append((base.elapsed_time(e0) / 1e3, base.elapsed_time(e1) / 1e3, "device", f"pass {nj} jobs"))
if os.environ.get("TL_QUIET"):
    rows = []
for a, b, th, lb in sorted(rows):
    print("%8.2f %8.2f  %7.2f ms  %-16s %s" % (1e3 * a, 1e3 * b, 1e3 * (b - a), th[:16], lb))
