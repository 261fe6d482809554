"""What slows the mask upload of the pipeline down (21.7 ms alone, 23-34 ms inside optimize_videos)?
Runs a3d_upload_masks for 6 clips back to back from a helper thread while the main thread (a) sleeps,
(b) spins in pure Python, (c) runs numpy copies, (d) runs small torch float64 ops."""
import os, sys, threading, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import opt_utils, synth, workloads

wl = workloads.WORKLOADS["c3"]
cfg = wl.cfg()
preds, _ = synth.make_video(2020, wl.tracks, wl.frames, cfg, kinds=[synth.KIND_ROT] * wl.tracks, device="cuda:0")
for p in preds:
    p.pred_masks = p.pred_masks.cpu().pin_memory()
planes = opt_utils.track_planes(preds, cfg)
torch.cuda.synchronize()


def uploads(out):
    for _ in range(6):
        t0 = time.perf_counter()
        s = opt_utils._Session([(preds, [planes['trans'], planes['rot']])], cfg, "cuda:0")
        s.pool
        torch.cuda.synchronize()
        out.append(1e3 * (time.perf_counter() - t0))


def busy(kind, stop):
    a = np.random.rand(1 << 20)
    b = np.empty_like(a)
    t = torch.rand(932 * 180, 9, dtype=torch.float64)
    x = 0
    while not stop[0]:
        if kind == "sleep":
            time.sleep(0.001)
        elif kind == "python":
            for i in range(10000):
                x += i
        elif kind == "numpy":
            np.copyto(b, a)
        elif kind == "torch":
            (t * 2.0 + 1.0).sum()


for env in ({"A3D_UPLOAD_DEPTH": "0"}, {"A3D_UPLOAD_DEPTH": "8"}, {"A3D_UPLOAD_DEPTH": "32"}, {"A3D_UPLOAD_DEPTH": "64"},
            {"A3D_UPLOAD": "python", "A3D_UPLOAD_DEPTH": "0"}, {"A3D_UPLOAD": "python", "A3D_UPLOAD_DEPTH": "8"}):
    os.environ.pop("A3D_UPLOAD", None)
    os.environ.update(env)
    out = []
    uploads(out)
    print(env, " ".join("%.1f" % v for v in out))
os.environ.pop("A3D_UPLOAD", None)
os.environ["A3D_UPLOAD_DEPTH"] = "8"

# breakdown of one upload (nothing else running)
from articulation3d_b200 import engine, _lib
LOG = []


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            LOG.append((label, 1e3 * (t - T0[0]), 1e3 * (time.perf_counter() - T0[0])))
    setattr(obj, name, w)


T0 = [0.0]
wrap(opt_utils._Session, "__init__", "Session.__init__")
wrap(opt_utils._Session, "_upload_worker", "_upload_worker")
wrap(opt_utils._Session, "_upload", "_upload")
wrap(engine.PoolBuilder, "__init__", "PoolBuilder.__init__")
wrap(engine.PoolBuilder, "finish", "PoolBuilder.finish")
wrap(engine, "mask_meta", "mask_meta")
lib = _lib.load()
wrap(lib, "a3d_upload_masks", "a3d_upload_masks")
_empty = torch.empty


def empty(*a, **k):
    t = time.perf_counter()
    r = _empty(*a, **k)
    if r.numel() > (1 << 20):
        LOG.append(("torch.empty %d MB on %s" % (r.numel() * r.element_size() >> 20, r.device), 1e3 * (t - T0[0]), 1e3 * (time.perf_counter() - T0[0])))
    return r


torch.empty = empty
for rep in range(2):
    del LOG[:]
    torch.cuda.synchronize()
    T0[0] = time.perf_counter()
    s = opt_utils._Session([(preds, [planes['trans'], planes['rot']])], cfg, "cuda:0")
    s.pool
    torch.cuda.synchronize()
    LOG.append(("total", 0.0, 1e3 * (time.perf_counter() - T0[0])))
    for lb, a, b in sorted(LOG, key=lambda r: r[1]):
        print("%8.2f %8.2f %7.2f ms  %s" % (a, b, b - a, lb))
    del s
