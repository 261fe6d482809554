"""A/B of the two projection kernels (A3D_PROJECT_KERNEL = exact | filter) on one workload: a3d_project
alone (k_unproject + k_project), L2 flushed between steps, GPU run-ahead so that launch gaps of the host
do not count, results compared bit for bit.    python tools/project_ab.py c2 [mode]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import _lib, engine, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else _lib.MODE_SEQ
steps = int(os.environ.get("AB_ITERS", "20"))
dev = torch.device("cuda:0")
inp = workloads.build_pass(workloads.WORKLOADS[name], 2020, dev, mode=mode)
lib = _lib.load()
cfg, pool, db = inp.cfg, inp.pool, inp.dbatch
H, W, pitch = cfg.height, cfg.width, _lib.pitch_words(cfg.width)
nc = db.n_cand_total
proj_bits = torch.empty((nc, H, pitch), dtype=torch.int32, device=dev)
proj_popc = torch.empty((nc,), dtype=torch.int32, device=dev)
proj_bbox = torch.empty((nc, 4), dtype=torch.int32, device=dev)
pcd_ws = torch.empty((max(_lib.PCD_PLANES * db.pcd_total, 32),), dtype=torch.float32, device=dev)
pcd_count = torch.empty((db.n_jobs + 1,), dtype=torch.int32, device=dev)
hom_ws = torch.empty((nc, _lib.HOM_FLOATS), dtype=torch.float32, device=dev)
cam = engine.camera_struct(cfg)
tile, tmap = db.tile_plan(cfg)                      # A3D_TILE_PLAN=uniform: tiles of equal size
if os.environ.get("AB_TILE"):
    tile, tmap = int(os.environ["AB_TILE"]), None
tmap_ptr, n_tiles = (tmap.data_ptr(), int(tmap.shape[0])) if tmap is not None else (None, 0)
print("tile", tile, "planned tiles", n_tiles, "points per job", inp.batch.jobs["pcd_cap"][:8].tolist())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def project():
    _lib.check(lib.a3d_project(C.byref(cam), db.jobs.data_ptr(), db.n_jobs, db.max_cand, tile,
                               pool.source_bits.data_ptr(), pool.source_bbox.data_ptr(), db.xform.data_ptr(),
                               pcd_ws.data_ptr(), pcd_count.data_ptr(), hom_ws.data_ptr(), tmap_ptr, n_tiles, proj_bits.data_ptr(),
                               proj_popc.data_ptr(),
                               proj_bbox.data_ptr(), int(os.environ.get("AB_OUT_MODE", "0")), torch.cuda.current_stream().cuda_stream),
               "a3d_project")


ref = None
for kernel in ("exact", "filter", "exact", "filter"):
    os.environ["A3D_PROJECT_KERNEL"] = kernel
    for _ in range(3):
        project()
    torch.cuda.synchronize()
    out = (proj_bits.clone(), proj_popc.clone(), proj_bbox.clone())
    if ref is None:
        ref = out
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    torch.cuda._sleep(int(0.04 * 1.9e9))
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        project()
        ev[k][1].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    print(f"{name} mode {mode} {kernel:6s}: median {t[len(t) // 2] * 1e3:8.1f} us  min {t[0] * 1e3:8.1f} us  "
          f"results {'same' if same else 'DIFFERENT'} (points {int(pcd_count[:-1].sum())}, candidates {nc})", flush=True)
