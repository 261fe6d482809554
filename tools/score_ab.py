"""A/B timing of the scoring kernels on one workload pass: a3d_project runs once, then
a3d_score is timed alone (CUDA events, L2 flushed between launches) for every kernel in
A3D_SCORE_KERNEL = ldg | tma | mma, and their outputs are compared with the first one.

    python tools/score_ab.py c3_shard ldg mma
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import _lib  # noqa: E402

if os.environ.get("A3D_LIB"):              # a tuning build of csrc/a3d.cu (tools/_build/*.so)
    _lib.LIB_PATH = os.path.abspath(os.environ["A3D_LIB"])
from articulation3d_b200 import engine, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3_shard"
kernels = sys.argv[2:] or ["ldg", "mma"]
iters = int(os.environ.get("AB_ITERS", "10"))
dev = torch.device("cuda:0")
if "," in name:                            # custom: videos,tracks,frames,candidates
    v, t, f, c = (int(x) for x in name.split(","))
    wl = workloads.Workload(name, "custom", v, t, f, c)
else:
    wl = workloads.WORKLOADS[name]
inp = workloads.build_pass(wl, 2020, dev)
cfg, pool, db = inp.cfg, inp.pool, inp.dbatch
ws = engine.Workspace(dev)
os.environ["A3D_SCORE_KERNEL"] = kernels[0]
res = engine.run_pass(cfg, pool, db, ws, want_table=True)
torch.cuda.synchronize()
lib = _lib.load()
H, W = cfg.height, cfg.width
nt, nc = db.n_tgt_total, db.n_cand_total
key_ws = ws.get("key_ws", (nt,), torch.int64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = engine._stream_ptr()


def score(table):
    _lib.check(lib.a3d_score(H, W, db.jobs.data_ptr(), db.n_jobs, db.max_tgt, db.max_cand, nt, len(pool), nc,
                             pool.bits.data_ptr(), pool.popc.data_ptr(), pool.bbox.data_ptr(),
                             db.tgt_index.data_ptr(), res.proj_bits.data_ptr(), res.proj_popc.data_ptr(),
                             res.proj_bbox.data_ptr(), key_ws.data_ptr(),
                             res.inter_tab.data_ptr() if table else None,
                             res.best_cand.data_ptr(), res.best_inter.data_ptr(), res.best_union.data_ptr(),
                             res.best_iou.data_ptr(), stream), "a3d_score")


ref = None
for kern in kernels:
    if ":" in kern:                       # e.g. mma:pf=0
        kern, opt = kern.split(":")
        os.environ["A3D_MMA_PF"] = opt.split("=")[1]
    os.environ["A3D_SCORE_KERNEL"] = kern
    res.inter_tab.zero_()
    score(True)
    torch.cuda.synchronize()
    out = (res.inter_tab.clone(), res.best_cand.clone(), res.best_inter.clone(), res.best_union.clone(),
           res.best_iou.clone().view(torch.int32))
    if ref is None:
        ref = out
        same = "reference"
    else:
        same = "identical" if all(torch.equal(a, b) for a, b in zip(ref, out)) else "DIFFERENT"
    ts = []
    for _ in range(iters):
        if not os.environ.get("AB_NOFLUSH"):
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        score(False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name} {kern}: a3d_score median {ts[len(ts) // 2] * 1e3:.1f} us  min {ts[0] * 1e3:.1f} us  "
          f"({inp.units / ts[len(ts) // 2] / 1e3:.3e} units/s)  results {same}", flush=True)
