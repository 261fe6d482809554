"""A/B of one whole pass: a3d_pass (keys cleared first, programmatic dependent launches) against the
a3d_project + a3d_score pair, same inputs, L2 flushed between steps, GPU run-ahead so that launch gaps
of the host do not count.    python tools/pass_ab.py c2"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import engine, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(os.environ.get("AB_ITERS", "20"))
dev = torch.device("cuda:0")
inp = workloads.build_pass(workloads.WORKLOADS[name], 2020, dev)
ws = engine.Workspace(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for api in ("split", "pass", "split", "pass"):
    os.environ["A3D_PASS_API"] = api
    for _ in range(3):
        res = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, ws)
    torch.cuda.synchronize()
    out = torch.stack([res.best_cand.clone(), res.best_inter.clone(), res.best_union.clone()])
    if ref is None:
        ref = out
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    torch.cuda._sleep(int(0.04 * 1.9e9))
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        engine.run_pass(inp.cfg, inp.pool, inp.dbatch, ws)
        ev[k][1].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    print(f"{name} {api:5s}: median {t[len(t) // 2] * 1e3:8.1f} us  min {t[0] * 1e3:8.1f} us  "
          f"results {'same' if torch.equal(ref, out) else 'DIFFERENT'}", flush=True)
