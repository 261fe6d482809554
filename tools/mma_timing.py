"""Per-role cycle breakdown of k_score_mma (CTA 0), from a library built with -DA3D_MMA_TIMING:
    nvcc <flags of articulation3d_b200/build.py> -DA3D_MMA_TIMING -o tools/_build/liba3d_timing.so articulation3d_b200/csrc/a3d.cu
    python tools/mma_timing.py c3_shard"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from articulation3d_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "tools", "_build", "liba3d_timing.so")
from articulation3d_b200 import engine, workloads  # noqa: E402

os.environ["A3D_SCORE_KERNEL"] = "mma"
name = sys.argv[1] if len(sys.argv) > 1 else "c3_shard"
dev = torch.device("cuda:0")
inp = workloads.build_pass(workloads.WORKLOADS[name], 2020, dev)
ws = engine.Workspace(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    if not os.environ.get("AB_NOFLUSH"):
        flush.zero_()
    engine.run_pass(inp.cfg, inp.pool, inp.dbatch, ws)
    torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 64)()
lib.a3d_debug_mma_timing.argtypes = [C.c_void_p]
assert lib.a3d_debug_mma_timing(buf) == 0
t = list(buf)
total, nsteps, nmask = t[32], t[33], t[34]
print(f"{name}: CTA 0 ran {total} cycles, {nsteps} steps ({total / max(nsteps, 1):.0f} cycles/step), {nmask} masks")
for row, (who, names) in enumerate([
        ("expander warp 0", ["wait raw_full", "lds", "wait empty", "expand+sts", "proxy fence", "syncwarp+arrive", "steps"]),
        ("expander warp 15", ["wait raw_full", "lds", "wait empty", "expand+sts", "proxy fence", "syncwarp+arrive", "steps"]),
        ("issuer", ["wait full", "issue+commit", "steps"]),
        ("loader warp 0", ["wait raw_empty", "issue loads", "first data", "sts+arrive", "turns"])]):
    vals = t[row * 8: row * 8 + len(names)]
    n = max(vals[-1], 1)
    print(f"  {who:17s} " + "  ".join(f"{nm} {v / n:.0f}" for nm, v in zip(names[:-1], vals[:-1])) + f"  (per {names[-1][:-1]}, {vals[-1]} {names[-1]})")
