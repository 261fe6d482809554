"""How often the filtered projection falls back to the exact chain (debug build with -DA3D_FILTER_STATS:
`nvcc ... -DA3D_FILTER_STATS -o tools/_build/liba3d_stats.so`; A3D_LIB points _lib at it)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import _lib, engine, workloads  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["c2"]:
    for mode in (0, 1, 2):
        inp = workloads.build_pass(workloads.WORKLOADS[name], 2020, dev, mode=mode)
        out = (C.c_ulonglong * 4)()
        lib.a3d_debug_filter_stats(None, 1)
        engine.run_pass(inp.cfg, inp.pool, inp.dbatch)
        torch.cuda.synchronize()
        lib.a3d_debug_filter_stats(out, 0)
        n, unc, eo, it = [int(v) for v in out]
        print(f"{name} mode {mode}: pairs {n}, to exact chain {unc} ({unc / max(n, 1):.4%}), of which exact-only candidates "
              f"{eo} ({eo / max(n, 1):.4%}), unproven {(unc - eo) / max(n - eo, 1):.4%}; exact-loop warp iterations {it} "
              f"= {it * 32 / max(n, 1):.4f} per pair-lane", flush=True)
