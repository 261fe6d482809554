"""Runs a few device-resident passes of a workload (for ncu).  Usage:
    ncu --set full --clock-control none --import-source on -k regex:k_project -s 2 -c 1 \
        -o gpurun_out/prof_project python tools/profile_pass.py --workload c3_mini --passes 4"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import engine, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_mini")
ap.add_argument("--passes", type=int, default=4)
args = ap.parse_args()
dev = torch.device("cuda", 0)
inp = workloads.build_pass(workloads.WORKLOADS[args.workload], 2020, dev)
ws = engine.Workspace(dev)
for _ in range(args.passes):
    res = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, ws)
torch.cuda.synchronize()
print("units/pass", inp.units, "best_cand[:8]", res.best_cand[:8].tolist())
