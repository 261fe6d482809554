"""Minimal single pass through a3d_score (TMA kernel) for compute-sanitizer."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulation3d_b200 import OptConfig, _lib, engine, geometry, synth
cfg = OptConfig()
preds, _ = synth.make_video(77, 1, 10, kinds=[0])
masks = torch.stack([p.pred_masks[0] for p in preds])
pool = engine.pack_masks(masks.cuda())
geo = geometry.source_geometry(preds[0], 0, cfg, False)
xf = geometry.xforms_seq(geometry.rotation_matrices(cfg.rot_cluster_grid, geo.dir_vec))
b = engine.build_batch([0], [0], [geo.normal.numpy()], [float(geo.offset)], [geo.pivot], [xf], [list(range(10))],
                       pool.source_points)
res = engine.run_pass(cfg, pool, engine.DeviceBatch(b, "cuda:0"), want_table=True)
torch.cuda.synchronize()
print("ok", res.best_cand.tolist(), res.best_inter.tolist())
