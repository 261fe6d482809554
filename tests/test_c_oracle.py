"""Pins the fast C checker (oracle/a3d_oracle.c) against the torch/numpy oracle
(oracle/restated.py), which is itself pinned to the reference's outputs."""
import numpy as np
import pytest
import torch

from articulation3d_b200 import synth
from oracle import c_oracle, restated

MODES = {"seq": 0, "composed": 1, "translate": 2}


def _xform(g, name, grid, ocfg):
    a = g["axis3d"][0].astype(np.float32)
    xf = np.zeros((len(grid), 12), np.float32)
    if name == "translate":
        xf[:, [0, 4, 8]] = 1
        xf[:, 9:] = restated.translation_vectors(grid, g["dir_vec"])
    else:
        R = restated.rotation_matrices(grid, g["dir_vec"])
        xf[:, :9] = R.reshape(-1, 9)
        if name == "composed":
            xf[:, 9:] = restated.composed_last_row(a, R)
    return xf, a


@pytest.mark.parametrize("name", ["seq", "composed", "translate"])
def test_c_project_and_score_equal_restated(name):
    ocfg = restated.OracleConfig()
    preds, _ = synth.make_video(61, 2, 8, kinds=[0, 1])
    box = 1 if name == "translate" else 0
    grid = {"seq": ocfg.rot_cluster_grid, "composed": ocfg.rot_final_grid, "translate": ocfg.trans_grid}[name]
    masks = np.stack([p.pred_masks[box].numpy() for p in preds])
    bits = c_oracle.pack(masks)
    assert np.array_equal(c_oracle.unpack(bits, 640), masks > 0.5)
    for frame in (1, 5):
        want, _, g = restated.candidate_masks(preds[frame], box, ocfg, name, grid)
        xf, a = _xform(g, name, grid, ocfg)
        got = c_oracle.project(ocfg.K_inv(), ocfg.focal_length, 320, 240, 480, 640, bits[frame],
                               g["normal"].numpy(), float(g["offset"]), a, MODES[name], xf)
        assert np.array_equal(c_oracle.unpack(got, 640), want.numpy() > 0.5)
        inter, uni, best, iou = c_oracle.score(480, 640, bits, got)
        for t in range(len(preds)):
            i2, u2, iou2 = restated.score(preds[t].pred_masks[box], want)
            assert np.array_equal(inter[t], i2.numpy()) and np.array_equal(uni[t], u2.numpy())
            assert best[t] == int(iou2.argmax())
            assert np.array_equal(iou[t], iou2.max().numpy(), equal_nan=True)


def test_c_oracle_nan_and_clamp_semantics():
    ocfg = restated.OracleConfig()
    preds, _ = synth.make_video(62, 1, 4, kinds=[0])
    g = restated.source_geometry(preds[0], 0, ocfg, False)
    bits = c_oracle.pack(preds[0].pred_masks.numpy())
    for bad in (np.nan, np.inf, -np.inf, 1e30, -1e30):
        xf = np.zeros((3, 12), np.float32)
        xf[:, [0, 4, 8]] = 1
        xf[1, :9] = bad
        xf[2, 9:] = bad
        pts = restated.transform_composed(g["pcd"], xf[:, :9].reshape(3, 3, 3), xf[:, 9:])
        row, col = restated.project_pixels(pts, ocfg, 480, 640)
        want = restated.splat(row, col, 480, 640).numpy() > 0.5
        got = c_oracle.project(ocfg.K_inv(), ocfg.focal_length, 320, 240, 480, 640, bits[0], g["normal"].numpy(),
                               float(g["offset"]), np.zeros(3, np.float32), 1, xf)
        assert np.array_equal(c_oracle.unpack(got, 640), want), bad
    # empty target and empty candidate: 0/0 = NaN wins, first index
    z = np.zeros((1, 480, 20), np.uint32)
    inter, uni, best, iou = c_oracle.score(480, 640, z, np.zeros((5, 480, 20), np.uint32))
    assert best[0] == 0 and np.isnan(iou[0])
