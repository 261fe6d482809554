"""CPU tests of the on-disk contracts (records <-> Instances, track summaries, .obj)."""
import json
import os

import numpy as np
import torch

from articulation3d_b200 import OptConfig, io, opt_utils, synth


def test_records_round_trip(tmp_path):
    preds, _ = synth.make_video(3, 2, 12, kinds=[0, 1])
    recs = io.preds_to_records(preds, video_id="abcdefghijk_1_20")
    torch.save(recs, tmp_path / "p.pth")
    back = torch.load(tmp_path / "p.pth", weights_only=False)
    dense = io.records_to_preds(back, conf_threshold=0.7, masks="dense")
    lazy = io.records_to_preds(back, conf_threshold=0.7, masks="rle")
    for p, q, r in zip(preds, dense, lazy):
        assert torch.equal(p.pred_boxes.tensor, q.pred_boxes.tensor)
        assert torch.equal(p.pred_masks > 0.5, q.pred_masks > 0.5)
        assert torch.equal(p.pred_planes, q.pred_planes) and torch.equal(p.pred_rot_axis, q.pred_rot_axis)
        assert np.array_equal(p.pred_classes, q.pred_classes) and np.allclose(p.scores, q.scores)
        assert len(r.pred_rle) == len(q.pred_masks)
    # same tracks from both forms
    a, b = opt_utils.track_planes(preds), opt_utils.track_planes(lazy)
    assert [p['ids'] for p in a['rot']] == [p['ids'] for p in b['rot']]


def test_tracks_summary_and_obj(tmp_path):
    cfg = OptConfig()
    preds, _ = synth.make_video(3, 2, 12, kinds=[0, 1])
    planes = opt_utils.track_planes(preds)
    for cat in planes:
        for plane in planes[cat]:
            frames = list(plane['ids'].keys())
            n = len(frames)
            plane['has_rot'] = True
            plane['std_axis'] = torch.tensor([1, 2, 3, 4]) if cat == 'rot' else torch.tensor([0.6, 0.8])
            plane['fit'] = {'frames': frames, 'center_frame': frames[0], 'rsq': np.array([0.9, np.nan]),
                            'angle_id': np.arange(n), 'angle': np.linspace(0, 1, n), 'inter': np.arange(n),
                            'union': np.arange(n) + 5, 'iou': np.full(n, 0.5, np.float32)}
    summ = io.tracks_summary(planes)
    json.dumps(summ)
    assert len(summ) == 2 and summ[0]['kind'] == 'trans' and summ[1]['angle_track'][3]['angle_id'] == 3
    assert summ[0]['rsq'] == [0.9, None]
    nv = io.write_obj(str(tmp_path / "f.obj"), preds, planes, 5, cfg)
    text = open(tmp_path / "f.obj").read()
    assert nv == 12 and text.count("\nf ") == 2 and text.count("\nl ") == 2
    recs = io.preds_to_records(preds)
    out = io.opt_preds_to_records(preds, recs)
    assert out[0]['instances'][0]['bbox'][2] > 0 and 'segmentation' not in out[0]['instances'][0]
    io.save_results(str(tmp_path), "vid", out, planes)
    assert os.path.exists(tmp_path / "vid_tracks.json") and os.path.exists(tmp_path / "vid_predictions_opt.pth")
