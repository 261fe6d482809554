"""GPU tests of the two tools at the pipeline shapes of BASELINE.json configs[0] / configs[4]: the temporal
stage of ``tools/inference.py`` on the detector stub at 30, 90 and 300 frames (the R-CNN itself is out of
scope), key frames [0, 30, 60, 89] exported as the reference does (tools/inference.py:282)."""
import json
import os
import random

import numpy as np
import pytest
import torch

from articulation3d_b200 import OptConfig, opt_utils, synth
from articulation3d_b200.tools import inference
from oracle import restated

pytestmark = pytest.mark.gpu


def _load(out_dir):
    recs = torch.load(os.path.join(out_dir, "synthetic00_0_0_predictions_opt.pth"), weights_only=False)
    with open(os.path.join(out_dir, "synthetic00_0_0_tracks.json")) as f:
        return recs, json.load(f)


def test_inference_tool_c1_shape_matches_oracle(tmp_path):
    """configs[0] shape: a 30-frame 640x480 clip, temporal optimisation on; the saved records and fitted
    tracks equal the CPU oracle's on the same clip and seed."""
    out = str(tmp_path / "c1")
    inference.main(["--output", out, "--frames", "30", "--tracks", "3", "--seed", "2020", "--save-obj",
                    "--save-textured-obj"])
    recs, tracks = _load(out)
    assert len(recs) == 30
    cfg = OptConfig()
    preds, _ = synth.make_video(2020, 3, 30, cfg)
    random.seed(2020)
    planes = restated.track_planes(preds)
    want = restated.optimize_planes(preds, planes, "3dc")
    order = [p for cat in ("trans", "rot") for p in planes[cat]]
    assert len(tracks) == len(order)
    for t, p in zip(tracks, order):
        assert t["has_rot"] == bool(p["has_rot"]) and t["frames"] == list(p["ids"].keys())
        if t["has_rot"]:
            np.testing.assert_allclose(t["std_axis"], torch.as_tensor(p["std_axis"]).reshape(-1).tolist(), rtol=1e-6)
    for r, w in zip(recs, want):
        assert np.array_equal(np.array([i["score"] for i in r["instances"]]), np.asarray(w.scores))
        np.testing.assert_allclose(r["pred_rot_axis"].numpy(), w.pred_rot_axis.numpy(), rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(r["pred_tran_axis"].numpy(), w.pred_tran_axis.numpy(), rtol=1e-4, atol=1e-7)
    for k in (0, 29):                                                      # key frames clipped to the clip
        assert os.path.getsize(os.path.join(out, f"frame{k}.obj")) > 0
    for k in (0, 29):                                                      # the textured export, one folder per frame
        assert os.path.getsize(os.path.join(out, "frame_{:0>4}".format(k), "arti_pred.mtl")) > 0


@pytest.mark.parametrize("frames", [90, 300])
def test_inference_tool_c5_shapes_both_schedules(frames, tmp_path, monkeypatch):
    """configs[4] shape (300-frame clip; 90 frames is the shortest clip with all four key frames): the
    all-sources schedule (one device pass for the cluster phase + host replay) and the chained schedule
    (one pass per round) must save identical results; every tracked frame gets an angle; key frames
    [0, 30, 60, 89] are exported."""
    outs = {}
    for sched in ("table", "chain"):
        monkeypatch.setenv("A3D_SCHEDULE", sched)
        out = str(tmp_path / sched)
        inference.main(["--output", out, "--frames", str(frames), "--tracks", "3", "--seed", "7", "--save-obj"])
        outs[sched] = _load(out)
        for k in (0, 30, 60, 89):
            assert os.path.getsize(os.path.join(out, f"frame{k}.obj")) > 0
    (ra, ta), (rb, tb) = outs["table"], outs["chain"]
    assert ta == tb
    assert len(ra) == frames == len(rb)
    for a, b in zip(ra, rb):
        assert [i["score"] for i in a["instances"]] == [i["score"] for i in b["instances"]]
        assert torch.equal(a["pred_rot_axis"], b["pred_rot_axis"]) and torch.equal(a["pred_tran_axis"], b["pred_tran_axis"])
    fitted = [t for t in ta if t["has_rot"]]
    assert fitted, "no track was fitted"
    for t in fitted:
        assert [e["frame"] for e in t["angle_track"]] == t["frames"] and len(t["frames"]) >= 10
        assert all(0 <= e["inter"] <= e["union"] for e in t["angle_track"])
