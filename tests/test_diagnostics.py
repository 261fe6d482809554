"""Row a15 of SURVEY.md §8a: the diagnostics check_axis / check_monotonic / fit_plane_from_normals
against outputs of the reference itself (tests/golden/diag/*.npz, oracle/gen_golden.py --diag)."""
import os

import numpy as np
import pytest
import torch

from articulation3d_b200 import diagnostics, synth
from articulation3d_b200.structures import Boxes, Instances
from tests import golden_util as gu

DIAG = os.path.join(gu.ROOT, "tests", "golden", "diag")


def _load(name):
    z = gu.load(name)
    preds = gu.arrays_to_preds(z, Instances, Boxes)
    opt = synth.clone_preds(preds)
    for t, p in enumerate(opt):
        p.scores = z[f"o{t}_scores"]
        p.pred_rot_axis = torch.from_numpy(z[f"o{t}_rot_axis"])
        p.pred_tran_axis = torch.from_numpy(z[f"o{t}_tran_axis"])
        p.pred_planes = torch.from_numpy(z[f"o{t}_planes"])
    planes = []
    for i in range(int(z["rot_n"])):
        ids = dict(z[f"rot{i}_ids"].tolist())
        planes.append({"ids": {int(f): int(ids[int(f)]) for f in z[f"rot{i}_ids_order"]}})
    return preds, opt, planes


@pytest.mark.parametrize("name", ["clip_a", "clip_b", "clip_d"])
def test_check_axis_and_monotonic_match_reference(name):
    want = np.load(os.path.join(DIAG, f"{name}.npz"))
    preds, opt, planes = _load(name)
    s0, s1 = diagnostics.check_axis(preds, opt, planes, "3dc")
    assert len(s0) == len(want["axis_scores"]) and len(s1) == len(want["axis_scores_opt"])
    np.testing.assert_allclose(np.array([float(v) for v in s0]), want["axis_scores"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.array([float(v) for v in s1]), want["axis_scores_opt"], rtol=1e-5, atol=1e-6)
    c0, c1 = diagnostics.check_monotonic(preds, opt, planes, "3dc")
    np.testing.assert_allclose(np.array([float(v[0]) for v in c0]), want["fit_scores"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(np.array([float(v[0]) for v in c1]), want["fit_scores_opt"], rtol=1e-4, atol=1e-6)


def test_line_metrics_and_plane_fit():
    a, b = diagnostics.Line([0, 0, 10, 10]), diagnostics.Line([0, 10, 10, 0])
    assert diagnostics.EA_metric(a, a) == 1.0
    assert diagnostics.sa_metric(a.angle(), b.angle()) == 0.0            # perpendicular
    assert abs(diagnostics.se_metric([0, 0, 0, 0], [0, 64, 0, 64]) - 0.81) < 1e-12
    assert diagnostics.Line([3, 5, 9, 5]).angle() == -np.pi / 2          # vertical in x
    with pytest.raises(AssertionError):
        diagnostics.Line([1, 2, 1, 2])
    # normals swept about the z axis lie in the xy plane: the fitted direction is +-z
    t = torch.linspace(0, 1.2, 9)
    n = torch.stack([torch.cos(t), torch.sin(t), torch.zeros_like(t)], 1)
    v = diagnostics.fit_plane_from_normals(n)
    assert abs(abs(float(v[2])) - 1) < 1e-6
