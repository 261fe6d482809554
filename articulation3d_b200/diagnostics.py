"""Diagnostics of the temporal stage: how consistent the predicted
rotation axes and plane normals of a track are before and after ``optimize_planes``.

Host-only, a few hundred flops per track; mirrors the reference's names and return values
(utils/opt_utils.py:49-72 ``fit_plane_from_normals``, :977-1065 ``check_axis``,
:1068-1152 ``check_monotonic``; utils/metrics.py:52-102 ``sa_metric``, ``se_metric``,
``EA_metric``, ``Line``).  Pinned by tests/golden/diag_*.npz (outputs of the reference itself).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .axis import angle_offset_to_axis


# ---------------------------------------------------------------------------
# line metrics (utils/metrics.py)
# ---------------------------------------------------------------------------
class Line:
    """A 2-D segment given as [y0, x0, y1, x1] (metrics.py:70-102)."""

    def __init__(self, coordinates=(0, 0, 1, 1)):
        c = list(coordinates)
        if len(c) != 4:
            raise AssertionError("a line needs four coordinates")
        if c[0] == c[2] and c[1] == c[3]:
            raise AssertionError("degenerate line")
        self._c = c

    @property
    def coord(self):
        return self._c

    @property
    def length(self):
        y0, x0, y1, x1 = (float(v) for v in self._c)
        return float(np.hypot(y0 - y1, x0 - x1))

    def angle(self):
        y0, x0, y1, x1 = (float(v) for v in self._c)
        if x0 == x1:
            return -np.pi / 2
        return float(np.arctan((y0 - y1) / (x0 - x1)))

    def __repr__(self):
        return str(self._c)


def sa_metric(angle_p, angle_g) -> float:
    """Angle similarity in [0, 1]: (1 - normalised angle difference)^2."""
    d = abs(float(angle_p) - float(angle_g))
    d = min(d, np.pi - d) * 2 / np.pi
    return max(0.0, 1 - d) ** 2


def se_metric(coord_p, coord_g, size=(640, 480)) -> float:
    """Mid-point similarity in [0, 1]: (1 - distance / longer image side)^2."""
    p = [float(v) for v in coord_p]
    g = [float(v) for v in coord_g]
    d = np.hypot((p[0] + p[2]) / 2 - (g[0] + g[2]) / 2, (p[1] + p[3]) / 2 - (g[1] + g[3]) / 2)
    return max(0.0, 1 - d / max(size)) ** 2


def EA_metric(l_pred: Line, l_gt: Line, size=(640, 480)) -> float:
    return sa_metric(l_pred.angle(), l_gt.angle()) * se_metric(l_pred.coord, l_gt.coord, size=size)


# ---------------------------------------------------------------------------
# track diagnostics (utils/opt_utils.py)
# ---------------------------------------------------------------------------
def fit_plane_from_normals(normals: torch.Tensor) -> torch.Tensor:
    """Unit vector spanning the LAST right-singular direction of S^T S for the (N, 3) normals S —
    the direction the normals are, together, most perpendicular to (the hinge direction of a
    rotating plane).  The reference's docstring says "largest", its code takes index 2."""
    sts = normals.transpose(0, 1) @ normals
    return torch.linalg.svd(sts).Vh[2]


def _track_axes(preds, plane):
    """(n, 4) int64 axis end-points [x1, y1, x2, y2] and the box scores of a track's frames."""
    axes, scores = [], []
    for frame, box_id in plane['ids'].items():
        p = preds[frame]
        centers = p.pred_boxes.get_centers()
        # the reference hands ALL box centres with the one axis row; row 0 of the result pairs the
        # axis with the centre of box 0 (opt_utils.py:1003-1004), which is what is reproduced here
        axes.append(angle_offset_to_axis(p.pred_rot_axis[box_id:box_id + 1], centers)[:1])
        scores.append(float(p.scores[box_id]))
    return torch.cat(axes, dim=0), torch.tensor(scores, dtype=torch.float32)


def _axis_distance(axes: torch.Tensor) -> torch.Tensor:
    """EA score of every ordered pair (i, j), i != j, of a track's axes; 0 for degenerate lines."""
    n = axes.shape[0]
    a = axes.tolist()
    out = []
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            try:
                li = Line([a[i][1], a[i][0], a[i][3], a[i][2]])
                lj = Line([a[j][1], a[j][0], a[j][3], a[j][2]])
                out.append(EA_metric(li, lj))
            except AssertionError:
                out.append(0.0)
    return torch.tensor(out, dtype=torch.float32)


def check_axis(preds, opt_preds, planes, method=None, frames=None):
    """Pairwise axis agreement inside every track, before and after optimisation.  Tracks whose
    mean box score dropped by 0.1 or more (the optimiser rejected them) are left out of both lists."""
    scores_all, opt_scores_all = [], []
    for plane in planes:
        axes, box_scores = _track_axes(preds, plane)
        opt_axes, opt_box_scores = _track_axes(opt_preds, plane)
        scores, opt_scores = _axis_distance(axes), _axis_distance(opt_axes)
        if box_scores.mean() - opt_box_scores.mean() < 0.1:
            scores_all.extend(scores)
            opt_scores_all.extend(opt_scores)
    return scores_all, opt_scores_all


def _track_normals(preds, plane) -> torch.Tensor:
    out = []
    for frame, box_id in plane['ids'].items():
        p = preds[frame].pred_planes[box_id:box_id + 1].clone()
        p[:, [1, 2]] = p[:, [2, 1]]
        p[:, 1] = -p[:, 1]
        out.append(F.normalize(p, p=2))
    return torch.cat(out, dim=0)


def check_monotonic(preds, opt_preds, planes, method=None, frames=None):
    """Per track: mean |normal . hinge direction| with the hinge fitted from the track's own normals
    (0 = the normals sweep a perfect plane), before and after optimisation, as [[score]] lists."""
    corrs, opt_corrs = [], []
    for plane in planes:
        for src, dst in ((preds, corrs), (opt_preds, opt_corrs)):
            normals = _track_normals(src, plane)
            n = fit_plane_from_normals(normals)
            dst.append([(normals @ n.unsqueeze(1)).abs().mean()])
    return corrs, opt_corrs
