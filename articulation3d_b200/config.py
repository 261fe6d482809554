"""Every literal of the reference's temporal optimizer, lifted into one object.

The reference hard-codes all of these (SURVEY.md §5 "Config / flags"); the
defaults below are those literals, so ``OptConfig()`` reproduces the reference
and the BASELINE configs (90/180/720-angle grids, 1024x768) are expressed by
overriding fields.  Citations are to /root/reference/articulation3d/articulation3d/.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch


def _rot_cluster_grid():
    return np.arange(-np.pi / 2, np.pi, np.pi / 30)          # utils/opt_utils.py:425-427 (45)


def _rot_final_grid():
    return np.arange(-np.pi / 2, np.pi / 2, np.pi / 30)      # utils/opt_utils.py:560-562 (30)


def _legacy_grid():
    return np.arange(-np.pi / 2, 0.1, np.pi / 30)            # utils/opt_utils.py:144-145 (16)


def _trans_grid():
    return torch.arange(-1, 1, 0.1)                          # utils/opt_utils.py:724,863 (20)


@dataclass
class OptConfig:
    # camera model of get_pcd / project2D (utils/vis.py:62-68, 86-95)
    height: int = 480
    width: int = 640
    focal_length: float = 517.97
    # candidate grids
    rot_cluster_grid: np.ndarray = field(default_factory=_rot_cluster_grid)
    rot_final_grid: np.ndarray = field(default_factory=_rot_final_grid)
    legacy_grid: np.ndarray = field(default_factory=_legacy_grid)
    trans_grid: torch.Tensor = field(default_factory=_trans_grid)
    # clustering / model selection (utils/opt_utils.py:392, 484, 505, 517)
    rounds: int = 5
    inlier_iou: float = 0.5
    min_inliers: int = 5
    rsq_thresh: float = 0.3
    # soft filter (utils/opt_utils.py:670, 947; legacy :368)
    score_decay: float = 0.6
    legacy_score_decay: float = 0.8
    # tracker (utils/opt_utils.py:1177, 1181, 1203)
    track_max_gap: int = 5
    track_iou: float = 0.5
    track_min_len: int = 10
    # mask binarisation of the scoring stage (utils/opt_utils.py:471)
    mask_thresh: float = 0.5
    # Pearson r of a cluster whose angle list is constant (a static plane: zero variance AND zero
    # covariance).  scipy.stats.linregress of the reference's era (<= 1.8) returns r = 0.0 there
    # (-> R^2 = 0 < 0.3 -> has_rot False, score x0.6); scipy >= 1.9 returns NaN (-> `NaN < 0.3` is
    # False -> has_rot True).  None = whatever the scipy installed beside this package does, i.e. what
    # the unmodified reference would do in the same environment; 0.0 / float('nan') force either.
    constant_track_r: float | None = None
    # host schedule of the cluster phase: 'table' = one all-sources device pass + host replay,
    # 'chain' = one device pass per round, 'auto' = table while it costs at most table_max_units
    # (track-frame x candidate evaluations) of device work
    schedule: str = "auto"
    table_max_units: int = 96_000_000
    # optimize_videos: videos optimised concurrently (worker threads, each with its own stream and pass
    # buffers; results are independent of it: every video draws from its own seeded generator).  Measured on
    # the 8-track, 120-frame clips: 2 workers 247-830 ms per 6 clips against 266-270 ms for one (the host side is
    # Python under one interpreter lock), so one is the default.
    pipeline_workers: int = 1

    @property
    def cx(self) -> float:
        return self.width / 2

    @property
    def cy(self) -> float:
        return self.height / 2

    def K(self) -> np.ndarray:
        return np.array([[self.focal_length, 0, self.cx],
                         [0, self.focal_length, self.cy],
                         [0, 0, 1]])

    def K_inv(self) -> np.ndarray:
        """float64 inverse exactly as numpy returns it (utils/vis.py:95): it is
        passed to the device as data, never re-derived in closed form
        (the two differ by one ulp in [0,2])."""
        key = (self.focal_length, self.width, self.height)
        if getattr(self, "_kinv_key", None) != key:
            object.__setattr__(self, "_kinv", np.linalg.inv(self.K()))
            object.__setattr__(self, "_kinv_key", key)
        return self._kinv

    @staticmethod
    def scaled(width: int, height: int, **kw) -> "OptConfig":
        """Intrinsics scaled from the reference's 640x480 camera, e.g. the
        1024x768 dense-sweep config (f = 517.97 * width / 640)."""
        return OptConfig(height=height, width=width,
                         focal_length=517.97 * width / 640.0, **kw)


def rot_grid(n: int, lo: float = -np.pi / 2, hi: float = np.pi) -> np.ndarray:
    """n-point rotation grid lo + k*(hi-lo)/n: the 90/180/720
    angle grids of the BASELINE configs over the reference's cluster range."""
    return lo + np.arange(n) * ((hi - lo) / n)
