"""articulation3d_b200 — B200-native temporal articulation optimizer.

Drop-in for the post-detection temporal stage of JasonQSY/Articulation3D
(``track_planes`` + ``optimize_planes``, reference utils/opt_utils.py:962-974,
1156-1208).  Host code is Python/PyTorch; the hot path runs in hand-written
sm_100a CUDA kernels behind a C-ABI shared library (include/a3d.h).
"""
from .config import OptConfig, rot_grid
from .structures import Boxes, Instances, pairwise_iou

__all__ = ["OptConfig", "rot_grid", "Boxes", "Instances", "pairwise_iou"]
__version__ = "0.1.0"
