"""Named synthetic workloads (BASELINE.json ``configs``) as
device-resident scoring passes and as host-side clips.

A *pass* = one source frame per track x T target frames x A candidates; one
unit of work = one (target frame, candidate) IoU evaluation.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, engine, geometry, synth
from .config import OptConfig, rot_grid
from .structures import Boxes, Instances


@dataclass
class Workload:
    name: str
    description: str
    videos: int          # per GPU
    tracks: int          # per video
    frames: int
    cand: int            # candidate grid of the scoring pass
    width: int = 640
    height: int = 480
    mode: int = _lib.MODE_SEQ     # transform of the pass: three-step rotations (cluster phase), composed, or translations

    def cfg(self) -> OptConfig:
        final = max(1, int(round(self.cand * 2 / 3)))
        return OptConfig.scaled(self.width, self.height,
                                rot_cluster_grid=rot_grid(self.cand),
                                rot_final_grid=rot_grid(final, -np.pi / 2, np.pi / 2)) \
            if (self.width, self.height) != (640, 480) else \
            OptConfig(rot_cluster_grid=rot_grid(self.cand),
                      rot_final_grid=rot_grid(final, -np.pi / 2, np.pi / 2))

    @property
    def units_per_pass(self) -> int:
        return self.videos * self.tracks * self.frames * self.cand

    def alg_bytes_per_pass(self) -> int:
        """per track (T+1) packed masks + 16 B/frame of results + 48 B/candidate."""
        pitch = _lib.pitch_words(self.width)
        per_track = (self.frames + 1) * self.height * pitch * 4 + 16 * self.frames + 48 * self.cand
        return self.videos * self.tracks * per_track


WORKLOADS = {
    # configs[1]: the single-GPU configuration the metric is quoted on
    "c2": Workload("c2", "opt_arti synthetic: 1 video, 4 plane tracks x 60 frames, 90-angle grid, 640x480",
                   1, 4, 60, 90),
    # configs[2] per-GPU shard at 8 GPUs (256 videos / 8), and the whole thing on one GPU
    "c3_shard": Workload("c3_shard", "batched opt_arti shard: 32 videos x 8 tracks x 120 frames, 180-angle grid",
                         32, 8, 120, 180),
    "c3_mini": Workload("c3_mini", "profiling slice: 8 videos x 8 tracks x 120 frames, 180-angle grid",
                        8, 8, 120, 180),
    "c3": Workload("c3", "batched opt_arti: 256 videos x 8 tracks x 120 frames, 180-angle grid",
                   256, 8, 120, 180),
    # configs[3] per-GPU shard: 1024x768, 720 rotation candidates (translation candidates are a second pass)
    "c4_shard": Workload("c4_shard", "dense sweep shard (1 of 8 GPUs): 8 videos x 8 tracks x 120 frames, 720-angle grid, 1024x768",
                         8, 8, 120, 720, 1024, 768),
    # ... and its translation candidates: the same tracks against arange(-1, 1, 0.1) along the axis direction
    "c4_trans": Workload("c4_trans", "dense sweep shard, translation candidates: 8 videos x 8 tracks x 120 frames, "
                         "20 offsets, 1024x768", 8, 8, 120, 20, 1024, 768, _lib.MODE_TRANSLATE),
    # configs[3] whole: 64 videos (512 tracks) on one GPU
    "c4": Workload("c4", "dense sweep: 64 videos x 8 tracks x 120 frames, 720-angle grid, 1024x768",
                   64, 8, 120, 720, 1024, 768),
}


@dataclass
class PassInputs:
    cfg: OptConfig
    pool: engine.MaskPool
    batch: engine.JobBatch
    dbatch: engine.DeviceBatch
    units: int


def build_pass(wl: Workload, seed0: int, device, source_frame: int | None = None,
               progress=None, mode: int | None = None, video_ids=None) -> PassInputs:
    """Render every track of ``wl`` on the device, pack the masks, and describe one pass with the
    middle frame of each track as source: cluster-phase three-step rotations (``MODE_SEQ``, the
    default), final-phase composed rotations or translations along the axis direction.
    ``video_ids``: the videos of the workload this device holds (default all); video v is always the
    scene of seed ``seed0 + v``, so shards of any world size add up to the same workload."""
    cfg = wl.cfg()
    mode = wl.mode if mode is None else mode
    T = wl.frames
    s = T // 2 if source_frame is None else source_frame
    bits, srcs, normals, offsets, pivots, dirs = [], [], [], [], [], []
    n = 0
    for v in (range(wl.videos) if video_ids is None else video_ids):
        scene = synth.make_scene(seed0 + v, wl.tracks, T, cfg, kinds=[synth.KIND_ROT] * wl.tracks)
        for k in range(wl.tracks):
            masks = synth.render_track_masks(scene, k, cfg, device=device)            # (T,H,W) bool
            boxes, planes, rot_axis, tran_axis = synth.track_predictions(scene, k, cfg, masks, frames=[s])
            inst = Instances((cfg.height, cfg.width))
            inst.pred_boxes = Boxes(boxes[s:s + 1])
            inst.pred_planes = planes[s:s + 1]
            inst.pred_rot_axis = rot_axis[s:s + 1]
            inst.pred_tran_axis = tran_axis[s:s + 1]
            geo = geometry.source_geometry(inst, 0, cfg, False)
            p = engine.pack_masks(masks)
            bits.append(p.bits)
            srcs.append(n + s)
            normals.append(geo.normal.numpy())
            offsets.append(float(geo.offset))
            pivots.append(geo.pivot)
            dirs.append(geo.dir_vec)
            n += T
        if progress:
            progress(v)
    H, pitch = cfg.height, _lib.pitch_words(cfg.width)
    all_bits = torch.empty(n, H, pitch, dtype=torch.int32, device=device)      # no second copy of the pool
    o = 0
    while bits:
        b = bits.pop(0)
        all_bits[o:o + b.shape[0]] = b
        o += b.shape[0]
    pool = engine.pool_from_bits(all_bits, cfg.height, cfg.width)
    if mode == _lib.MODE_TRANSLATE:
        xf = geometry.xforms_translate(np.linspace(-1.0, 1.0, wl.cand, endpoint=False), np.stack(dirs))
    else:
        R = geometry.rotation_matrices(cfg.rot_cluster_grid, np.stack(dirs))         # (S,A,3,3)
        xf = geometry.xforms_seq(R) if mode == _lib.MODE_SEQ else geometry.xforms_composed(R, np.stack(pivots))
    S = len(srcs)
    batch = engine.build_batch(srcs, [mode] * S, normals, offsets, pivots, list(xf),
                               [np.arange(i * T, (i + 1) * T, dtype=np.int32) for i in range(S)],
                               pool.source_points)
    dbatch = engine.DeviceBatch(batch, device)
    return PassInputs(cfg, pool, batch, dbatch, batch.units)


def make_clip(wl: Workload, seed: int, tracks: int | None = None, frames: int | None = None,
              kinds=None):
    """Host-side clip of the workload's shape (list[Instances] with fp32 CPU masks)."""
    cfg = wl.cfg()
    tracks = tracks or wl.tracks
    frames = frames or wl.frames
    kinds = kinds if kinds is not None else [synth.KIND_ROT] * tracks
    preds, _ = synth.make_video(seed, tracks, frames, cfg, kinds=kinds)
    return preds, cfg
