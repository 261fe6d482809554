"""Multi-GPU plumbing: shard independent videos over ranks, gather fixed-size records.

The hot path has no exchange step: every track of every video is
independent and a video's RNG chain stays on one GPU.  Ranks take contiguous blocks
of ``ceil(n / world)`` videos — the reference's SLURM-array split
(tools/opt_arti.py:116-123) — and the only collective is one ``all_gather`` of the
per-track-frame records ``{video, track, frame, angle_id, inter, union}`` and
per-track records at the end (NCCL over NVLink on GPUs, gloo on CPU for tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

FRAME_REC = 6     # video, kind(0 trans / 1 rot), track, frame, angle_id, inter, (union in col 6)
FRAME_COLS = 7
TRACK_COLS = 10   # video, kind, track, has_rot, center_frame, rsq bits, std_axis[4] (ints or fp32 bits)


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block of rank ``rank``: [rank*ceil(n/world), min(n, (rank+1)*ceil(n/world)))."""
    per = -(-n_items // world) if n_items else 0
    return range(min(n_items, rank * per), min(n_items, (rank + 1) * per))


def pack_records(video_ids, planes_list):
    """Flatten the ``plane['fit']`` results of optimised videos into two int32 tables."""
    frames, tracks = [], []
    for vid, planes in zip(video_ids, planes_list):
        for kind, cat in enumerate(("trans", "rot")):
            for ti, plane in enumerate(planes[cat]):
                fit = plane.get("fit", {})
                rsq = np.float32(np.nanmax(fit["rsq"]) if len(fit.get("rsq", [])) and not np.all(np.isnan(fit["rsq"]))
                                 else np.nan)
                std = np.zeros(4, np.int32)
                if plane.get("has_rot"):
                    a = torch.as_tensor(plane["std_axis"]).reshape(-1)
                    if a.dtype.is_floating_point:
                        std[:a.numel()] = a.to(torch.float32).numpy().view(np.int32)[:4]
                    else:
                        std[:min(4, a.numel())] = a.numpy().astype(np.int32)[:4]
                tracks.append([vid, kind, ti, int(bool(plane.get("has_rot"))), int(fit.get("center_frame", -1)),
                               int(rsq.view(np.int32))] + std.tolist())
                if plane.get("has_rot"):
                    for f, a, i, u in zip(fit["frames"], fit["angle_id"], fit["inter"], fit["union"]):
                        frames.append([vid, kind, ti, int(f), int(a), int(i), int(u)])
    fr = torch.tensor(frames, dtype=torch.int32).reshape(-1, FRAME_COLS)
    tr = torch.tensor(tracks, dtype=torch.int32).reshape(-1, TRACK_COLS)
    return fr, tr


def all_gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate 2-D int32 tables of different lengths from all ranks (rank order).
    One size exchange + ONE padded ``all_gather_into_tensor``; works with nccl (CUDA tensors) and gloo."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = sizes.tolist()
    mx = max(sizes + [1])
    cols = local.shape[1]
    padded = torch.zeros(mx, cols, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty(world * mx, cols, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx: r * mx + s] for r, s in enumerate(sizes)])


def optimize_videos_sharded(videos, seeds, cfg=None, device=None, optimize_fn=None, group=None):
    """Each rank optimises its contiguous block of videos, then every rank receives the
    records of all videos.  ``optimize_fn(videos, seeds, cfg=, device=)`` defaults to
    ``opt_utils.optimize_videos``.  Returns (local outputs, local video ids, frame records,
    track records)."""
    if optimize_fn is None:
        from .opt_utils import optimize_videos as optimize_fn
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = list(shard_range(len(videos), rank, world))
    local = [videos[i] for i in mine]          # entries (preds, None) are tracked by optimize_videos, in place
    outs = optimize_fn(local, [seeds[i] for i in mine], cfg=cfg, device=device) if mine else []
    fr, tr = pack_records(mine, [pl for _, pl in local])
    if device is not None and dist.is_initialized() and dist.get_backend(group) == "nccl":
        fr, tr = fr.to(device), tr.to(device)
    return outs, mine, all_gather_rows(fr, group), all_gather_rows(tr, group)
