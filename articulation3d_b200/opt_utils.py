"""Drop-in temporal articulation optimizer: ``track_planes`` / ``optimize_planes``.

Same names, argument meaning, return values, in-place side effects and global
``random`` consumption as the reference's utils/opt_utils.py (:962-974 dispatcher,
:382-682 ``optimize_planes_3dc``, :685-959 ``optimize_planes_3d_trans``,
:112-379 legacy ``'3d'``, :75-109 ``'average'``, :1156-1208 ``track_planes``),
with the pixel work moved to the device:

    reference (per source frame)                     here
    -------------------------------------------     ---------------------------------
    nonzero + get_pcd + Transform3d x A + project2D  a3d_project   (one launch per pass)
      + A python-loop scatters
    per target: >0.5, &, |, sum, sum, /, argmax      a3d_score     (one launch per pass)
    proj_masks[angle_id].cpu() per frame             lazy a3d_emit_masks

The control flow that depends on Python's global RNG and on list mutation
 is kept on the host, written once as a generator per
track list: it *yields* a job (source frame, candidate transforms, target frames)
and is *sent* the per-target arg-max back.  One video is driven job by job; many
videos are driven in lock-step so each device pass carries one job per video
(``optimize_videos``).

Additions over the reference (never removals): every optimised track gets
``plane['fit']`` — the per-frame angle index / inter / union / iou of the final
assignment, cluster R^2 values and the centre frame — i.e. the "angle-per-frame
track" the reference only holds implicitly; ``plane['reg_masks']`` is a lazy
mapping that materialises the fp32 masks on access instead of copying T x 1.2 MB
to the host eagerly.
"""
from __future__ import annotations

import random as _global_random
from collections.abc import Mapping
from dataclasses import dataclass, field

import numpy as np
import torch
from scipy.stats import linregress as _scipy_linregress

# scipy wraps linregress in an axis/nan-policy decorator that costs more than the regression;
# for 1-D finite inputs it forwards unchanged to this inner function (tests/test_host_logic.py).
linregress = getattr(_scipy_linregress, "__wrapped__", _scipy_linregress)



def _rvalue(y) -> np.float64:
    """Pearson r of ``y`` against 0..n-1, exactly as ``scipy.stats.linregress(range(n), y).rvalue``
    computes it (float64; mean-removed dot products through ``np.vecdot``; clip to [-1, 1]; NaN when both
    the covariance and a variance vanish, 0 when only a variance does) without the array-API plumbing
    around it, which costs several times the arithmetic.  tests/test_host_logic.py checks bit equality."""
    y = np.asarray(y).astype(np.float64)
    n = y.shape[0]
    x = np.arange(n, dtype=np.float64)
    x_ = x - np.mean(x, keepdims=True)
    y_ = y - np.mean(y, keepdims=True)
    ssxm = np.vecdot(x_, x_) / n
    ssym = np.vecdot(y_, y_) / n
    ssxym = np.vecdot(x_, y_) / n
    if ssxm == 0.0 or ssym == 0.0:
        return np.float64(np.nan) if ssxym == 0 else np.float64(0.0)
    return np.clip(ssxym / np.sqrt(ssxm * ssym), -1.0, 1.0)


from . import _lib, engine, geometry
from .axis import angle_offset_to_axis, axis_to_angle_offset
from .config import OptConfig
from .structures import Instances as _OwnInstances
from .diagnostics import check_axis, check_monotonic, fit_plane_from_normals  # noqa: F401  (reference names, row a15)

__all__ = ["track_planes", "optimize_planes", "optimize_planes_3dc", "optimize_planes_3d_trans",
           "optimize_planes_3d", "optimize_planes_average", "optimize_videos", "RegMasks"]


# ---------------------------------------------------------------------------
# tracker (host only; reference utils/opt_utils.py:1156-1208)
# ---------------------------------------------------------------------------
def _box_iou_f32(a, b) -> float:
    """IoU of two XYXY boxes with every operation rounded to fp32, in the order
    detectron2's ``pairwise_iou`` evaluates it on 1x1 inputs (area1 + area2 - inter)."""
    f = np.float32
    iw = f(min(a[2], b[2]) - max(a[0], b[0]))
    ih = f(min(a[3], b[3]) - max(a[1], b[1]))
    if not (iw > 0 and ih > 0):
        return 0.0
    inter = f(iw * ih)
    area_a = f(f(a[2] - a[0]) * f(a[3] - a[1]))
    area_b = f(f(b[2] - b[0]) * f(b[3] - b[1]))
    return float(f(inter / f(f(area_a + area_b) - inter)))


def track_planes(preds, cfg: OptConfig | None = None):
    """Greedy online box tracker -> {'rot': [...], 'trans': [...]}; each track is
    {'bbox', 'ids': {frame: box_id}, 'latest_frame'}.  A box joins the FIRST live
    track of its class (class 1 -> 'trans') whose latest box overlaps it with
    IoU > 0.5 and whose gap is <= 5 frames; tracks shorter than 10 frames are
    dropped.  Box IoUs are evaluated on fp32 scalars (same roundings as the
    reference's tensor ops) instead of one tiny tensor program per pair."""
    cfg = cfg or OptConfig()
    planes = {'rot': [], 'trans': []}
    for idx, p_instance in enumerate(preds):
        pred_classes = p_instance.pred_classes
        pred_boxes = p_instance.pred_boxes
        rows = pred_boxes.tensor.detach().cpu().numpy().astype(np.float32, copy=False)
        for box_id in range(rows.shape[0]):
            cur = rows[box_id]
            cat = 'trans' if pred_classes[box_id] == 1 else 'rot'
            matched = False
            for plane in planes[cat]:
                if idx - plane['latest_frame'] > cfg.track_max_gap:
                    continue
                if _box_iou_f32(cur, plane['_row']) > cfg.track_iou:
                    plane['ids'][idx] = box_id
                    plane['bbox'] = pred_boxes[box_id]
                    plane['_row'] = cur
                    plane['latest_frame'] = idx
                    matched = True
                    break
            if not matched:
                planes[cat].append({'bbox': pred_boxes[box_id], 'ids': {idx: box_id}, 'latest_frame': idx,
                                    '_row': cur})
    for cat in planes:
        planes[cat] = [p for p in planes[cat] if len(p['ids']) >= cfg.track_min_len]
        for p in planes[cat]:
            del p['_row']
    return planes


# ---------------------------------------------------------------------------
# lazy reg_masks
# ---------------------------------------------------------------------------
class RegMasks(Mapping):
    """``plane['reg_masks']``: frame -> (H, W) fp32 CPU mask of the best candidate
    (reference utils/opt_utils.py:614, 906), unpacked from the device-resident
    bit-packed copy on first access."""

    def __init__(self, frames, packed: torch.Tensor, H: int, W: int):
        self._frames = list(frames)
        self._row = {f: i for i, f in enumerate(self._frames)}
        self.packed = packed            # (T, H, pitch) int32 on the device
        self._H, self._W = H, W
        self._cache = {}

    def __getitem__(self, frame):
        i = self._row[frame]
        if i not in self._cache:
            idx = torch.tensor([i], dtype=torch.int32, device=self.packed.device)
            self._cache[i] = engine.emit_masks(self.packed, idx, self._H, self._W)[0].cpu()
        return self._cache[i]

    def __iter__(self):
        return iter(self._frames)

    def __len__(self):
        return len(self._frames)

    def dense(self, dtype=torch.float32) -> torch.Tensor:
        """All masks as one (T, H, W) device tensor."""
        return engine.emit_masks(self.packed, None, self._H, self._W, dtype=dtype)


# ---------------------------------------------------------------------------
# the host control flow, as a generator of device jobs
# ---------------------------------------------------------------------------
@dataclass
class JobSpec:
    source: int                    # pool index of the source mask
    mode: int
    normal: np.ndarray
    offset: float
    pivot: np.ndarray
    xform: np.ndarray              # (A, 12) fp32
    targets: list                  # pool indices, in id_list order
    keep_masks: bool = False       # final phase: return the winning packed masks


@dataclass
class JobResult:
    best_cand: np.ndarray          # (n_tgt,) int32
    best_iou: np.ndarray           # (n_tgt,) fp32
    best_inter: np.ndarray
    best_union: np.ndarray
    masks: torch.Tensor | None = None     # (n_tgt, H, pitch) int32 device, if keep_masks


@dataclass
class Stats:
    """What was evaluated, in the reference's accounting: one unit =
    one (visited frame, candidate) IoU."""
    units_visited: int = 0         # units the reference's loops evaluate
    units_computed: int = 0        # units the device evaluated (all of id_list each round)
    passes: int = 0
    jobs: int = 0
    h2d_bytes: int = 0             # masks + job descriptors copied host -> device
    d2h_bytes: int = 0             # per-target results copied device -> host
    detail: list = field(default_factory=list)


def _phase_setup(translation: bool, legacy: bool, cfg: OptConfig):
    if translation:
        return cfg.trans_grid, cfg.trans_grid, _lib.MODE_TRANSLATE, _lib.MODE_TRANSLATE
    if legacy:
        return cfg.legacy_grid, cfg.legacy_grid, _lib.MODE_SEQ, _lib.MODE_SEQ
    return cfg.rot_cluster_grid, cfg.rot_final_grid, _lib.MODE_SEQ, _lib.MODE_COMPOSED


def _candidates(geo: geometry.SourceGeometry, grid, mode: int):
    """(A,12) transforms, the (A,1) fp32 angle tensor the reference indexes, and R."""
    if mode == _lib.MODE_TRANSLATE:
        return geometry.xforms_translate(grid, geo.dir_vec), torch.as_tensor(grid, dtype=torch.float32).unsqueeze(1), None
    R = geometry.rotation_matrices(grid, geo.dir_vec)
    angles = torch.FloatTensor(np.asarray(grid)[:, np.newaxis])
    if mode == _lib.MODE_SEQ:
        return geometry.xforms_seq(R), angles, R
    return geometry.xforms_composed(R, geo.pivot), angles, R


def _tracks_gen(preds, planes, cfg: OptConfig, translation: bool, rng, pool_of, stats: Stats,
                legacy: bool = False):
    """Cluster rounds, model selection and final assignment for a list of tracks
    (reference :386-622 / :689-908 / legacy :113-340).  Mutates ``planes``."""
    cgrid, fgrid, cmode, fmode = _phase_setup(translation, legacy, cfg)
    remove_inliers = not legacy
    finals = []
    for plane in planes:
        ids = plane['ids']
        id_list = list(ids.keys())
        clusters = []
        geo_of = {}                       # source frame -> geometry (the final phase reuses its centre frame's)
        for _ in range(cfg.rounds):
            if len(id_list) == 0 and remove_inliers:
                break
            select_idx = rng.choice(id_list)
            box_id = ids[select_idx]
            geo = geo_of.get(select_idx)
            if geo is None:
                geo = geo_of[select_idx] = geometry.source_geometry(preds[select_idx], box_id, cfg, translation)
            xf, angles, _ = _candidates(geo, cgrid, cmode)
            order = list(id_list)
            res = yield JobSpec(pool_of[(select_idx, box_id)], cmode, geo.normal.numpy(), float(geo.offset),
                                geo.pivot, xf, [pool_of[(i, ids[i])] for i in order])
            row = {f: k for k, f in enumerate(order)}
            # plain python lists: fp32 IoUs are exact as doubles, so `> 0.5` decides as the fp32 compare does
            ious, cands = res.best_iou.tolist(), res.best_cand.tolist()
            thr, n_cand = float(np.float32(cfg.inlier_iou)), len(xf)
            inliers, c_ids, c_ious = [], [], []
            it = 0
            while it < len(id_list):            # `for idx in id_list` with in-loop removal
                idx = id_list[it]
                it += 1
                k = row[idx]
                stats.units_visited += n_cand
                if ious[k] > thr:
                    inliers.append(idx)
                    if remove_inliers:
                        id_list.remove(idx)
                    c_ids.append(cands[k])
                    c_ious.append(ious[k])
            clusters.append({'center_id': select_idx, 'inliners': inliers,
                             'angles': angles[c_ids, 0].clone() if c_ids else torch.FloatTensor([]),
                             'ious': c_ious})

        rsqs = []
        for cluster in clusters:
            if len(cluster['inliners']) < cfg.min_inliers:
                rsqs.append(0.0)
                continue
            rsqs.append(_rvalue(cluster['angles'].numpy()) ** 2)
        rsqs = np.array(rsqs)
        if rsqs.max() < cfg.rsq_thresh:
            plane['has_rot'] = False
            plane['fit'] = {'rsq': rsqs, 'clusters': clusters}
            continue
        plane['has_rot'] = True

        # final assignment: it draws no random numbers, so the final jobs of all tracks of this list
        # are deferred and scored together in ONE device pass after the last cluster round
        final_cluster = clusters[rsqs.argmax()]
        select_idx = final_cluster['center_id']
        box_id = ids[select_idx]
        p_instance = preds[select_idx]
        geo = geo_of[select_idx] if not legacy else \
            geometry.source_geometry(p_instance, box_id, cfg, translation, all_boxes=True)
        xf, angles, R = _candidates(geo, fgrid, fmode)
        frames = list(ids.keys())
        spec = JobSpec(pool_of[(select_idx, box_id)], fmode, geo.normal.numpy(), float(geo.offset),
                       geo.pivot, xf, [pool_of[(i, ids[i])] for i in frames], keep_masks=True)
        finals.append((plane, spec, geo, xf, angles, R, frames, select_idx, box_id, p_instance, rsqs, clusters))

    if not finals:
        return
    results = yield [f[1] for f in finals]
    for (plane, spec, geo, xf, angles, R, frames, select_idx, box_id, p_instance, rsqs, clusters), res in zip(finals, results):
        stats.units_visited += len(xf) * len(frames)
        H, W = cfg.height, cfg.width
        plane['reg_masks'] = RegMasks(frames, res.masks, H, W)
        if fmode == _lib.MODE_COMPOSED:
            normal_trans = geometry.transform_normals(geo.normal, R)
            plane['reg_normals'] = {}
            for k, idx in enumerate(frames):
                n = normal_trans[int(res.best_cand[k])].clone()
                n[1] = -n[1]
                n[[1, 2]] = n[[2, 1]]
                plane['reg_normals'][idx] = n
        if translation:
            plane['std_axis'] = p_instance.pred_tran_axis[box_id]
        elif legacy:
            plane['std_axis'] = geo.pts.clone()
        else:
            plane['std_axis'] = geo.pts[box_id]
        plane['fit'] = {
            'frames': frames, 'center_frame': select_idx, 'rsq': rsqs, 'clusters': clusters,
            'angle_id': res.best_cand.copy(), 'angle': angles[:, 0].numpy()[res.best_cand],
            'inter': res.best_inter.copy(), 'union': res.best_union.copy(), 'iou': res.best_iou.copy(),
        }


# ---------------------------------------------------------------------------
# write-back (host; reference :624-682, :910-959, legacy :342-379)
# ---------------------------------------------------------------------------
def _rebuild(p_instance, scores):
    if type(p_instance) is _OwnInstances:          # same fields, same order, without eight length checks
        f = p_instance.get_fields()
        fields = {"scores": scores, "pred_boxes": f["pred_boxes"], "pred_planes": f["pred_planes"],
                  "pred_rot_axis": f["pred_rot_axis"], "pred_tran_axis": f["pred_tran_axis"]}
        for name in ("pred_masks", "pred_rle"):
            if name in f:
                fields[name] = f[name]
        fields["pred_classes"] = f["pred_classes"]
        out = _OwnInstances.__new__(_OwnInstances)
        object.__setattr__(out, "_image_size", p_instance.image_size)
        object.__setattr__(out, "_fields", fields)
        return out
    out = type(p_instance)(p_instance.image_size)
    out.scores = scores
    out.pred_boxes = p_instance.pred_boxes
    out.pred_planes = p_instance.pred_planes
    out.pred_rot_axis = p_instance.pred_rot_axis
    out.pred_tran_axis = p_instance.pred_tran_axis
    if _has(p_instance, 'pred_masks'):
        out.pred_masks = p_instance.pred_masks
    if _has(p_instance, 'pred_rle'):
        out.pred_rle = p_instance.pred_rle
    out.pred_classes = p_instance.pred_classes
    return out


def _has(p_instance, name: str) -> bool:
    return p_instance.has(name) if hasattr(p_instance, 'has') else hasattr(p_instance, name)


def _write_back(preds, planes, cfg: OptConfig, kind: str):
    """kind: 'rot' | 'trans' | 'legacy'."""
    # rotation tracks: the fitted axis re-expressed relative to every frame's box centre.  The
    # reference does this one (frame, track) at a time (:646-651); the arithmetic is elementwise
    # fp32, so one call per track over all its frames gives the same bits.
    new_axis = {}
    if kind == 'rot':
        for ti, plane in enumerate(planes):
            if not plane['has_rot'] or not plane['ids']:
                continue
            frames = list(plane['ids'].keys())
            centers = torch.stack([preds[f].pred_boxes.tensor[plane['ids'][f]] for f in frames])
            centers = (centers[:, :2] + centers[:, 2:]) / 2
            line = plane['std_axis'].unsqueeze(0).numpy().tolist()
            new_axis[ti] = dict(zip(frames, axis_to_angle_offset(line * len(frames), centers)[:, :3]))
    opt_preds = []
    for idx, p_instance in enumerate(preds):
        pred_boxes = p_instance.pred_boxes
        pred_classes = p_instance.pred_classes
        chosen = [False] * pred_boxes.tensor.shape[0]
        if kind != 'legacy':
            keep_class = 1 if kind == 'rot' else 0          # the other articulation type is never filtered
            for i in range(pred_classes.size):
                if pred_classes[i] == keep_class:
                    chosen[i] = True
        if kind == 'rot':
            p_instance.pred_rot_axis = p_instance.pred_rot_axis.clone()
            p_instance.pred_planes = p_instance.pred_planes.clone()
        for ti, plane in enumerate(planes):
            if idx not in plane['ids']:
                continue
            box_id = plane['ids'][idx]
            if not plane['has_rot']:
                chosen[box_id] = False
                continue
            chosen[box_id] = True
            if kind == 'rot':
                p_instance.pred_rot_axis[box_id] = new_axis[ti][idx]
            elif kind == 'trans':
                p_instance.pred_tran_axis[box_id] = plane['std_axis']     # in place, like the reference
        chosen = np.array(chosen, dtype=bool)
        scores = np.copy(p_instance.scores)
        decay = cfg.legacy_score_decay if kind == 'legacy' else cfg.score_decay
        scores[~chosen] = scores[~chosen] * decay
        opt_preds.append(_rebuild(p_instance, scores))
    return opt_preds


# ---------------------------------------------------------------------------
# device session: mask pool + job driver
# ---------------------------------------------------------------------------
class _Session:
    """Packed masks of every tracked box of a set of videos, resident on one GPU."""

    def __init__(self, videos, cfg: OptConfig, device):
        self.cfg = cfg
        self.device = torch.device(device)
        self.ws = engine.Workspace(self.device)
        self.staging = engine.Staging(self.device)
        self.pool_of = []                   # per video: {(frame, box_id): pool index}
        chunks, rles, n = [], [], 0
        for preds, plane_lists in videos:
            need = {}
            for planes in plane_lists:
                for plane in planes:
                    for f, b in plane['ids'].items():
                        need.setdefault(f, set()).add(b)
            index = {}
            for f in sorted(need):
                boxes = sorted(need[f])
                if _has(preds[f], 'pred_masks'):
                    m = preds[f].pred_masks
                    sel = m if len(boxes) == m.shape[0] else m[boxes]
                    if tuple(sel.shape[1:]) != (cfg.height, cfg.width):
                        raise ValueError(f"mask shape {tuple(sel.shape[1:])} != camera {cfg.height}x{cfg.width}")
                    chunks.append(sel)
                else:                        # run-length masks, decoded on the device
                    rles.extend(preds[f].pred_rle[b] for b in boxes)
                for b in boxes:
                    index[(f, b)] = n
                    n += 1
            self.pool_of.append(index)
        self.h2d_bytes = 0
        if n == 0:
            self.pool = None
            return
        if rles:
            if chunks:
                raise ValueError("mixing dense pred_masks and pred_rle frames is not supported")
            self.pool = engine.rle_to_pool(rles, cfg.height, cfg.width, self.device)
            self.h2d_bytes += sum(len(r["counts"]) for r in rles)
            return
        dev_chunks = []
        for c in chunks:
            if c.dtype not in (torch.float32, torch.uint8, torch.bool):
                c = c.float()
            if not c.is_cuda:
                self.h2d_bytes += c.numel() * c.element_size()
            dev_chunks.append(c.to(self.device, non_blocking=True))
        dt = dev_chunks[0].dtype
        dense = torch.cat([c if c.dtype == dt else c.to(dt) for c in dev_chunks])
        self.pool = engine.pack_masks(dense, cfg.mask_thresh, with_nonzero=(dt == torch.float32))
        del dense, dev_chunks

    def run(self, specs, stats: Stats | None = None):
        """One device pass over a list of JobSpec -> list of JobResult."""
        batch = engine.build_batch([s.source for s in specs], [s.mode for s in specs],
                                   [s.normal for s in specs], [s.offset for s in specs],
                                   [s.pivot for s in specs], [s.xform for s in specs],
                                   [s.targets for s in specs], self.pool.source_points)
        dbatch = engine.DeviceBatch(batch, self.device, self.staging, self.cfg)
        res = engine.run_pass(self.cfg, self.pool, dbatch, self.ws)
        host = self.staging.results_host(res.block.numel())
        host.view(4, -1).copy_(res.block, non_blocking=True)                 # one D2H into pinned memory
        torch.cuda.current_stream().synchronize()
        packed = host.numpy().reshape(4, -1)
        if stats is not None:
            stats.h2d_bytes += batch.jobs.nbytes + batch.xform.nbytes + batch.tgt_index.nbytes
            stats.d2h_bytes += packed.nbytes
        out = []
        for j, s in enumerate(specs):
            a, n = int(batch.jobs[j]["tgt_begin"]), int(batch.jobs[j]["n_tgt"])
            r = JobResult(packed[0, a:a + n].copy(), packed[3, a:a + n].copy().view(np.float32),
                          packed[1, a:a + n].copy(), packed[2, a:a + n].copy())
            if s.keep_masks:
                gidx = (res.best_cand[a:a + n].long() + int(batch.jobs[j]["cand_begin"]))
                r.masks = res.proj_bits.index_select(0, gidx)       # copy out of the workspace
            out.append(r)
        return out, batch.units


def _drive(gens, session: _Session, stats: Stats):
    """Run generators in lock-step: every device pass carries the pending job of
    each still-active generator."""
    pending = {}
    for i, g in enumerate(gens):
        try:
            pending[i] = next(g)
        except StopIteration:
            pass
    while pending:
        keys = list(pending.keys())
        specs, spans = [], []
        for k in keys:                       # a request is one JobSpec or a list of them
            req = pending[k]
            group = req if isinstance(req, list) else [req]
            spans.append((len(specs), len(group), isinstance(req, list)))
            specs.extend(group)
        results, units = session.run(specs, stats)
        stats.passes += 1
        stats.jobs += len(specs)
        stats.units_computed += units
        for k, (lo, n, is_list) in zip(keys, spans):
            try:
                pending[k] = gens[k].send(results[lo:lo + n] if is_list else results[lo])
            except StopIteration:
                del pending[k]


def _default_device(device):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise _lib.A3DError("articulation3d_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------------------
# public API (reference signatures)
# ---------------------------------------------------------------------------
def optimize_planes_3d_trans(preds, planes, frames=None, cfg=None, device=None, rng=None,
                             _session=None, stats=None):
    """Translation tracks (reference utils/opt_utils.py:685-959)."""
    cfg, stats = cfg or OptConfig(), stats or Stats()
    session = _session or _Session([(preds, [planes])], cfg, _default_device(device))
    _drive([_tracks_gen(preds, planes, cfg, True, rng or _global_random, session.pool_of[0], stats)],
           session, stats)
    return _write_back(preds, planes, cfg, 'trans')


def optimize_planes_3dc(preds, planes, frames=None, cfg=None, device=None, rng=None,
                        _session=None, stats=None):
    """Rotation tracks with 3-D clustering (reference utils/opt_utils.py:382-682)."""
    cfg, stats = cfg or OptConfig(), stats or Stats()
    session = _session or _Session([(preds, [planes])], cfg, _default_device(device))
    _drive([_tracks_gen(preds, planes, cfg, False, rng or _global_random, session.pool_of[0], stats)],
           session, stats)
    return _write_back(preds, planes, cfg, 'rot')


def optimize_planes_3d(preds, planes, cfg=None, device=None, rng=None, stats=None):
    """Legacy method '3d' (reference utils/opt_utils.py:112-379)."""
    cfg, stats = cfg or OptConfig(), stats or Stats()
    session = _Session([(preds, [planes])], cfg, _default_device(device))
    _drive([_tracks_gen(preds, planes, cfg, False, rng or _global_random, session.pool_of[0], stats,
                        legacy=True)], session, stats)
    return _write_back(preds, planes, cfg, 'legacy')


def optimize_planes_average(preds, planes):
    """Method 'average' (reference utils/opt_utils.py:75-109); host only."""
    for plane in planes:
        std_axes = []
        for idx in plane['ids']:
            box_id = plane['ids'][idx]
            p_instance = preds[idx]
            pts = angle_offset_to_axis(p_instance.pred_rot_axis, p_instance.pred_boxes.get_centers())
            img_centers = torch.FloatTensor(np.array([[320, 240]]))
            std_axis = axis_to_angle_offset(pts.numpy().tolist(), img_centers)[:, :3]
            std_axes.append(std_axis[box_id:(box_id + 1)])
        plane['std_axis'] = torch.cat(std_axes).mean(axis=0)
    for idx, p_instance in enumerate(preds):
        for plane in planes:
            if idx in plane['ids']:
                p_instance.pred_rot_axis[plane['ids'][idx]] = plane['std_axis']
    return list(preds)


def optimize_planes(preds, planes, method, frames=None, cfg=None, device=None, stats=None):
    """Reference dispatcher (utils/opt_utils.py:962-974).  ``planes`` is the dict
    ``track_planes`` returns for '3dc', a list of tracks for 'average' / '3d'."""
    if method == 'average':
        return optimize_planes_average(preds, planes)
    elif method == '3d':
        return optimize_planes_3d(preds, planes, cfg=cfg, device=device, stats=stats)
    elif method == '3dc':
        cfg = cfg or OptConfig()
        stats = stats if stats is not None else Stats()
        session = _Session([(preds, [planes['trans'], planes['rot']])], cfg, _default_device(device))
        stats.h2d_bytes += session.h2d_bytes
        opt_preds = optimize_planes_3d_trans(preds, planes['trans'], frames=frames, cfg=cfg,
                                             _session=session, stats=stats)
        return optimize_planes_3dc(opt_preds, planes['rot'], frames=frames, cfg=cfg,
                                   _session=session, stats=stats)
    else:
        raise NotImplementedError


def optimize_videos(videos, seeds, cfg=None, device=None, stats=None):
    """Batched '3dc' over independent videos: ``videos`` is a list of
    ``(preds, planes)``; video i draws its source frames from
    ``random.Random(seeds[i])`` (the reference seeds one process per video,
    tools/inference.py:172), so the result of every video equals
    ``random.seed(seeds[i]); optimize_planes(preds, planes, '3dc')``.  All videos
    advance in lock-step: one device pass per chain step carries one job per
    video."""
    cfg = cfg or OptConfig()
    stats = stats if stats is not None else Stats()
    session = _Session([(p, [pl['trans'], pl['rot']]) for p, pl in videos], cfg, _default_device(device))
    stats.h2d_bytes += session.h2d_bytes
    rngs = [_global_random.Random(s) for s in seeds]

    def video_gen(i):
        preds, planes = videos[i]
        yield from _tracks_gen(preds, planes['trans'], cfg, True, rngs[i], session.pool_of[i], stats)
        mid = _write_back(preds, planes['trans'], cfg, 'trans')
        outs[i] = mid
        yield from _tracks_gen(mid, planes['rot'], cfg, False, rngs[i], session.pool_of[i], stats)
        outs[i] = _write_back(mid, planes['rot'], cfg, 'rot')

    outs = [None] * len(videos)
    _drive([video_gen(i) for i in range(len(videos))], session, stats)
    return outs
