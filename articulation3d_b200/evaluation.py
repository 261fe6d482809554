"""Detection metrics of the temporal stage's output: the AP numbers
``tools/opt_arti.py`` prints before / after optimisation.

Host-only.  Mirrors ``evaluate_for_arti_axis`` and ``evaluate_for_recognition`` of the reference
(evaluation/arti_evaluation.py:262-665, :669-757) and ``VOCap.compute_ap`` (utils/VOCap.py) — the same
matching rules, including the ones that look accidental (noted inline) — without pycocotools or
detectron2: ground truth is a plain COCO-style dict wrapped by ``CocoGT``.  Pinned by
tests/golden/eval/*.json, outputs of the reference itself under oracle/ref_shim.py.

Prediction records are the ``instances_predictions.pth`` dicts (``io.preds_to_records`` /
``io.opt_preds_to_records``): ``{image_id, instances: [{bbox XYWH, score, category_id}], pred_plane (n,3),
pred_rot_axis (n,3), pred_tran_axis (n,2)}`` with ``category_id`` the contiguous class index.
Ground-truth annotations carry ``bbox`` XYWH, ``category_id`` (dataset id), ``rot_axis`` / ``tran_axis``
(``[x1, y1, x2, y2]`` or None) and optionally ``normal``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

from .axis import angle_offset_to_axis, axis_to_angle_offset
from .diagnostics import EA_metric, Line
from .structures import Boxes, pairwise_iou

METRICS = ("bbox", "bbox+axis", "bbox+normal", "bbox+normal+axis")


class CocoGT:
    """The five pycocotools.COCO calls the evaluators make, over a dict
    ``{'images': [...], 'annotations': [...], 'categories': [...]}``."""

    def __init__(self, dataset: dict):
        self.dataset = dataset
        self._anns = {a["id"]: a for a in dataset["annotations"]}
        self._by_img = {}
        for a in dataset["annotations"]:
            self._by_img.setdefault(a["image_id"], []).append(a["id"])
        self._cats = {c["id"]: c for c in dataset["categories"]}
        self._imgs = {i["id"]: i for i in dataset.get("images", [])}

    def getCatIds(self):
        return list(self._cats.keys())

    def getAnnIds(self, imgIds):
        return [i for img in imgIds for i in self._by_img.get(img, [])]

    def loadAnns(self, ids):
        return [self._anns[i] for i in ids]

    def loadCats(self, ids):
        return [self._cats[i] for i in ids]

    def loadImgs(self, ids):
        return [self._imgs[i] for i in ids]


@dataclass
class Metadata:
    """The two fields of detectron2's MetadataCatalog entry the evaluators read."""
    thing_classes: list = field(default_factory=lambda: ["arti_rot", "arti_tran"])
    thing_dataset_id_to_contiguous_id: dict = field(default_factory=lambda: {1: 0, 2: 1})


def compute_ap(scores: torch.Tensor, labels: torch.Tensor, npos: float):
    """VOC average precision, all-point interpolation: area under the monotone envelope of the
    precision/recall curve of the score-ranked list.  0.0 for an empty list."""
    if len(scores) == 0:
        return 0.0
    order = torch.sort(scores, descending=True)[1]
    tp = torch.cumsum((labels == 1)[order].to(torch.float32), dim=0)
    fp = torch.cumsum((labels == 0)[order].to(torch.float32), dim=0)
    rec = tp / npos
    prec = tp / (fp + tp)
    zero, one = rec.new_zeros(1), rec.new_ones(1)
    mrec = torch.cat((zero, rec, one))
    mpre = torch.cat((zero, prec, zero))
    mpre = torch.flip(torch.cummax(torch.flip(mpre, [0]), 0)[0], [0])       # running max from the right
    step = (mrec[1:] != mrec[:-1]).nonzero()[:, 0] + 1
    ap = mrec.new_zeros(())
    for i in step.tolist():                       # summed in rank order, as the reference's loop does
        ap = ap + (mrec[i] - mrec[i - 1]) * mpre[i]
    return ap


def _xywh_to_xyxy(b) -> np.ndarray:
    b = np.array(b, dtype=np.float64).reshape(-1, 4)
    b[:, 2] += b[:, 0]
    b[:, 3] += b[:, 1]
    return b


def _ea_matrix(pred_coord: torch.Tensor, gt_coord: torch.Tensor, on_degenerate) -> np.ndarray:
    out = np.zeros((len(pred_coord), len(gt_coord)))
    pc, gc = pred_coord.tolist(), gt_coord.tolist()
    for p in range(len(pc)):
        for g in range(len(gc)):
            a = pc[p]
            if a[0] == a[2] and a[1] == a[3]:
                on_degenerate(p, g)
                continue
            b = gc[g]
            out[p][g] = EA_metric(Line([a[1], a[0], a[3], a[2]]), Line([b[1], b[0], b[3], b[2]]))
    return out


def evaluate_for_arti_axis(predictions, dataset: CocoGT, metadata: Metadata, filter_iou, iou_thresh=0.5,
                           normal_threshold=30, offset_threshold=100, device=None) -> dict:
    """Per category and per criterion in METRICS: AP of the detections, a detection being a true positive
    when its class matches the ground truth box it overlaps most, box IoU > ``iou_thresh``, that box is not
    yet claimed under the criterion, and — for the '+axis' / '+normal' criteria — the EA score of the
    predicted axis against the ground-truth axis exceeds ``iou_thresh`` / the normal error is below
    ``normal_threshold`` degrees.  Keys: ``'<criterion> - <category name>'``."""
    cat_ids = sorted(dataset.getCatIds())
    to_dataset_id = {v: k for k, v in metadata.thing_dataset_id_to_contiguous_id.items()}
    ap_scores = {m: {c: [torch.zeros(0, dtype=torch.float32)] for c in cat_ids} for m in METRICS}
    ap_labels = {m: {c: [torch.zeros(0, dtype=torch.uint8)] for c in cat_ids} for m in METRICS}
    npos = {c: 0.0 for c in cat_ids}
    for ann in dataset.dataset["annotations"]:
        npos[ann["category_id"]] += 1.0

    for prediction in predictions:
        if "instances" not in prediction or len(prediction["instances"]) == 0:
            continue
        n_pred = len(prediction["instances"])
        scores = [ins["score"] for ins in prediction["instances"]]
        labels = [ins["category_id"] for ins in prediction["instances"]]
        boxes = Boxes(torch.tensor(_xywh_to_xyxy([ins["bbox"] for ins in prediction["instances"]])))
        axis_rot, axis_tran = prediction["pred_rot_axis"], prediction["pred_tran_axis"]
        try:
            pred_normals = F.normalize(prediction["pred_plane"], p=2)
        except Exception:
            pred_normals = F.normalize(torch.ones(n_pred, 3), p=2)
        pred_normals[:, [1, 2]] = pred_normals[:, [2, 1]]          # detector frame -> camera frame
        pred_normals[:, 1] = -pred_normals[:, 1]

        gt_anns = dataset.loadAnns(dataset.getAnnIds(imgIds=[prediction["image_id"]]))
        if len(gt_anns) == 0:
            continue
        gt_labels = [a["category_id"] for a in gt_anns]
        gt_boxes = _xywh_to_xyxy([a["bbox"] for a in gt_anns])
        gt_normals = torch.FloatTensor([a["normal"] if a.get("normal") is not None else [-1, -1, -1] for a in gt_anns])
        gt_normals[:, 1] = -gt_normals[:, 1]
        gt_centers = Boxes(gt_boxes).get_centers()
        gt_rot_ao = axis_to_angle_offset([a["rot_axis"] for a in gt_anns], gt_centers)
        gt_tran_ao = axis_to_angle_offset([a["tran_axis"] for a in gt_anns], gt_centers)
        valid_gt_rot, valid_gt_tran = gt_rot_ao[:, 3].ge(0.5), gt_tran_ao[:, 3].ge(0.5)
        gt_rot_coord = angle_offset_to_axis(gt_rot_ao[:, :3], gt_centers)
        gt_tran_ao[:, 2] = 0
        gt_tran_coord = angle_offset_to_axis(gt_tran_ao[:, :3], gt_centers)

        centers = boxes.get_centers()
        pred_rot_coord = angle_offset_to_axis(axis_rot, centers)
        pred_tran_coord = angle_offset_to_axis(torch.cat((axis_tran, torch.zeros(len(axis_tran), 1)), 1), centers)

        axis_rot_metrics = _ea_matrix(pred_rot_coord, gt_rot_coord, lambda p, g: None)

        def zero_rot_entry(p, g):          # the reference clears the ROTATION entry for a degenerate translation line
            axis_rot_metrics[p][g] = 0
        axis_tran_metrics = _ea_matrix(pred_tran_coord, gt_tran_coord, zero_rot_entry)

        boxiou = pairwise_iou(boxes, Boxes(torch.tensor(gt_boxes, dtype=torch.float32)))
        valid_pred = boxiou > filter_iou
        scores_t = torch.tensor(np.array(scores), dtype=torch.float32)
        order = torch.sort(scores_t, descending=True)[1]
        covered = {m: [] for m in METRICS}
        for rank in range(n_pred):
            idx = int(order[rank])
            # the reference tests `valid_pred[idx] == 0` as a scalar, which only works with one ground-truth
            # box per image (its dataset); with several, a detection is kept when it passes for any of them
            if not bool(valid_pred[idx].any()):
                continue
            gt_id = int(torch.argmax(boxiou[idx]))
            gt_label = gt_labels[gt_id]
            pred_label = to_dataset_id[labels[idx]]
            pred_biou = float(boxiou[idx, gt_id])
            gt_class = metadata.thing_classes[metadata.thing_dataset_id_to_contiguous_id[gt_label]]
            if "rot" in gt_class:
                pred_ea = float(axis_rot_metrics[idx, gt_id]) if bool(valid_gt_rot[gt_id]) else 0
            elif "tran" in gt_class:
                pred_ea = float(axis_tran_metrics[idx, gt_id]) if bool(valid_gt_tran[gt_id]) else 0
            else:
                raise NotImplementedError(gt_class)
            # the normal is taken at the RANK, not at the detection the rank points to (as the reference does)
            normal_error = float(torch.acos(torch.dot(pred_normals[rank], gt_normals[gt_id]))) / np.pi * 180.0
            if float(torch.norm(gt_normals[gt_id])) > 1.1:          # no ground-truth normal
                normal_error = 180.0
            for m in METRICS:
                is_tp = pred_label == gt_label and pred_biou > iou_thresh and gt_id not in covered[m]
                if "axis" in m:
                    is_tp = is_tp and pred_ea > iou_thresh
                if "normal" in m:
                    is_tp = is_tp and normal_error < normal_threshold
                if is_tp:
                    covered[m].append(gt_id)
                ap_scores[m][pred_label].append(scores_t[idx].view(1))
                ap_labels[m][pred_label].append(torch.tensor([1 if is_tp else 0], dtype=torch.uint8))

    out = {}
    for c in cat_ids:
        if npos[c] == 0:
            continue
        name = dataset.loadCats([c])[0]["name"]
        for m in METRICS:
            out[f"{m} - {name}"] = compute_ap(torch.cat(ap_scores[m][c]), torch.cat(ap_labels[m][c]), npos[c])
    return out


def evaluate_for_recognition(predictions, dataset: CocoGT, metadata: Metadata = None, filter_iou=None, **_) -> dict:
    """Image-level "is anything articulated here": AUROC of the top detection score against "the image has
    ground truth", and the accuracy of thresholding that score at 0.95; -1 when undefined."""
    from sklearn.metrics import roc_auc_score
    preds, gts = [], []
    for prediction in predictions:
        scores = [ins["score"] for ins in prediction["instances"]]
        preds.append(np.array(scores).max() if len(scores) > 0 else 0)
        gts.append(len(dataset.getAnnIds(imgIds=[prediction["image_id"]])) > 0)
    preds, gts = np.array(preds), np.array(gts)
    try:
        return {"auroc": roc_auc_score(gts, preds),
                "accuracy": ((preds > 0.95) == gts).sum() / (preds == preds).sum()}
    except Exception:
        return {"auroc": -1, "accuracy": -1}
