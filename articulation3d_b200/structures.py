"""Light stand-ins for the two detectron2 types the temporal optimizer consumes.

The reference takes ``list[detectron2.structures.Instances]`` (one per frame,
built by ``create_instances``, reference utils/arti_vis.py:152-194) and only
touches: attribute get/set of fields, ``image_size``, ``pred_boxes.tensor``,
``pred_boxes.get_centers()``, ``pred_boxes[i]`` and ``pairwise_iou``
(reference utils/opt_utils.py:406,540,646,668,1168,1180).  detectron2 is not a
dependency of this package; real detectron2 objects work unchanged because only
that duck-typed surface is used.
"""
from __future__ import annotations

import torch


class Boxes:
    """XYXY fp32 boxes, (N, 4)."""

    def __init__(self, tensor):
        device = tensor.device if isinstance(tensor, torch.Tensor) else torch.device("cpu")
        tensor = torch.as_tensor(tensor, dtype=torch.float32, device=device)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4))
        if tensor.dim() != 2 or tensor.size(-1) != 4:
            raise ValueError(f"Boxes expects (N, 4), got {tuple(tensor.shape)}")
        self.tensor = tensor

    def area(self) -> torch.Tensor:
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def get_centers(self) -> torch.Tensor:
        return (self.tensor[:, :2] + self.tensor[:, 2:]) / 2

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self) -> int:
        return self.tensor.shape[0]

    def __repr__(self) -> str:
        return f"Boxes({self.tensor!r})"


def pairwise_iou(boxes1: Boxes, boxes2: Boxes) -> torch.Tensor:
    """IoU matrix (N, M); 0 where the boxes do not overlap."""
    a1, a2 = boxes1.area(), boxes2.area()
    b1, b2 = boxes1.tensor, boxes2.tensor
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter),
                       torch.zeros(1, dtype=inter.dtype, device=inter.device))


class Instances:
    """Per-frame container of equally long fields (scores, pred_boxes, pred_classes,
    pred_planes, pred_rot_axis, pred_tran_axis, pred_masks)."""

    def __init__(self, image_size, **kwargs):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return self._fields[name]

    def set(self, name, value):
        n = len(value)
        if len(self._fields) and len(self) != n:
            raise AssertionError(f"Adding a field of length {n} to Instances of length {len(self)}")
        self._fields[name] = value

    def has(self, name) -> bool:
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __len__(self) -> int:
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __repr__(self) -> str:
        return f"Instances(n={len(self) if self._fields else 0}, fields={list(self._fields)})"
