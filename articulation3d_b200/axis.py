"""Axis parametrisation used on the hot path (host side, per source frame).

``(sin, cos, offset/100)`` relative to a box centre  <->  integer end-points
``[x1, y1, x2, y2]`` on the image border.  Mirrors the behaviour of the
reference's ``angle_offset_to_axis`` / ``get_boundary_point`` /
``axis_to_angle_offset`` (data/planercnn_transforms.py:101-176, 31-68):
all scalar work is IEEE fp32 (the reference holds 0-dim fp32 tensors and runs
numpy's fp32 ``arctan``/``tan`` on them), border hits are probed in the order
left, right, top, bottom with truncation toward zero, and a line that misses
the image falls back to ``[0, 0, 1, 1]``.
"""
from __future__ import annotations

import numpy as np
import torch

_F = np.float32
_NEG_HALF_PI = _F(-np.pi / 2)
_INT64_MIN = -(2 ** 63)


def _long(v) -> int:
    """fp32 -> int64 the way ``Tensor.long()`` does on x86: truncation; NaN, inf
    and out-of-range give INT64_MIN."""
    v = float(v)
    if not (abs(v) < 9.223372036854775807e18):
        return _INT64_MIN
    return int(v)


def _border_hits(y, x, angle, H, W):
    """Up to two distinct border points of the line through (x, y) with slope
    tan(angle); None when the line misses the image."""
    if angle == _NEG_HALF_PI:                       # vertical line
        return (x, 0), (x, H - 1)
    if angle == 0.0:                                # horizontal line
        return (0, y), (W - 1, y)
    k = np.tan(_F(angle))
    b = y - k * x                                   # intercept at x = 0
    top = x - y / k                                 # x at y = 0
    probes = (
        (b, H, lambda v: (0, int(v))),                                  # left
        (k * _F(W - 1) + y - k * x, H, lambda v: (W - 1, int(v))),      # right
        (top, W, lambda v: (int(v), 0)),                                # top
        (top + _F(H - 1) / k, W, lambda v: (int(v), H - 1)),            # bottom
    )
    first = second = None
    for value, limit, make in probes:
        if not (value >= 0 and value < limit):
            continue
        if first is None:
            first = make(value)
        elif second is None:
            cand = make(value)
            if cand != first:
                second = cand
    if second is None:
        second = first
    if first is None:
        return None
    return first, second


def angle_offset_to_axis(angle_offsets, centers, H: int = 480, W: int = 640) -> torch.Tensor:
    """(n,3) [sin, cos, offset] + (n,2) centres -> (n,4) int64 [x1,y1,x2,y2]."""
    ao = torch.as_tensor(angle_offsets).detach().cpu().numpy().astype(np.float32).reshape(-1, 3)
    ce = torch.as_tensor(centers).detach().cpu().numpy().astype(np.float32).reshape(-1, 2)
    out = np.empty((len(ao), 4), dtype=np.int64)
    with np.errstate(all="ignore"):
        for i in range(len(ao)):
            s, c, p = ao[i]
            x0, y0 = ce[i]
            p = _F(p * _F(100))
            angle = _NEG_HALF_PI if s == 0 else _F(-np.arctan(_F(c / s)))
            x = _F(_F(p * c) + x0)
            y = _F(_F(p * s) + y0)
            hits = _border_hits(y, x, angle, H, W)
            if hits is None:
                out[i] = (0, 0, 1, 1)
            else:
                (ax, ay), (bx, by) = hits
                out[i] = (_long(ax), _long(ay), _long(bx), _long(by))
    return torch.from_numpy(out)


def _long_np(v: np.ndarray) -> np.ndarray:
    """Vector form of ``_long``."""
    v = np.asarray(v, dtype=np.float64)
    ok = np.abs(v) < 9.223372036854775807e18
    out = np.full(v.shape, _INT64_MIN, dtype=np.int64)
    out[ok] = np.trunc(v[ok]).astype(np.int64)
    return out


def angle_offset_to_axis_rows(ao: np.ndarray, ce: np.ndarray, H: int = 480, W: int = 640) -> np.ndarray:
    """``angle_offset_to_axis`` for many independent (line, centre) rows at once: the same fp32
    operations in the same order, as numpy array arithmetic (IEEE per element, so the bits equal the
    scalar loop's; tests/test_host_logic.py compares the two on random and edge rows).
    ao (n,3) fp32 [sin, cos, offset], ce (n,2) fp32 -> (n,4) int64."""
    ao = np.ascontiguousarray(ao, dtype=np.float32).reshape(-1, 3)
    ce = np.ascontiguousarray(ce, dtype=np.float32).reshape(-1, 2)
    n = len(ao)
    out = np.empty((n, 4), dtype=np.int64)
    if n == 0:
        return out
    with np.errstate(all="ignore"):
        s, c, p = ao[:, 0], ao[:, 1], ao[:, 2]
        x0, y0 = ce[:, 0], ce[:, 1]
        p = p * _F(100)
        angle = np.where(s == 0, _NEG_HALF_PI, -np.arctan(c / s)).astype(np.float32)
        x = p * c + x0
        y = p * s + y0
        vert = angle == _NEG_HALF_PI
        horiz = (angle == 0.0) & ~vert
        k = np.tan(angle)
        b = y - k * x
        top = x - y / k
        v = (b, k * _F(W - 1) + y - k * x, top, top + _F(H - 1) / k)
        lim = (H, H, W, W)
        have1 = np.zeros(n, dtype=bool)
        have2 = np.zeros(n, dtype=bool)
        p1 = np.zeros((n, 2), dtype=np.int64)
        p2 = np.zeros((n, 2), dtype=np.int64)
        for i in range(4):
            ok = (v[i] >= 0) & (v[i] < lim[i])
            t = np.zeros(n, dtype=np.int64)
            t[ok] = np.trunc(v[i][ok]).astype(np.int64)
            if i == 0:
                px, py = np.zeros(n, dtype=np.int64), t
            elif i == 1:
                px, py = np.full(n, W - 1, dtype=np.int64), t
            elif i == 2:
                px, py = t, np.zeros(n, dtype=np.int64)
            else:
                px, py = t, np.full(n, H - 1, dtype=np.int64)
            take2 = ok & have1 & ~have2 & ((px != p1[:, 0]) | (py != p1[:, 1]))
            take1 = ok & ~have1
            p1[take1, 0], p1[take1, 1] = px[take1], py[take1]
            p2[take2, 0], p2[take2, 1] = px[take2], py[take2]
            have1 |= take1
            have2 |= take2
        p2[~have2] = p1[~have2]
        out[:, :2], out[:, 2:] = p1, p2
        out[~have1] = (0, 0, 1, 1)
        if vert.any():
            xv = _long_np(x[vert])
            out[vert] = np.stack([xv, np.zeros_like(xv), xv, np.full_like(xv, H - 1)], 1)
        if horiz.any():
            yh = _long_np(y[horiz])
            out[horiz] = np.stack([np.zeros_like(yh), yh, np.full_like(yh, W - 1), yh], 1)
    return out


def axis_to_angle_offset(axis, center: torch.Tensor) -> torch.Tensor:
    """[[x1,y1,x2,y2] | None, ...] + (n,2) centres -> (n,4) fp32
    [sin, cos, offset/100, valid] of the line relative to the centre."""
    rows, valid = [], []
    for a in axis:
        rows.append([0, 0, 1, 1] if a is None else list(a))
        valid.append([0.0 if a is None else 1.0])
    pts = torch.tensor(rows, dtype=torch.float32) - torch.cat((center, center), dim=1)
    valid = torch.tensor(valid, dtype=torch.float32)
    x1, y1, x2, y2 = pts[:, 0:1], pts[:, 1:2], pts[:, 2:3], pts[:, 3:4]
    A = y1 - y2
    B = x2 - x1
    C = x1 * y2 - x2 * y1
    length = torch.sqrt(A * A + B * B)
    offset = torch.abs(C) / length / 100
    sgn = torch.sign(C)
    cos = -A * sgn / length
    sin = -B * sgn / length
    return torch.cat((sin, cos, offset, valid), dim=1)
