"""Textured Wavefront .obj of one frame's articulation: the optional output of
``tools/inference.py --save-obj``.

Host-only.  Follows ``save_obj_model`` (reference tools/inference.py:44-168), ``get_single_image_mesh_arti``
(utils/vis.py:256-378) and ``save_obj`` / ``_save`` (utils/mesh_utils.py:126-266): the most confident box of
the frame becomes a planar mesh textured with the rectified image, followed by its copies rotated about the
predicted axis by ``arange(-1.8, 0.1, 0.45)`` rad (``axis_dir='l'``; ``arange(0, 1.8, 0.45)`` for ``'r'``),
two small icosahedra at the axis end-points and the background plane (the complement of the mask); every
mesh gets its own 300x300 texture, tinted as the reference tints them; vertices are written with ten
decimals, faces double-sided, one material per texture in ``<prefix>.mtl``.

What is the reference's and what is not.  Geometry, file layout, tints and the quirks noted inline are the
reference's.  Two third-party pieces are absent from this image and from /root/reference and are replaced
[3P-unverified]: ``skimage.measure.find_contours`` (mask outline; here ``cv2.findContours`` on pixel
centres: the outline runs half a pixel inside skimage's) and ``mapbox_earcut`` (here a plain ear-clipping
triangulation: the same polygon, a different but equally valid set of triangles).  pytorch3d's ``Meshes`` /
``ico_sphere`` are replaced by arrays and the 12-vertex icosahedron.

Parity: tests/test_export.py compares every written file — .obj, .mtl, the 300x300 textures — byte for byte
with what the reference's OWN ``save_obj_model`` / ``get_single_image_mesh_arti`` / ``save_obj`` write for the
same predictions and frame (oracle/ref_export.py executes them unmodified under the import shim, with the
outline and triangulation routines above bound in place of skimage / earcut and small stand-ins for the
pytorch3d containers; fixtures tests/golden/export/*.json, plus a live comparison where /root/reference
exists).  That pins the plane conversion, the mesh camera, the rectifying homography and its textures, uv
coordinates, winding, rotated copies, axis markers, tints, mesh order and the file text; which outline and
which triangulation skimage / earcut would have chosen stays [3P-unverified].
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

from .axis import angle_offset_to_axis
from .config import OptConfig
from .geometry import rotation_matrices

TARGET = 300                     # side of the rectified texture (utils/vis.py:313)
OBJ_TINT = np.array([252 / 255, 116 / 255, 81 / 255])      # tools/inference.py:150
AXIS_TINT = np.array([56 / 255, 207 / 255, 252 / 255])     # tools/inference.py:154-157


def _cv2():
    import cv2
    return cv2


# ---------------------------------------------------------------------------------------------------
# camera of the mesh code
# ---------------------------------------------------------------------------------------------------
def _mesh_K(cfg: OptConfig, focal_length: float = 571.623718) -> np.ndarray:
    """The reference calls ``get_pcd(verts, normal, offset, focal_length)`` and ``project2D(pts,
    focal_length)`` with the focal length in the position of ``h`` (utils/vis.py:300,312,334): the camera of
    the mesh code is ``f = 517.97, cx = w/2, cy = 571.623718/2``.  Reproduced as is."""
    return np.array([[cfg.focal_length, 0.0, cfg.width / 2], [0.0, cfg.focal_length, focal_length / 2], [0.0, 0.0, 1.0]])


def _get_pcd(verts_xy, normal, offset, K) -> np.ndarray:
    """utils/vis.py:86-102 in float64 with an explicit K."""
    v = np.asarray(verts_xy, dtype=np.float64).reshape(-1, 2)
    ray = np.linalg.inv(K) @ np.hstack((v, np.ones((len(v), 1)))).T
    depth = offset / np.dot(normal, ray)                # fp32 normal / offset promote to float64 here
    return depth.reshape(-1, 1) * ray.T


def _project2D(pcd, K) -> np.ndarray:
    p = (K @ np.asarray(pcd, dtype=np.float64).T).T
    return p[:, :2] / p[:, 2:3]


# ---------------------------------------------------------------------------------------------------
# mask -> rings -> triangles
# ---------------------------------------------------------------------------------------------------
def mask_to_polygons(mask) -> list:
    """Closed rings ``[(x, y), ...]`` of a binary mask, outer boundaries and holes alike (the reference
    triangulates every ring on its own, holes included: utils/vis.py:329-350).  Rings with fewer than
    three points are dropped (pycococreatortools.py:52)."""
    cv2 = _cv2()
    m = (np.asarray(mask) > 0.5).astype(np.uint8)
    contours, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    rings = []
    for c in contours:
        ring = c.reshape(-1, 2).astype(np.float64)
        if len(ring) >= 3:
            rings.append(ring)
    return rings


def _area2(p) -> float:
    x, y = p[:, 0], p[:, 1]
    return float(np.dot(x, np.roll(y, -1)) - np.dot(np.roll(x, -1), y))


def triangulate(ring) -> np.ndarray:
    """Ear clipping of one simple ring -> (n_tri, 3) indices into ``ring`` (counter-clockwise in image
    coordinates).  Collinear and repeated points are kept as vertices (every triangle that would have zero
    area is skipped), so the vertex list is the ring itself, as with earcut."""
    pts = np.asarray(ring, dtype=np.float64)
    n = len(pts)
    if n < 3:
        return np.zeros((0, 3), np.int64)
    idx = list(range(n))
    if _area2(pts) < 0:
        idx.reverse()
    tris = []
    guard = 0
    while len(idx) > 3 and guard < 4 * n * n:
        m = len(idx)
        clipped = False
        P = pts[idx]
        prev, nxt = np.roll(P, 1, axis=0), np.roll(P, -1, axis=0)
        cross = (P[:, 0] - prev[:, 0]) * (nxt[:, 1] - prev[:, 1]) - (P[:, 1] - prev[:, 1]) * (nxt[:, 0] - prev[:, 0])
        for i in np.argsort(-(cross > 0).astype(np.int8), kind="stable"):
            guard += 1
            if cross[i] < 0:
                break                                         # only reflex corners left: degenerate ring
            a, b, c = P[i - 1], P[i], P[(i + 1) % m]
            if cross[i] == 0:                                 # collinear point: drop it without a triangle
                del idx[i]
                clipped = True
                break
            # no other vertex strictly inside the ear
            d1 = (b[0] - a[0]) * (P[:, 1] - a[1]) - (b[1] - a[1]) * (P[:, 0] - a[0])
            d2 = (c[0] - b[0]) * (P[:, 1] - b[1]) - (c[1] - b[1]) * (P[:, 0] - b[0])
            d3 = (a[0] - c[0]) * (P[:, 1] - c[1]) - (a[1] - c[1]) * (P[:, 0] - c[0])
            inside = (d1 >= 0) & (d2 >= 0) & (d3 >= 0)
            inside[[i - 1, i, (i + 1) % m]] = False
            # vertices coinciding with a corner of the ear do not block it
            for q in (a, b, c):
                inside &= ~((P[:, 0] == q[0]) & (P[:, 1] == q[1]))
            if inside.any():
                continue
            tris.append((idx[i - 1], idx[i], idx[(i + 1) % m]))
            del idx[i]
            clipped = True
            break
        if not clipped:
            break
    if len(idx) == 3:
        a, b, c = pts[idx[0]], pts[idx[1]], pts[idx[2]]
        if (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) != 0:
            tris.append(tuple(idx))
    return np.asarray(tris, dtype=np.int64).reshape(-1, 3)


# ---------------------------------------------------------------------------------------------------
# one plane -> mesh + texture   (get_single_image_mesh_arti, one plane at a time)
# ---------------------------------------------------------------------------------------------------
def plane_mesh(plane_param, mask, image, cfg: OptConfig, webvis: bool = False):
    """-> (verts (V,3) fp32, faces (F,3) int64, uvs (V,2) fp32, texture (300,300,3) uint8) or None when the
    mask has no ring.  ``plane_param`` is the detector's ``[a, b, c]`` (utils/vis.py:258-261)."""
    cv2 = _cv2()
    # fp32 as in the reference (the detector's planes are fp32 tensors, utils/vis.py:257-261); get_pcd then
    # promotes the fp32 normal / offset to float64
    p = np.array(plane_param, dtype=np.float32).reshape(1, 3)
    p[:, [1, 2]] = p[:, [2, 1]]
    p[:, 1] = -p[:, 1]
    offsets = np.linalg.norm(p, ord=2, axis=1)
    normal, offset = (p / offsets.reshape(-1, 1))[0], offsets[0]
    rings = mask_to_polygons(mask)
    if not rings:
        return None
    K = _mesh_K(cfg)
    all_v = np.concatenate(rings)
    pcd = _get_pcd(all_v, normal, offset, K)
    # rectifying homography from four control points on the plane (utils/vis.py:300-326)
    p0 = pcd[0]
    p1 = pcd[np.argmax(((pcd - p0) ** 2).sum(1))]
    d1 = (p1 - p0) / np.linalg.norm(p1 - p0)
    d2 = np.cross(d1, normal)
    ctrl = _project2D(np.stack([p0, p0 + d1, p0 + d2, p0 + d1 + d2]), K).astype(np.float32)
    fake = np.array([[0, 0], [0, TARGET], [TARGET, 0], [TARGET, TARGET]], dtype=np.float32)
    Hm = cv2.getPerspectiveTransform(ctrl, fake)
    P = cv2.perspectiveTransform(all_v.reshape(1, -1, 2), Hm)[0]
    x_t, y_t = P[:, 0].min(), P[:, 1].min()
    scale = max(P[:, 0].max() - P[:, 0].min(), P[:, 1].max() - P[:, 1].min())
    scale = scale if scale > 0 else 1.0
    Hs = np.array([[TARGET / scale, 0, -x_t * TARGET / scale], [0, TARGET / scale, -y_t * TARGET / scale], [0, 0, 1]])
    Hu = Hs @ Hm
    texture = cv2.warpPerspective(np.ascontiguousarray(image), Hu, (TARGET, TARGET))
    verts, faces, uvs = [], [], []
    n_verts = 0
    for ring in rings:
        tri = triangulate(ring)
        if len(tri) == 0:
            continue                                           # utils/vis.py:357-358
        pts3 = _get_pcd(ring, normal, offset, K)
        if webvis:                                             # utils/vis.py:341
            pts3 = (np.diag([-1.0, 1.0, -1.0]) @ np.diag([-1.0, -1.0, 1.0]) @ pts3.T).T
        rect = cv2.perspectiveTransform(ring.astype(np.float32).reshape(1, -1, 2), Hu)[0]
        uvs.append(np.array([0.0, 1.0]) + np.array([1.0, -1.0]) * rect / TARGET)
        verts.append(pts3)
        faces.append(tri[:, ::-1] + n_verts)                   # winding flipped (utils/vis.py:354)
        n_verts += len(pts3)
    if not verts:
        return None
    return (np.concatenate(verts).astype(np.float32), np.concatenate(faces).astype(np.int64),
            np.concatenate(uvs).astype(np.float32), texture)


def icosahedron(radius: float = 0.1, centre=(0.0, 0.0, 0.0)):
    """``ico_sphere(0)`` scaled and moved (tools/inference.py:75-86): 12 vertices, 20 faces."""
    a, b = 0.5257, 0.8507                                   # pytorch3d's level-0 table, four decimals [3P-unverified]
    v = np.array([[-a, b, 0], [a, b, 0], [-a, -b, 0], [a, -b, 0], [0, -a, b], [0, a, b], [0, -a, -b], [0, a, -b],
                  [b, 0, -a], [b, 0, a], [-b, 0, -a], [-b, 0, a]], dtype=np.float32)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    # Scale(0.1) then Translate(centre), each an fp32 step (tools/inference.py:72-86)
    return (v * np.float32(radius)) + np.asarray(centre, dtype=np.float64).astype(np.float32), f


# ---------------------------------------------------------------------------------------------------
# writer   (save_obj / _save, utils/mesh_utils.py:126-266)
# ---------------------------------------------------------------------------------------------------
def _write(folder, prefix, meshes, textures, decimal_places=10):
    cv2 = _cv2()
    os.makedirs(os.path.join(folder, "uv_maps"), exist_ok=True)
    fmt = "%." + str(decimal_places) + "f"
    names = []
    for k, tex in enumerate(textures):
        name = f"{prefix}_uv_plane_{k}"
        cv2.imwrite(os.path.join(folder, "uv_maps", name + ".png"), np.ascontiguousarray(tex[:, :, ::-1]))   # RGB -> BGR
        names.append(name)
    with open(os.path.join(folder, prefix + ".mtl"), "w") as f:
        for name in names:
            f.write(f"newmtl {name}\nmap_Kd {os.path.join('uv_maps', name + '.png')}\n# Test colors\n"
                    "Ka 1.000 1.000 1.000  # white\nKd 1.000 1.000 1.000  # white\nKs 0.000 0.000 0.000  # black\nNs 10.0\n")
    base = 0
    with open(os.path.join(folder, prefix + ".obj"), "w") as f:
        f.write(f"mtllib {prefix}.mtl\n\n")
        for k, ((verts, faces, uvs), name) in enumerate(zip(meshes, names)):
            lines = [f"# mesh {k}"]
            lines += ["v " + " ".join(fmt % c for c in v) for v in verts.tolist()]
            lines += ["vt " + " ".join(fmt % c for c in uv) for uv in uvs[:len(verts)].tolist()]
            lines.append(f"usemtl {name}")
            off = base + 1
            for tri in faces.tolist():
                lines.append("f " + " ".join(f"{i + off}/{i + off}" for i in tri))
                lines.append("f " + " ".join(f"{i + off}/{i + off}" for i in reversed(tri)))        # double-sided
            f.write("\n".join(lines) + "\n")
            base += len(verts)
    return os.path.join(folder, prefix + ".obj")


def write_textured_obj(folder: str, prefix: str, meshes, textures, decimal_places: int = 10):
    """``meshes``: list of (verts (V,3), faces (F,3), uvs (V,2)); ``textures``: one (h, w, 3) uint8 RGB image
    per mesh.  Writes ``<prefix>.obj`` (``# mesh k`` blocks of ``v`` / ``vt`` / ``usemtl`` / double-sided
    ``f a/a b/b c/c``), ``<prefix>.mtl`` and ``uv_maps/<prefix>_uv_plane_<k>.png``.  Face indices count over
    the whole file (the OBJ rule) — pytorch3d's ``faces_packed`` gives the reference's writer the same."""
    return _write(folder, prefix, meshes, textures, decimal_places)


# ---------------------------------------------------------------------------------------------------
# save_obj_model   (tools/inference.py:44-168)
# ---------------------------------------------------------------------------------------------------
def _tint(tex, colour, weight=1.0):
    out = (tex / 255.0 + colour.reshape(1, 1, 3) * weight) / 2
    return (out * 255.0).astype(np.uint8)


def articulation_meshes(preds, frame_id: int, image=None, cfg: OptConfig | None = None, axis_dir: str = "l",
                        webvis: bool = False):
    """-> (meshes, textures) of frame ``frame_id``: the most confident box, its rotated copies, the two axis
    markers and the background; None when the frame has no prediction."""
    cfg = cfg or OptConfig()
    p = preds[frame_id]
    scores = np.asarray(p.scores, dtype=np.float64)
    if scores.shape[0] == 0:
        return None
    box_id = int(scores.argmax())
    if image is None:
        image = np.full((cfg.height, cfg.width, 3), 160, dtype=np.uint8)
    image = np.asarray(image)
    mask = p.pred_masks[box_id]
    mask = (mask.detach().cpu().numpy() if torch.is_tensor(mask) else np.asarray(mask)) > 0.5
    # rotation axis in 3-D (tools/inference.py:55-70): get_pcd's own camera, not the mesh camera
    plane = p.pred_planes[box_id:(box_id + 1)].clone()
    plane[:, [1, 2]] = plane[:, [2, 1]]
    plane[:, 1] = -plane[:, 1]
    normal = F.normalize(plane, p=2)[0].numpy().astype(np.float64)
    offset = float(torch.norm(plane, p=2))
    centers = p.pred_boxes.get_centers()
    pts = angle_offset_to_axis(p.pred_rot_axis, centers, H=cfg.height, W=cfg.width)
    axis3d = _get_pcd(np.asarray(pts[box_id], dtype=np.float64).reshape(-1, 2), normal, offset, cfg.K())
    if webvis:
        axis3d = (np.diag([-1.0, 1.0, -1.0]) @ np.diag([-1.0, -1.0, 1.0]) @ axis3d.T).T
    d = axis3d[1] - axis3d[0]
    dir_vec = d / np.linalg.norm(d)

    fg = plane_mesh(p.pred_planes[box_id].numpy(), mask, image, cfg, webvis)
    if fg is None:
        return None
    verts, faces, uvs, tex = fg
    bg = plane_mesh(p.pred_planes[box_id].numpy(), ~mask, image, cfg, webvis)
    grid = np.arange(-1.8, 0.1, 1.8 / 4) if axis_dir == "l" else np.arange(0.0, 1.8, 1.8 / 4)
    if axis_dir not in ("l", "r"):
        raise NotImplementedError(axis_dir)
    R = rotation_matrices(grid, dir_vec)                      # (A, 3, 3) fp32, row-vector convention
    a = axis3d[0].astype(np.float32)
    meshes, textures = [(verts, faces, uvs)], [tex]
    q = verts - a                                              # t3, t2, t1 as three fp32 steps (:121-127)
    for i in range(len(R)):
        moved = ((q[:, 0:1] * R[i, 0] + q[:, 1:2] * R[i, 1]) + q[:, 2:3] * R[i, 2]) + a
        meshes.append((moved.astype(np.float32), faces.copy(), uvs))
        textures.append(tex)
    for end in axis3d:
        sv, sf = icosahedron(0.1, end)
        meshes.append((sv, sf, np.ones((len(sv), 2), dtype=np.float32)))
        textures.append(tex)
    # tints (tools/inference.py:149-157): the first five textures towards the door colour, the markers blue
    for i in range(min(5, len(textures))):
        textures[i] = _tint(textures[i], OBJ_TINT, i / 10 + 1 / 2)
    textures[-1] = _tint(textures[-1], AXIS_TINT)
    textures[-2] = _tint(textures[-2], AXIS_TINT)
    if bg is not None:
        meshes.append(bg[:3])
        textures.append(bg[3])
    return meshes, textures


def save_obj_model(output: str, preds, frame_id: int, image=None, cfg: OptConfig | None = None,
                   axis_dir: str = "l", webvis: bool = False):
    """``<output>/frame_<id>/arti_pred.obj`` (+ .mtl, uv_maps/) for one frame; returns the .obj path or None."""
    got = articulation_meshes(preds, frame_id, image, cfg, axis_dir, webvis)
    if got is None:
        print("no prediction!")
        return None
    meshes, textures = got
    return write_textured_obj(os.path.join(output, "frame_{:0>4}".format(frame_id)), "arti_pred", meshes, textures,
                              decimal_places=10)
