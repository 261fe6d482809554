"""TEST INFRASTRUCTURE — runs the reference's OWN .obj export (row f3) on a CPU box.

``save_obj_model`` (tools/inference.py:44-168) is compiled from the function's source text as it lies in
/root/reference (the tool module itself cannot be imported: it pulls in the detector), and executed with
the reference's unmodified ``get_single_image_mesh_arti`` (utils/vis.py:256-393) and ``save_obj`` / ``_save``
(utils/mesh_utils.py:126-266) imported under oracle/ref_shim.py.  Nothing of the reference is copied into
the repo.

Third-party pieces that are absent from this image AND from /root/reference are stood in for
[3P-unverified]:

* ``skimage.measure.find_contours`` / ``approximate_polygon`` (inside ``binary_mask_to_polygon``) and
  ``mapbox_earcut.triangulate_float32``: rebound to the product's own outline and ear-clipping routines
  (``export.mask_to_polygons`` / ``export.triangulate``).  The fixture therefore pins everything the
  reference does AROUND those two algorithms — plane conversion, the mesh camera, the rectifying homography
  and its textures, uv coordinates, winding, the rotated copies, the axis markers, tints, mesh order, the
  .obj / .mtl text — and says nothing about which outline / which triangulation of it skimage and earcut
  would have produced;
* ``pytorch3d.structures.Meshes`` / ``join_meshes_as_batch``, ``pytorch3d.renderer.Textures``,
  ``pytorch3d.utils.ico_sphere`` and ``pytorch3d.transforms.Scale``: minimal containers with the members the
  export touches, written from the published behaviour of pytorch3d 0.4-0.7 (list accessors return the
  stored lists; ``faces_packed`` offsets faces by the vertices before them; the deprecated ``Textures``
  hands padded uv rows back unsplit; ``ico_sphere(0)`` is the 12-vertex icosahedron).
* ``imageio.imwrite``: cv2 (RGB -> BGR).
"""
from __future__ import annotations

import ast
import importlib
import os
import types

import numpy as np
import torch

from . import ref_shim

TOOL = os.path.join(os.path.dirname(ref_shim.REFERENCE_ROOT.rstrip("/")), "articulation3d", "tools", "inference.py")


# --------------------------------------------------------------------------------------------------
# pytorch3d containers [3P-unverified]
# --------------------------------------------------------------------------------------------------
class Textures:
    """The deprecated ``pytorch3d.renderer.Textures`` as far as the export uses it."""

    def __init__(self, maps=None, faces_uvs=None, verts_uvs=None, verts_rgb=None, _lists=None):
        self._maps = maps
        if _lists is not None:
            self._verts_uvs_list, self._faces_uvs_list = _lists
            return
        # padded (N, V, 2) tensors come back as N full rows (no per-mesh sizes are kept); lists as they are
        self._verts_uvs_list = list(verts_uvs.unbind(0)) if torch.is_tensor(verts_uvs) else list(verts_uvs)
        self._faces_uvs_list = list(faces_uvs.unbind(0)) if torch.is_tensor(faces_uvs) else list(faces_uvs)

    def verts_uvs_list(self):
        return self._verts_uvs_list

    def faces_uvs_list(self):
        return self._faces_uvs_list

    def clone(self):
        return Textures(maps=self._maps, _lists=([v.clone() for v in self._verts_uvs_list],
                                                 [f.clone() for f in self._faces_uvs_list]))

    def join_batch(self, others):
        vu, fu = list(self._verts_uvs_list), list(self._faces_uvs_list)
        for o in others:
            vu += o._verts_uvs_list
            fu += o._faces_uvs_list
        return Textures(maps=self._maps, _lists=(vu, fu))

    def cuda(self):
        return self

    cpu = cuda


class Meshes:
    def __init__(self, verts, faces, textures=None):
        self._verts_list, self._faces_list = list(verts), list(faces)
        self.textures = textures

    def verts_list(self):
        return self._verts_list

    def faces_list(self):
        return self._faces_list

    def num_faces_per_mesh(self):
        return torch.tensor([len(f) for f in self._faces_list], dtype=torch.int64)

    def faces_packed(self):
        out, base = [], 0
        for v, f in zip(self._verts_list, self._faces_list):
            out.append(f.to(torch.int64) + base)
            base += len(v)
        return torch.cat(out) if out else torch.zeros(0, 3, dtype=torch.int64)

    def cuda(self):
        return self

    cpu = cuda


def join_meshes_as_batch(meshes):
    verts = [v for m in meshes for v in m.verts_list()]
    faces = [f for m in meshes for f in m.faces_list()]
    tex = meshes[0].textures.join_batch([m.textures for m in meshes[1:]])
    return Meshes(verts, faces, tex)


def ico_sphere(level: int = 0):
    assert level == 0
    a, b = 0.5257, 0.8507                                   # pytorch3d's level-0 vertex table (four decimals)
    v = torch.tensor([[-a, b, 0], [a, b, 0], [-a, -b, 0], [a, -b, 0], [0, -a, b], [0, a, b], [0, -a, -b], [0, a, -b],
                      [b, 0, -a], [b, 0, a], [-b, 0, -a], [-b, 0, a]], dtype=torch.float32)
    f = torch.tensor([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                      [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                      [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=torch.int64)
    return Meshes([v], [f])


class Scale(ref_shim.Transform3d):
    def __init__(self, s, dtype=torch.float32, device="cpu"):
        super().__init__(dtype=dtype, device=device)
        m = torch.eye(4, dtype=dtype)
        m[0, 0] = m[1, 1] = m[2, 2] = float(s)
        self._matrix = m[None]


# --------------------------------------------------------------------------------------------------
def _function_source(path: str, name: str) -> str:
    with open(path) as f:
        src = f.read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def load_export():
    """-> the reference's ``save_obj_model`` function object, bound to the reference's own helpers."""
    import cv2
    import torch.nn.functional as F
    from articulation3d_b200 import export
    ou = ref_shim.load_reference()
    vis = importlib.import_module("articulation3d.utils.vis")
    mesh_utils = importlib.import_module("articulation3d.utils.mesh_utils")
    pt = importlib.import_module("articulation3d.data.planercnn_transforms")

    def binary_mask_to_polygon(binary_mask, tolerance=0):                    # [3P] outline: the product's
        return [ring.ravel().tolist() for ring in export.mask_to_polygons(np.asarray(binary_mask))]

    class _Earcut:                                                            # [3P] triangulation: the product's
        @staticmethod
        def triangulate_float32(verts, rings):
            return export.triangulate(np.asarray(verts, dtype=np.float64)).reshape(-1).astype(np.uint32)

    vis.binary_mask_to_polygon = binary_mask_to_polygon
    vis.earcut = _Earcut
    vis.Meshes, vis.Textures = Meshes, Textures
    mesh_utils.imageio = types.SimpleNamespace(
        imwrite=lambda path, img: cv2.imwrite(path, np.ascontiguousarray(np.asarray(img)[:, :, ::-1])))
    p3d = types.SimpleNamespace(
        transforms=types.SimpleNamespace(Scale=Scale, Transform3d=ref_shim.Transform3d, Rotate=ref_shim.Rotate,
                                         axis_angle_to_matrix=ref_shim.axis_angle_to_matrix),
        structures=types.SimpleNamespace(join_meshes_as_batch=join_meshes_as_batch, Meshes=Meshes))
    ns = {"np": np, "torch": torch, "F": F, "os": os, "pytorch3d": p3d, "Meshes": Meshes, "Textures": Textures,
          "ico_sphere": ico_sphere, "ArtiVisualizer": lambda *a, **k: None,
          "angle_offset_to_axis": pt.angle_offset_to_axis, "get_pcd": vis.get_pcd,
          "get_single_image_mesh_arti": vis.get_single_image_mesh_arti, "save_obj": mesh_utils.save_obj}
    exec(compile(_function_source(TOOL, "save_obj_model"), TOOL, "exec"), ns)
    return ns["save_obj_model"]


def run_reference_export(preds, frames, frame_id: int, out_dir: str, axis_dir: str = "l", webvis: bool = False) -> str:
    """The reference's ``save_obj_model(args, preds, frames, frame_id)`` -> folder of the written files."""
    fn = load_export()
    os.makedirs(out_dir, exist_ok=True)
    fn(types.SimpleNamespace(output=out_dir, webvis=webvis), preds, frames, frame_id, axis_dir=axis_dir)
    return os.path.join(out_dir, "frame_{:0>4}".format(frame_id))
