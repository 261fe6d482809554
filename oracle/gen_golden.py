"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED
reference (under oracle/ref_shim.py) on seeded synthetic clips.

    python -m oracle.gen_golden            # only works where /root/reference exists

The reference has no golden vectors of its own, so these
fixtures ARE the pin: inputs (bit-packed masks + per-box predictions) and every
observable output of ``track_planes`` + ``optimize_planes('3dc')`` — RNG source
choices, the cluster angle lists handed to ``linregress``, ``has_rot``,
``std_axis``, ``reg_masks`` (bit-packed), ``reg_normals``, output scores and
axes.  The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (seed, n_tracks, n_frames, kinds, drop_prob)
CASES = {
    "clip_a": (2020, 4, 24, [0, 0, 1, 2], 0.0),
    "clip_b": (12, 3, 20, [3, 3, 1], 0.05),
    "clip_c": (5, 2, 16, [2, 0], 0.0),
    "clip_d": (7, 5, 40, None, 0.05),
    "clip_e": (21, 3, 30, [1, 1, 0], 0.0),
}


def pack_bits(m: np.ndarray) -> np.ndarray:
    """(..., H, W) bool -> (..., H, ceil(W/8)) uint8, bit i of byte j = pixel 8j+i."""
    return np.packbits(m.astype(bool), axis=-1, bitorder="little")


def unpack_bits(b: np.ndarray, W: int) -> np.ndarray:
    return np.unpackbits(b, axis=-1, bitorder="little")[..., :W].astype(bool)


def preds_to_arrays(preds) -> dict:
    out = {"n_frames": np.int64(len(preds)), "image_size": np.array(preds[0].image_size)}
    for t, p in enumerate(preds):
        out[f"f{t}_scores"] = np.asarray(p.scores)
        out[f"f{t}_boxes"] = p.pred_boxes.tensor.numpy()
        out[f"f{t}_classes"] = np.asarray(p.pred_classes)
        out[f"f{t}_planes"] = p.pred_planes.numpy()
        out[f"f{t}_rot_axis"] = p.pred_rot_axis.numpy()
        out[f"f{t}_tran_axis"] = p.pred_tran_axis.numpy()
        out[f"f{t}_masks"] = pack_bits(p.pred_masks.numpy() > 0.5)
    return out


def arrays_to_preds(z, instances_cls, boxes_cls, prefix="f"):
    H, W = [int(v) for v in z["image_size"]]
    preds = []
    for t in range(int(z["n_frames"])):
        p = instances_cls((H, W))
        p.scores = np.array(z[f"{prefix}{t}_scores"])
        p.pred_boxes = boxes_cls(torch.from_numpy(np.array(z[f"{prefix}{t}_boxes"])))
        p.pred_classes = np.array(z[f"{prefix}{t}_classes"])
        p.pred_planes = torch.from_numpy(np.array(z[f"{prefix}{t}_planes"]))
        p.pred_rot_axis = torch.from_numpy(np.array(z[f"{prefix}{t}_rot_axis"]))
        p.pred_tran_axis = torch.from_numpy(np.array(z[f"{prefix}{t}_tran_axis"]))
        bits = np.array(z[f"{prefix}{t}_masks"])
        p.pred_masks = torch.from_numpy(unpack_bits(bits, W).astype(np.float32)).reshape(-1, H, W)
        preds.append(p)
    return preds


class _RecordingRandom:
    """Stands in for the ``random`` module inside the reference's opt_utils:
    same global Mersenne Twister, but every ``choice`` is logged."""

    def __init__(self):
        self.choices = []

    def choice(self, seq):
        c = random.choice(seq)
        self.choices.append((int(c), len(seq)))
        return c

    def __getattr__(self, name):
        return getattr(random, name)


def run_reference(preds, seed: int, method: str = "3dc") -> dict:
    """Run the unmodified reference on ``preds`` (shim-typed) and collect
    every observable."""
    ou = ref_shim.load_reference()
    rec = _RecordingRandom()
    lin_calls = []
    real_linregress = ou.linregress

    def linregress(x, y):
        lin_calls.append(np.asarray(y, dtype=np.float32).copy())
        return real_linregress(x, y)

    old_random = ou.random
    ou.random, ou.linregress = rec, linregress
    try:
        random.seed(seed)
        planes = ou.track_planes(preds)
        out = ou.optimize_planes(preds, planes, method)
    finally:
        ou.random, ou.linregress = old_random, real_linregress

    res = {"seed": np.int64(seed),
           "choices": np.array(rec.choices, dtype=np.int64).reshape(-1, 2),
           "n_lin": np.int64(len(lin_calls))}
    for i, a in enumerate(lin_calls):
        res[f"lin{i}"] = a
    for cat in ("trans", "rot"):
        res[f"{cat}_n"] = np.int64(len(planes[cat]))
        for i, p in enumerate(planes[cat]):
            k = f"{cat}{i}"
            res[f"{k}_ids"] = np.array(sorted(p["ids"].items()), dtype=np.int64).reshape(-1, 2)
            res[f"{k}_ids_order"] = np.array(list(p["ids"].keys()), dtype=np.int64)
            res[f"{k}_has_rot"] = np.bool_(p["has_rot"])
            if p["has_rot"]:
                res[f"{k}_std_axis"] = torch.as_tensor(p["std_axis"]).numpy()
                frames = list(p["reg_masks"].keys())
                res[f"{k}_reg_frames"] = np.array(frames, dtype=np.int64)
                res[f"{k}_reg_masks"] = pack_bits(
                    np.stack([p["reg_masks"][f].numpy() > 0.5 for f in frames]))
                if "reg_normals" in p:
                    res[f"{k}_reg_normals"] = np.stack([p["reg_normals"][f].numpy() for f in frames])
    for t, p in enumerate(out):
        res[f"o{t}_scores"] = np.asarray(p.scores)
        res[f"o{t}_rot_axis"] = p.pred_rot_axis.numpy()
        res[f"o{t}_tran_axis"] = p.pred_tran_axis.numpy()
        res[f"o{t}_planes"] = p.pred_planes.numpy()
    return res


def run_reference_diag(preds, seed: int) -> dict:
    """The reference's own diagnostics (``check_axis`` / ``check_monotonic``, utils/opt_utils.py:977-1152)
    on a clip before / after its ``optimize_planes('3dc')``."""
    from articulation3d_b200 import synth
    ou = ref_shim.load_reference()
    before = synth.clone_preds(preds, ref_shim.Instances, ref_shim.Boxes)
    work = synth.clone_preds(preds, ref_shim.Instances, ref_shim.Boxes)
    random.seed(seed)
    planes = ou.track_planes(work)
    out = ou.optimize_planes(work, planes, "3dc")
    s0, s1 = ou.check_axis(before, out, planes["rot"], "3dc")
    c0, c1 = ou.check_monotonic(before, out, planes["rot"], "3dc")
    return {"axis_scores": np.array([float(v) for v in s0], dtype=np.float64),
            "axis_scores_opt": np.array([float(v) for v in s1], dtype=np.float64),
            "fit_scores": np.array([float(v[0]) for v in c0], dtype=np.float64),
            "fit_scores_opt": np.array([float(v[0]) for v in c1], dtype=np.float64)}


def main_diag():
    from articulation3d_b200 import synth
    if not ref_shim.available():
        raise SystemExit("reference not present; fixtures can only be generated in the build container")
    os.makedirs(os.path.join(GOLDEN_DIR, "diag"), exist_ok=True)
    for name in ("clip_a", "clip_b", "clip_d"):
        seed, n_tracks, n_frames, kinds, drop = CASES[name]
        preds, _ = synth.make_video(seed, n_tracks, n_frames, kinds=kinds, drop_prob=drop)
        res = run_reference_diag(preds, seed)
        path = os.path.join(GOLDEN_DIR, "diag", f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"wrote {path}: {len(res['axis_scores'])} axis pairs, fit {res['fit_scores']} -> {res['fit_scores_opt']}")


def build_eval_case(seed: int, n_tracks: int, n_frames: int, kinds):
    """Synthetic evaluation input: prediction records of a clip and a COCO-style ground truth with ONE
    annotation per image (the reference's matching loop needs that), derived from the predictions by
    seeded perturbations so that every criterion has hits and misses."""
    from articulation3d_b200 import synth
    from articulation3d_b200.axis import angle_offset_to_axis
    rng = np.random.RandomState(seed)
    preds, _ = synth.make_video(seed, n_tracks, n_frames, kinds=kinds)
    records, anns, images = [], [], []
    for t, p in enumerate(preds):
        boxes = p.pred_boxes.tensor.numpy().astype(np.float64)
        n = len(boxes)
        scores = np.clip(0.95 - 0.1 * rng.rand(n) * (rng.rand(n) < 0.5), 0, 1)
        inst = [{"image_id": t, "category_id": int(p.pred_classes[k]),
                 "bbox": [boxes[k, 0], boxes[k, 1], boxes[k, 2] - boxes[k, 0], boxes[k, 3] - boxes[k, 1]],
                 "score": float(scores[k])} for k in range(n)]
        records.append({"image_id": t, "instances": inst, "pred_plane": p.pred_planes.clone(),
                        "pred_rot_axis": p.pred_rot_axis.clone(), "pred_tran_axis": p.pred_tran_axis.clone()})
        images.append({"id": t, "width": 640, "height": 480})
        if n == 0 or rng.rand() < 0.15:
            continue                                           # image without ground truth
        k = int(rng.randint(n))
        cls = int(p.pred_classes[k])
        if rng.rand() < 0.1:
            cls = 1 - cls                                      # wrong class
        jit = rng.randn(4) * (3 if rng.rand() < 0.7 else 60)   # mostly IoU > 0.5, sometimes not
        x0, y0, x1, y1 = boxes[k] + jit
        centers = p.pred_boxes.get_centers()
        rot_line = angle_offset_to_axis(p.pred_rot_axis[k:k + 1], centers[k:k + 1])[0].numpy().astype(np.float64)
        tr = torch.cat((p.pred_tran_axis[k:k + 1], torch.zeros(1, 1)), 1)
        tran_line = angle_offset_to_axis(tr, centers[k:k + 1])[0].numpy().astype(np.float64)
        ajit = rng.randn(4) * (4 if rng.rand() < 0.6 else 150)
        plane = p.pred_planes[k].numpy().astype(np.float64)
        nrm = np.array([plane[0], -plane[2], plane[1]]) / max(np.linalg.norm(plane), 1e-9)   # camera frame
        nrm = nrm + rng.randn(3) * (0.1 if rng.rand() < 0.6 else 1.0)
        nrm /= np.linalg.norm(nrm)
        ann = {"id": len(anns) + 1, "image_id": t, "category_id": cls + 1,
               "bbox": [float(x0), float(y0), float(max(x1 - x0, 2.0)), float(max(y1 - y0, 2.0))],
               "rot_axis": (rot_line + ajit).tolist() if cls == 0 and rng.rand() < 0.9 else None,
               "tran_axis": (tran_line + ajit).tolist() if cls == 1 and rng.rand() < 0.9 else None,
               "normal": [float(nrm[0]), float(-nrm[1]), float(nrm[2])] if rng.rand() < 0.85 else None}
        anns.append(ann)
    gt = {"images": images, "annotations": anns,
          "categories": [{"id": 1, "name": "arti_rot"}, {"id": 2, "name": "arti_tran"}]}
    return records, gt


def main_eval():
    """tests/golden/eval/*.json: inputs + the reference's own evaluate_for_arti_axis / _recognition outputs."""
    import importlib
    import json
    import types
    if not ref_shim.available():
        raise SystemExit("reference not present; fixtures can only be generated in the build container")
    ref_shim.load_reference()
    ev = importlib.import_module("articulation3d.evaluation.arti_evaluation")

    class _Boxes(ref_shim.Boxes):
        def to(self, device):
            return self

    class _BoxMode:                                            # [3P-unverified] detectron2 BoxMode.convert, XYWH -> XYXY
        XYWH_ABS, XYXY_ABS = 1, 0

        @staticmethod
        def convert(box, from_mode, to_mode):
            assert (from_mode, to_mode) == (_BoxMode.XYWH_ABS, _BoxMode.XYXY_ABS)
            arr = np.array(box, dtype=np.float64).reshape(-1, 4)
            arr[:, 2] += arr[:, 0]
            arr[:, 3] += arr[:, 1]
            return arr

    ev.Boxes, ev.BoxMode, ev.pairwise_iou = _Boxes, _BoxMode, ref_shim.pairwise_iou
    ev.create_small_table = lambda d: str(d)
    from articulation3d_b200.evaluation import CocoGT
    meta = types.SimpleNamespace(thing_classes=["arti_rot", "arti_tran"], thing_dataset_id_to_contiguous_id={1: 0, 2: 1})
    os.makedirs(os.path.join(GOLDEN_DIR, "eval"), exist_ok=True)
    for name, (seed, n_tracks, n_frames, kinds) in {"case_a": (31, 3, 40, [0, 1, 0]), "case_b": (32, 2, 60, [1, 0]),
                                                    "case_c": (33, 4, 30, [0, 0, 1, 2])}.items():
        records, gt = build_eval_case(seed, n_tracks, n_frames, kinds)
        want = ev.evaluate_for_arti_axis(records, CocoGT(gt), meta, 0.0)
        rec = ev.evaluate_for_recognition(records, CocoGT(gt), meta, 0.0)
        out = {"records": [{"image_id": r["image_id"], "instances": r["instances"],
                            "pred_plane": r["pred_plane"].tolist(), "pred_rot_axis": r["pred_rot_axis"].tolist(),
                            "pred_tran_axis": r["pred_tran_axis"].tolist()} for r in records],
               "gt": gt, "filter_iou": 0.0,
               "arti_axis": {k: float(v) for k, v in want.items()},
               "recognition": {k: float(v) for k, v in rec.items()}}
        path = os.path.join(GOLDEN_DIR, "eval", f"{name}.json")
        with open(path, "w") as f:
            json.dump(out, f)
        print(f"wrote {path} ({os.path.getsize(path) // 1024} KB):", out["arti_axis"], out["recognition"])


# --------------------------------------------------------------------------
# row f1: the reference's own override_depth / get_K_inv_dot_xy_1  (utils/arti_vis.py:101-149)
# --------------------------------------------------------------------------
DEPTH_CASES = {"depth_a": (41, 3, 12, [0, 1, 0], (0, 5, 11)), "depth_b": (42, 2, 10, [1, 0], (3, 9))}


def depth_case_inputs(name: str):
    """Seeded inputs of a depth fixture, regenerated identically wherever the fixture is read: per frame
    the detections' RLE masks (+ one empty mask) with their plane rows, and a depth map of integer
    millimetres (``RandomState.randint`` is platform independent; the fixture stores a checksum)."""
    from articulation3d_b200 import rle, synth
    seed, n_tracks, n_frames, kinds, frames = DEPTH_CASES[name]
    preds, _ = synth.make_video(seed, n_tracks, n_frames, kinds=kinds)
    rng = np.random.RandomState(seed)
    H, W = preds[0].image_size
    records, depths = [], []
    for f in frames:
        p = preds[f]
        n = int(p.pred_masks.shape[0])
        dets = [{"segmentation": rle.encode(p.pred_masks[k].numpy() > 0.5)} for k in range(n)]
        dets.append({"segmentation": rle.encode(np.zeros((H, W), bool))})            # empty mask keeps its plane
        planes = torch.cat([p.pred_planes, torch.tensor([[0.3, 1.0, -0.2]])])
        records.append({"instances": dets, "pred_plane": planes})
        # a tilted surface with millimetre noise, 1.5 .. 4.5 m
        yy, xx = np.mgrid[0:H, 0:W]
        mm = 1500 + (2 * xx + 3 * yy) % 2000 + rng.randint(0, 1000, size=(H, W))
        depths.append((mm.astype(np.float64) / 1000.0).astype(np.float32))
    return records, np.stack(depths)


def main_depth():
    """tests/golden/depth/*.npz: outputs of the reference's static ``PlaneRCNN_Branch.override_depth`` and of
    its ``get_K_inv_dot_xy_1`` run under the shim; ``mask_util.decode`` (pycocotools, absent) is rebound to
    the oracle's RLE decoder [3P-unverified]."""
    import importlib
    from oracle import restated
    if not ref_shim.available():
        raise SystemExit("reference not present; fixtures can only be generated in the build container")
    ref_shim.load_reference()
    av = importlib.import_module("articulation3d.utils.arti_vis")
    av.mask_util.decode = restated.rle_decode
    os.makedirs(os.path.join(GOLDEN_DIR, "depth"), exist_ok=True)
    rays64 = av.PlaneRCNN_Branch.get_K_inv_dot_xy_1(None)                      # (3, 480, 640) float64
    rays = torch.FloatTensor(rays64)                                           # arti_vis.py:52
    for name in DEPTH_CASES:
        records, depths = depth_case_inputs(name)
        res = {"depth_checksum": np.float64(depths.astype(np.float64).sum()),
               "rays_probe": rays64[:, ::37, ::41].copy(), "n_frames": np.int64(len(records))}
        for i, rec in enumerate(records):
            xyz = rays * torch.from_numpy(depths[i])                           # depth2XYZ, arti_vis.py:90-99
            inst = {"instances": rec["instances"], "pred_plane": rec["pred_plane"].clone()}
            out = av.PlaneRCNN_Branch.override_depth(xyz, inst)
            res[f"f{i}_pred_plane_in"] = rec["pred_plane"].numpy()
            res[f"f{i}_pred_plane_out"] = out["pred_plane"].numpy()
        path = os.path.join(GOLDEN_DIR, "depth", f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"wrote {path}:", res["f0_pred_plane_out"].tolist())


# --------------------------------------------------------------------------
# row f3: the reference's own .obj export  (tools/inference.py:44-168, utils/vis.py:256-393,
# utils/mesh_utils.py:126-266) run by oracle/ref_export.py
# --------------------------------------------------------------------------
EXPORT_CASES = {"export_a": (5, 2, 12, [0, 1], 3, "l", False), "export_b": (6, 3, 10, [0, 0, 1], 7, "r", True)}


def export_case_inputs(name: str):
    """Seeded inputs of an export fixture: the clip's predictions and a random RGB frame."""
    from articulation3d_b200 import synth
    seed, n_tracks, n_frames, kinds, frame_id, axis_dir, webvis = EXPORT_CASES[name]
    preds, _ = synth.make_video(seed, n_tracks, n_frames, kinds=kinds)
    image = np.random.RandomState(seed).randint(0, 256, size=(480, 640, 3)).astype(np.uint8)
    return preds, image, frame_id, axis_dir, webvis


def export_digest(folder: str) -> dict:
    """sha256 of every file an export wrote, plus the element counts of the .obj (for a readable failure)."""
    import hashlib
    out = {}
    for root, _, files in os.walk(folder):
        for fn in sorted(files):
            path = os.path.join(root, fn)
            with open(path, "rb") as f:
                out[os.path.relpath(path, folder)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(folder, "arti_pred.obj")) as f:
        lines = f.read().splitlines()
    out["_counts"] = {k: sum(1 for ln in lines if ln.startswith(k + " ")) for k in ("v", "vt", "f", "usemtl")}
    out["_meshes"] = sum(1 for ln in lines if ln.startswith("# mesh"))
    return out


def main_export():
    import json
    import tempfile
    from articulation3d_b200 import synth
    from oracle import ref_export
    if not ref_shim.available():
        raise SystemExit("reference not present; fixtures can only be generated in the build container")
    os.makedirs(os.path.join(GOLDEN_DIR, "export"), exist_ok=True)
    for name in EXPORT_CASES:
        preds, image, frame_id, axis_dir, webvis = export_case_inputs(name)
        rp = synth.clone_preds(preds, ref_shim.Instances, ref_shim.Boxes)
        with tempfile.TemporaryDirectory() as tmp:
            folder = ref_export.run_reference_export(rp, [image] * len(preds), frame_id, tmp, axis_dir=axis_dir, webvis=webvis)
            digest = export_digest(folder)
        path = os.path.join(GOLDEN_DIR, "export", f"{name}.json")
        with open(path, "w") as f:
            json.dump(digest, f, indent=1, sort_keys=True)
        print(f"wrote {path}:", digest["_counts"], digest["_meshes"], "meshes")


def main():
    from articulation3d_b200 import synth
    if not ref_shim.available():
        raise SystemExit("reference not present; fixtures can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, (seed, n_tracks, n_frames, kinds, drop) in CASES.items():
        preds, _ = synth.make_video(seed, n_tracks, n_frames, kinds=kinds, drop_prob=drop)
        arrays = preds_to_arrays(preds)
        rp = synth.clone_preds(preds, ref_shim.Instances, ref_shim.Boxes)
        res = run_reference(rp, seed)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **arrays, **res)
        print(name, os.path.getsize(path) // 1024, "KiB",
              {c: [bool(res[f"{c}{i}_has_rot"]) for i in range(int(res[f"{c}_n"]))] for c in ("trans", "rot")},
              "choices", res["choices"].tolist())


if __name__ == "__main__":
    if "--diag" in sys.argv:
        main_diag()
    elif "--eval" in sys.argv:
        main_eval()
    elif "--depth" in sys.argv:
        main_depth()
    elif "--export" in sys.argv:
        main_export()
    else:
        main()
