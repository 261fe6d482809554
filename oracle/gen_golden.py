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
    main_diag() if "--diag" in sys.argv else main()
