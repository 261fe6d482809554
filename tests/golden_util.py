"""Helpers to read tests/golden/*.npz (written by oracle/gen_golden.py)."""
import glob
import os

import numpy as np
import torch

from oracle.gen_golden import arrays_to_preds, unpack_bits  # noqa: F401  (test infrastructure)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))


def check_against_golden(z, planes, out, choices=None, lin=None, float_rtol=1e-4):
    """Compare the observable outputs of an optimize_planes('3dc') run with a
    golden record.  Integer / index / mask outputs must be bit-exact; floats
    within 1e-4 relative (BASELINE.json north_star)."""
    W = int(z["image_size"][1])
    if choices is not None:
        assert np.array_equal(np.asarray(choices, dtype=np.int64).reshape(-1, 2), z["choices"])
    if lin is not None:
        assert len(lin) == int(z["n_lin"])
        for i, a in enumerate(lin):
            assert np.array_equal(np.asarray(a, dtype=np.float32), z[f"lin{i}"]), f"cluster angles {i}"
    for cat in ("trans", "rot"):
        assert len(planes[cat]) == int(z[f"{cat}_n"])
        for i, p in enumerate(planes[cat]):
            k = f"{cat}{i}"
            assert np.array_equal(np.array(list(p["ids"].keys())), z[f"{k}_ids_order"])
            assert np.array_equal(np.array(sorted(p["ids"].items()), dtype=np.int64).reshape(-1, 2),
                                  z[f"{k}_ids"])
            assert bool(p["has_rot"]) == bool(z[f"{k}_has_rot"]), k
            if not p["has_rot"]:
                continue
            std = torch.as_tensor(p["std_axis"]).numpy()
            if std.dtype.kind == "i":
                assert np.array_equal(std, z[f"{k}_std_axis"])
            else:
                np.testing.assert_allclose(std, z[f"{k}_std_axis"], rtol=float_rtol)
            frames = list(p["reg_masks"].keys())
            assert np.array_equal(np.array(frames), z[f"{k}_reg_frames"])
            gold = unpack_bits(z[f"{k}_reg_masks"], W)
            for j, f in enumerate(frames):
                got = np.asarray(p["reg_masks"][f]) > 0.5
                assert got.sum() == gold[j].sum(), f"{k} frame {f}: mask-pixel count"
                assert np.array_equal(got, gold[j]), f"{k} frame {f}: reg_mask pixels"
            if f"{k}_reg_normals" in z.files:
                got = np.stack([np.asarray(p["reg_normals"][f]) for f in frames])
                np.testing.assert_allclose(got, z[f"{k}_reg_normals"], rtol=float_rtol, atol=1e-6)
    for t, p in enumerate(out):
        np.testing.assert_allclose(np.asarray(p.scores), z[f"o{t}_scores"], rtol=1e-12)
        np.testing.assert_allclose(p.pred_rot_axis.numpy(), z[f"o{t}_rot_axis"], rtol=float_rtol, atol=1e-7)
        np.testing.assert_allclose(p.pred_tran_axis.numpy(), z[f"o{t}_tran_axis"], rtol=float_rtol, atol=1e-7)
        np.testing.assert_allclose(p.pred_planes.numpy(), z[f"o{t}_planes"], rtol=float_rtol, atol=1e-7)
