"""GPU tests of the rows either side of the hot path: on-device
RLE decode and override_depth, against the CPU oracle."""
import numpy as np
import pytest
import torch

from articulation3d_b200 import adapter, engine, rle, synth
from oracle import restated

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _unpack(bits, W):
    b = bits.cpu().numpy().view(np.uint32)
    u8 = b.view(np.uint8).reshape(b.shape[0], b.shape[1], -1)
    return np.unpackbits(u8, axis=-1, bitorder="little")[..., :W].astype(bool)


@pytest.mark.parametrize("H,W", [(480, 640), (37, 100), (768, 1024)])
def test_rle_to_pool_matches_dense_decode(H, W):
    rng = np.random.RandomState(H)
    masks = []
    m = np.zeros((H, W), bool); masks.append(m.copy())                 # empty
    m[:] = True; masks.append(m.copy())                                  # full (one run of H*W ones)
    m[:] = False; m[H // 4: H // 2, W // 5: W // 2] = True; masks.append(m.copy())
    masks.append(rng.rand(H, W) < 0.5)                                   # ~H*W/2 runs
    m = np.zeros((H, W), bool); m[0, 0] = m[-1, -1] = True; masks.append(m)
    rles = [rle.encode(x) for x in masks]
    rles[2] = {"size": [H, W], "counts": rle.rle_counts(rles[2]).tolist()}     # uncompressed variant
    pool = engine.rle_to_pool(rles, H, W, DEV)
    got = _unpack(pool.bits, W)
    for i, x in enumerate(masks):
        assert np.array_equal(got[i], x), i
        assert np.array_equal(restated.rle_decode(rles[i]).astype(bool), x)
    assert pool.popc.cpu().tolist() == [int(x.sum()) for x in masks]


def test_override_depth_matches_oracle():
    H, W = 480, 640
    preds, _ = synth.make_video(9, 3, 12, kinds=[0, 1, 0])
    g = torch.Generator().manual_seed(0)
    rays64 = adapter.get_K_inv_dot_xy_1(H, W)
    assert np.array_equal(rays64[:, ::37, ::41], restated.get_K_inv_dot_xy_1(H, W)[:, ::37, ::41])
    rays = torch.FloatTensor(rays64)
    frames = [0, 5, 11]
    depths = 1.5 + torch.rand(len(frames), H, W, generator=g) * 2
    inst_a, inst_b = [], []
    for f in frames:
        p = preds[f]
        dets = [{"segmentation": rle.encode(p.pred_masks[k].numpy() > 0.5)} for k in range(3)]
        dets.append({"segmentation": rle.encode(np.zeros((H, W), bool))})          # empty mask keeps its plane
        planes = torch.cat([p.pred_planes, torch.tensor([[0.3, 1.0, -0.2]])])
        inst_a.append({"instances": dets, "pred_plane": planes.clone()})
        inst_b.append({"instances": dets, "pred_plane": planes.clone()})
    out = adapter.override_depth_batch(inst_a, depths=depths.to(DEV), rays=rays.to(DEV))
    for i, f in enumerate(frames):
        xyz = rays * depths[i]
        want = restated.override_depth(xyz, inst_b[i])["pred_plane"]
        np.testing.assert_allclose(out[i]["pred_plane"].numpy(), want.numpy(), rtol=1e-4, atol=1e-6)
    # reference signature: one frame, XYZ already formed
    one = {"instances": inst_b[0]["instances"], "pred_plane": torch.cat([preds[0].pred_planes, torch.tensor([[0.3, 1.0, -0.2]])])}
    two = {"instances": one["instances"], "pred_plane": one["pred_plane"].clone()}
    a = adapter.override_depth(one, xyz=(rays * depths[0]).to(DEV))
    b = restated.override_depth(rays * depths[0], two)
    np.testing.assert_allclose(a["pred_plane"].numpy(), b["pred_plane"].numpy(), rtol=1e-4, atol=1e-6)


def test_optimize_planes_from_rle_masks_equals_dense():
    import random
    from articulation3d_b200 import opt_utils
    preds, _ = synth.make_video(17, 3, 14, kinds=[0, 1, 0])
    dense = synth.clone_preds(preds)
    rl = synth.clone_preds(preds)
    for p in rl:
        p.pred_rle = [rle.encode(m.numpy() > 0.5) for m in p.pred_masks]
        del p._fields["pred_masks"]
    random.seed(4)
    pa = opt_utils.track_planes(dense)
    oa = opt_utils.optimize_planes(dense, pa, '3dc', device=DEV)
    random.seed(4)
    pb = opt_utils.track_planes(rl)
    ob = opt_utils.optimize_planes(rl, pb, '3dc', device=DEV)
    for cat in ("trans", "rot"):
        for p, q in zip(pa[cat], pb[cat]):
            assert p['has_rot'] == q['has_rot']
            if p['has_rot']:
                assert np.array_equal(p['fit']['angle_id'], q['fit']['angle_id'])
                assert np.array_equal(p['fit']['inter'], q['fit']['inter'])
                assert np.array_equal(p['fit']['union'], q['fit']['union'])
    for x, y in zip(oa, ob):
        assert np.array_equal(x.scores, y.scores) and torch.equal(x.pred_rot_axis, y.pred_rot_axis)


def test_opt_arti_tool_end_to_end(tmp_path):
    """synthetic .pth in -> optimised .pth + tracks.json + .obj out, equal to the API run on dense masks."""
    import json
    import random
    from articulation3d_b200 import io, opt_utils
    from articulation3d_b200.tools import opt_arti
    inp, out = str(tmp_path / "pred.pth"), str(tmp_path / "out")
    opt_arti.main(["--input", inp, "--output", out, "--synthetic", "2", "--tracks", "3", "--frames", "14",
                   "--seed", "2020", "--save-obj", "--device", DEV])
    for v in range(2):
        vid = f"synthetic{v:02d}_0_0"
        tracks = json.load(open(f"{out}/{vid}_tracks.json"))
        recs = torch.load(f"{out}/{vid}_predictions_opt.pth", weights_only=False)
        assert len(recs) == 14 and any(f.endswith(".obj") for f in __import__("os").listdir(out))
        preds, _ = synth.make_video(2020 + v, 3, 14)
        random.seed(2020 + v)
        planes = opt_utils.track_planes(preds)
        ref = opt_utils.optimize_planes(preds, planes, '3dc', device=DEV)
        want = io.tracks_summary(planes)
        assert [t["has_rot"] for t in tracks] == [t["has_rot"] for t in want]
        for a, b in zip(tracks, want):
            if a["has_rot"]:
                assert [x["angle_id"] for x in a["angle_track"]] == [x["angle_id"] for x in b["angle_track"]]
                assert [x["inter"] for x in a["angle_track"]] == [x["inter"] for x in b["angle_track"]]
        for r, p in zip(recs, ref):
            assert torch.equal(r["pred_rot_axis"], p.pred_rot_axis)
            assert [i["score"] for i in r["instances"]] == list(p.scores)


def test_opt_arti_tool_prints_ap_before_and_after(tmp_path, capsys):
    """--gt-json: the AP table of tools/opt_arti.py (evaluation.evaluate_for_arti_axis) before and after the
    temporal optimisation, on a ground truth derived from the synthetic predictions themselves."""
    import json
    import re
    from articulation3d_b200.axis import angle_offset_to_axis
    from articulation3d_b200.tools import opt_arti
    inp, out, gtp = str(tmp_path / "pred.pth"), str(tmp_path / "out"), str(tmp_path / "gt.json")
    opt_arti.main(["--input", inp, "--output", out, "--synthetic", "1", "--tracks", "1", "--frames", "14",
                   "--seed", "2021", "--device", DEV])
    records = torch.load(inp, weights_only=False)
    images, anns = [], []
    for r in records:
        images.append({"id": r["image_id"], "width": 640, "height": 480})
        if not r["instances"]:
            continue
        ins = r["instances"][0]
        x, y, w, h = ins["bbox"]
        centre = torch.tensor([[x + w / 2, y + h / 2]], dtype=torch.float32)
        line = angle_offset_to_axis(torch.as_tensor(r["pred_rot_axis"][:1]).float(), centre)[0].tolist()
        cls = int(ins["category_id"])
        anns.append({"id": len(anns) + 1, "image_id": r["image_id"], "category_id": cls + 1, "bbox": [x, y, w, h],
                     "rot_axis": line if cls == 0 else None, "tran_axis": None, "normal": None})
    json.dump({"images": images, "annotations": anns,
               "categories": [{"id": 1, "name": "arti_rot"}, {"id": 2, "name": "arti_tran"}]}, open(gtp, "w"))
    capsys.readouterr()
    opt_arti.main(["--input", inp, "--output", out, "--device", DEV, "--gt-json", gtp])
    text = capsys.readouterr().out
    lines = [l for l in text.splitlines() if l.startswith("AP ")]
    assert len(lines) == 2 and lines[0].startswith("AP before") and lines[1].startswith("AP after")
    for l in lines:
        vals = [float(v) for v in re.findall(r"- arti_\w+ ([0-9.]+)", l)]
        assert vals and all(0.0 <= v <= 1.0 for v in vals)
    # the boxes are the ground truth's own: detection AP of the untouched predictions is perfect
    assert re.search(r"bbox - arti_\w+ 1\.0000", lines[0])
