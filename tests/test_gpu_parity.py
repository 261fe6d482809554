"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
committed golden fixtures.  Bit-exact for masks, pixel counts, inter/union and
arg-max indices; 1e-4 relative for floats (BASELINE.json north_star)."""
import random

import numpy as np
import pytest
import torch

from articulation3d_b200 import OptConfig, _lib, engine, geometry, opt_utils, synth
from articulation3d_b200.structures import Boxes, Instances
from oracle import restated
from tests import golden_util as gu

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(autouse=True, params=["exact", "filter"])
def projection_kernel(request, monkeypatch):
    """Every test of this module runs with both projection kernels (the library's default picks one from
    the grid size): the reference chain for every point, and the homography filter with proven truncation."""
    monkeypatch.setenv("A3D_PROJECT_KERNEL", request.param)
    # fresh pass buffers hold 0xAB bytes: nothing may depend on what a workspace buffer held before
    monkeypatch.setenv("A3D_WS_POISON", "1")
    return request.param


def _unpack(bits: torch.Tensor, W: int) -> np.ndarray:
    """(n,H,pitch) int32 device -> (n,H,W) bool numpy"""
    b = bits.cpu().numpy().view(np.uint32)
    u8 = b.view(np.uint8).reshape(b.shape[0], b.shape[1], -1)
    return np.unpackbits(u8, axis=-1, bitorder="little")[..., :W].astype(bool)


def _bbox_ref(m: np.ndarray):
    if not m.any():
        return [0, -1, 0, -1]
    rows = np.where(m.any(1))[0]
    cols = np.where(m.any(0))[0]
    return [rows[0], rows[-1], cols[0] // 32, cols[-1] // 32]


@pytest.mark.parametrize("H,W,dtype", [(480, 640, torch.float32), (75, 100, torch.float32),
                                       (33, 1, torch.float32), (17, 1057, torch.uint8),
                                       (40, 64, torch.bool)])
def test_pack_and_meta(H, W, dtype):
    g = torch.Generator().manual_seed(H * W)
    n = 7
    dense = torch.rand(n, H, W, generator=g)
    dense[0] = 0                                   # empty mask
    dense[1] = 1                                   # full mask
    dense[2, :, : W // 2] = 0
    if dtype == torch.float32:
        m = torch.where(dense > 0.7, dense, torch.zeros(()))   # non-binary: 0 or (0.7,1]
        m[3] = torch.where(dense[3] > 0.5, 0.3, 0.0)           # nonzero but below threshold
    else:
        m = (dense > 0.7).to(dtype)
    pool = engine.pack_masks(m.to(DEV), 0.5, with_nonzero=(dtype == torch.float32))
    want = (m.float() > 0.5).numpy()
    got = _unpack(pool.bits, W)
    assert np.array_equal(got, want)
    # padding bits are zero
    raw = pool.bits.cpu().numpy().view(np.uint32)
    assert raw.shape[2] == _lib.pitch_words(W)
    full = np.unpackbits(raw.view(np.uint8).reshape(n, H, -1), axis=-1, bitorder="little")
    assert not full[..., W:].any()
    assert np.array_equal(pool.popc.cpu().numpy(), want.reshape(n, -1).sum(1))
    assert pool.bbox.cpu().numpy().tolist() == [_bbox_ref(w) for w in want]
    if dtype == torch.float32:
        assert pool.bits_nz is not None
        assert np.array_equal(_unpack(pool.bits_nz, W), (m != 0).numpy())
    back = engine.emit_masks(pool.bits, None, H, W)
    assert np.array_equal(back.cpu().numpy() > 0.5, want)
    idx = torch.tensor([2, 0, 1], dtype=torch.int32, device=DEV)
    assert np.array_equal(engine.emit_masks(pool.bits, idx, H, W, torch.uint8).cpu().numpy(),
                          want[[2, 0, 1]].astype(np.uint8))


def _oracle_job(preds, frame, box, ocfg, mode, grid):
    name = {_lib.MODE_SEQ: "seq", _lib.MODE_COMPOSED: "composed", _lib.MODE_TRANSLATE: "translate"}[mode]
    masks, _, g = restated.candidate_masks(preds[frame], box, ocfg, name, grid)
    return masks.numpy() > 0.5


def _device_job(preds, frame, box, cfg, mode, grid, pool, src_idx, targets, tile=None, want_table=True):
    geo = geometry.source_geometry(preds[frame], box, cfg, mode == _lib.MODE_TRANSLATE)
    if mode == _lib.MODE_TRANSLATE:
        xf = geometry.xforms_translate(grid, geo.dir_vec)
    else:
        R = geometry.rotation_matrices(grid, geo.dir_vec)
        xf = geometry.xforms_seq(R) if mode == _lib.MODE_SEQ else geometry.xforms_composed(R, geo.pivot)
    batch = engine.build_batch([src_idx], [mode], [geo.normal.numpy()], [float(geo.offset)], [geo.pivot],
                               [xf], [targets], pool.source_points)
    res = engine.run_pass(cfg, pool, engine.DeviceBatch(batch, DEV), want_table=want_table, tile_cand=tile)
    torch.cuda.synchronize()
    return res, batch


def _check_pass(res, batch, proj_want, tgt_masks, W):
    """proj_want (A,H,W) bool, tgt_masks (T,H,W) bool."""
    A, T = proj_want.shape[0], tgt_masks.shape[0]
    got = _unpack(res.masks()[:A], W)
    assert got.shape == proj_want.shape
    diff = int((got != proj_want).sum())
    assert diff == 0, f"{diff} projected pixels differ"
    assert np.array_equal(res.proj_popc[:A].cpu().numpy(), proj_want.reshape(A, -1).sum(1))
    assert res.proj_bbox[:A].cpu().numpy().tolist() == [_bbox_ref(p) for p in proj_want]
    inter = (tgt_masks[:, None] & proj_want[None]).reshape(T, A, -1).sum(-1)
    union = (tgt_masks[:, None] | proj_want[None]).reshape(T, A, -1).sum(-1)
    assert np.array_equal(res.inter_tab[:T * A].cpu().numpy().reshape(T, A), inter)
    iou = torch.from_numpy(inter) / torch.from_numpy(union)            # int64/int64 -> fp32, as torch does
    best = iou.argmax(1).numpy()
    assert np.array_equal(res.best_cand[:T].cpu().numpy(), best)
    assert np.array_equal(res.best_inter[:T].cpu().numpy(), inter[np.arange(T), best])
    assert np.array_equal(res.best_union[:T].cpu().numpy(), union[np.arange(T), best])
    assert np.array_equal(res.best_iou[:T].cpu().numpy(), iou.numpy()[np.arange(T), best], equal_nan=True)


@pytest.mark.parametrize("mode", [_lib.MODE_SEQ, _lib.MODE_COMPOSED, _lib.MODE_TRANSLATE])
@pytest.mark.parametrize("tile", [1, 4, None])
@pytest.mark.parametrize("kernel", ["tma", "ldg", "mma"])
def test_project_and_score_match_oracle(mode, tile, kernel, monkeypatch):
    monkeypatch.setenv("A3D_SCORE_KERNEL", kernel)
    cfg, ocfg = OptConfig(), restated.OracleConfig()
    preds, _ = synth.make_video(77, 3, 10, kinds=[0, 1, 0])
    box = 1 if mode == _lib.MODE_TRANSLATE else 0
    grid = {_lib.MODE_SEQ: cfg.rot_cluster_grid, _lib.MODE_COMPOSED: cfg.rot_final_grid,
            _lib.MODE_TRANSLATE: cfg.trans_grid}[mode]
    T = len(preds)
    masks = torch.stack([p.pred_masks[box] for p in preds])
    pool = engine.pack_masks(masks.to(DEV))
    for frame in (0, 6):
        want = _oracle_job(preds, frame, box, ocfg, mode, grid)
        res, batch = _device_job(preds, frame, box, cfg, mode, grid, pool, frame, list(range(T)), tile)
        _check_pass(res, batch, want, masks.numpy() > 0.5, cfg.width)


@pytest.mark.parametrize("kernel,key", [("tma", "packed"), ("ldg", "packed"), ("ldg", "wide"), ("tma", "wide"),
                                        ("mma", "packed"), ("mma", "wide")])
def test_odd_resolution_and_many_candidates(kernel, key, monkeypatch):
    """W not a multiple of 32, scaled intrinsics, a 97-candidate grid, ragged targets."""
    monkeypatch.setenv("A3D_SCORE_KERNEL", kernel)
    monkeypatch.setenv("A3D_SCORE_KEY", key)
    H, W = 150, 200
    cfg = OptConfig.scaled(W, H)
    ocfg = restated.OracleConfig(height=H, width=W, focal_length=cfg.focal_length)
    preds, _ = synth.make_video(5, 2, 9, cfg, kinds=[0, 0])
    grid = np.linspace(-1.2, 2.0, 97)
    masks = torch.cat([torch.stack([p.pred_masks[b] for p in preds]) for b in (0, 1)])
    pool = engine.pack_masks(masks.to(DEV))
    want = _oracle_job(preds, 4, 1, ocfg, _lib.MODE_SEQ, grid)
    tg = [0, 3, 9, 10, 11, 17, 4]
    res, batch = _device_job(preds, 4, 1, cfg, _lib.MODE_SEQ, grid, pool, 9 + 4, tg)
    _check_pass(res, batch, want, (masks.numpy() > 0.5)[tg], W)


@pytest.mark.parametrize("kernel", ["tma", "ldg", "mma"])
def test_edge_cases_empty_source_degenerate_axis_behind_camera(kernel, monkeypatch):
    monkeypatch.setenv("A3D_SCORE_KERNEL", kernel)
    cfg, ocfg = OptConfig(), restated.OracleConfig()
    preds, _ = synth.make_video(8, 1, 10, kinds=[0])
    T = len(preds)
    # (1) empty source mask: every candidate is empty, IoU = 0/|T|; with an empty target 0/0 = NaN -> index 0
    p = synth.clone_preds(preds)
    p[2].pred_masks[0].zero_()
    p[3].pred_masks[0].zero_()
    masks = torch.stack([q.pred_masks[0] for q in p])
    pool = engine.pack_masks(masks.to(DEV))
    want = _oracle_job(p, 2, 0, ocfg, _lib.MODE_SEQ, cfg.rot_cluster_grid)
    assert not want.any()
    res, batch = _device_job(p, 2, 0, cfg, _lib.MODE_SEQ, cfg.rot_cluster_grid, pool, 2, list(range(T)))
    _check_pass(res, batch, want, masks.numpy() > 0.5, cfg.width)
    assert np.isnan(res.best_iou[3].item()) and res.best_cand[3].item() == 0
    # (2a) axis line that misses the image -> fallback end-points [0,0,1,1]
    p = synth.clone_preds(preds)
    p[1].pred_rot_axis[0] = torch.tensor([0.6, 0.8, 50.0])
    geo = geometry.source_geometry(p[1], 0, cfg, False)
    assert geo.pts[0].tolist() == [0, 0, 1, 1]
    want = _oracle_job(p, 1, 0, ocfg, _lib.MODE_COMPOSED, cfg.rot_final_grid)
    masks = torch.stack([q.pred_masks[0] for q in p])
    pool = engine.pack_masks(masks.to(DEV))
    res, batch = _device_job(p, 1, 0, cfg, _lib.MODE_COMPOSED, cfg.rot_final_grid, pool, 1, list(range(T)))
    _check_pass(res, batch, want, masks.numpy() > 0.5, cfg.width)
    # (2b) coincident end-points give a NaN direction, hence NaN transforms: every point
    #      lands on pixel (0,0) (NaN -> INT64_MIN -> clamp 0); same for +-inf / huge entries
    g = restated.source_geometry(p[0], 0, ocfg, False)
    for bad in (np.nan, np.inf, 1e30):
        xf = np.zeros((3, 12), np.float32)
        xf[:, [0, 4, 8]] = 1.0
        xf[1, :9] = bad
        xf[2, 9:] = bad
        pts = restated.transform_composed(g["pcd"], xf[:, :9].reshape(3, 3, 3), xf[:, 9:])
        row, col = restated.project_pixels(pts, ocfg, cfg.height, cfg.width)
        want = restated.splat(row, col, cfg.height, cfg.width).numpy() > 0.5
        b = engine.build_batch([0], [_lib.MODE_COMPOSED], [g["normal"].numpy()], [float(g["offset"])],
                               [np.zeros(3, np.float32)], [xf], [list(range(T))], pool.source_points)
        res = engine.run_pass(cfg, pool, engine.DeviceBatch(b, DEV), want_table=True)
        _check_pass(res, b, want, masks.numpy() > 0.5, cfg.width)
    # (3) plane nearly edge-on: points cross Z <= 0 and project mirrored / clamp to the border
    p = synth.clone_preds(preds)
    p[4].pred_planes[0] = torch.tensor([1.2, 0.05, 0.02])
    for mode, grid in ((_lib.MODE_SEQ, cfg.rot_cluster_grid), (_lib.MODE_TRANSLATE, cfg.trans_grid)):
        want = _oracle_job(p, 4, 0, ocfg, mode, grid)
        res, batch = _device_job(p, 4, 0, cfg, mode, grid, pool, 4, list(range(T)))
        _check_pass(res, batch, want, masks.numpy() > 0.5, cfg.width)


@pytest.mark.parametrize("key", ["packed", "wide"])
def test_many_jobs_one_pass_equals_single_jobs(key, monkeypatch):
    monkeypatch.setenv("A3D_SCORE_KEY", key)        # "wide": arg-max key without the packed count
    cfg = OptConfig()
    preds, _ = synth.make_video(13, 4, 12, kinds=[0, 0, 1, 0])
    masks = torch.cat([torch.stack([p.pred_masks[b] for p in preds]) for b in range(4)])
    pool = engine.pack_masks(masks.to(DEV))
    specs, singles = [], []
    rng = np.random.RandomState(0)
    for b in range(4):
        trans = (b == 2)
        frame = int(rng.randint(12))
        geo = geometry.source_geometry(preds[frame], b, cfg, trans)
        if trans:
            mode, xf = _lib.MODE_TRANSLATE, geometry.xforms_translate(cfg.trans_grid, geo.dir_vec)
        else:
            mode, xf = _lib.MODE_SEQ, geometry.xforms_seq(geometry.rotation_matrices(cfg.rot_cluster_grid, geo.dir_vec))
        tg = sorted(rng.choice(12, size=int(rng.randint(1, 12)), replace=False) + 12 * b)
        specs.append((12 * b + frame, mode, geo.normal.numpy(), float(geo.offset), geo.pivot, xf, list(tg)))
    cols = list(zip(*specs))
    batch = engine.build_batch(*cols, pool.source_points)
    res = engine.run_pass(cfg, pool, engine.DeviceBatch(batch, DEV), want_table=True)
    all_cand, all_inter, all_tab = res.best_cand.cpu().numpy().copy(), res.best_inter.cpu().numpy().copy(), res.inter_tab.cpu().numpy().copy()
    for j, s in enumerate(specs):
        b1 = engine.build_batch(*[[c] for c in s], pool.source_points)
        r1 = engine.run_pass(cfg, pool, engine.DeviceBatch(b1, DEV), want_table=True)
        a, n = int(batch.jobs[j]["tgt_begin"]), int(batch.jobs[j]["n_tgt"])
        assert np.array_equal(all_cand[a:a + n], r1.best_cand.cpu().numpy())
        assert np.array_equal(all_inter[a:a + n], r1.best_inter.cpu().numpy())
        t0, tn = int(batch.jobs[j]["tab_begin"]), n * int(batch.jobs[j]["n_cand"])
        assert np.array_equal(all_tab[t0:t0 + tn], r1.inter_tab.cpu().numpy())


@pytest.fixture(params=["table", "chain"])
def schedule(request, monkeypatch):
    """Both host schedules of the cluster phase: one all-sources device pass + host replay, and one
    device pass per round.  They must give identical results (same RNG draws, same clusters)."""
    monkeypatch.setenv("A3D_SCHEDULE", request.param)
    return request.param


@pytest.mark.parametrize("name", gu.golden_cases())
def test_optimize_planes_matches_reference_golden(name, schedule):
    """End to end through the drop-in API against outputs of the unmodified reference."""
    z = gu.load(name)
    preds = gu.arrays_to_preds(z, Instances, Boxes)
    random.seed(int(z["seed"]))
    planes = opt_utils.track_planes(preds)
    stats = opt_utils.Stats()
    out = opt_utils.optimize_planes(preds, planes, '3dc', device=DEV, stats=stats)
    gu.check_against_golden(z, planes, out)
    assert stats.units_visited > 0 and stats.units_computed >= stats.units_visited
    assert stats.schedule == schedule
    if schedule == "table":                      # cluster table + one final pass per track list
        assert stats.passes <= 4


@pytest.mark.parametrize("seed,n_tracks,n_frames,drop", [(41, 4, 30, 0.0), (42, 5, 24, 0.1)])
def test_optimize_planes_matches_oracle(seed, n_tracks, n_frames, drop, schedule):
    preds, _ = synth.make_video(seed, n_tracks, n_frames, drop_prob=drop)
    a, b = synth.clone_preds(preds), synth.clone_preds(preds)
    random.seed(seed)
    planes = opt_utils.track_planes(a)
    out = opt_utils.optimize_planes(a, planes, '3dc', device=DEV)
    random.seed(seed)
    planes_o = restated.track_planes(b)
    trace = []
    out_o = restated.optimize_planes(b, planes_o, '3dc', trace=trace)
    finals = {(t['kind'], t['track']): t for t in trace if t['phase'] == 'final'}
    for cat in ("trans", "rot"):
        assert len(planes[cat]) == len(planes_o[cat])
        for i, (p, q) in enumerate(zip(planes[cat], planes_o[cat])):
            assert p['ids'] == q['ids'] and p['has_rot'] == q['has_rot']
            np.testing.assert_allclose(p['fit']['rsq'], q['rsqs'], rtol=1e-12, equal_nan=True)
            if not p['has_rot']:
                continue
            assert torch.equal(torch.as_tensor(p['std_axis']), torch.as_tensor(q['std_axis']))
            fin = finals[(cat, i)]
            assert p['fit']['center_frame'] == fin['source']
            assert [v['frame'] for v in fin['visits']] == p['fit']['frames']
            for k, v in enumerate(fin['visits']):
                assert p['fit']['angle_id'][k] == v['angle_id']
                assert p['fit']['inter'][k] == v['inter'][v['angle_id']]
                assert p['fit']['union'][k] == v['union'][v['angle_id']]
                assert np.array_equal(p['fit']['iou'][k], v['iou'][v['angle_id']], equal_nan=True)
            dense = p['reg_masks'].dense(torch.uint8).cpu().numpy().astype(bool)
            for k, f in enumerate(p['fit']['frames']):
                assert np.array_equal(dense[k], q['reg_masks'][f].numpy() > 0.5)
            if 'reg_normals' in q:
                for f in q['reg_normals']:
                    np.testing.assert_allclose(p['reg_normals'][f].numpy(), q['reg_normals'][f].numpy(),
                                               rtol=1e-4, atol=1e-6)
    for x, y in zip(out, out_o):
        assert np.array_equal(x.scores, y.scores)
        np.testing.assert_allclose(x.pred_rot_axis.numpy(), y.pred_rot_axis.numpy(), rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(x.pred_tran_axis.numpy(), y.pred_tran_axis.numpy(), rtol=1e-4, atol=1e-7)


def test_legacy_and_average_methods_match_oracle(schedule):
    preds, _ = synth.make_video(51, 2, 14, kinds=[0, 0])
    for method in ("3d", "average"):
        a, b = synth.clone_preds(preds), synth.clone_preds(preds)
        random.seed(1)
        pa = opt_utils.track_planes(a)['rot']
        oa = opt_utils.optimize_planes(a, pa, method, device=DEV)
        random.seed(1)
        pb = restated.track_planes(b)['rot']
        ob = restated.optimize_planes(b, pb, method)
        for p, q in zip(pa, pb):
            assert torch.equal(torch.as_tensor(p['std_axis']), torch.as_tensor(q['std_axis']))
            if method == "3d":
                assert p['has_rot'] == q['has_rot']
                for f in q.get('reg_masks', {}):
                    assert torch.equal(p['reg_masks'][f] > 0.5, q['reg_masks'][f] > 0.5)
        for x, y in zip(oa, ob):
            assert np.array_equal(np.asarray(x.scores), np.asarray(y.scores))
            assert torch.equal(x.pred_rot_axis, y.pred_rot_axis)


def test_optimize_videos_equals_per_video_calls(schedule):
    vids, seeds = [], [5, 6, 7]
    for s in seeds:
        preds, _ = synth.make_video(100 + s, 3, 16, kinds=[0, 1, 0])
        vids.append(preds)
    singles = []
    for s, preds in zip(seeds, vids):
        p = synth.clone_preds(preds)
        random.seed(s)
        planes = opt_utils.track_planes(p)
        singles.append((opt_utils.optimize_planes(p, planes, '3dc', device=DEV), planes))
    batch_in = []
    for preds in vids:
        p = synth.clone_preds(preds)
        batch_in.append((p, opt_utils.track_planes(p)))
    stats = opt_utils.Stats()
    outs = opt_utils.optimize_videos(batch_in, seeds, device=DEV, stats=stats)
    assert stats.passes < sum(1 for _ in range(3)) * 40
    for (o1, pl1), o2, (_, pl2) in zip(singles, outs, batch_in):
        for cat in ("trans", "rot"):
            for p, q in zip(pl1[cat], pl2[cat]):
                assert p['has_rot'] == q['has_rot']
                if p['has_rot']:
                    assert np.array_equal(p['fit']['angle_id'], q['fit']['angle_id'])
                    assert np.array_equal(p['fit']['inter'], q['fit']['inter'])
        for x, y in zip(o1, o2):
            assert np.array_equal(x.scores, y.scores)
            assert torch.equal(x.pred_rot_axis, y.pred_rot_axis)
            assert torch.equal(x.pred_tran_axis, y.pred_tran_axis)


@pytest.mark.parametrize("variant", ["lazy_tracking", "python_upload", "dma_descriptors", "uint8_masks"])
def test_optimize_videos_host_paths_agree(variant, monkeypatch):
    """The batch API's host-side variants — videos tracked inside the pipeline (``(preds, None)``), the upload
    loop in Python instead of a3d_upload_masks, the pass descriptors by DMA instead of a3d_fetch_host_block,
    uint8 masks — give the records of the default path."""
    seeds = [11, 12, 13, 14]
    clips = [synth.make_video(300 + s, 3, 14, kinds=[0, 1, 0])[0] for s in seeds]

    def run(lazy, as_u8=False):
        vids = []
        for c in clips:
            p = synth.clone_preds(c)
            for q in p:
                m = q.pred_masks.cpu()
                q.pred_masks = ((m > 0.5).to(torch.uint8) if as_u8 else m).pin_memory()
            vids.append((p, None if lazy else opt_utils.track_planes(p)))
        outs = opt_utils.optimize_videos(vids, seeds, device=DEV)
        return outs, [pl for _, pl in vids]

    base_o, base_pl = run(False)
    if variant == "python_upload":
        monkeypatch.setenv("A3D_UPLOAD", "python")
    elif variant == "dma_descriptors":
        monkeypatch.setenv("A3D_DESC_COPY", "dma")
    o, pl = run(variant == "lazy_tracking", as_u8=(variant == "uint8_masks"))
    for pa, pb in zip(base_pl, pl):
        for cat in ("trans", "rot"):
            assert len(pa[cat]) == len(pb[cat])
            for a, b in zip(pa[cat], pb[cat]):
                assert a['ids'] == b['ids'] and a['has_rot'] == b['has_rot']
                if a['has_rot']:
                    for k in ("angle_id", "inter", "union"):
                        assert np.array_equal(a['fit'][k], b['fit'][k])
    for oa, ob in zip(base_o, o):
        for x, y in zip(oa, ob):
            assert np.array_equal(np.asarray(x.scores), np.asarray(y.scores))
            assert torch.equal(x.pred_rot_axis, y.pred_rot_axis) and torch.equal(x.pred_tran_axis, y.pred_tran_axis)


def _table_properties(res, batch, pool):
    """Size-independent invariants of one pass (used where the oracle is too slow)."""
    popc = pool.popc.cpu().numpy()
    tab = res.inter_tab.cpu().numpy()
    cand, inter, union, iou = (res.best_cand.cpu().numpy(), res.best_inter.cpu().numpy(),
                               res.best_union.cpu().numpy(), res.best_iou.cpu().numpy())
    ppop = res.proj_popc.cpu().numpy()
    for j in range(batch.n_jobs):
        jb = batch.jobs[j]
        T, A = int(jb["n_tgt"]), int(jb["n_cand"])
        t = tab[int(jb["tab_begin"]): int(jb["tab_begin"]) + T * A].reshape(T, A).astype(np.int64)
        pt = popc[batch.tgt_index[int(jb["tgt_begin"]): int(jb["tgt_begin"]) + T]].astype(np.int64)
        pp = ppop[int(jb["cand_begin"]): int(jb["cand_begin"]) + A].astype(np.int64)
        assert (t >= 0).all() and (t <= np.minimum(pt[:, None], pp[None, :])).all()
        u = pt[:, None] + pp[None, :] - t
        want_iou = (torch.from_numpy(t) / torch.from_numpy(u))
        best = want_iou.argmax(1).numpy()
        sl = slice(int(jb["tgt_begin"]), int(jb["tgt_begin"]) + T)
        assert np.array_equal(cand[sl], best)
        assert np.array_equal(inter[sl], t[np.arange(T), best]) and np.array_equal(union[sl], u[np.arange(T), best])
        assert np.array_equal(iou[sl], want_iou.numpy()[np.arange(T), best], equal_nan=True)
        # the splat can only lose pixels (collisions), never create them
        assert (pp <= int(pool.source_points[int(jb["src_mask"])])).all()


@pytest.mark.parametrize("wl_name", ["c3_mini", "c4_probe"])
def test_full_size_properties_and_kernel_agreement(wl_name, monkeypatch):
    """BASELINE-sized grids (180 candidates x 120 frames; 720 candidates at 1024x768): invariants of
    the result tables, identical tables from both scoring kernels and any candidate tiling, and the
    C oracle on a sampled job."""
    from articulation3d_b200 import workloads
    from oracle import c_oracle
    wl = workloads.WORKLOADS.get(wl_name) or workloads.Workload("c4_probe", "2 videos x 2 tracks x 24 frames, 720 candidates, 1024x768",
                                                                2, 2, 24, 720, 1024, 768)
    if wl_name == "c3_mini":
        wl = workloads.Workload("c3_probe", "2 videos x 8 tracks x 120 frames, 180 candidates", 2, 8, 120, 180)
    inp = workloads.build_pass(wl, 77, DEV)
    runs = {}
    for kernel, tile in (("ldg", None), ("tma", 1), ("ldg", 2), ("mma", None)):
        monkeypatch.setenv("A3D_SCORE_KERNEL", kernel)
        res = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, want_table=True, tile_cand=tile)
        torch.cuda.synchronize()
        runs[(kernel, tile)] = (res.inter_tab.cpu().numpy().copy(), res.best_cand.cpu().numpy().copy(),
                                res.proj_popc.cpu().numpy().copy())
        if kernel == "ldg" and tile is None:
            _table_properties(res, inp.batch, inp.pool)
            # C oracle on the last job
            j = inp.batch.n_jobs - 1
            jb = inp.batch.jobs[j]
            cfg = inp.cfg
            A, T = int(jb["n_cand"]), int(jb["n_tgt"])
            bits = inp.pool.bits.cpu().numpy().view(np.uint32)
            xf = inp.batch.xform[int(jb["cand_begin"]): int(jb["cand_begin"]) + A]
            proj = c_oracle.project(cfg.K_inv(), cfg.focal_length, cfg.cx, cfg.cy, cfg.height, cfg.width,
                                    bits[int(jb["src_mask"])], jb["normal"], float(jb["offset"]), jb["pivot"],
                                    int(jb["mode"]), xf)
            got = res.masks(torch.arange(int(jb["cand_begin"]), int(jb["cand_begin"]) + A, device=DEV)).cpu().numpy().view(np.uint32)
            assert np.array_equal(got, proj)
            tg = inp.batch.tgt_index[int(jb["tgt_begin"]): int(jb["tgt_begin"]) + T]
            inter, uni, best, iou = c_oracle.score(cfg.height, cfg.width, bits[tg], proj)
            tab = res.inter_tab[int(jb["tab_begin"]): int(jb["tab_begin"]) + T * A].cpu().numpy().reshape(T, A)
            assert np.array_equal(tab, inter)
            assert np.array_equal(res.best_cand[int(jb["tgt_begin"]): int(jb["tgt_begin"]) + T].cpu().numpy(), best)
    ref = runs[("ldg", None)]
    for k, v in runs.items():
        assert np.array_equal(v[0], ref[0]) and np.array_equal(v[1], ref[1]) and np.array_equal(v[2], ref[2]), k


def test_c2_shape_with_90_angle_grid_matches_oracle():
    """BASELINE config 2 shape (60 frames, 90-candidate cluster grid / 60-candidate final grid) through
    the drop-in API with a lifted config, against the CPU oracle with the same config."""
    from articulation3d_b200 import workloads
    wl = workloads.WORKLOADS["c2"]
    preds, cfg = workloads.make_clip(wl, 2021, tracks=2, kinds=[synth.KIND_ROT, synth.KIND_TRANS])
    assert len(cfg.rot_cluster_grid) == 90 and len(preds) == 60
    ocfg = restated.OracleConfig(rot_cluster_grid=cfg.rot_cluster_grid, rot_final_grid=cfg.rot_final_grid,
                                 trans_grid=cfg.trans_grid)
    a, b = synth.clone_preds(preds), synth.clone_preds(preds)
    random.seed(9)
    planes = opt_utils.track_planes(a, cfg)
    stats = opt_utils.Stats()
    out = opt_utils.optimize_planes(a, planes, '3dc', cfg=cfg, device=DEV, stats=stats)
    random.seed(9)
    planes_o = restated.track_planes(b)
    trace = []
    out_o = restated.optimize_planes(b, planes_o, '3dc', cfg=ocfg, trace=trace)
    units = sum(len(v['iou']) for t in trace for v in t['visits'])
    assert stats.units_visited == units                                  # same accounting as the reference loops
    finals = {(t['kind'], t['track']): t for t in trace if t['phase'] == 'final'}
    for cat in ("trans", "rot"):
        for i, (p, q) in enumerate(zip(planes[cat], planes_o[cat])):
            assert p['has_rot'] == q['has_rot']
            if p['has_rot']:
                fin = finals[(cat, i)]
                assert [v['angle_id'] for v in fin['visits']] == p['fit']['angle_id'].tolist()
                assert [int(v['inter'][v['angle_id']]) for v in fin['visits']] == p['fit']['inter'].tolist()
    for x, y in zip(out, out_o):
        assert np.array_equal(x.scores, y.scores)
        np.testing.assert_allclose(x.pred_rot_axis.numpy(), y.pred_rot_axis.numpy(), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("kernel", ["ldg", "mma"])
@pytest.mark.parametrize("seed", range(8))
def test_randomised_shapes_against_c_oracle(seed, kernel, monkeypatch):
    """Random image sizes (W not a multiple of 32, tiny and wide), random blob masks, random planes,
    pivots and rigid transforms in all three modes, ragged target lists: the CUDA pass must equal the
    C restatement bit for bit (projected masks, inter/union tables, arg-max)."""
    from oracle import c_oracle
    monkeypatch.setenv("A3D_SCORE_KERNEL", kernel)
    rng = np.random.RandomState(1000 + seed)
    H = int(rng.choice([7, 33, 64, 120, 200]))
    W = int(rng.choice([5, 31, 32, 33, 100, 257, 640]))
    cfg = OptConfig.scaled(W, H) if rng.rand() < 0.5 else OptConfig(height=H, width=W)
    n = 6
    masks = np.zeros((n, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(n):
        for _ in range(rng.randint(1, 4)):
            cy, cx = rng.rand() * H, rng.rand() * W
            ry, rx = 1 + rng.rand() * H / 2, 1 + rng.rand() * W / 2
            masks[i] += (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1)
    masks = (masks > 0).astype(np.float32)
    masks[n - 1] = 0                                             # an empty mask (source P = 0 / empty target)
    pool = engine.pack_masks(torch.from_numpy(masks).to(DEV))
    bits_ref = c_oracle.pack(masks)
    assert np.array_equal(pool.bits.cpu().numpy().view(np.uint32), bits_ref)

    def rand_rot(k):
        ax = rng.randn(k, 3)
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        ang = rng.uniform(-np.pi, np.pi, k)
        return geometry._axis_angle_to_matrix(torch.from_numpy(ax * ang[:, None])).to(torch.float32).numpy().reshape(-1, 3, 3)

    specs = []
    for j in range(5):
        mode = int(rng.randint(3))
        A = int(rng.randint(1, 23))
        xf = np.zeros((A, 12), np.float32)
        xf[:, :9] = rand_rot(A).reshape(A, 9)
        xf[:, 9:] = rng.randn(A, 3).astype(np.float32) * 0.3
        normal = rng.randn(3)
        normal[2] += 2.0 * (1 if rng.rand() < 0.8 else -1)       # mostly facing the camera, sometimes behind
        normal = (normal / np.linalg.norm(normal)).astype(np.float32)
        offset = np.float32(rng.uniform(0.5, 3.0))
        pivot = rng.randn(3).astype(np.float32)
        src = int(rng.randint(n)) if j else n - 1                # job 0: empty source
        tg = [int(t) for t in rng.choice(n, size=rng.randint(1, n + 1), replace=True)]
        specs.append((src, mode, normal, float(offset), pivot, xf, tg))
    batch = engine.build_batch(*zip(*specs), pool.source_points)
    res = engine.run_pass(cfg, pool, engine.DeviceBatch(batch, DEV), want_table=True)
    torch.cuda.synchronize()
    got_bits = res.masks().cpu().numpy().view(np.uint32)
    tab = res.inter_tab.cpu().numpy()
    for j, (src, mode, normal, offset, pivot, xf, tg) in enumerate(specs):
        jb = batch.jobs[j]
        A, T = len(xf), len(tg)
        want = c_oracle.project(cfg.K_inv(), cfg.focal_length, cfg.cx, cfg.cy, H, W, bits_ref[src], normal, offset,
                                pivot, mode, xf)
        c0 = int(jb["cand_begin"])
        assert np.array_equal(got_bits[c0:c0 + A], want), (seed, j, mode)
        inter, uni, best, iou = c_oracle.score(H, W, bits_ref[tg], want)
        t0, b0 = int(jb["tgt_begin"]), int(jb["tab_begin"])
        assert np.array_equal(tab[b0:b0 + T * A].reshape(T, A), inter)
        assert np.array_equal(res.best_cand[t0:t0 + T].cpu().numpy(), best)
        assert np.array_equal(res.best_union[t0:t0 + T].cpu().numpy(), uni[np.arange(T), best])
        assert np.array_equal(res.best_iou[t0:t0 + T].cpu().numpy(), iou, equal_nan=True)


@pytest.mark.parametrize("key", ["packed", "wide"])
def test_mma_kernel_tiles_targets_and_candidates(key, monkeypatch):
    """More targets than one tensor-core tile holds (128) and more candidates than one tile (240):
    two target tiles x two candidate tiles per job meet in the per-target arg-max key; a second job
    with an empty source shares the pass.  Checked against the C restatement."""
    from oracle import c_oracle
    monkeypatch.setenv("A3D_SCORE_KERNEL", "mma")
    monkeypatch.setenv("A3D_SCORE_KEY", key)
    rng = np.random.RandomState(42)
    H, W = 64, 100
    cfg = OptConfig.scaled(W, H)
    n = 7
    yy, xx = np.mgrid[0:H, 0:W]
    masks = np.zeros((n, H, W), np.float32)
    for i in range(n - 1):
        cy, cx = rng.rand() * H, rng.rand() * W
        masks[i] = (((yy - cy) / (4 + rng.rand() * H / 2)) ** 2 + ((xx - cx) / (4 + rng.rand() * W / 2)) ** 2 < 1)
    pool = engine.pack_masks(torch.from_numpy(masks).to(DEV))
    bits_ref = c_oracle.pack(masks)
    A = 250
    ax = rng.randn(A, 3)
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rng.uniform(-0.6, 0.6, A)
    xf = np.zeros((A, 12), np.float32)
    xf[:, :9] = geometry._axis_angle_to_matrix(torch.from_numpy(ax * ang[:, None])).to(torch.float32).numpy().reshape(A, 9)
    xf[:, 9:] = rng.randn(A, 3).astype(np.float32) * 0.1
    normal = np.array([0.1, -0.2, 1.0]) / np.linalg.norm([0.1, -0.2, 1.0])
    specs = []
    for src, T in ((0, 150), (n - 1, 131)):
        tg = [int(t) for t in rng.choice(n, size=T, replace=True)]
        specs.append((src, _lib.MODE_COMPOSED, normal.astype(np.float32), 2.0, np.zeros(3, np.float32), xf, tg))
    batch = engine.build_batch(*zip(*specs), pool.source_points)
    res = engine.run_pass(cfg, pool, engine.DeviceBatch(batch, DEV), want_table=True)
    torch.cuda.synchronize()
    tab = res.inter_tab.cpu().numpy()
    for j, (src, mode, nrm, offset, pivot, xf_j, tg) in enumerate(specs):
        jb = batch.jobs[j]
        T = len(tg)
        want = c_oracle.project(cfg.K_inv(), cfg.focal_length, cfg.cx, cfg.cy, H, W, bits_ref[src], nrm, offset,
                                pivot, mode, xf_j)
        inter, uni, best, iou = c_oracle.score(H, W, bits_ref[tg], want)
        t0, b0 = int(jb["tgt_begin"]), int(jb["tab_begin"])
        assert np.array_equal(tab[b0:b0 + T * A].reshape(T, A), inter)
        assert np.array_equal(res.best_cand[t0:t0 + T].cpu().numpy(), best)
        assert np.array_equal(res.best_inter[t0:t0 + T].cpu().numpy(), inter[np.arange(T), best])
        assert np.array_equal(res.best_union[t0:t0 + T].cpu().numpy(), uni[np.arange(T), best])
        assert np.array_equal(res.best_iou[t0:t0 + T].cpu().numpy(), iou, equal_nan=True)


@pytest.mark.parametrize("pdl", ["0", "1"])
def test_single_call_pass_equals_project_then_score(pdl, monkeypatch):
    """a3d_pass (keys cleared first, optionally programmatic dependent launches) must give exactly what
    a3d_project followed by a3d_score gives, including the full intersection table."""
    from articulation3d_b200 import workloads
    wl = workloads.Workload("pass_probe", "3 videos x 4 tracks x 24 frames, 45 candidates", 3, 4, 24, 45)
    inp = workloads.build_pass(wl, 5, DEV)
    monkeypatch.setenv("A3D_PASS_API", "split")
    a = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, want_table=True)
    torch.cuda.synchronize()
    want = [t.clone() for t in (a.best_cand, a.best_inter, a.best_union, a.best_iou.view(torch.int32), a.inter_tab,
                                a.proj_popc, a.proj_bbox)]
    monkeypatch.delenv("A3D_PASS_API")
    monkeypatch.setenv("A3D_PDL", pdl)
    for _ in range(3):                                   # back to back: dependent launches across passes
        b = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, want_table=True)
    torch.cuda.synchronize()
    got = [b.best_cand, b.best_inter, b.best_union, b.best_iou.view(torch.int32), b.inter_tab, b.proj_popc, b.proj_bbox]
    for w, g in zip(want, got):
        assert torch.equal(w, g)


@pytest.mark.parametrize("mode", [_lib.MODE_SEQ, _lib.MODE_COMPOSED, _lib.MODE_TRANSLATE])
@pytest.mark.parametrize("shape", ["640x480", "1024x768"])
def test_filtered_projection_equals_exact_chain(mode, shape, monkeypatch):
    """k_project<filter> (homography + proven truncation, exact chain on demand) against k_project<exact>
    (the reference's fp32 chain for every point): projected masks, counts, boxes and the scores bit for bit,
    over ~10^8 (point, candidate) pairs per case, any candidate tiling."""
    from articulation3d_b200 import workloads
    wl = workloads.Workload("probe", "4 videos x 4 tracks x 16 frames, 96 candidates", 4, 4, 16, 96) if shape == "640x480" \
        else workloads.Workload("probe", "2 videos x 2 tracks x 12 frames, 144 candidates, 1024x768", 2, 2, 12, 144, 1024, 768)
    inp = workloads.build_pass(wl, 300 + mode, DEV, mode=mode)
    out = {}
    for kernel, tile in (("exact", None), ("filter", None), ("filter", 1), ("filter", 2), ("filter/persistent", 1),
                         ("exact/persistent", 1)):
        # ".../persistent": CTAs of two 512-thread groups that fetch tiles from a counter (A3D_PROJECT_SCHED)
        monkeypatch.setenv("A3D_PROJECT_KERNEL", kernel.split("/")[0])
        monkeypatch.setenv("A3D_PROJECT_SCHED", "persistent" if "/" in kernel else "cta")
        res = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, want_table=True, tile_cand=tile)
        torch.cuda.synchronize()
        out[(kernel, tile)] = [t.cpu().numpy().copy() for t in (res.masks(), res.proj_popc, res.proj_bbox,
                                                                 res.inter_tab, res.best_cand)]
    ref = out[("exact", None)]
    assert ref[1].sum() > 0
    for k, v in out.items():
        for a, b in zip(v, ref):
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("seed", range(6))
def test_filter_kernel_adversarial_geometry(seed, monkeypatch):
    """Filter against exact chain where the error bound is under stress: grazing planes, planes a few
    centimetres from the camera, pivots far off the surface so that points swing through w = 0 and behind
    the camera, full-circle rotations, translations of metres, huge and tiny focal lengths, big masks.
    Bit-identical projected masks, counts and boxes are required for every candidate."""
    rng = np.random.RandomState(7000 + seed)
    H, W = [(480, 640), (768, 1024), (120, 200)][seed % 3]
    cfg = OptConfig.scaled(W, H) if seed % 2 == 0 else OptConfig(height=H, width=W,
                                                                  focal_length=float(rng.choice([40.0, 517.97, 6000.0])))
    n = 5
    yy, xx = np.mgrid[0:H, 0:W]
    masks = np.zeros((n, H, W), np.float32)
    for i in range(n):
        cy, cx = rng.uniform(0.2, 0.8) * H, rng.uniform(0.2, 0.8) * W
        ry, rx = rng.uniform(0.1, 0.6) * H, rng.uniform(0.1, 0.6) * W
        masks[i] = (np.abs(yy - cy) < ry) & (np.abs(xx - cx) < rx)
    masks[1, ::2] = 0                                             # ragged rows
    masks[2] *= (rng.rand(H, W) < 0.5)                            # salt and pepper
    pool = engine.pack_masks(torch.from_numpy(masks).to(DEV))
    specs = []
    for j in range(10):
        mode = j % 3
        A = 24
        ax = rng.randn(A, 3)
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        ang = rng.uniform(-np.pi, np.pi, A) * (1e-3 if j == 3 else 1.0)     # job 3: rotations of a milliradian
        xf = np.zeros((A, 12), np.float32)
        xf[:, :9] = geometry._axis_angle_to_matrix(torch.from_numpy(ax * ang[:, None])).to(torch.float32).numpy().reshape(A, 9)
        xf[:, 9:] = (rng.randn(A, 3) * rng.choice([1e-3, 0.3, 5.0])).astype(np.float32)
        kind = j % 5
        normal = rng.randn(3)
        if kind == 0:
            normal[2] = 1e-3 * rng.randn()                         # grazing: the plane contains the viewing direction
        elif kind == 1:
            normal = np.array([0.0, 0.0, 1.0]) + 1e-4 * rng.randn(3)   # fronto-parallel
        normal = (normal / np.linalg.norm(normal)).astype(np.float32)
        offset = np.float32(rng.choice([0.02, 0.5, 2.0, 50.0]))
        pivot = (rng.randn(3) * rng.choice([0.1, 3.0, 100.0])).astype(np.float32)
        specs.append((int(rng.randint(n)), mode, normal, float(offset), pivot, xf, [0]))
    batch = engine.build_batch(*zip(*specs), pool.source_points)
    out = {}
    for kernel in ("exact", "filter"):
        monkeypatch.setenv("A3D_PROJECT_KERNEL", kernel)
        res = engine.run_pass(cfg, pool, engine.DeviceBatch(batch, DEV))
        torch.cuda.synchronize()
        out[kernel] = [t.cpu().numpy().copy() for t in (res.masks(), res.proj_popc, res.proj_bbox, res.best_inter)]
    for a, b in zip(out["exact"], out["filter"]):
        assert np.array_equal(a, b)
    assert out["exact"][1].max() > 0


def test_bbox_rows_output_equals_full_output(monkeypatch):
    """A3D_OUT_BBOX_ROWS (the engine's default: a workspace keeps its projected-mask buffers zero outside each
    slot's box, and a pass writes only the rows of the slot's old and new box) against A3D_OUT_FULL on fresh
    buffers: identical masks — every word — statistics and scores, also when the same workspace serves passes
    of different geometry, of fewer and then of more candidate slots one after the other."""
    from articulation3d_b200 import workloads
    ws = engine.Workspace(DEV)
    shapes = [(3, 3, 14, 40, 77, _lib.MODE_SEQ), (3, 3, 14, 40, 78, _lib.MODE_COMPOSED), (1, 2, 12, 16, 79, _lib.MODE_SEQ),
              (4, 3, 10, 56, 80, _lib.MODE_TRANSLATE), (3, 3, 14, 40, 77, _lib.MODE_SEQ)]
    for videos, tracks, frames, cand, seed, mode in shapes:
        wl = workloads.Workload("probe", "probe", videos, tracks, frames, cand)
        inp = workloads.build_pass(wl, seed, DEV, mode=mode)
        full = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, engine.Workspace(DEV), want_table=True, out_mode=_lib.OUT_FULL)
        rows = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, ws, want_table=True, out_mode=_lib.OUT_BBOX_ROWS)
        torch.cuda.synchronize()
        assert rows.rows_only and not full.rows_only
        for name in ("proj_bits", "proj_popc", "proj_bbox", "inter_tab", "best_cand", "best_inter", "best_union"):
            assert torch.equal(getattr(full, name), getattr(rows, name)), (name, seed)
        assert torch.equal(full.best_iou.view(torch.int32), rows.best_iou.view(torch.int32))
        assert int(full.proj_popc.sum()) > 0
    # slots beyond the last pass still hold masks that are zero outside their boxes
    key, bits, popc, bbox = ws._bufs["_proj"]
    b, bb = bits.cpu().numpy(), bbox.cpu().numpy()
    for k in range(len(bb)):
        inside = np.zeros(b.shape[1], bool)
        if bb[k, 1] >= bb[k, 0]:
            inside[bb[k, 0]: bb[k, 1] + 1] = True
        assert not b[k][~inside].any(), k


@pytest.mark.parametrize("name,videos,tracks,frames,cand,W,H,mode", [
    ("c3_slice", 4, 8, 120, 180, 640, 480, _lib.MODE_SEQ),              # 32 jobs of configs[2]'s shape
    ("c4_slice", 2, 4, 24, 720, 1024, 768, _lib.MODE_SEQ),              # configs[3]: 720 rotations at 1024x768
    ("c4_trans_slice", 2, 4, 24, 20, 1024, 768, _lib.MODE_TRANSLATE),   # ... and its 20 translation candidates
])
def test_c_oracle_on_every_job_of_baseline_shaped_slices(name, videos, tracks, frames, cand, W, H, mode):
    """BASELINE-shaped passes against the C oracle on EVERY job (not a sample): projected masks, the full
    intersection table, unions, arg-max and the IoU bits — exact.  The library picks the kernels it would
    pick in production (filter / tensor-core scoring on the big grids) unless the module fixture forces one."""
    from articulation3d_b200 import workloads
    from oracle import c_oracle
    wl = workloads.Workload(name, name, videos, tracks, frames, cand, W, H)
    inp = workloads.build_pass(wl, 500 + cand, DEV, mode=mode)
    res = engine.run_pass(inp.cfg, inp.pool, inp.dbatch, want_table=True)
    torch.cuda.synchronize()
    cfg = inp.cfg
    bits = inp.pool.bits.cpu().numpy().view(np.uint32)
    got_all = res.masks().cpu().numpy().view(np.uint32)
    tab_all = res.inter_tab.cpu().numpy()
    cand_all, inter_all, union_all = (t.cpu().numpy() for t in (res.best_cand, res.best_inter, res.best_union))
    iou_all = res.best_iou.cpu().numpy()
    assert inp.batch.n_jobs == videos * tracks
    nonempty = 0
    for jb in inp.batch.jobs:
        A, T = int(jb["n_cand"]), int(jb["n_tgt"])
        c0, t0, b0 = int(jb["cand_begin"]), int(jb["tgt_begin"]), int(jb["tab_begin"])
        proj = c_oracle.project(cfg.K_inv(), cfg.focal_length, cfg.cx, cfg.cy, cfg.height, cfg.width,
                                bits[int(jb["src_mask"])], jb["normal"], float(jb["offset"]), jb["pivot"],
                                int(jb["mode"]), inp.batch.xform[c0:c0 + A])
        assert np.array_equal(got_all[c0:c0 + A], proj)
        nonempty += int(proj.any())
        tg = inp.batch.tgt_index[t0:t0 + T]
        inter, uni, best, iou = c_oracle.score(cfg.height, cfg.width, bits[tg], proj)
        assert np.array_equal(tab_all[b0:b0 + T * A].reshape(T, A), inter)
        assert np.array_equal(cand_all[t0:t0 + T], best)
        assert np.array_equal(inter_all[t0:t0 + T], inter[np.arange(T), best])
        assert np.array_equal(union_all[t0:t0 + T], uni[np.arange(T), best])
        assert np.array_equal(iou_all[t0:t0 + T].view(np.uint32), iou.view(np.uint32))
    assert nonempty == inp.batch.n_jobs
