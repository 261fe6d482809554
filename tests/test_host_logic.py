"""CPU tests of the host-side logic and of the C-ABI library's surface (no GPU)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from articulation3d_b200 import OptConfig, _lib, axis, engine, geometry, opt_utils, synth
from articulation3d_b200.structures import Boxes, Instances, pairwise_iou
from oracle import restated

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "a3d.h")).read()
    declared = set(re.findall(r"\b(a3d_[a-z_0-9]+)\s*\(", header))
    declared = {d for d in declared if not d.endswith("_t")}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.a3d_version() == 100
    for W in (1, 31, 32, 33, 100, 640, 1024, 1025):
        assert lib.a3d_pitch_words(W) == _lib.pitch_words(W)
        assert _lib.pitch_words(W) % 4 == 0 and _lib.pitch_words(W) * 32 >= W
    assert ctypes.sizeof(_lib.Camera) == 9 * 8 + 3 * 4 + 2 * 4 + 4   # padded to 8


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    preds, _ = synth.make_video(1, 1, 12, kinds=[0])
    planes = opt_utils.track_planes(preds)
    with pytest.raises(_lib.A3DError):
        opt_utils.optimize_planes(preds, planes, '3dc')
    with pytest.raises(_lib.A3DError):
        engine.pack_masks(torch.zeros(1, 8, 8))
    # the host-path entry points of the batch API fail with a status and a message, they do not crash or hang
    import ctypes as C
    lib = _lib.load()
    buf = np.zeros(64, np.uint8)
    assert lib.a3d_fetch_host_block(buf.ctypes.data, buf.ctypes.data, 64, None) < 0
    assert lib.a3d_last_error_string()
    ptrs, counts = (C.c_void_p * 1)(buf.ctypes.data), (C.c_int64 * 1)(1)
    assert lib.a3d_upload_masks(ptrs, counts, 1, _lib.A3D_U8, 8, 8, 0.5, buf.ctypes.data, 1, buf.ctypes.data, None, 2, None) < 0
    with pytest.raises(_lib.A3DError):
        opt_utils.optimize_videos([(preds, None)], [1])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "articulation3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


@pytest.mark.parametrize("seed", range(6))
def test_axis_functions_match_oracle(seed):
    g = torch.Generator().manual_seed(seed)
    n = 64
    ang = torch.rand(n, generator=g) * 2 * np.pi
    ao = torch.stack([torch.sin(ang), torch.cos(ang), torch.rand(n, generator=g) * 3 - 0.5], 1)
    ao[0, 0] = 0.0            # sin == 0 branch
    ao[1, 1] = 0.0            # horizontal line
    ao[2] = torch.tensor([0.6, 0.8, 50.0])   # misses the image -> fallback
    ce = torch.rand(n, 2, generator=g) * torch.tensor([640.0, 480.0])
    got = axis.angle_offset_to_axis(ao, ce)
    want = restated.angle_offset_to_axis(ao.numpy(), ce.numpy())
    assert torch.equal(got, want)
    lines = got.tolist()
    a = axis.axis_to_angle_offset(lines, ce)
    b = restated.axis_to_angle_offset(lines, ce)
    assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))


def test_track_planes_matches_oracle():
    for seed, drop in ((4, 0.0), (9, 0.15), (10, 0.3)):
        preds, _ = synth.make_video(seed, 5, 26, drop_prob=drop)
        a = opt_utils.track_planes(preds)
        b = restated.track_planes(preds)
        for cat in ("rot", "trans"):
            assert [p['ids'] for p in a[cat]] == [p['ids'] for p in b[cat]]
            assert [p['latest_frame'] for p in a[cat]] == [p['latest_frame'] for p in b[cat]]


def test_pairwise_iou_and_instances():
    b1 = Boxes(torch.tensor([[0., 0., 10., 10.], [5., 5., 6., 6.]]))
    b2 = Boxes(torch.tensor([[5., 5., 15., 15.], [20., 20., 30., 30.]]))
    iou = pairwise_iou(b1, b2)
    assert iou.shape == (2, 2) and abs(iou[0, 0].item() - 25 / 175) < 1e-7 and iou[0, 1] == 0
    inst = Instances((480, 640))
    inst.scores = np.ones(2)
    with pytest.raises(AssertionError):
        inst.pred_classes = np.ones(3)
    with pytest.raises(AttributeError):
        inst.nope


def test_source_geometry_and_transforms_match_oracle():
    cfg, ocfg = OptConfig(), restated.OracleConfig()
    preds, _ = synth.make_video(21, 3, 12, kinds=[0, 1, 0])
    for frame, box, trans in ((0, 0, False), (5, 2, False), (7, 1, True)):
        geo = geometry.source_geometry(preds[frame], box, cfg, trans, all_boxes=True)
        ref = restated.source_geometry(preds[frame], box, ocfg, trans)
        assert torch.equal(geo.normal, ref["normal"]) and torch.equal(geo.offset, ref["offset"])
        assert torch.equal(geo.pts, ref["pts"])
        one = geometry.source_geometry(preds[frame], box, cfg, trans)          # only the needed row
        assert torch.equal(one.pts[box], ref["pts"][box]) and np.array_equal(one.dir_vec, geo.dir_vec)
        assert np.array_equal(geo.axis3d, ref["axis3d"]) and np.array_equal(geo.dir_vec, ref["dir_vec"])
        R = geometry.rotation_matrices(cfg.rot_cluster_grid, geo.dir_vec)
        assert np.array_equal(R, restated.rotation_matrices(ocfg.rot_cluster_grid, ref["dir_vec"]))
        a = ref["axis3d"][0].astype(np.float32)
        xf = geometry.xforms_composed(R, geo.pivot)
        assert np.array_equal(xf[:, 9:], restated.composed_last_row(a, R))
        xt = geometry.xforms_translate(cfg.trans_grid, geo.dir_vec)
        assert np.array_equal(xt[:, 9:], restated.translation_vectors(ocfg.trans_grid, ref["dir_vec"]))
        assert torch.equal(geometry.transform_normals(geo.normal, R), restated.transform_normals(ref["normal"], R))
    # batched form used by the benchmark: (S, A, ...) equals S single calls
    d = np.stack([geometry.source_geometry(preds[f], 0, cfg, False).dir_vec for f in range(4)])
    Rb = geometry.rotation_matrices(cfg.rot_final_grid, d)
    for i in range(4):
        assert np.array_equal(Rb[i], geometry.rotation_matrices(cfg.rot_final_grid, d[i]))
    piv = np.random.RandomState(0).randn(4, 3).astype(np.float32)
    xb = geometry.xforms_composed(Rb, piv)
    for i in range(4):
        assert np.array_equal(xb[i], geometry.xforms_composed(Rb[i], piv[i]))


def test_build_batch_layout():
    xf = [np.zeros((3, 12), np.float32), np.ones((5, 12), np.float32)]
    pts = np.zeros(10, np.int64)
    pts[7], pts[9] = 33, 64
    b = engine.build_batch([7, 9], [0, 2], [np.arange(3), np.ones(3)], [1.5, 2.5],
                           [np.zeros(3), np.ones(3)], xf, [[1, 2], [3, 4, 5, 6]], pts)
    assert b.jobs.dtype.itemsize == 72 and b.jobs.tobytes().__len__() == 144
    assert list(b.jobs["pcd_cap"]) == [64, 64] and list(b.jobs["pcd_begin"]) == [0, 64]
    assert list(b.jobs["cand_begin"]) == [0, 3] and list(b.jobs["tgt_begin"]) == [0, 2]
    assert list(b.jobs["tab_begin"]) == [0, 6] and b.units == 3 * 2 + 5 * 4
    assert b.xform.shape == (8, 12) and b.tgt_index.tolist() == [1, 2, 3, 4, 5, 6]
    raw = np.frombuffer(b.jobs.tobytes(), dtype=np.int32).reshape(2, 18)
    assert raw[1, 0] == 9 and raw[1, 1] == 2 and raw[1, 3] == 5 and raw[1, 5] == 4 and raw[1, 13] == 64
    assert np.frombuffer(b.jobs.tobytes(), dtype=np.float32).reshape(2, 18)[1, 9] == 2.5
    assert np.frombuffer(b.jobs.tobytes(), dtype=np.int64).reshape(2, 9)[1, 8] == 64


def test_removal_quirk_known_answer():
    """App. A #18: with every visited frame an inlier a T-frame track is visited
    ceil(n/2) times per round: 60 -> 30, 15, 8, 4, 2 = 59 visits."""
    cfg = OptConfig()
    T = 60
    preds, _ = synth.make_video(2, 1, 12, kinds=[0])
    preds = [preds[i % 12] for i in range(T)]
    plane = {'ids': {i: 0 for i in range(T)}, 'bbox': None, 'latest_frame': T - 1}
    stats = opt_utils.Stats()
    pool_of = {(i, 0): i for i in range(T)}
    gen = opt_utils._tracks_gen(preds, [plane], cfg, False, random.Random(0), pool_of, stats)
    spec = next(gen)
    visits = []
    def answer(sp):
        n = len(sp.targets)
        visits.append(n)
        if isinstance(sp, opt_utils.SourceReq):         # a cluster round: the driver also hands back the geometry
            n_cand, geo = len(sp.grid), geometry.source_geometry(preds[sp.frame], sp.box, cfg, False)
        else:
            n_cand, geo = len(sp.xform), None
        return opt_utils.JobResult(np.arange(n, dtype=np.int32) % n_cand,
                                   np.full(n, 0.9, np.float32), np.ones(n, np.int32), np.ones(n, np.int32),
                                   masks=torch.zeros(n, 1, 4, dtype=torch.int32), geo=geo)

    try:
        while True:
            spec = gen.send([answer(x) for x in spec] if isinstance(spec, list) else answer(spec))
    except StopIteration:
        pass
    assert visits[:5] == [60, 30, 15, 7, 3]          # id_list sizes handed to the device
    A = len(cfg.rot_cluster_grid)
    assert stats.units_visited == (30 + 15 + 8 + 4 + 2) * A + T * len(cfg.rot_final_grid)


def test_source_geometry_matches_oracle_at_other_resolutions():
    for (W, H) in ((200, 150), (1024, 768)):
        cfg = OptConfig.scaled(W, H)
        ocfg = restated.OracleConfig(height=H, width=W, focal_length=cfg.focal_length)
        preds, _ = synth.make_video(5, 2, 9, cfg, kinds=[0, 1])
        for box, trans in ((0, False), (1, True)):
            geo = geometry.source_geometry(preds[4], box, cfg, trans)
            ref = restated.source_geometry(preds[4], box, ocfg, trans)
            assert torch.equal(geo.pts[box], ref["pts"][box])
            assert np.array_equal(geo.axis3d, ref["axis3d"]) and np.array_equal(geo.dir_vec, ref["dir_vec"])


def test_linregress_inner_function_equals_scipy():
    from scipy.stats import linregress as sp
    rng = np.random.RandomState(0)
    for n in (5, 8, 30, 120):
        for y in (rng.rand(n).astype(np.float32), np.full(n, 0.1047, np.float32),
                  (np.arange(n) * 0.1047).astype(np.float32)):
            a = sp(range(n), torch.from_numpy(y))
            b = opt_utils.linregress(range(n), torch.from_numpy(y))
            assert (a.rvalue == b.rvalue) or (np.isnan(a.rvalue) and np.isnan(b.rvalue))


def test_fast_rvalue_is_bit_equal_to_scipy():
    """opt_utils._rvalue replaces scipy.stats.linregress(range(n), y).rvalue in the model selection:
    same float64 bits on random, near-linear, constant (NaN) and grid-valued angle lists."""
    from scipy.stats import linregress as sp
    rng = np.random.RandomState(1)
    grid = np.arange(-np.pi / 2, np.pi, np.pi / 30).astype(np.float32)
    for trial in range(400):
        n = int(rng.randint(1, 90))
        kind = trial % 4
        if kind == 0:
            y = rng.randn(n).astype(np.float32)
        elif kind == 1:
            y = (np.arange(n) * 0.1047 + rng.randn(n) * 0.05).astype(np.float32)
        elif kind == 2:
            y = np.full(n, rng.randn(), np.float32)
        else:
            y = rng.choice(grid, n)
        with np.errstate(all="ignore"):
            a = np.float64(sp(range(n), torch.from_numpy(y)).rvalue)
            b = np.float64(opt_utils._rvalue(y))
        assert (np.isnan(a) and np.isnan(b)) or a.tobytes() == b.tobytes(), (n, kind, a, b)


@pytest.mark.parametrize("seed", range(6))
def test_plan_tiles_covers_every_candidate_once(seed):
    """engine.plan_tiles: role-0 entries partition each job's candidates, no tile exceeds the shared-memory
    limit, one role-1 entry per job, the grid fits the wave(s) it was planned for, and larger source masks
    never get larger tiles."""
    rng = np.random.RandomState(seed)
    n = int(rng.randint(1, 9))
    jobs = np.zeros(n, dtype=_lib.JOB_DTYPE)
    jobs["n_cand"] = rng.randint(0, 120, size=n)
    jobs["pcd_cap"] = (rng.randint(0, 60000, size=n) + 31) & ~31
    jobs["cand_begin"] = np.cumsum(np.r_[0, jobs["n_cand"][:-1]])
    max_tile = int(rng.randint(1, 7))
    tile, tmap = engine.plan_tiles(jobs, max_tile, sm_count=148)
    # the library's host planner (a3d_plan_tiles) is the same statement in C
    tile_c, tmap_c = engine.plan_tiles_native(jobs, max_tile, sm_count=148)
    assert tile_c == tile and (tmap is None) == (tmap_c is None)
    if tmap is not None:
        assert np.array_equal(tmap, tmap_c)
    if tmap is None:
        assert tile == max_tile and int((-(-jobs["n_cand"].astype(np.int64) // max_tile)).sum()) + n > 148
        return
    assert tmap.dtype == np.int32 and tmap.shape[1] == 4 and len(tmap) <= 2 * 148
    assert 1 <= tile <= max_tile and tmap[:, 2].max() == tile
    sizes = {}
    for j in range(n):
        assert int(((tmap[:, 0] == j) & (tmap[:, 3] == 1)).sum()) == 1
        r = tmap[(tmap[:, 0] == j) & (tmap[:, 3] == 0)]
        r = r[np.argsort(r[:, 1])]
        assert int(r[:, 2].sum()) == int(jobs["n_cand"][j])
        assert r[:, 1].tolist() == np.cumsum(np.r_[0, r[:, 2]])[:-1].tolist()
        if len(r):
            assert r[:, 2].min() >= 1
            sizes[j] = int(r[:, 2].max())
    for a in sizes:
        for b in sizes:
            if jobs["pcd_cap"][a] > jobs["pcd_cap"][b] and jobs["n_cand"][a] >= tile and jobs["n_cand"][b] >= tile:
                assert sizes[a] <= sizes[b] + 1


def test_plan_tiles_argument_errors_and_degenerate_inputs():
    """a3d_plan_tiles (host planner in the library): bad arguments give A3D_EINVAL with a message, empty and
    many-wave inputs ask for uniform tiles (0 tiles, tile_cand = the largest tile)."""
    lib = _lib.load()
    tile = ctypes.c_int(-1)
    jobs = np.zeros(3, dtype=_lib.JOB_DTYPE)
    jobs["n_cand"] = [45, 45, 45]
    jobs["pcd_cap"] = [1024, 2048, 4096]
    out = np.zeros((2 * 148, 4), np.int32)
    assert lib.a3d_plan_tiles(jobs.ctypes.data, 3, 6, 148, out.ctypes.data, 4, ctypes.byref(tile)) == -1     # map too small
    assert b"tile_map_out" in lib.a3d_last_error_string()
    assert lib.a3d_plan_tiles(jobs.ctypes.data, 3, 0, 148, out.ctypes.data, len(out), ctypes.byref(tile)) == -1   # tile_max < 1
    assert lib.a3d_plan_tiles(None, 3, 6, 148, out.ctypes.data, len(out), ctypes.byref(tile)) == -1              # no jobs
    bad = jobs.copy()
    bad["n_cand"][1] = -5
    assert lib.a3d_plan_tiles(bad.ctypes.data, 3, 6, 148, out.ctypes.data, len(out), ctypes.byref(tile)) == -1
    assert lib.a3d_plan_tiles(jobs.ctypes.data, 0, 6, 148, out.ctypes.data, len(out), ctypes.byref(tile)) == 0 and tile.value == 6
    many = np.zeros(400, dtype=_lib.JOB_DTYPE)
    many["n_cand"] = 180
    many["pcd_cap"] = 12000
    assert lib.a3d_plan_tiles(many.ctypes.data, 400, 6, 148, out.ctypes.data, len(out), ctypes.byref(tile)) == 0 and tile.value == 6
    # a valid small plan: every candidate once, one extra entry per job, most expensive first
    n = lib.a3d_plan_tiles(jobs.ctypes.data, 3, 6, 148, out.ctypes.data, len(out), ctypes.byref(tile))
    assert 3 < n <= 148 and 1 <= tile.value <= 6
    m = out[:n]
    assert int((m[:, 3] == 1).sum()) == 3 and int(m[m[:, 3] == 0][:, 2].sum()) == 135
    assert m[0, 0] == 2                                    # the job with the largest source mask leads


def test_batched_source_geometry_is_bit_equal_to_per_source():
    """geometry.source_geometry_rows / axis.angle_offset_to_axis_rows (array arithmetic over many sources)
    against the per-source functions they vectorise, on synthetic clips and on perturbed axes that hit
    the vertical / horizontal / missed-image branches."""
    from articulation3d_b200 import axis as axis_mod
    cfg = OptConfig()
    rng = np.random.default_rng(1)
    for seed in range(4):
        preds, _ = synth.make_video(seed, 4, 12, cfg, kinds=[0, 1, 2, 0])
        for trans in (False, True):
            P, A, Cn, refs = [], [], [], []
            for p in preds:
                if seed >= 2:
                    n = len(p.pred_rot_axis)
                    p.pred_rot_axis = p.pred_rot_axis.clone()
                    p.pred_rot_axis[:, 0] *= torch.tensor(rng.choice([0.0, 1.0, 1e-9, -1.0], size=n), dtype=torch.float32)
                    p.pred_rot_axis[:, 1] *= torch.tensor(rng.choice([0.0, 1.0, 1.0], size=n), dtype=torch.float32)
                    p.pred_rot_axis[:, 2] += torch.tensor(rng.standard_normal(n) * 3, dtype=torch.float32)
                for b in range(len(p.pred_boxes)):
                    refs.append((geometry.source_geometry(p, b, cfg, trans), b))
                    P.append(p.pred_planes[b])
                    Cn.append(p.pred_boxes.get_centers()[b])
                    A.append(torch.cat((p.pred_tran_axis[b], torch.zeros(1))) if trans else p.pred_rot_axis[b])
            R = geometry.source_geometry_rows(torch.stack(P), torch.stack(A), torch.stack(Cn), cfg)
            for i, (g, b) in enumerate(refs):
                assert np.array_equal(g.normal.numpy(), R.normal[i]) and float(g.offset) == float(R.offset[i])
                assert np.array_equal(g.pts[b].numpy(), R.pts[i])
                assert np.array_equal(g.axis3d, R.axis3d[i], equal_nan=True)
                assert np.array_equal(g.dir_vec, R.dir_vec[i], equal_nan=True)
                assert np.array_equal(g.pivot, R.pivot[i], equal_nan=True)
                one = R.row(i, b, 4)
                assert torch.equal(one.pts[b], g.pts[b]) and torch.equal(one.normal, g.normal)
    # random lines, including degenerate ones, through the axis function alone
    ao = rng.standard_normal((4000, 3)).astype(np.float32)
    ao[::7, 0] = 0
    ao[::11, 1] = 0
    ao[::13, 2] = np.inf
    ao[::17, 2] = np.nan
    ce = (rng.random((4000, 2)) * [640, 480]).astype(np.float32)
    want = axis_mod.angle_offset_to_axis(torch.from_numpy(ao), torch.from_numpy(ce)).numpy()
    assert np.array_equal(axis_mod.angle_offset_to_axis_rows(ao, ce), want)


def test_batched_candidate_transforms_equal_per_source():
    cfg = OptConfig()
    rng = np.random.default_rng(3)
    d = rng.standard_normal((9, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    piv = rng.standard_normal((9, 3)).astype(np.float32)
    R = geometry.rotation_matrices(cfg.rot_cluster_grid, d)
    C = geometry.xforms_composed(geometry.rotation_matrices(cfg.rot_final_grid, d), piv)
    Tr = geometry.xforms_translate(cfg.trans_grid, d)
    for i in range(9):
        assert np.array_equal(R[i], geometry.rotation_matrices(cfg.rot_cluster_grid, d[i]))
        assert np.array_equal(geometry.xforms_seq_from_dirs(cfg.rot_cluster_grid, d)[i], geometry.xforms_seq(R[i]))
        assert np.array_equal(C[i], geometry.xforms_composed(geometry.rotation_matrices(cfg.rot_final_grid, d[i]), piv[i]))
        assert np.array_equal(Tr[i], geometry.xforms_translate(cfg.trans_grid, d[i]))


def test_constant_track_r_switch():
    """a constant angle list (static plane) is the degenerate case of linregress: r = 0.0 with
    the scipy of the reference's era, NaN with scipy >= 1.9.  The default follows the installed scipy
    (what the unmodified reference would do here); OptConfig.constant_track_r forces either."""
    from scipy.stats import linregress as sp
    y = np.full(7, 0.1047, np.float32)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        installed = float(sp(range(7), y).rvalue)
    got = float(opt_utils._rvalue(y))
    assert (np.isnan(installed) and np.isnan(got)) or installed == got
    assert float(opt_utils._rvalue(y, 0.0)) == 0.0 and np.isnan(float(opt_utils._rvalue(y, float("nan"))))
    # what it decides: with r = 0 a static track is filtered (has_rot False), with NaN it is kept
    for r, keep in ((0.0, False), (float("nan"), True)):
        rsqs = np.array([float(opt_utils._rvalue(y, r)) ** 2])
        assert (not (rsqs.max() < OptConfig().rsq_thresh)) == keep


def test_build_batch_rows_equals_build_batch():
    rng = np.random.default_rng(5)
    S = 5
    n_cand, n_tgt = np.array([3, 3, 3, 3, 3]), np.array([2, 4, 1, 3, 2])
    pts = rng.integers(1, 500, size=20)
    src = rng.integers(0, 20, size=S)
    normals, pivots = rng.standard_normal((S, 3)).astype(np.float32), rng.standard_normal((S, 3)).astype(np.float32)
    offs = rng.random(S).astype(np.float32)
    xf = rng.standard_normal((S, 3, 12)).astype(np.float32)
    tg = [rng.integers(0, 20, size=k).astype(np.int32) for k in n_tgt]
    a = engine.build_batch(list(src), [1] * S, list(normals), list(offs), list(pivots), list(xf), tg, pts)
    b = engine.build_batch_rows(src, np.ones(S, np.int32), normals, offs, pivots, xf.reshape(-1, 12), n_cand,
                                np.concatenate(tg), n_tgt, pts)
    assert a.jobs.tobytes() == b.jobs.tobytes()
    assert np.array_equal(a.xform, b.xform) and np.array_equal(a.tgt_index, b.tgt_index)


def test_track_planes_stress_matches_oracle():
    """Frames with duplicated / overlapping boxes (a second box of a frame matching a track the first one
    just joined), gaps up to and beyond the limit, empty frames, both classes."""
    rng = np.random.default_rng(7)
    for trial in range(6):
        T = 40
        preds = []
        anchors = rng.random((4, 2)) * [400, 300] + 40
        for t in range(T):
            rows, cls = [], []
            for k in range(4):
                if rng.random() < 0.25:
                    continue
                for _ in range(1 + int(rng.random() < 0.3)):        # sometimes the same box twice (jittered)
                    c = anchors[k] + rng.normal(0, 4 + 6 * trial / 5, 2)
                    w, h = 80 + rng.normal(0, 6), 60 + rng.normal(0, 6)
                    rows.append([c[0], c[1], c[0] + w, c[1] + h])
                    cls.append(k % 2)
            if t % 13 == 12:
                rows, cls = [], []
            inst = Instances((480, 640))
            inst.pred_boxes = Boxes(torch.tensor(rows, dtype=torch.float32).reshape(-1, 4))
            inst.pred_classes = np.array(cls, dtype=np.int64)
            preds.append(inst)
        a = opt_utils.track_planes(preds)
        b = restated.track_planes(preds)
        for cat in ("rot", "trans"):
            assert [p['ids'] for p in a[cat]] == [p['ids'] for p in b[cat]]
            assert [p['latest_frame'] for p in a[cat]] == [p['latest_frame'] for p in b[cat]]
            for p, q in zip(a[cat], b[cat]):
                assert torch.equal(p['bbox'].tensor, q['bbox'].tensor)
        assert sum(len(a[c]) for c in a) > 0


def test_write_back_matches_oracle(monkeypatch):
    """opt_utils._write_back (flat-array form) against the oracle's frame x track loops, with the same
    decided tracks: scores, rewritten rotation axes, in-place translation axes, untouched fields."""
    cfg = OptConfig()
    preds, _ = synth.make_video(31, 5, 16, cfg, kinds=[0, 1, 0, 1, 2], drop_prob=0.15)
    rng = np.random.default_rng(2)

    def decide(planes, translation, src_preds):
        for plane in planes:
            plane['has_rot'] = bool(rng.random() < 0.6)
            if plane['has_rot']:
                f0 = next(iter(plane['ids']))
                plane['std_axis'] = (src_preds[f0].pred_tran_axis[plane['ids'][f0]] if translation
                                     else torch.tensor(rng.integers(0, 480, 4), dtype=torch.int64))

    for kind, translation in (("trans", True), ("rot", False)):
        a, b = synth.clone_preds(preds), synth.clone_preds(preds)
        pa, pb = opt_utils.track_planes(a)[kind], restated.track_planes(b)[kind]
        state = rng.bit_generator.state
        decide(pa, translation, a)
        rng.bit_generator.state = state
        decide(pb, translation, b)
        monkeypatch.setattr(restated, "_optimize_tracks", lambda *args, **kw: None)
        want = (restated.optimize_planes_3d_trans if translation else restated.optimize_planes_3dc)(b, pb)
        got = opt_utils._write_back(a, pa, cfg, kind)
        assert len(got) == len(want)
        for x, y, xin, yin in zip(got, want, a, b):
            assert np.array_equal(x.scores, y.scores) and x.scores.dtype == y.scores.dtype
            assert torch.equal(x.pred_rot_axis, y.pred_rot_axis) and torch.equal(x.pred_tran_axis, y.pred_tran_axis)
            assert torch.equal(x.pred_planes, y.pred_planes)
            assert x.pred_boxes is xin.pred_boxes and x.pred_masks is xin.pred_masks
            assert torch.equal(xin.pred_tran_axis, yin.pred_tran_axis) and torch.equal(xin.pred_rot_axis, yin.pred_rot_axis)


def test_axis_angle_matrices_equal_oracle_form():
    """geometry._axis_angle_to_matrix (component-plane form) against the oracle's statement of
    pytorch3d's construction, incl. the small-angle branch and a zero vector: same float64 bits."""
    rng = np.random.default_rng(0)
    aa = torch.from_numpy(rng.standard_normal((300, 45, 3)) * np.exp(rng.standard_normal((300, 45, 1)) * 3))
    aa[::9] *= 1e-8
    aa[5, 3] = 0
    a, b = geometry._axis_angle_to_matrix(aa).reshape(300, 45, 3, 3), restated.axis_angle_to_matrix64(aa)
    assert torch.equal(torch.nan_to_num(a, nan=7.0), torch.nan_to_num(b, nan=7.0))


def test_fused_candidate_rows_equal_the_torch_form():
    """``geometry._rotation_xforms`` (quaternion entries in one fused pass of the library's host helper) must be
    bit-equal to the all-torch form the oracle comparison above pins, including tiny angles (Taylor branch),
    the zero angle and angles past pi."""
    from articulation3d_b200 import OptConfig
    rng = np.random.RandomState(3)
    d = rng.randn(300, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[0] = [0.0, 0.0, 1.0]
    d[1] = [1e-9, 0.0, 0.0]                                         # |axis_angle| below the 1e-6 switch
    cfg = OptConfig()
    for grid in (cfg.rot_cluster_grid, cfg.rot_final_grid, cfg.legacy_grid, np.array([0.0, 1e-8, -1e-7, 3.5, -4.0])):
        want = geometry._rotation_entries(grid, d).to(torch.float32).numpy()
        got = geometry._rotation_xforms(grid, d)
        assert got.shape == want.shape[:-1] + (12,) and got.dtype == np.float32
        assert np.array_equal(got[..., :9].view(np.uint32), want.view(np.uint32))
        assert not got[..., 9:].any()
        R = geometry.rotation_matrices(grid, d)
        assert np.array_equal(R.reshape(want.shape).view(np.uint32), want.view(np.uint32))
    one = geometry._rotation_xforms(cfg.rot_final_grid, d[7])       # a single axis (final phase of one track)
    assert np.array_equal(one, geometry._rotation_xforms(cfg.rot_final_grid, d[7:8])[0])
