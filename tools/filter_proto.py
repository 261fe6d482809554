"""Prototype (numpy, CPU) of the filtered projection, written before the kernel: per candidate a
plane-induced homography gives an approximate pixel coordinate; a running error bound decides whether
truncating it is guaranteed to equal the reference's fp32 chain (oracle/restated.py).  Prints the largest
|q_fast - q_exact| / bound, the share of coordinates the bound cannot decide, and how many decided ones
differ from the oracle (must be 0).  The kernel's bound has the same structure with slightly larger constants
(saturating-FMA clamp, per-item coefficient): k_project<filter> in articulation3d_b200/csrc/a3d.cu.
    python tools/filter_proto.py             # synthetic clips, ~2 min
    python tools/filter_proto.py --random 40 # adversarial random geometry (grazing planes, w through 0, ...)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import restated as R_
from articulation3d_b200 import synth

U = 2.0 ** -24
f32 = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def homography(kinv, normal, off, A, b, f, cx, cy):
    """H (3 rows u,v,w x 3 cols x,y,1) in fp64."""
    n = normal.astype(np.float64)
    G = off * A.astype(np.float64) + np.outer(n, b)              # G_kj
    L = kinv.T @ G                                               # L_mj
    Hu = f * L[:, 0] + cx * L[:, 2]
    Hv = f * L[:, 1] + cy * L[:, 2]
    Hw = L[:, 2]
    return np.stack([Hu, Hv, Hw])


def new_stats():
    return dict(n=0, unc=0, maxratio=0.0, wrong=0, eps=[])


def evaluate(cfg, verts, pcd, normal, off, a, mode, pts, As, bs, stats):
    """One source mask under a list of candidates: pts = the oracle's transformed points (A, P, 3);
    (As[c], bs[c]) the real-valued map s = p A + b of candidate c."""
    kinv = cfg.K_inv()
    f, cx, cy = f32(cfg.focal_length), f32(cfg.width / 2), f32(cfg.height / 2)
    Wd, Hd = cfg.width, cfg.height
    Dmax, cmax = max(Wd, Hd), max(cx, cy)
    X, Y, Z = pts[..., 0], pts[..., 1], pts[..., 2]
    with np.errstate(all="ignore"):
        u = (f * X) + cx * Z
        v = (f * Y) + cy * Z
        qx_e, qy_e = u / Z, v / Z
    row_e, col_e = R_.project_pixels(pts, cfg, Hd, Wd)
    n = np.asarray(normal, dtype=np.float32).astype(np.float64)
    ray = (kinv @ np.concatenate([verts, np.ones((len(verts), 1))], 1).T).T
    dot = ray @ n
    dmag = np.abs(ray * n[None]).sum(1)
    finite = np.isfinite(pcd).all(1)
    p1 = np.abs(np.nan_to_num(pcd.astype(np.float64))).sum(1)
    if mode == "seq":
        pp1 = np.abs(np.nan_to_num(pcd.astype(np.float64)) - a[None].astype(np.float64)).sum(1)
        Sig = 1.001 * (1.01 * p1 + 5 * pp1) + np.abs(a).max()
        M = 1.001 * pp1 + np.abs(a).max()
        tmax = 0.0
    else:
        Sig = 1.001 * 5.01 * p1
        M = 1.001 * p1
        tmax = max(float(np.abs(b).max()) for b in bs)           # the kernel uses the job's largest |t|
    x0, y0 = np.floor((verts[:, 0].min() + verts[:, 0].max()) / 2), np.floor((verts[:, 1].min() + verts[:, 1].max()) / 2)
    xs, ys = (verts[:, 0] - x0).astype(f32), (verts[:, 1] - y0).astype(f32)
    xm, ym = np.abs(xs).max(), np.abs(ys).max()
    with np.errstate(all="ignore"):
        Cpt = U * 1.25 * np.abs(dot) * ((f + cmax) * (Sig + tmax + 2 * (M + tmax)) + Dmax * (Sig + tmax))
    Cpt = np.where(finite & (np.abs(dot) >= 1e-6 * dmag) & np.isfinite(Cpt), Cpt, np.inf)
    for c in range(len(As)):
        A = np.asarray(As[c], dtype=np.float64)
        if mode != "translate" and not (np.abs(A) <= 1.001).all():
            continue                                             # exact-only candidate in the kernel
        H = homography(kinv, n, off, A, np.asarray(bs[c], dtype=np.float64), float(f), float(cx), float(cy))
        Hc = H.copy()
        Hc[:, 2] = H[:, 0] * x0 + H[:, 1] * y0 + H[:, 2]         # recentred
        Hf = Hc.astype(f32)
        if not np.isfinite(Hf).all():
            continue
        mag = np.abs(Hf[:, 0]).astype(np.float64) * xm + np.abs(Hf[:, 1]).astype(np.float64) * ym + np.abs(Hf[:, 2])
        Ec = 1.25 * 3 * U * (max(mag[0], mag[1]) + (Dmax + 1) * mag[2])
        full = lambda r, k: np.full_like(xs, Hf[r, k])           # noqa: E731
        UF = fma32(full(0, 0), xs, fma32(full(0, 1), ys, full(0, 2)))
        VF = fma32(full(1, 0), xs, fma32(full(1, 1), ys, full(1, 2)))
        WF = fma32(full(2, 0), xs, fma32(full(2, 1), ys, full(2, 2)))
        with np.errstate(all="ignore"):
            r = (f32(1) / WF).astype(f32)
            qx_f = (UF * r).astype(f32)
            qy_f = (VF * r).astype(f32)
            eps = (Cpt + Ec) * np.abs(r).astype(np.float64) + (Dmax + 1) * (2.0 ** -22 + 3 * U) + U * Dmax
        for qf, qe, idx_e, hi in ((qx_f, qx_e[c], col_e[c], Wd - 1), (qy_f, qy_e[c], row_e[c], Hd - 1)):
            with np.errstate(all="ignore"):
                zc = np.clip(np.nan_to_num(qf.astype(np.float64) - 0.5, nan=0.0), 0, hi)
                nfast = np.rint(zc)
                d = np.abs(zc - nfast)
                unc = ~(d <= 0.5 - eps)
            ok = ~unc
            stats["wrong"] += int((nfast[ok] != idx_e[ok]).sum())
            stats["unc"] += int(unc.sum())
            stats["n"] += len(qf)
            inr = ok & (qe > 0.5) & (qe < hi + 0.5) & np.isfinite(qe)
            if inr.any():
                ratio = np.abs(qf[inr].astype(np.float64) - qe[inr].astype(np.float64)) / eps[inr]
                stats["maxratio"] = max(stats["maxratio"], float(ratio.max()))
        stats["eps"].append(float(np.median(eps[np.isfinite(eps)])) if np.isfinite(eps).any() else np.inf)


def candidates(mode, pcd, a, dir_vec, rot_grid, trans_grid):
    """Oracle-transformed points and the real-valued (A, b) of every candidate."""
    if mode == "translate":
        vecs = R_.translation_vectors(trans_grid, dir_vec)
        return R_.transform_translate(pcd, vecs), [np.eye(3)] * len(vecs), [v.astype(np.float64) for v in vecs]
    Rm = R_.rotation_matrices(rot_grid, dir_vec)
    if mode == "seq":
        a64 = a.astype(np.float64)
        return R_.transform_seq(pcd, a, Rm), list(Rm), [a64 - a64 @ r.astype(np.float64) for r in Rm]
    m3 = R_.composed_last_row(a, Rm)
    return R_.transform_composed(pcd, Rm, m3), list(Rm), [m.astype(np.float64) for m in m3]


def run(seed, mode, n_frames=20, frame_step=5, quiet=False):
    cfg = R_.OracleConfig()
    preds, _ = synth.make_video(seed, 3, n_frames, kinds=[0, 1, 0])
    stats = new_stats()
    for t in range(0, n_frames, frame_step):
        p = preds[t]
        for b_id in range(len(p.pred_boxes)):
            g = R_.source_geometry(p, b_id, cfg, mode == "translate")
            if len(g["pcd"]) == 0:
                continue
            a = g["axis3d"][0].astype(f32)
            grid = cfg.rot_cluster_grid if mode == "seq" else cfg.rot_final_grid
            pts, As, bs = candidates(mode, g["pcd"], a, g["dir_vec"], grid, cfg.trans_grid)
            evaluate(cfg, g["verts"].numpy().astype(np.float64), g["pcd"], g["normal"].numpy(), float(g["offset"]), a,
                     mode, pts, As, bs, stats)
    if not quiet:
        report(f"seed {seed} mode {mode}", stats)
    return stats


def run_random(seed, quiet=False):
    """Geometry that stresses the bound, as tests/test_gpu_parity.py::test_filter_kernel_adversarial_geometry:
    grazing and fronto-parallel planes, planes centimetres from the camera, pivots far off the surface, full-circle
    rotations, translations of metres, tiny and huge focal lengths."""
    rng = np.random.RandomState(9000 + seed)
    H, W = [(480, 640), (120, 200), (96, 128)][seed % 3]
    cfg = R_.OracleConfig(height=H, width=W, focal_length=float(rng.choice([40.0, 517.97 * W / 640, 6000.0])))
    stats = new_stats()
    yy, xx = np.mgrid[0:H, 0:W]
    for j in range(6):
        cy, cx = rng.uniform(0.2, 0.8) * H, rng.uniform(0.2, 0.8) * W
        ry, rx = rng.uniform(0.05, 0.4) * H, rng.uniform(0.05, 0.4) * W
        m = (np.abs(yy - cy) < ry) & (np.abs(xx - cx) < rx)
        if j % 2:
            m &= rng.rand(H, W) < 0.5
        verts = np.stack(np.nonzero(m)[::-1], axis=1).astype(np.float64)          # (x, y), row-major like nonzero()
        if len(verts) == 0:
            continue
        mode = ("seq", "composed", "translate")[j % 3]
        normal = rng.randn(3)
        if j % 5 == 0:
            normal[2] = 1e-3 * rng.randn()
        elif j % 5 == 1:
            normal = np.array([0.0, 0.0, 1.0]) + 1e-4 * rng.randn(3)
        normal = (normal / np.linalg.norm(normal)).astype(f32)
        off = float(f32(rng.choice([0.02, 0.5, 2.0, 50.0])))
        a = (rng.randn(3) * rng.choice([0.1, 3.0, 100.0])).astype(f32)
        pcd = R_.pcd_to_f32(R_.get_pcd(verts, normal, off, cfg))
        A = 12
        if mode == "translate":
            vecs = (rng.randn(A, 3) * rng.choice([1e-3, 0.3, 5.0])).astype(f32)
            pts, As, bs = R_.transform_translate(pcd, vecs), [np.eye(3)] * A, [v.astype(np.float64) for v in vecs]
        else:
            ax = rng.randn(A, 3)
            ax /= np.linalg.norm(ax, axis=1, keepdims=True)
            ang = rng.uniform(-np.pi, np.pi, A) * (1e-3 if j == 3 else 1.0)
            Rm = R_.axis_angle_to_matrix64(torch.from_numpy(ax * ang[:, None])).to(torch.float32).numpy()
            if mode == "seq":
                a64 = a.astype(np.float64)
                pts, As, bs = R_.transform_seq(pcd, a, Rm), list(Rm), [a64 - a64 @ r.astype(np.float64) for r in Rm]
            else:
                m3 = (rng.randn(A, 3) * rng.choice([1e-3, 0.3, 5.0])).astype(f32)
                pts, As, bs = R_.transform_composed(pcd, Rm, m3), list(Rm), [t.astype(np.float64) for t in m3]
        evaluate(cfg, verts, pcd, normal, off, a, mode, pts, As, bs, stats)
    if not quiet:
        report(f"random {seed} ({W}x{H}, f {cfg.focal_length:.0f})", stats)
    return stats


def report(tag, stats):
    n = max(stats["n"], 1)
    print(f"{tag}: coords {stats['n']}, uncertain {stats['unc'] / n:.5f}, wrong-certified {stats['wrong']}, "
          f"max |dq|/eps {stats['maxratio']:.4f}, median eps {np.median(stats['eps']) if stats['eps'] else float('nan'):.2e}",
          flush=True)


if __name__ == "__main__":
    if "--random" in sys.argv:
        k = int(sys.argv[sys.argv.index("--random") + 1])
        tot = new_stats()
        for s in range(k):
            st = run_random(s)
            tot["wrong"] += st["wrong"]
            tot["n"] += st["n"]
            tot["maxratio"] = max(tot["maxratio"], st["maxratio"])
        print(f"total: coords {tot['n']}, wrong-certified {tot['wrong']}, max |dq|/eps {tot['maxratio']:.4f}")
    else:
        for mode in ("seq", "composed", "translate"):
            for seed in (1, 2):
                run(seed, mode)
