"""Prototype (numpy, CPU) of the filtered projection, written before the kernel: per candidate a
plane-induced homography gives an approximate pixel coordinate; a running error bound decides whether
truncating it is guaranteed to equal the reference's fp32 chain (oracle/restated.py).  Prints the largest
|q_fast - q_exact| / bound, the share of coordinates the bound cannot decide, and how many decided ones
differ from the oracle (must be 0).  The kernel's bound has the same structure with slightly larger constants
(saturating-FMA clamp, per-item coefficient): k_project<filter> in articulation3d_b200/csrc/a3d.cu.
    python tools/filter_proto.py          # ~2 min"""
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
    """H (3 rows u,v,w x 3 cols x,y,1) in fp64 and its abs-magnitude version."""
    n = normal.astype(np.float64)
    G = off * A.astype(np.float64) + np.outer(n, b)              # G_kj
    L = kinv.T @ G                                               # L_mj
    Hu = f * L[:, 0] + cx * L[:, 2]
    Hv = f * L[:, 1] + cy * L[:, 2]
    Hw = L[:, 2]
    return np.stack([Hu, Hv, Hw])


def run(seed, mode, n_frames=20, frame_step=5, quiet=False):
    cfg = R_.OracleConfig()
    preds, _ = synth.make_video(seed, 3, n_frames, kinds=[0, 1, 0])
    kinv = cfg.K_inv()
    f, cx, cy = f32(cfg.focal_length), f32(cfg.width / 2), f32(cfg.height / 2)
    Wd, Hd = cfg.width, cfg.height
    Dmax, cmax = max(Wd, Hd), max(cx, cy)
    stats = dict(n=0, unc=0, maxratio=0.0, wrong=0, eps=[])
    for t in range(0, n_frames, frame_step):
        p = preds[t]
        for b_id in range(len(p.pred_boxes)):
            translation = mode == "translate"
            g = R_.source_geometry(p, b_id, cfg, translation)
            a = g["axis3d"][0].astype(f32)
            pcd = g["pcd"]
            verts = g["verts"].numpy().astype(np.float64)
            if len(pcd) == 0:
                continue
            if mode == "translate":
                vecs = R_.translation_vectors(cfg.trans_grid, g["dir_vec"])
                pts = R_.transform_translate(pcd, vecs)
                As = [np.eye(3)] * len(vecs)
                bs = [v.astype(np.float64) for v in vecs]
            else:
                grid = cfg.rot_cluster_grid if mode == "seq" else cfg.rot_final_grid
                Rm = R_.rotation_matrices(grid, g["dir_vec"])
                if mode == "seq":
                    pts = R_.transform_seq(pcd, a, Rm)
                    As = list(Rm)
                    bs = [a.astype(np.float64) - a.astype(np.float64) @ r.astype(np.float64) for r in Rm]
                else:
                    m3 = R_.composed_last_row(a, Rm)
                    pts = R_.transform_composed(pcd, Rm, m3)
                    As = list(Rm)
                    bs = [m.astype(np.float64) for m in m3]
            X, Y, Z = pts[..., 0], pts[..., 1], pts[..., 2]
            with np.errstate(all="ignore"):
                u = (f * X) + cx * Z
                v = (f * Y) + cy * Z
                qx_e, qy_e = u / Z, v / Z
            row_e, col_e = R_.project_pixels(pts, cfg, Hd, Wd)
            # per-point constants
            n = g["normal"].numpy().astype(np.float64)
            off = float(g["offset"])
            ray = (kinv @ np.concatenate([verts, np.ones((len(verts), 1))], 1).T).T
            dot = ray @ n
            p1 = np.abs(pcd.astype(np.float64)).sum(1)
            if mode == "seq":
                pp1 = np.abs(pcd - a[None]).astype(np.float64).sum(1)
                Sig = 1.001 * (1.01 * p1 + 5 * pp1) + np.abs(a).max()
                M = 1.001 * pp1 + np.abs(a).max()
            else:
                Sig = 1.001 * 5.01 * p1
                M = 1.001 * p1
            x0, y0 = np.floor(verts[:, 0].mean()), np.floor(verts[:, 1].mean())
            xs, ys = (verts[:, 0] - x0).astype(f32), (verts[:, 1] - y0).astype(f32)
            xm, ym = np.abs(xs).max(), np.abs(ys).max()
            for c in range(len(As)):
                tinf = 0.0 if mode == "seq" else np.abs(bs[c]).max()
                Cpt = U * np.abs(dot) * ((f + cmax) * (Sig + tinf + 2 * (M + tinf)) + Dmax * (Sig + tinf))
                H = homography(kinv, n, off, As[c], bs[c], float(f), float(cx), float(cy))
                Hc = H.copy()
                Hc[:, 2] = H[:, 0] * x0 + H[:, 1] * y0 + H[:, 2]         # recentred
                Hf = Hc.astype(f32)
                mag = np.abs(Hf[:, 0]) * xm + np.abs(Hf[:, 1]) * ym + np.abs(Hf[:, 2])
                Ec = 3 * U * (max(mag[0], mag[1]) + (Dmax + 1) * mag[2])
                UF = fma32(np.full_like(xs, Hf[0, 0]), xs, fma32(np.full_like(xs, Hf[0, 1]), ys, np.full_like(xs, Hf[0, 2])))
                VF = fma32(np.full_like(xs, Hf[1, 0]), xs, fma32(np.full_like(xs, Hf[1, 1]), ys, np.full_like(xs, Hf[1, 2])))
                WF = fma32(np.full_like(xs, Hf[2, 0]), xs, fma32(np.full_like(xs, Hf[2, 1]), ys, np.full_like(xs, Hf[2, 2])))
                with np.errstate(all="ignore"):
                    r = (f32(1) / WF).astype(f32)
                    qx_f = (UF * r).astype(f32)
                    qy_f = (VF * r).astype(f32)
                    eps = (Cpt + Ec) * np.abs(r) * 1.25 + (Dmax + 1) * (2.0 ** -22 + 3 * U) + U * Dmax
                for qf, qe, idx_e, hi in ((qx_f, qx_e[c], col_e[c], Wd - 1), (qy_f, qy_e[c], row_e[c], Hd - 1)):
                    zc = np.clip(qf.astype(np.float64) - 0.5, 0, hi)
                    nfast = np.rint(zc)
                    d = np.abs(zc - nfast)
                    unc = ~(d <= 0.5 - eps)
                    ok = ~unc
                    stats["wrong"] += int((nfast[ok] != idx_e[ok]).sum())
                    stats["unc"] += int(unc.sum())
                    stats["n"] += len(qf)
                    inr = (qe > 0.5) & (qe < hi + 0.5) & np.isfinite(qe)
                    if inr.any():
                        ratio = np.abs(qf[inr].astype(np.float64) - qe[inr].astype(np.float64)) / eps[inr]
                        stats["maxratio"] = max(stats["maxratio"], float(ratio.max()))
                stats["eps"].append(float(np.median(eps)))
    if not quiet:
        print(f"seed {seed} mode {mode}: coords {stats['n']}, uncertain {stats['unc'] / stats['n']:.5f}, "
              f"wrong-certified {stats['wrong']}, max |dq|/eps {stats['maxratio']:.4f}, median eps {np.median(stats['eps']):.2e}")
    return stats


if __name__ == "__main__":
    for mode in ("seq", "composed", "translate"):
        for seed in (1, 2):
            run(seed, mode)
