#!/usr/bin/env python
"""Benchmark of the temporal articulation optimizer hot path (contract: task spec §④).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3_shard|c3|c4_shard]
    python bench.py --impl reference ...     # the CPU oracle port on the host cores
    torchrun ... bench.py --gpus N ...       # one rank per GPU, weak scaling (one workload per rank)

A *step* is one scoring pass over the workload: per track one source frame is
unprojected, moved by every candidate transform, re-projected and scored against
every frame of the track (project + score + arg-max kernels).  ``value`` counts
track-frame x candidate IoU evaluations per second with all inputs resident in
HBM; ``e2e`` is the same metric through the public ``optimize_planes`` API with
HOST fp32 masks (H2D, packing, every round's D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "track_frames_x_angles_per_sec"
UNIT = "track-frame*angle IoU evaluations/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# CPU oracle port timing (cpu_baseline / --impl reference)
# --------------------------------------------------------------------------
def _oracle_cfg(cfg):
    from oracle import restated
    return restated.OracleConfig(height=cfg.height, width=cfg.width, focal_length=cfg.focal_length,
                                 rot_cluster_grid=cfg.rot_cluster_grid, rot_final_grid=cfg.rot_final_grid,
                                 trans_grid=cfg.trans_grid)


def _count_units(trace, cfg):
    n = 0
    for step in trace:
        if step["kind"] == "trans":
            a = len(cfg.trans_grid)
        else:
            a = len(cfg.rot_cluster_grid) if step["phase"] == "cluster" else len(cfg.rot_final_grid)
        n += a * len(step["visits"])
    return n


def run_oracle_sample(wl, seed, tracks, frames):
    """One run of the CPU port on a (tracks x frames) sample of the workload -> (seconds, units)."""
    from articulation3d_b200 import workloads
    from oracle import restated
    preds, cfg = workloads.make_clip(wl, seed, tracks=tracks, frames=frames)
    random.seed(seed)
    t0 = time.perf_counter()
    planes = restated.track_planes(preds)
    trace = []
    restated.optimize_planes(preds, planes, "3dc", cfg=_oracle_cfg(cfg), trace=trace)
    dt = time.perf_counter() - t0
    return dt, _count_units(trace, cfg)


def pick_sample(wl, budget_s: float):
    """Largest (tracks, frames) sample of the workload whose CPU run should fit ``budget_s``."""
    dt, units = run_oracle_sample(wl, 2020, 1, 24)            # calibration (also warms torch)
    rate = max(units, 1) / dt
    ladder = [(wl.tracks, wl.frames), (2, wl.frames), (1, wl.frames), (1, wl.frames * 2 // 3),
              (1, wl.frames // 2), (1, wl.frames // 3), (1, 24)]
    cand = len(wl.cfg().rot_cluster_grid)
    final = len(wl.cfg().rot_final_grid)
    for tr, fr in ladder:
        if fr < 24:
            continue
        est_units = tr * (fr * cand + fr * final)             # ~T visits in the cluster rounds + T final
        if est_units / rate <= budget_s:
            return tr, fr, rate
    return 1, 24, rate


def reference_arm(args, wl):
    """--impl reference: the reference's CPU implementation (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    total = args.steps + args.warmup
    tr, fr, _ = pick_sample(wl, budget_s=max(2.0, 150.0 / max(total, 1)))
    for _ in range(args.warmup):
        run_oracle_sample(wl, 2020, tr, fr)
    secs, units = 0.0, 0
    for _ in range(args.steps):
        dt, u = run_oracle_sample(wl, 2020, tr, fr)
        secs += dt
        units += u
    value = units / secs
    sample = f"{tr} track(s) x {fr} frames of {wl.name}, full optimize_planes('3dc') incl. cluster rounds"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 popcount / fp32+fp64 geometry",
        "data": "synthetic", "config": {"workload": wl.description, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def measure_pass(wl, steps, warmup, dev, rank, world, dist, sample_clocks=True):
    """Device-resident scoring passes of one workload: K timed steps with CUDA events on the
    launching stream -> dict(value, ms, roofline, ...).  Every rank runs its own workload."""
    from articulation3d_b200 import engine, workloads

    inp = workloads.build_pass(wl, seed0=2020 + 1000 * rank, device=dev)
    # two sets of pass buffers when ranks exchange results: the gather of step k reads the result block of
    # step k in place while step k+1 writes the other set (no staging copy in the step)
    wss = [engine.Workspace(dev) for _ in range(2 if world > 1 else 1)]
    ws = wss[0]
    packed_bytes = inp.pool.bits.numel() * 4
    l2_bytes = 126 * 2 ** 20
    flush = None if packed_bytes > 2 * l2_bytes else torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
    # multi-GPU: the only exchange of the path is the gather of fixed-size per-track-frame
    # records {angle_id, inter, union}.  It runs on a side stream, double-buffered, so the
    # collective of step k overlaps the kernels of step k+1.
    recs = [torch.zeros(3, inp.dbatch.n_tgt_total, dtype=torch.int32, device=dev) for _ in range(2)]
    gathered = [[torch.zeros_like(recs[0]) for _ in range(world)] for _ in range(2)] if world > 1 else None
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    comm_done = [None, None]
    step_no = [0]

    def one_step(evs=None, split=False):
        """One pass through the C ABI.  split=False: a3d_pass, the call the package makes (keys cleared
        first, dependent launches on small passes).  split=True: a3d_project | a3d_score with an event
        between the two launch groups, for the per-kernel durations of the roofline."""
        if flush is not None:
            flush.zero_()
        if evs:
            evs[0].record()
        b = step_no[0] & 1 if (world > 1 and not split) else 0
        if world > 1 and not split and comm_done[b] is not None:
            torch.cuda.current_stream().wait_event(comm_done[b])         # buffer set b free again
        res = _run_split(engine, inp, ws, evs) if split else engine.run_pass(inp.cfg, inp.pool, inp.dbatch, wss[b])
        if world > 1 and not split:
            if res.block is not None:
                rec = res.block[:3]                       # the rows are contiguous in the result block
            else:
                rec = recs[b]
                rec[0], rec[1], rec[2] = res.best_cand, res.best_inter, res.best_union
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ready)
                dist.all_gather(gathered[b], rec)
                comm_done[b] = torch.cuda.Event(enable_timing=True)
                comm_done[b].record()
            step_no[0] += 1
        if evs:
            evs[2].record()
        return res

    for _ in range(max(warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index).start() if (rank == 0 and sample_clocks) else None
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    wall0 = time.perf_counter()
    # Let the host run ahead of the device: the GPU spins ~40 ms while all K steps are
    # enqueued, so the event intervals below contain kernel time only, never launch gaps.
    torch.cuda._sleep(int(0.04 * 1.9e9))
    for k in range(steps):
        one_step(events[k])
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    t_step = sum(e[0].elapsed_time(e[2]) for e in events) / steps          # ms, the timed K steps
    if world > 1:
        # exposed tail of the last (un-overlapped) gather, amortised over the K steps
        last = comm_done[(step_no[0] - 1) & 1]
        t_step += max(0.0, events[-1][2].elapsed_time(last)) / steps
    # the same K steps again as a3d_project | a3d_score, for the kernel durations of the roofline
    split_events = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    torch.cuda._sleep(int(0.04 * 1.9e9))
    for k in range(steps):
        one_step(split_events[k], split=True)
    torch.cuda.synchronize()
    t_proj = sum(e[0].elapsed_time(e[1]) for e in split_events) / steps
    t_score = sum(e[1].elapsed_time(e[2]) for e in split_events) / steps
    t_split = sum(e[0].elapsed_time(e[2]) for e in split_events) / steps
    if world > 1:
        t = torch.tensor([t_step, t_proj, t_score], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step, t_proj, t_score = t.tolist()
    units_all = inp.units * world
    value = units_all / (t_step * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------
    peak, peak_src = _peaks()
    alg = wl.alg_bytes_per_pass()
    proj_dom = t_proj >= t_score
    dom, t_dom = ("a3d_project (k_unproject + k_project)", t_proj) if proj_dom else \
                 ("a3d_score (k_score + k_finalize)", t_score)
    achieved = alg / (t_dom * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(wl.name, {}).get("k_project" if proj_dom else "k_score")
    # second bound of SURVEY 8d: the instruction pipes, as ncu saw them on the committed captures
    pipes = None
    ppath = os.path.join(ROOT, "profiles", "pipes.json")
    if os.path.exists(ppath):
        with open(ppath) as f:
            pipes = json.load(f).get(wl.name)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg, "kernel_ms": t_dom,
                "kernels_ms": {"project": t_proj, "score": t_score, "step_two_calls": t_split, "step": t_step},
                "pipes_pct_of_peak": pipes,
                "note": "bit-packed masks make the pass ALU-bound (fp32 splat / AND+POPC), not HBM-bound; "
                        "achieved = SURVEY 8d algorithmic bytes / dominant-kernel time"}
    del inp, ws
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": t_step, "roofline": roofline, "units_per_step_per_gpu": units_all // world,
            "packed_mask_bytes_per_gpu": packed_bytes,
            "l2": "flushed between steps (256 MiB write)" if flush is not None else "inputs exceed L2",
            "clocks": clocks, "wall_s": wall}


def gpu_arm(args, wl):
    from articulation3d_b200 import opt_utils, workloads

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU oracle port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    m = measure_pass(wl, args.steps, args.warmup, dev, rank, world, dist)
    # the same pass at throughput scale (one 8-GPU shard of configs[2]) rides along as context
    batched = None
    if args.batched and wl.name == "c2":
        wb = workloads.WORKLOADS["c3_shard"]
        mb = measure_pass(wb, max(3, min(args.steps, 10)), 3, dev, rank, world, dist, sample_clocks=False)
        batched = {"workload": wb.description, "value": mb["value"], "unit": UNIT, "ms_per_step": mb["ms_per_step"],
                   "units_per_step_per_gpu": mb["units_per_step_per_gpu"], "l2": mb["l2"],
                   "roofline": mb["roofline"]}

    line = None
    if rank == 0:
        pack = _pack_stream(dev)
        # ---- e2e through the public API with host buffers ------------------------------
        e2e = _e2e(wl, dev, opt_utils, workloads, steps=max(1, min(args.steps, 5)))
        # ---- CPU port on a bounded sample (N=1 only) -----------------------------------
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            tr, fr, _ = pick_sample(wl, budget_s=20.0)
            dt, u = run_oracle_sample(wl, 2020, tr, fr)
            cpu = {"value": u / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{tr} track(s) x {fr} frames of {wl.name}, optimize_planes('3dc'), {dt:.1f} s"}
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 popcount / fp32+fp64 geometry",
            "data": "synthetic",
            "config": {"workload": wl.description, "name": wl.name, "per_gpu": True,
                       "units_per_step_per_gpu": m["units_per_step_per_gpu"],
                       "packed_mask_bytes_per_gpu": m["packed_mask_bytes_per_gpu"], "l2": m["l2"],
                       "step": "one a3d_pass call (k_unproject, k_project, scoring kernel, k_finalize) on device-resident "
                               "inputs; roofline.kernels_ms from the same K steps issued as a3d_project | a3d_score",
                       "kernels": "chosen by the library from the grid: projection = reference chain per point with "
                                  "planned tiles up to two waves of CTAs (this workload), homography filter with proven "
                                  "truncation beyond (the batched shard); scoring = integer-pipe AND+POPC here, "
                                  "tcgen05 kind::i8 on the batched shard; identical results either way",
                       "parallelism": (f"videos sharded x{world}; per step one NCCL all_gather of 12 B/track-frame "
                                       f"records on a side stream (overlaps the next step)") if world > 1 else "single GPU"},
            "roofline": m["roofline"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": 4 * args.steps, "clocks": m["clocks"], "wall_s": m["wall_s"], "batched": batched, "pack": pack,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return line


def _pack_stream(dev, n_masks=1024, H=480, W=640, iters=5):
    """The one pure HBM stream of the path: fp32 (n,H,W) -> packed bits (a3d_pack_masks), timed
    alone with CUDA events; input (1.26 GB) exceeds L2.  Reported against the measured HBM peak."""
    from articulation3d_b200 import _lib
    lib = _lib.load()
    src = (torch.rand(n_masks, H, W, device=dev) > 0.7).float()
    pitch = _lib.pitch_words(W)
    bits = torch.empty(n_masks, H, pitch, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    for k in range(3 + iters):
        if k >= 3:
            ev[2 * (k - 3)].record()
        _lib.check(lib.a3d_pack_masks(src.data_ptr(), _lib.A3D_F32, n_masks, H, W, 0.5, bits.data_ptr(), None, stream),
                   "a3d_pack_masks")
        if k >= 3:
            ev[2 * (k - 3) + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters))
    nbytes = src.numel() * 4 + bits.numel() * 4
    peak, src_name = _peaks()
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "a3d_pack_masks (k_pack_vec<float>)", "bytes": nbytes, "ms": ms, "achieved": gbs, "unit": "GB/s",
            "peak": peak, "frac": gbs / peak, "peak_source": src_name, "bound": "hbm"}


def _run_split(engine, inp, ws, evs):
    """engine.run_pass with an event between a3d_project and a3d_score."""
    import ctypes as C

    from articulation3d_b200 import _lib
    lib = _lib.load()
    cfg, pool, db = inp.cfg, inp.pool, inp.dbatch
    H, W = cfg.height, cfg.width
    pitch = _lib.pitch_words(W)
    nc, nt = db.n_cand_total, db.n_tgt_total
    proj_bits = ws.get("proj_bits", (nc, H, pitch), torch.int32)
    proj_popc = ws.get("proj_popc", (nc,), torch.int32)
    proj_bbox = ws.get("proj_bbox", (nc, 4), torch.int32)
    pcd_ws = ws.get("pcd_ws", (max(_lib.PCD_PLANES * db.pcd_total, 32),), torch.float32)
    pcd_count = ws.get("pcd_count", (db.n_jobs + 1,), torch.int32)
    hom_ws = ws.get("hom_ws", (max(nc, 1), _lib.HOM_FLOATS), torch.float32)
    key_ws = ws.get("key_ws", (nt,), torch.int64)
    outs = [ws.get(n, (nt,), dt) for n, dt in (("best_cand", torch.int32), ("best_inter", torch.int32),
                                               ("best_union", torch.int32), ("best_iou", torch.float32))]
    cam = engine.camera_struct(cfg)
    stream = torch.cuda.current_stream().cuda_stream
    tile, tmap = db.tile_plan(cfg)
    tmap_ptr, n_tiles = (tmap.data_ptr(), int(tmap.shape[0])) if tmap is not None else (None, 0)
    _lib.check(lib.a3d_project(C.byref(cam), db.jobs.data_ptr(), db.n_jobs, db.max_cand, tile,
                               pool.source_bits.data_ptr(), pool.source_bbox.data_ptr(), db.xform.data_ptr(),
                               pcd_ws.data_ptr(), pcd_count.data_ptr(), hom_ws.data_ptr(), tmap_ptr, n_tiles,
                               proj_bits.data_ptr(), proj_popc.data_ptr(), proj_bbox.data_ptr(), stream),
               "a3d_project")
    if evs:
        evs[1].record()
    _lib.check(lib.a3d_score(H, W, db.jobs.data_ptr(), db.n_jobs, db.max_tgt, db.max_cand, nt,
                             len(pool), nc, pool.bits.data_ptr(), pool.popc.data_ptr(), pool.bbox.data_ptr(),
                             db.tgt_index.data_ptr(), proj_bits.data_ptr(), proj_popc.data_ptr(),
                             proj_bbox.data_ptr(), key_ws.data_ptr(), None, outs[0].data_ptr(),
                             outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), stream), "a3d_score")
    return engine.PassResult(outs[0], outs[1], outs[2], outs[3], proj_bits, proj_popc, proj_bbox, None)


def _e2e(wl, dev, opt_utils, workloads, steps):
    """Public API, host buffers: fp32 masks in pinned host memory -> optimize_planes('3dc')."""
    n_videos = min(wl.videos, 4)
    clips, seeds = [], []
    for v in range(n_videos):
        preds, cfg = workloads.make_clip(wl, 2020 + v)
        for p in preds:
            p.pred_masks = p.pred_masks.pin_memory()
        clips.append(preds)
        seeds.append(2020 + v)

    def run():
        vids = [(preds, opt_utils.track_planes(preds, cfg)) for preds in clips]
        st = opt_utils.Stats()
        if n_videos == 1:
            random.seed(seeds[0])
            opt_utils.optimize_planes(vids[0][0], vids[0][1], "3dc", cfg=cfg, device=dev, stats=st)
        else:
            opt_utils.optimize_videos(vids, seeds, cfg=cfg, device=dev, stats=st)
        torch.cuda.synchronize()
        return st

    # optimize_planes mutates pred_tran_axis in place; restore between runs
    saved = [[(p.pred_tran_axis.clone(), p.pred_rot_axis.clone()) for p in preds] for preds in clips]

    def restore():
        for preds, sv in zip(clips, saved):
            for p, (ta, ra) in zip(preds, sv):
                p.pred_tran_axis = ta.clone()
                p.pred_rot_axis = ra.clone()

    run()
    restore()
    secs, units, h2d, d2h, passes = 0.0, 0, 0, 0, 0
    for _ in range(steps):
        t0 = time.perf_counter()
        st = run()
        secs += time.perf_counter() - t0
        units += st.units_visited
        h2d += st.h2d_bytes
        d2h += st.d2h_bytes
        passes += st.passes
        restore()
    return {"value": units / secs, "unit": UNIT, "h2d_bytes_per_step": h2d // steps,
            "d2h_bytes_per_step": d2h // steps, "ms_per_step": 1e3 * secs / steps,
            "api": f"optimize_planes(preds, planes, '3dc') on {n_videos} video(s), fp32 host masks (pinned), "
                   f"units = visited (frame, candidate) pairs as the reference counts them",
            "device_passes_per_step": passes // steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", dest="batched", action="store_false",
                    help="skip the extra c3_shard measurement reported under 'batched'")
    args = ap.parse_args()
    from articulation3d_b200 import workloads
    wl = workloads.WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, wl)
    else:
        gpu_arm(args, wl)


if __name__ == "__main__":
    main()
