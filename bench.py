#!/usr/bin/env python
"""bench.py — the SAME hot path on B200: candidate pairs/s and triangle checks/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--tiles T]

Workload (BASELINE.json configs[3]): a synthetic section of 2,500 tiles
(~1.03 M reference / ~0.93 M query cells, K=3, coordinates x50), radius=250, knn=8,
window_size=5000, overlap=250 -> a 7x7 sliding-window grid processed as ONE batch per GPU.
A step = one pass of the whole hot path over that batch: window subsetting, candidate search +
pair costs, triangle remap/filter/tables, constraint grouping, lazy separation of one incumbent,
post-solve analysis.  `value` = candidate pairs emitted / time of the candidate stage
(subset + bin + search + compaction + cost), the unit BASELINE.md defines; triangle checks/s,
the per-stage times and the full-pass time ride along in the same JSON line.

N > 1 (torchrun): weak scaling — every rank owns one strip of N x 2,500 tiles; the cells of the
neighbouring strip that its last window row reaches into are exchanged once with an NCCL
all_gather before the timed region (the only collective on the path).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RADIUS, KNN, WINDOW, OVERLAP, MIN_ANGLE, MIN_CELLS = 250.0, 8, 5000, 250, 15.0, 10
SCALE, N_TYPES = 50.0, 3
TILES_PER_ROW = 50


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="same_b200", choices=["same_b200", "reference"])
    ap.add_argument("--tiles", type=int, default=2500, help="tiles per GPU (2500 = the 1M-cell section)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work the cpu_baseline leg aims for")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
def make_workload(tiles, rank, world):
    """This rank's strip of the section: frames as arrays + the global window list restricted to its rows."""
    from same_b200 import datagen
    per_row = TILES_PER_ROW if tiles >= TILES_PER_ROW else max(1, int(np.sqrt(tiles)))
    ref, qry, ct = datagen.make_section_pair(n_tiles=tiles, n_types=N_TYPES, seed=2 + 101 * rank, scale=SCALE, tiles_per_row=per_row)
    rows = -(-tiles // per_row)
    strip_h = rows * datagen.TILE_PITCH * SCALE
    for df in (ref, qry):
        df["Y"] = df["Y"] + rank * strip_h
    lut = {c: i for i, c in enumerate(ct)}
    t_prep = time.perf_counter()
    W = dict(
        a_xy=np.ascontiguousarray(qry[["X", "Y"]].to_numpy(np.float64)), r_xy=np.ascontiguousarray(ref[["X", "Y"]].to_numpy(np.float64)),
        a_prob=np.ascontiguousarray(qry[ct].to_numpy(np.float64)), r_prob=np.ascontiguousarray(ref[ct].to_numpy(np.float64)),
        a_type=qry["cell_type"].map(lut).to_numpy(np.int32), r_type=ref["cell_type"].map(lut).to_numpy(np.int32),
        strip_h=strip_h, strip_lo=rank * strip_h, ct=ct)
    W["prep_ms"] = (time.perf_counter() - t_prep) * 1e3
    return W


def exchange_halo(W, rank, world, device):
    """The one collective of the weak-scaling arm (same_b200.sharding.exchange_halo: a neighbour send/recv over NCCL): cells of the
    NEXT strip within `WINDOW` of the strip border, so that windows starting in this strip see every cell of their rectangle.
    Device-resident: the strip's columns are on the GPU (where the section lives anyway); selection, packing, transfer and
    unpacking run there.  Timed after the communicator is warm (NCCL builds its channels on the first point-to-point call)."""
    import torch
    import torch.distributed as dist
    from same_b200 import sharding as S
    S.exchange_halo({"y": np.full((8, 1), -1.0)}, "y", 0.0, device=device)          # communicator warm-up: 8 rows each way
    dev_frames = {}
    for name in ("a", "r"):
        dev_frames[name] = {"xy": torch.from_numpy(W[f"{name}_xy"]).to(device), "prob": torch.from_numpy(W[f"{name}_prob"]).to(device),
                            "type": torch.from_numpy(W[f"{name}_type"]).to(device)}
    for rep in range(2):     # the first full-size exchange also loads torch's selection kernels and sizes NCCL's buffers: timed is the second
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        nbytes = rows = 0
        out = {}
        for name in ("a", "r"):
            out[name], info = S.exchange_halo(dev_frames[name], ("xy", 1), W["strip_lo"] + WINDOW)
            nbytes += info["bytes"]; rows += info["rows"]
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
    for name in ("a", "r"):          # (the bench's end-to-end arm starts from HOST frames, so the received rows are also appended there)
        halo = {k: v.cpu().numpy() for k, v in out[name].items()}
        W[f"{name}_xy"] = np.ascontiguousarray(np.concatenate([W[f"{name}_xy"], halo["xy"]]))
        W[f"{name}_prob"] = np.ascontiguousarray(np.concatenate([W[f"{name}_prob"], halo["prob"]]))
        W[f"{name}_type"] = np.concatenate([W[f"{name}_type"], halo["type"][:, 0].astype(np.int32)])
    return dict(ms=ms, bytes=int(nbytes), halo_cells=int(rows), note="device-resident neighbour send/recv of packed rows (native dtypes): "
                "selection and packing on the GPU, NCCL point-to-point over NVLink, unpacking on the GPU; second of two identical exchanges "
                "(the first one loads kernels and sizes NCCL's buffers); rank 0's clock")


def window_rects(W, rank, world):
    """Window grid over the global section (same.py:481-488); this rank runs the window rows that START in its strip."""
    x_min = min(W["a_xy"][:, 0].min(), W["r_xy"][:, 0].min())
    x_max = max(W["a_xy"][:, 0].max(), W["r_xy"][:, 0].max())
    y_lo, y_hi = W["strip_lo"], W["strip_lo"] + W["strip_h"]
    step = WINDOW - OVERLAP
    xw = list(range(int(x_min), int(x_max), step))
    y_max_global = world * W["strip_h"]
    yw_all = list(range(0, int(y_max_global), step))
    yw = [y for y in yw_all if y_lo <= y < y_hi]
    return np.asarray([[x, x + WINDOW, y, y + WINDOW] for x in xw for y in yw], dtype=np.float64), (len(xw), len(yw))


def triangulate(a_xy):
    from scipy.spatial import Delaunay
    return Delaunay(a_xy).simplices.astype(np.int64)


def incumbent(pairs, seed=0):
    """x = 1 on one pseudo-randomly chosen pair of 90 % of the aligned rows."""
    rng = np.random.default_rng(seed)
    x = np.zeros(len(pairs))
    i = pairs[:, 0].astype(np.int64)
    new_row = np.r_[True, (i[1:] != i[:-1])] if len(i) else np.zeros(0, bool)
    starts = np.flatnonzero(new_row)
    counts = np.diff(np.r_[starts, len(i)])
    pick = rng.integers(0, 1 << 30, size=len(starts)) % np.maximum(counts, 1)
    sel = starts + pick
    x[sel[rng.uniform(size=len(starts)) < 0.9]] = 1.0
    return x


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def oracle_candidates(W, rects, sel):
    """CPU arm, candidate stage (window subsetting + KNN + pair cost) of the windows `sel`: the C oracle with OpenMP on
    all host cores.  -> (pairs, seconds)"""
    from oracle import oracle as O
    pairs, t = 0, 0.0
    for w in sel:
        rect = rects[w]
        t0 = time.perf_counter()
        ra, rr = O.subset(W["a_xy"], *rect), O.subset(W["r_xy"], *rect)
        axy, rxy = W["a_xy"][ra], W["r_xy"][rr]
        keepA, keepR, pr = O.find_knn_within_radius(axy, rxy, RADIUS, KNN)
        O.pair_cost(pr, axy[keepA], rxy[keepR], W["a_prob"][ra][keepA], W["r_prob"][rr][keepR], 1.0)
        t += time.perf_counter() - t0
        pairs += len(pr)
    return pairs, t


def oracle_separation(W, rects, sel):
    """CPU arm, one separation call per window of `sel` (matching from x + orientation test per triangle); the
    triangulation and tables it runs on are built untimed.  -> (triangles, seconds)"""
    from oracle import oracle as O
    from scipy.spatial import Delaunay
    n_tri, t = 0, 0.0
    for w in sel:
        rect = rects[w]
        ra, rr = O.subset(W["a_xy"], *rect), O.subset(W["r_xy"], *rect)
        axy, rxy = W["a_xy"][ra], W["r_xy"][rr]
        keepA, keepR, pr = O.find_knn_within_radius(axy, rxy, RADIUS, KNN)
        if len(keepA) < 4:
            continue
        tri = Delaunay(axy[keepA]).simplices.astype(np.int32)
        kept, _, _ = O.filter_triangles(axy[keepA], tri, RADIUS, MIN_ANGLE, W["a_type"][ra][keepA], True)
        tri = tri[kept]
        tt = O.tri_tables(axy[keepA], np.ones(len(keepA)), tri)
        x = incumbent(pr, seed=int(w))
        t0 = time.perf_counter()
        mj, _ = O.matching_from_x(x, pr, len(keepA))
        O.separation(tri, tt["sign"], mj, rxy[keepR])
        t += time.perf_counter() - t0
        n_tri += len(tri)
    return n_tri, t


def cpu_arm(W, rects, target_s):
    """Candidate stage of the WHOLE window list, repeated until about `target_s` seconds of CPU work; separation on 8 windows."""
    all_w = np.arange(len(rects))
    oracle_candidates(W, rects, all_w[:2])                       # warm caches / OpenMP pool
    p, t = oracle_candidates(W, rects, all_w)
    reps = 1
    while t < target_s and reps < 200:
        p2, t2 = oracle_candidates(W, rects, all_w)
        p += p2; t += t2; reps += 1
    tri, ts = oracle_separation(W, rects, np.linspace(0, len(rects) - 1, min(8, len(rects))).astype(int))
    return dict(pairs=p, seconds=t, reps=reps, tri=tri, tri_seconds=ts)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port — the Python reference cannot travel to the GPU
    box and takes ~2 ms/pair in its cost loop) on all host threads; a step = the candidate stage of the whole window list."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = O.set_threads(len(os.sched_getaffinity(0)))   # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    W = make_workload(args.tiles, 0, 1)
    rects, grid = window_rects(W, 0, 1)
    all_w = np.arange(len(rects))
    for _ in range(max(args.warmup, 1)):
        oracle_candidates(W, rects, all_w[:4])
    tot_pairs, tot_t = 0, 0.0
    for _ in range(args.steps):
        p, t = oracle_candidates(W, rects, all_w)
        tot_pairs += p; tot_t += t
    tri, ts = oracle_separation(W, rects, np.linspace(0, len(rects) - 1, min(8, len(rects))).astype(int))
    value = tot_pairs / tot_t
    line = {
        "impl": "reference", "metric": "candidate_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, grid, len(rects)),
        "triangle_checks": {"value": tri / max(ts, 1e-12), "unit": "triangle checks/s"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"all {len(rects)} windows per step x {args.steps} steps, candidate stage (subset + KNN + cost); OpenMP C oracle; {tot_t:.1f} s"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_same_wall(tiles, seed=1):
    """`run_same` wall clock, solver excluded (BASELINE metric, second half; configs[0] = 5 tiles, configs[1] = 25 tiles): the public
    API end to end on host DataFrames — frames prepared, GPU stages, model arrays fetched, ONE separation call on a seeded
    incumbent, post-solve analysis, matches frame + var_out built — with `IncumbentBackend` in place of Gurobi.  The parameters
    are examples/synthetic/run_same.sh's; tools/time_reference_c2.py times the unmodified reference on the same input."""
    import same_b200
    from same_b200 import datagen
    from same_b200.solver import IncumbentBackend
    ref, qry, ct = datagen.make_section_pair(n_tiles=tiles, n_types=3, seed=seed)
    optim = dict(radius=1.0, knn=8, max_matches=2, min_angle_deg=5, cell_id_col="Cell_Num_Old", dist_ct_coeff=1, ignore_same_type_triangles=False,
                 delaunay_penalty=10, no_match_penalty=10000, penalty_coeff=100, lazy_constraints=True)
    gurobi = dict(mip_gap=0.025, lazy_allowed_flip_fraction=0.0, time_limit=7200, mip_focus=2)

    def inc(spec):
        rp = np.asarray(spec.row_ptr, dtype=np.int64)
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        return incumbent(np.column_stack([rows, rows]), seed=seed)

    import contextlib, io
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            matches, var_out = same_b200.run_same(ref, qry, list(ct), outprefix=None, optim_params=dict(optim), gurobi_params=dict(gurobi),
                                                  solver=IncumbentBackend(inc))
        times.append(time.perf_counter() - t0)
    return {"cells": [len(ref), len(qry)], "pairs": len(var_out["x"]), "triangles": len(var_out["triangle_data"]["triangles"]),
            "cuts": int(var_out["lazy_cuts_added"]), "matches": len(matches), "seconds": min(times), "seconds_first_call": times[0]}


def strong_scaling_arm(args, rank, world, local_rank, stream):
    """The product's own multi-GPU partition (sharding.distributed_sliding_window_matching, src/same.py:507-593 run per block):
    EVERY rank holds the whole configs[3] section (replicated frames, identical seed) and runs one contiguous block of its window
    list; strong scaling — the total work is fixed.  Device-timed candidate stage (max over ranks) and the same thing end to end
    through CandidateStream (each rank uploads the whole section, downloads its block's results)."""
    import gc
    import torch
    import torch.distributed as dist
    from same_b200 import _lib as L
    from same_b200.device import CandidateStream, Section
    from same_b200.windows import shard_windows
    W1 = make_workload(args.tiles, 0, 1)
    rects_all, grid = window_rects(W1, 0, 1)
    lo, hi = shard_windows(len(rects_all), world, rank)
    mine = np.ascontiguousarray(rects_all[lo:hi])
    frames = tuple(torch.from_numpy(W1[k]).pin_memory().numpy() for k in ("a_xy", "r_xy", "a_prob", "r_prob", "a_type", "r_type"))
    device = torch.device("cuda", local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    sec = Section(*frames, device=local_rank, stream=stream)

    def once(ev=None):
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
        b = sec.batch(mine)
        b.candidates(RADIUS, KNN, False, 1.0)
        if ev is not None:
            e1 = torch.cuda.Event(enable_timing=True); e1.record(); ev.append((e0, e1))
        n = b.length(L.PAIRS)
        b.close()
        return n
    for _ in range(3):
        pairs = once()
    torch.cuda.synchronize(); dist.barrier()
    ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        ev = []
        pairs = once(ev)
        torch.cuda.synchronize()
        ms += ev[0][0].elapsed_time(ev[0][1])
    ms /= args.steps
    sec.close()
    # end to end: the whole section goes up on every rank, the block's results come back
    n_e2e = max(2, min(args.steps, 10))
    L.set_host_wait(world * 3 > len(os.sched_getaffinity(0)) // 2)
    with CandidateStream(RADIUS, KNN, False, 1.0, device=local_rank, j16=True) as cs:
        def stream_n(n):
            prev = None
            for _ in range(n):
                h = cs.submit(frames, mine)
                if prev is not None:
                    out = prev.result()
                prev = h
            out = prev.result()
            return sum(out[w].nbytes for w in CandidateStream.ARRAYS)
        stream_n(8)
        cs.reserve()
        torch.cuda.synchronize(); dist.barrier()
        gc.collect(); gc.disable()
        t0 = time.perf_counter()
        d2h = stream_n(n_e2e)
        torch.cuda.synchronize(); dist.barrier()
        e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        gc.enable()
    # the same thing the B200 way: every rank uploads 1/world of the rows over its own PCIe link, the shards are all-gathered over
    # NVLink (sharding.section_from_row_shards), the block's results come back
    from same_b200 import sharding as S
    torch.cuda.set_stream(torch.cuda.Stream(device=device))

    def sharded_once():
        s2 = S.section_from_row_shards(frames, device_index=local_rank)
        b = s2.batch(mine)
        b.candidates(RADIUS, KNN, False, 1.0)
        got = b.get_many(list(CandidateStream.ARRAYS))
        nb = sum(v.nbytes for v in got.values())
        b.close()
        s2.close()
        return nb
    for _ in range(4):
        sharded_once()
    torch.cuda.synchronize(); dist.barrier()
    gc.collect(); gc.disable()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        d2h_s = sharded_once()
    torch.cuda.synchronize(); dist.barrier()
    g_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    gc.enable()
    L.set_host_wait(False)
    t = torch.tensor([ms, e_ms, float(pairs), g_ms], dtype=torch.float64, device=device)
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    allr = torch.stack(allr).cpu().numpy()
    tot_pairs = float(allr[:, 2].sum())
    h2d_all = int(sum(f.nbytes for f in frames))
    return {"scaling": "strong", "value": tot_pairs / (allr[:, 0].max() * 1e-3), "unit": "pairs/s", "ms_per_step": float(allr[:, 0].max()),
            "ms_per_step_per_rank": [float(v) for v in allr[:, 0]], "pairs_per_step": int(tot_pairs), "windows_total": int(len(rects_all)),
            "windows_per_rank": [int(shard_windows(len(rects_all), world, r)[1] - shard_windows(len(rects_all), world, r)[0]) for r in range(world)],
            "e2e": {"value": tot_pairs / (allr[:, 3].max() * 1e-3), "unit": "pairs/s", "ms_per_step": float(allr[:, 3].max()),
                    "ms_per_step_per_rank": [float(v) for v in allr[:, 3]], "h2d_bytes_per_step": int(-(-h2d_all // world)),
                    "d2h_bytes_per_step": int(d2h_s), "nvlink_allgather_bytes_per_step": h2d_all,
                    "how": "row-sharded upload (1/world of every frame per rank over its own PCIe link) + all_gather over NVLink "
                           "(sharding.section_from_row_shards), kernels on the rank's window block, download of the block's results"},
            "e2e_replicated_upload": {"value": tot_pairs / (allr[:, 1].max() * 1e-3), "unit": "pairs/s", "ms_per_step": float(allr[:, 1].max()),
                                      "ms_per_step_per_rank": [float(v) for v in allr[:, 1]],
                                      "h2d_bytes_per_step": int(sum(f.nbytes for f in frames[:4])),    # (CandidateStream leaves the type codes on the host)
                                      "d2h_bytes_per_step": int(d2h),
                                      "how": "every rank uploads the whole section (CandidateStream), no collective"},
            "note": "ONE 2,500-tile section held by every rank, window list in contiguous blocks (same_b200.windows.shard_windows, the "
                    "partition sharding.distributed_sliding_window_matching uses); the only collective is the all_gather of the uploaded row shards"}


def luad_shape_arm(rank, world, window_size):
    """BASELINE configs[4] (examples/luad/run_same.sh:92-108) at metacell scale: 40,000 reference / 37,600 query cells uniform over
    13,000^2 units, K = 5 Dirichlet probabilities, sizes 1..3 as after greedy_triangle_collapse(max_metacell_size=3) (the collapse
    itself: tools/time_collapse_luad.py), the script's optim parameters, through the PUBLIC sliding_window_matching with
    window_shard=(rank, world) and IncumbentBackend in place of Gurobi.  Wall clock per rank."""
    import contextlib, io
    import same_b200
    from same_b200 import datagen
    from same_b200.solver import IncumbentBackend
    ref, qry, ct = datagen.make_uniform_pair(40000, 37600, 13000.0, n_types=5, seed=4, id_col="metacell_id")
    rng = np.random.default_rng(44)
    ref["size"] = rng.integers(1, 4, len(ref)).astype(float)
    qry["size"] = rng.integers(1, 4, len(qry)).astype(float)
    optim = dict(window_size=window_size, overlap=250, min_cells_per_window=30, max_matches=1, radius=250, knn=8, no_match_penalty=10000,
                 dist_ct_coeff=1, penalty_coeff=100, delaunay_penalty=10, cell_id_col="metacell_id", ref_metacell_match_multiplier=3)
    gurobi = dict(mip_gap=0.05, lazy_allowed_flip_fraction=0.05)

    def inc(spec):
        rp = np.asarray(spec.row_ptr, dtype=np.int64)
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        return incumbent(np.column_stack([rows, rows]), seed=len(rows))
    best, out = None, None
    for _ in range(2):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out = same_b200.sliding_window_matching(ref, qry, commonCT=list(ct), optim_params=dict(optim), gurobi_params=dict(gurobi),
                                                    solver=IncumbentBackend(inc), window_shard=(rank, world))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, (0 if out is None else len(out)), (0 if out is None or not len(out) else int(out["window_id"].nunique()))


def separation_latency_arm(local_rank):
    """Per-incumbent latency of the lazy separation at real window sizes: configs[0] (~2 k cells), configs[1]
    (~10 k cells) and one configs[3] window (~30 k cells).  Gurobi hands the callback a HOST vector, so the call is H2D of x +
    k_match_rows + k_separation + one synchronisation + D2H of counts and cuts; each part is also timed on its own."""
    import torch
    from scipy.spatial import Delaunay
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section, pinned_empty
    out = {}
    for label, tiles in (("configs[0] ~2k cells", 5), ("configs[1] ~10k cells", 25), ("one configs[3] window ~30k cells", 75)):
        ref, qry, ct = datagen.make_section_pair(n_tiles=tiles, n_types=N_TYPES, seed=7, scale=SCALE)
        lut = {c: i for i, c in enumerate(ct)}
        a_xy, r_xy = qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy()
        with Section(a_xy, r_xy, qry[ct].to_numpy(), ref[ct].to_numpy(), qry["cell_type"].map(lut).to_numpy(np.int32),
                     ref["cell_type"].map(lut).to_numpy(np.int32), device=local_rank) as sec, sec.batch() as b:
            b.candidates(RADIUS, KNN, False, 1.0)
            keepA = b.get(L.KEEP_A)
            tri = Delaunay(a_xy[keepA]).simplices.astype(np.int32)
            b.triangles_set(tri, [0, len(tri)])
            b.tri_classify(RADIUS, MIN_ANGLE, True)
            b.tri_finalize(True, True, False)
            pairs = b.get(L.PAIRS)
            x_host = pinned_empty(len(pairs), np.float64)
            x_host[:] = incumbent(pairs, seed=tiles)
            x_dev = torch.from_numpy(np.asarray(x_host)).to(torch.device("cuda", local_rank))
            cap = 1000

            def timed(fn, n=300):
                for _ in range(20):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(n):
                    t0 = time.perf_counter()
                    fn()
                    ts.append(time.perf_counter() - t0)
                return float(np.median(ts) * 1e6), float(np.percentile(ts, 95) * 1e6)
            host_med, host_p95 = timed(lambda: b.separation(x_host, cap=cap))
            dev_med, dev_p95 = timed(lambda: b.separation(int(x_dev.data_ptr()), cap=cap))

            def h2d_only():
                x_dev.copy_(torch.from_numpy(np.asarray(x_host)), non_blocking=True)
                torch.cuda.synchronize()
            h2d_med, _ = timed(h2d_only)
            L.profile_enable(True)
            for _ in range(50):
                b.separation(int(x_dev.data_ptr()), cap=cap)
            prof = L.profile_report()
            L.profile_enable(False)
            kern = {k: v[1] / v[0] * 1e3 for k, v in prof.items() if k in ("k_match_rows", "k_copy_words") or k.startswith("k_separation")}
            nv, nc, _ = b.separation(x_host, cap=cap)
            out[label] = {"cells": [len(ref), len(qry)], "pairs": int(len(pairs)), "triangles": int(b.length(L.TRI)), "checked": int(nc[0]),
                          "violated": int(nv[0]), "call_us_host_x": host_med, "call_us_host_x_p95": host_p95, "call_us_device_x": dev_med,
                          "call_us_device_x_p95": dev_p95, "h2d_of_x_alone_us": h2d_med, "x_bytes": int(x_host.nbytes),
                          "d2h_bytes": int(cap * 16 + 12), "kernel_us": kern}
    out["note"] = ("median wall time of same_batch_separation per call (300 calls): with the solution vector in page-locked HOST memory "
                   "(what the MIPSOL callback has), with a device-resident vector, the H2D of x alone (cudaMemcpyAsync + synchronise), "
                   "and the device time of each kernel of the call (events, separate pass).  cap = 1000 cuts come back per call")
    return out


def host_link_probe(device, world):
    """The ceiling of any transfer-bound end-to-end number on THIS box: every rank copies 128 MiB host->device and 128 MiB
    device->host at the same time (page-locked memory, two streams, plain cudaMemcpyAsync), all ranks together.
    -> GB/s per rank (both directions summed, slowest rank) and the aggregate over ranks."""
    import torch
    import torch.distributed as dist
    n = 128 << 20
    h_up, h_dn = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up, d_dn = torch.empty(n, dtype=torch.uint8, device=device), torch.zeros(n, dtype=torch.uint8, device=device)
    s_up, s_dn = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

    def once():
        with torch.cuda.stream(s_up):
            d_up.copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_dn.copy_(d_dn, non_blocking=True)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 8
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    per_rank = 2 * n * reps / dt / 1e9
    return {"per_rank_gb_per_s": per_rank, "aggregate_gb_per_s": per_rank * world,
            "note": "128 MiB up + 128 MiB down per rank at once, page-locked memory, all ranks together, slowest rank's clock"}


def workload_config(args, world, grid, n_windows):
    return {"workload": f"BASELINE configs[3]: synthetic {args.tiles}-tile section per GPU (~{args.tiles * 411 / 1e6:.2f}M ref / ~{args.tiles * 372 / 1e6:.2f}M query cells, K={N_TYPES}), "
                        f"candidate+cost+triangle+separation kernels, sliding window {grid[0]}x{grid[1]} per GPU",
            "radius": RADIUS, "knn": KNN, "window_size": WINDOW, "overlap": OVERLAP, "min_angle_deg": MIN_ANGLE, "windows_per_gpu": n_windows,
            "tiles_per_gpu": args.tiles, "sharding": f"window rows by strip, {world} rank(s)", "l2": "flushed between timed steps (512 MiB write)"}


# ----------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything else any library writes to fd 1 (NCCL prints its version banner there when
    NCCL_DEBUG is set in the environment) is sent to stderr; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from same_b200 import _lib as L
    from same_b200.device import Section

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: same_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    if world > 1:
        # one process per GPU: run on the CPUs next to this rank's GPU so that its pinned buffers are first-touched on the local
        # NUMA node (eight ranks pushing ~230 MB per step each through one socket halve the PCIe rate otherwise)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception as e:      # affinity is an optimisation only
            sys.stderr.write(f"[bench] rank {rank}: CPU affinity not set ({e})\n")
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout = the one JSON line
        dist.init_process_group("nccl", device_id=device)

    # ---- workload (untimed setup) ----
    W = make_workload(args.tiles, rank, world)
    halo = exchange_halo(W, rank, world, device) if world > 1 else None
    t_h = time.perf_counter()
    rects, grid = window_rects(W, rank, world)
    host_ms = {"window_grid": (time.perf_counter() - t_h) * 1e3}
    t_h = time.perf_counter()
    tri_vid = triangulate(W["a_xy"])            # global (per-strip) Delaunay, the "precomputed triangulation" of the examples
    host_ms["qhull_delaunay_of_the_aligned_frame"] = (time.perf_counter() - t_h) * 1e3
    if rank == 0:
        # the other route of the product (no precomputed triangulation): one Delaunay per window, which run_same spreads over the
        # host's cores (Qhull runs outside the GIL) — here on the cells of each window rectangle
        from concurrent.futures import ThreadPoolExecutor
        from scipy.spatial import Delaunay
        from same_b200.same import host_threads
        a = W["a_xy"]
        t_h = time.perf_counter()
        sel = [a[(a[:, 0] >= r[0]) & (a[:, 0] < r[1]) & (a[:, 1] >= r[2]) & (a[:, 1] < r[3])] for r in rects]
        with ThreadPoolExecutor(max_workers=host_threads()) as ex:
            n_tri = sum(ex.map(lambda p: len(Delaunay(p).simplices) if len(p) >= 3 else 0, sel))
        host_ms["qhull_delaunay_per_window_on_host_threads"] = {"ms": (time.perf_counter() - t_h) * 1e3, "threads": host_threads(),
                                                                "windows": len(rects), "triangles": int(n_tri)}
    host_ms["frames_to_arrays"] = W["prep_ms"]
    host_ms["note"] = ("host work of the path that is NOT on the GPU and NOT inside any timed region above: scipy/Qhull Delaunay of the whole "
                       "aligned frame (the precomputed triangulation every example hands to sliding_window_matching), the window grid, "
                       "DataFrame -> contiguous arrays.  e2e.full_path excludes them; a section's wall clock does not")
    # one explicit stream for everything: the library launches on it and the stage events are recorded on it
    # (the default stream has handle 0, which the library reads as "create your own stream")
    tstream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=device)

    def pinned(a):
        t = torch.from_numpy(a).pin_memory()
        return t, t.numpy()

    pins = {k: pinned(W[k]) for k in ("a_xy", "r_xy", "a_prob", "r_prob", "a_type", "r_type")}
    tri_pin = pinned(tri_vid)

    def make_section():
        sec = Section(pins["a_xy"][1], pins["r_xy"][1], pins["a_prob"][1], pins["r_prob"][1], pins["a_type"][1], pins["r_type"][1],
                      device=local_rank, stream=stream)
        sec.set_triangles(tri_pin[1], None)
        return sec

    sec = make_section()
    x_dev = {"t": None}
    STAGES = ["subset", "candidates", "triangles", "groups", "separation", "postsolve"]

    def one_pass(sec, ev=None, fetch=False):
        """The hot path over this rank's window batch.  ev: list to receive stage-boundary events."""
        def mark():
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append(e)
        mark()
        b = sec.batch(rects)
        mark()
        b.candidates(RADIUS, KNN, False, 1.0)
        mark()
        b.triangles_remap()
        b.tri_classify(RADIUS, MIN_ANGLE, True)
        b.tri_finalize(True, True, True)
        mark()
        b.groups(1, None)
        mark()
        if x_dev["t"] is None:       # first (warm-up) pass: build the incumbent once, keep it resident
            x_dev["t"] = torch.from_numpy(incumbent(b.get(L.PAIRS), seed=rank)).to(device)
        nv, nc, cuts = b.separation(int(x_dev["t"].data_ptr()), cap=1000)
        mark()
        b.postsolve(int(x_dev["t"].data_ptr()))
        mark()
        stats = dict(evals=b.stat(L.STAT_KNN_EVALUATIONS), P=b.length(L.PAIRS), T=b.length(L.TRI), Tin=b.length(L.TRI_IN), nKA=b.length(L.KEEP_A), nKR=b.length(L.KEEP_R),
                     nAi=b.length(L.WIN_A), nRi=b.length(L.WIN_R), G=b.length(L.REF_GROUP_NODE), viol=int(nv.sum()), checked=int(nc.sum()))
        d2h = 0
        if fetch:   # everything the host model builder consumes, into pinned buffers, one synchronisation
            got = b.get_many([L.KEEP_A, L.KEEP_R, L.PAIRS, L.COST, L.TRI, L.TRI_WEIGHT, L.TRI_SIGN, L.REF_GROUP_NODE, L.REF_GROUP_PTR,
                              L.REF_GROUP_IDX, L.REF_GROUP_LIMIT, L.ROW_PTR, L.FLIPPED, L.TRI_MASK])
            d2h = sum(v.nbytes for v in got.values()) + cuts.nbytes + nv.nbytes + nc.nbytes
        b.close()
        return stats, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        stats, _ = one_pass(sec)
    barrier()

    # ---- timed region: K steps, device-timed per stage, max over ranks ----
    sampler = ClockSampler(local_rank)
    launches0 = L.launch_count()
    stage_ms = np.zeros(len(STAGES))
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_buf.fill_(1)              # L2 flush between timed iterations
        ev = []
        stats, _ = one_pass(sec, ev)
        torch.cuda.synchronize()
        stage_ms += np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(len(STAGES))])
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    gpu_launches = L.launch_count() - launches0
    clocks = sampler.stop()
    stage_ms /= args.steps

    # ---- e2e: host buffers in, host results out, through the public device API ----
    # (1) the metric's own scope — the candidate stage, exactly what `--impl reference` times (subset + KNN + cost):
    #     pinned frames -> H2D -> subset, search, compaction, cost -> D2H of kept rows, pairs and costs;
    # (2) the whole hot path (adds the triangulation upload, the triangle/group/separation/post-solve stages and
    #     everything the host model builder consumes coming back), reported beside it as `full_path`.
    e2e = None
    if not args.no_e2e:
        sec.close()
        n_e2e = max(10, min(2 * args.steps, 20))    # sections in the stream (the first result pays the whole latency: pipeline fill)
        x_pin = pinned(x_dev["t"].cpu().numpy())                 # the incumbent comes from the host solver in real use
        saved = x_dev["t"]
        import gc

        # valid_pairs in compact form: i is implied by ROW_PTR, j is window-local and fits 16 bits (every window keeps < 65,536 rows)
        CAND_ARRAYS = [L.KEEP_A, L.KEEP_R, L.ROW_PTR, L.PAIR_J, L.COST]
        CAND_ARRAYS_ONCE = [L.KEEP_A, L.KEEP_R, L.ROW_PTR, L.PAIR_J16, L.COST]
        frames_pinned = (pins["a_xy"][1], pins["r_xy"][1], pins["a_prob"][1], pins["r_prob"][1], pins["a_type"][1], pins["r_type"][1])

        def cand_once():
            """ONE section, nothing overlapped: the latency of a single call (upload -> kernels -> download)."""
            s2 = Section(*frames_pinned, device=local_rank, stream=stream)
            b = s2.batch(rects)
            b.candidates(RADIUS, KNN, False, 1.0)
            got = b.get_many(CAND_ARRAYS_ONCE)
            npairs, nbytes = b.length(L.PAIRS), sum(v.nbytes for v in got.values())
            b.close()
            s2.close()
            return npairs, nbytes

        from same_b200.device import CandidateStream
        # several ranks x (main thread + two section threads) on a host with few cores: waiting threads sleep instead of spinning
        yield_wait = (os.environ.get("SAME_B200_HOST_WAIT") == "yield") if "SAME_B200_HOST_WAIT" in os.environ else world * 3 > len(all_cpus) // 2
        L.set_host_wait(yield_wait)
        E2E_DEPTH = int(os.environ.get("SAME_BENCH_E2E_DEPTH", "3"))   # measured: 2 -> 2.41, 3 -> 2.23, 4 -> no better (ms per section)
        cstream = CandidateStream(RADIUS, KNN, False, 1.0, device=local_rank, j16=True, depth=E2E_DEPTH)

        def cand_stream(n):
            """n sections back to back through CandidateStream: section k+1 is submitted (upload + kernels on its own stream)
            before the downloads of section k are awaited, so the two PCIe directions and the kernels overlap.  Every section's
            inputs go up from page-locked host memory and every section's results come down inside the timed region."""
            from collections import deque
            queue, npairs, nbytes, marks = deque(), 0, 0, [time.perf_counter()]
            for _ in range(n):
                if len(queue) == E2E_DEPTH:                     # at most `depth` sections outstanding: take the oldest result first
                    out = queue.popleft().result()
                    marks.append(time.perf_counter())
                queue.append(cstream.submit(frames_pinned, rects))
            while queue:
                out = queue.popleft().result()
                marks.append(time.perf_counter())
            from same_b200 import device as DV
            sys.stderr.write(f"[bench] e2e cand_stream results arrive after ms: {[round((b - a) * 1e3, 2) for a, b in zip(marks, marks[1:])]} "
                             f"(page-locked blocks obtained so far {DV.PINNED_ALLOCS[0]}, device pool {L.mempool_stats(local_rank)[0] >> 20} MiB)\n")
            return len(out[L.PAIR_J]), sum(out[w].nbytes for w in CAND_ARRAYS)

        def full_once():
            s2 = make_section()                                  # H2D of both frames + triangulation from pinned memory
            x_dev["t"] = x_pin[0].to(device, non_blocking=True)  # H2D of the solution vector
            st, nbytes = one_pass(s2, fetch=True)
            s2.close()
            return st["P"], nbytes

        def timed(fn):
            for _ in range(2):                                   # untimed: first touch of the pinned result pool and of the memory pool
                fn()
            barrier()
            gc.collect()
            gc.disable()                                         # a gen-2 collection inside the timed loop costs 10+ ms
            t0 = time.perf_counter()
            its = []
            for _ in range(n_e2e):
                t1 = time.perf_counter()
                npairs, nbytes = fn()
                its.append(round((time.perf_counter() - t1) * 1e3, 2))
            barrier()
            ms = (time.perf_counter() - t0) * 1e3 / n_e2e
            gc.enable()
            sys.stderr.write(f"[bench] e2e {fn.__name__} iterations ms: {its}\n")
            return ms, npairs, nbytes

        c_ms, c_P, c_d2h = timed(cand_once)
        # throughput of a stream of sections (the headline e2e): n_e2e sections through CandidateStream, timed as a whole
        gc.collect(); gc.disable()
        cand_stream(8)      # untimed: both streams' share of the memory pool and the page-locked result pool fill up
        cstream.reserve()   # head-room in the device pool: no cudaMalloc inside the timed stream
        barrier()
        t0 = time.perf_counter()
        s_P, s_d2h = cand_stream(n_e2e)
        barrier()
        s_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        gc.enable()
        sys.stderr.write(f"[bench] e2e cand_stream: {s_ms:.2f} ms per section over {n_e2e} sections\n")
        assert s_P == c_P
        L.set_host_wait(False)
        f_ms, f_P, f_d2h = timed(full_once)
        x_dev["t"] = saved
        frames = sum(p[1].nbytes for p in pins.values())
        cand_h2d = sum(p[1].nbytes for k, p in pins.items() if k not in ("a_type", "r_type"))   # CandidateStream leaves the type codes on the host
        e2e = dict(ms=s_ms, latency_ms=c_ms, h2d=cand_h2d, d2h=s_d2h, P=s_P, full_ms=f_ms, full_h2d=frames + tri_pin[1].nbytes + x_pin[1].nbytes,
                   full_d2h=f_d2h, full_P=f_P)
        sec = make_section()

    # ---- per-kernel device times (separate pass so the headline is unperturbed) ----
    L.profile_enable(True)
    for _ in range(3):
        flush_buf.fill_(1)
        pstats, _ = one_pass(sec)
    knn_evals = pstats["evals"]
    prof = L.profile_report()
    L.profile_enable(False)

    # ---- the product's own partition (strong scaling of ONE section) and the LUAD-shape section through the public API ----
    extra = {}

    def arm(name, fn):
        """Side arms never cost the main line: a failure is recorded under the arm's key (collectives inside an arm are entered by
        every rank, so a failure on one rank is a failure on all and nothing is left waiting)."""
        try:
            out = fn()
            if out is not None:
                extra[name] = out
        except Exception as e:
            extra[name] = {"error": repr(e)}
            sys.stderr.write(f"[bench] arm {name} failed: {e!r}\n")

    if world > 1 and not args.no_e2e:
        sec.close()
        arm("strong_scaling", lambda: strong_scaling_arm(args, rank, world, local_rank, stream))
        sec = make_section()
    if not args.no_e2e:
        arm("host_link_probe", lambda: host_link_probe(device, world))
    if not args.no_e2e and rank == 0:
        arm("separation_callback_latency", lambda: separation_latency_arm(local_rank))

    def luad():
        out = {}
        for ws in (13000, 4000):
            secs, n_match, n_win = luad_shape_arm(rank, world, ws)
            t = torch.tensor([secs, float(n_match), float(n_win)], dtype=torch.float64, device=device)
            if world > 1:
                allr = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allr, t)
                allr = torch.stack(allr).cpu().numpy()
            else:
                allr = t.cpu().numpy()[None]
            out[f"window_size_{ws}"] = {"seconds": float(allr[:, 0].max()), "seconds_per_rank": [float(v) for v in allr[:, 0]],
                                        "matches": int(allr[:, 1].sum()), "windows": int(allr[:, 2].sum()),
                                        "windows_per_rank": [int(v) for v in allr[:, 2]]}
        out["note"] = (
            "BASELINE configs[4] at metacell scale (40,000 / 37,600 cells over 13,000^2, K=5, sizes 1..3, examples/luad/run_same.sh:92-104 "
            "parameters) through the public sliding_window_matching(window_shard=(rank, world)) with IncumbentBackend (one seeded "
            "incumbent + one separation call per window, no MIP); wall clock, max over ranks.  window_size_13000 is the script's "
            "setting (one large window + border slivers: it cannot use more than a few GPUs); window_size_4000 cuts 4x4 windows")
        return out
    if not args.no_e2e:
        arm("configs4_luad_shape", luad)

    # ---- reduce over ranks ----
    def allmax(v):
        if world == 1:
            return np.asarray(v, dtype=np.float64)
        t = torch.tensor(np.asarray(v, dtype=np.float64), device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def allsum(v):
        if world == 1:
            return np.asarray(v, dtype=np.float64)
        t = torch.tensor(np.asarray(v, dtype=np.float64), device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    stage_max = allmax(stage_ms)
    if world > 1:
        tt = torch.tensor(stage_ms, dtype=torch.float64, device=device)
        allr = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allr, tt)
        stage_per_rank = [[float(v) for v in r.cpu().numpy()] for r in allr]
    else:
        stage_per_rank = [[float(v) for v in stage_ms]]
    tot = allsum([stats["P"], stats["T"], stats["nAi"], stats["nRi"], stats["checked"], stats["viol"], e2e["P"] if e2e else 0,
                  e2e["full_P"] if e2e else 0])
    e2e_max, e2e_full_max, e2e_lat_max = allmax([e2e["ms"] if e2e else 0.0, e2e["full_ms"] if e2e else 0.0, e2e["latency_ms"] if e2e else 0.0])
    cand_ms = stage_max[0] + stage_max[1]
    sep_ms = stage_max[4]
    full_ms = float(stage_max.sum())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        K = N_TYPES
        s = stats
        alg = {   # algorithmic bytes per launch, this rank's batch
            "k_knn<8>": 20 * s["nAi"] + 20 * s["nRi"] + (4 * KNN + 4) * s["nAi"],
            # candidate table in; one packed record (x, y, K probabilities padded to an even count) per kept row of both frames and
            # one (kept index, row) word per kept reference instance gathered; pair + cost out
            "k_emit_pairs": (4 * KNN + 4) * s["nAi"] + 8 * ((2 + K + 1) // 2 * 2) * (s["nKA"] + s["nKR"]) + 8 * s["nKR"] + 16 * s["P"],
            "k_separation": 13 * s["T"] + 4 * s["nKA"] + 16 * s["nKR"] + 16 * min(s["viol"], 1000 * len(rects)),
            "k_postsolve": 12 * s["T"] + 20 * s["nKA"] + 16 * s["nKR"] + 21 * s["T"],
            "k_tri_classify": 12 * s["Tin"] + 20 * s["nKA"] + 9 * s["Tin"],
            "k_tri_tables": 12 * s["T"] + 24 * s["nKA"] + (8 + 1 + 32 + 16) * s["T"],
            "k_match_rows": 8 * s["P"] + 4 * s["nKA"] + 8 * s["nKA"] + 8 * s["nKA"],   # x streamed; row_ptr; ONE pair read per row; two outputs
        }
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        ncu_metrics = {}
        try:
            ncu_metrics = json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics.json")))
        except Exception:
            pass
        kernels = {}
        for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            name = re.sub(r"^(k_emit_pairs|k_separation|k_remap_count|k_subset_count)<\d+>$", r"\1", name)   # (instantiated per record width / tile size)
            if name in kernels:
                continue                         # (a second instantiation of the same kernel: the larger share is already listed)
            per = ms / cnt
            ent = {"launches_per_step": cnt / 3.0, "avg_ms": per, "share_of_step": None}
            if name in alg:
                gbs = alg[name] / (per * 1e-3) / 1e9
                ent.update({"algorithmic_bytes": int(alg[name]), "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak})
            kernels[name] = ent
        total_k = sum(v["avg_ms"] * v["launches_per_step"] for v in kernels.values())
        for v in kernels.values():
            v["share_of_step"] = v["avg_ms"] * v["launches_per_step"] / total_k
        dom = max((n for n in kernels if n in alg), key=lambda n: kernels[n]["avg_ms"] * kernels[n]["launches_per_step"])
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic.get(dom), "peak_source": peak_src,
                    "avg_launch_ms": kernels[dom]["avg_ms"], "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                    "note": "the search kernel is bound by instruction issue (see `ncu`), not HBM (DESIGN.md §4); frac is against the HBM copy peak. "
                            "The largest HBM-bound kernel of the path is reported in `hbm_kernel`."}
        if dom in ncu_metrics:
            roofline["ncu"] = dict(ncu_metrics[dom], source="profiles/ncu_metrics.json (ncu --set full of this kernel, same command)")
        try:        # compute-side denominator, measured in this run: FP64 FMA micro-kernel
            roofline["fp64_peak_tflops_measured"] = L.fp64_peak_tflops(local_rank)
        except Exception as e:
            roofline["fp64_peak_tflops_measured"] = None
        if "k_knn<8>" in kernels and knn_evals and knn_evals > 0 and roofline["fp64_peak_tflops_measured"]:
            # compute side of the search kernel: every evaluated candidate costs 5 FP64 operations (2 subtractions, 2 products, 1 sum —
            # unfused by contract, so the ceiling is the FP64 INSTRUCTION rate, half of the FMA-counted peak)
            t_s = kernels["k_knn<8>"]["avg_ms"] * 1e-3
            inst_peak = roofline["fp64_peak_tflops_measured"] * 1e12 / 2.0
            kroof = {"distance_evaluations": int(knn_evals), "evaluations_per_query": knn_evals / max(s["nAi"], 1), "fp64_ops": int(5 * knn_evals),
                     "achieved_gflops": 5 * knn_evals / t_s / 1e9, "fp64_instruction_peak_gflops": inst_peak / 1e9,
                     "compute_frac": (5 * knn_evals / t_s) / inst_peak,
                     "hbm_frac": kernels["k_knn<8>"]["frac_of_hbm_peak"],
                     "bound": "neither: instruction issue under divergence (ncu: issue slots 70 % busy at 18 of 32 lanes)"}
            roofline["k_knn_compute"] = kroof
            if dom == "k_knn<8>":
                roofline["compute_frac"] = kroof["compute_frac"]
        hb = max((n for n in kernels if n in alg and n != dom), key=lambda n: kernels[n]["avg_ms"] * kernels[n]["launches_per_step"])
        roofline["hbm_kernel"] = {"kernel": hb, "bound": "hbm", "achieved": kernels[hb]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                                  "frac": kernels[hb]["frac_of_hbm_peak"], "traffic": traffic.get(hb), "avg_launch_ms": kernels[hb]["avg_ms"],
                                  "algorithmic_bytes_per_launch": kernels[hb]["algorithmic_bytes"]}
        value = tot[0] / (cand_ms * 1e-3)
        line = {
            "metric": "candidate_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": cand_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, grid, len(rects)),
            "triangle_checks": {"value": tot[1] / (sep_ms * 1e-3), "unit": "triangle checks/s", "ms_per_call": sep_ms,
                                "triangles": int(tot[1]), "checked": int(tot[4]), "violated": int(tot[5])},
            "full_pass_ms": full_ms, "stage_ms": {k: float(v) for k, v in zip(STAGES, stage_max)},
            "stage_ms_per_rank": {"stages": STAGES, "ms": stage_per_rank},
            "pairs_per_step": int(tot[0]), "cells_per_step": int(tot[2] + tot[3]), "wall_ms_per_step_incl_flush": wall_ms / args.steps,
            "gpu_launches": int(gpu_launches), "clocks": clocks,
            "roofline": roofline, "roofline_kernels": kernels, "host_ms": host_ms,
        }
        if e2e:
            line["e2e"] = {"value": tot[6] / (e2e_max * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(e2e["h2d"]),
                           "d2h_bytes_per_step": int(e2e["d2h"]), "ms_per_step": e2e_max,
                           "per_rank_gb_per_s": (e2e["h2d"] + e2e["d2h"]) / (e2e_max * 1e-3) / 1e9,
                           "aggregate_gb_per_s": world * (e2e["h2d"] + e2e["d2h"]) / (e2e_max * 1e-3) / 1e9,
                           "frac_of_host_link_probe": ((e2e["h2d"] + e2e["d2h"]) / (e2e_max * 1e-3) / 1e9) / extra["host_link_probe"]["per_rank_gb_per_s"]
                           if "per_rank_gb_per_s" in extra.get("host_link_probe", {}) else None,
                           "single_section_latency_ms": e2e_lat_max,
                           "host_wait": "yield (blocking-sync events)" if yield_wait else "spin",
                           "note": "candidate stage (the scope of `value` and of --impl reference: subset + KNN + cost) through the public API "
                                   "(same_b200.device.CandidateStream over the C-ABI) on a stream of sections: every section's frames go up from "
                                   "page-locked host memory and its kept rows, row pointers, pair reference indices (16-bit, window-local: SAME_ARR_PAIR_J16) and costs come back, all inside "
                                   "the timed region; section k+1 is submitted before section k's download is awaited, so both PCIe directions and "
                                   "the kernels overlap.  wall clock over all sections / sections, per rank, max over ranks.  "
                                   "single_section_latency_ms = one section alone, nothing overlapped",
                           "full_path": {"value": tot[7] / (e2e_full_max * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_full_max,
                                         "h2d_bytes_per_step": int(e2e["full_h2d"]), "d2h_bytes_per_step": int(e2e["full_d2h"]),
                                         "note": "every stage of the hot path (candidates, triangles, groups, one separation call, post-solve) "
                                                 "from pinned host frames + triangulation + incumbent to everything the host model builder reads"}}
        line.update(extra)
        if halo:
            line["halo_exchange"] = halo
        if world == 1 and not args.no_e2e:
            try:
                ref_cpu = {}
                for t in (5, 25):
                    try:
                        ref_cpu[t] = json.load(open(os.path.join(ROOT, "profiles", f"reference_cpu_c2_{t}tiles.json")))["seconds"]
                    except Exception:
                        ref_cpu[t] = None
                line["run_same_wall_clock"] = {
                    "note": "public run_same on host DataFrames, solver replaced by one seeded incumbent + one separation call (IncumbentBackend); "
                            "reference_seconds = the unmodified Python reference on the same input and rule, timed in the build container "
                            "(tools/time_reference_c2.py; it cannot travel to the GPU box)",
                    "configs[0] (5 tiles, ~2k cells)": dict(run_same_wall(5), reference_seconds=ref_cpu[5]),
                    "configs[1] (25 tiles, ~10k cells)": dict(run_same_wall(25), reference_seconds=ref_cpu[25])}
            except Exception as e:      # an extra, never the reason a bench line is missing
                line["run_same_wall_clock"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:   # (N > 1: the other ranks would spin in a barrier beside it; the reference arm is the CPU number there)
            from oracle import oracle as O
            O.build()
            os.sched_setaffinity(0, all_cpus)       # the CPU arm gets every host core back
            cores = O.set_threads(len(all_cpus))    # torchrun exports OMP_NUM_THREADS=1
            c = cpu_arm(W, rects, target_s=args.cpu_seconds)
            line["cpu_baseline"] = {"value": c["pairs"] / c["seconds"], "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": f"candidate stage of all {len(rects)} windows x {c['reps']} repetitions = {c['seconds']:.1f} s of CPU work; "
                                              f"OpenMP C oracle (the Python reference cannot travel to the GPU box)",
                                    "triangle_checks_per_s": c["tri"] / max(c["tri_seconds"], 1e-12)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
