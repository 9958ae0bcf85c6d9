"""`run_same` and `sliding_window_matching` with the reference's signatures and DataFrame contract
(src/same.py:297-595, 706-1489).  Host orchestration stays Python; every array loop of the reference —
candidate search, pair costs, constraint grouping, triangle remap/filter/tables, lazy separation, post-solve
analysis, window subsetting — runs in libsame_b200 on the GPU.  Gurobi stays the host solver.

There is no CPU fallback: without the CUDA library / a CUDA device these functions raise.
"""
from __future__ import annotations

import os
import time
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd
from scipy.spatial import Delaunay

from . import _lib as L
from . import helpers as H
from . import windows as WN
from .frames import build_section
from .params import init_gurobi_params, init_optim_params
from .solver import ModelSpec, get_backend
from .violationhelper import eager_report, violations_from_mask


# --------------------------------------------------------------------------------------------------------
def _as_triangle_array(delaunay_like):
    """Normalise a triangulation-like object to an int ndarray (n, 3) (src/same.py:245-259)."""
    if delaunay_like is None:
        return None
    if isinstance(delaunay_like, np.ndarray):
        tri = delaunay_like
    elif isinstance(delaunay_like, pd.DataFrame):
        tri = delaunay_like.iloc[:, :3].to_numpy()
    else:
        tri = np.asarray(delaunay_like)
    if tri.size == 0:
        return np.array([], dtype=int).reshape(0, 3)
    if tri.ndim != 2 or tri.shape[1] != 3:
        raise ValueError(f"aligned_delaunay must have shape (n, 3); got {tri.shape}")
    return tri.astype(int, copy=False)


def _remap_triangles_by_vertex_ids(triangles, vertex_ids):
    """Triangles in vertex-id space -> 0..n-1 row indices, dropping triangles with a missing vertex
    (src/same.py:262-290).  GPU: id resolution (radix sort + binary search) and the remap kernel."""
    tri = _as_triangle_array(triangles)
    if tri is None:
        return None
    if tri.size == 0:
        return tri
    from .device import Section
    vid = np.ascontiguousarray(vertex_ids, dtype=np.int64)
    n = len(vid)
    z = np.zeros((n, 2))
    z[:, 0] = np.arange(n)      # distinct points so that every row keeps itself as its only candidate
    with Section(z, z, np.zeros((n, 0)), np.zeros((n, 0))) as sec, sec.batch() as b:
        b.candidates(0.0, 1)
        sec.set_triangles(tri.astype(np.int64), vid)
        b.triangles_remap()
        return b.get(L.TRI_IN).astype(int)


def subset_data(df, x_min, x_max, y_min, y_max):
    """Half-open window (src/same.py:293-295) — host convenience; the driver subsets on the GPU."""
    return df[(df["X"] >= x_min) & (df["X"] < x_max) & (df["Y"] >= y_min) & (df["Y"] < y_max)]


def load_gurobi_config():
    """WLS credentials from `.gurobienv` next to this file (src/same.py:598-618)."""
    config = {}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), ".gurobienv")
    try:
        with open(path) as f:
            for line in f:
                line = line.strip()
                if line and not line.startswith("#") and "=" in line:
                    k, v = line.split("=", 1)
                    config[k.strip()] = v.strip()
    except FileNotFoundError:
        pass
    return config


def _env_options():
    cfg = load_gurobi_config()
    return {"WLSACCESSID": os.environ.get("GUROBI_WLSACCESSID", "") or cfg.get("WLSACCESSID", ""),
            "WLSSECRET": os.environ.get("GUROBI_WLSSECRET", "") or cfg.get("WLSSECRET", ""),
            "LICENSEID": int(os.environ.get("GUROBI_LICENSEID", 0)) or int(cfg.get("LICENSEID", 0))}


# --------------------------------------------------------------------------------------------------------
def host_threads():
    """Host threads for work that runs outside the GIL (per-window Qhull): SAME_B200_HOST_THREADS, else the cores this process
    may run on."""
    env = os.environ.get("SAME_B200_HOST_THREADS")
    if env:
        return max(1, int(env))
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class _Run:
    """A section + one batch of windows on the GPU, driven stage by stage; shared by run_same (one unbounded
    window) and sliding_window_matching (all runnable windows at once)."""

    def __init__(self, aligned_df, ref_df, commonCT, optim, rects=None, aligned_delaunay=None, vertex_ids=None,
                 ignore_precomputed=False, section=None):
        self.aligned_df, self.ref_df, self.commonCT, self.optim = aligned_df, ref_df, list(commonCT), optim
        o = optim
        have_types = "cell_type" in aligned_df.columns and "cell_type" in ref_df.columns
        if not have_types and (bool(o["ignore_same_type_triangles"]) or bool(o["ignore_knn_if_matched"])):
            # the reference reads df['cell_type'] in both cases (src/helpers.py:329, src/knn_utils.py:37) and fails with KeyError;
            # silently treating every cell as one type would drop every triangle / claim every nearest neighbour
            raise KeyError("cell_type")
        self.batch = None
        self.window_errors = {}
        self.section = section if section is not None else build_section(aligned_df, ref_df, self.commonCT)
        try:
            self._build(aligned_df, ref_df, o, rects, aligned_delaunay, vertex_ids, ignore_precomputed, have_types)
        except BaseException:
            # nothing leaks when a stage fails (CUDA out of memory, no GPU, a bad triangle array ...)
            if self.batch is not None:
                self.batch.close()
            self.section.close()
            raise

    def _build(self, aligned_df, ref_df, o, rects, aligned_delaunay, vertex_ids, ignore_precomputed, have_types):
        self.batch = self.section.batch(rects)
        self.W = self.batch.W
        self.batch.candidates(o["radius"], o["knn"], bool(o["ignore_knn_if_matched"]), o["dist_ct_coeff"])
        p_off = self.batch.offsets(L.PAIRS)
        self.n_pairs = np.diff(p_off)
        self.using_precomputed = aligned_delaunay is not None and not ignore_precomputed
        self._a_xy = aligned_df[["X", "Y"]].to_numpy(dtype=np.float64)
        self._a_type = None
        if self.using_precomputed:
            tri = _as_triangle_array(aligned_delaunay)
            self.section.set_triangles(tri.astype(np.int64), vertex_ids)
            self.batch.triangles_remap()                                              # same.py:1028-1031
        else:
            ka_off, keepA = self.batch.offsets(L.KEEP_A), self.batch.get(L.KEEP_A)

            def triangulate(w):
                if self.n_pairs[w] <= 0:
                    return np.zeros((0, 3), np.int32)
                try:
                    return Delaunay(self._a_xy[keepA[ka_off[w]:ka_off[w + 1]]]).simplices.astype(np.int32)   # same.py:1023 (Qhull stays on the host)
                except Exception as e:          # a degenerate window (QhullError): raised at that window's turn, like the reference's
                    self.window_errors[w] = e   # per-window loop, so the windows before it still run and checkpoint
                    return np.zeros((0, 3), np.int32)

            # Qhull runs outside the GIL (scipy wraps qh_new_qhull in `with nogil`): the windows are triangulated on the host's
            # cores at once — 8 windows of 30 k cells: 1.72 s one after the other, 0.42 s on 8 threads, identical simplices
            n_thr = min(self.W, host_threads())
            if n_thr > 1:
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(max_workers=n_thr, thread_name_prefix="same_b200-qhull") as ex:
                    tris = list(ex.map(triangulate, range(self.W)))
            else:
                tris = [triangulate(w) for w in range(self.W)]
            off = np.concatenate([[0], np.cumsum([len(t) for t in tris])]).astype(np.int64).tolist()
            self.batch.triangles_set(np.concatenate(tris) if tris else np.zeros((0, 3), np.int32), off)
        same = bool(o["ignore_same_type_triangles"])
        mad = o.get("min_angle_deg", 15)
        if self.batch.tri_classify(o["radius"], mad, same) > 0:
            if have_types:      # (only the guard band needs the codes on the host)
                from .frames import joint_type_codes
                self._a_type, _ = joint_type_codes(aligned_df["cell_type"].to_numpy(), ref_df["cell_type"].to_numpy())
            H.redecide_band(self.batch, self._a_xy, self._a_type, o["radius"], mad, same)
        self.batch.tri_finalize(same, True, remove_unconstrained=self.using_precomputed)   # same.py:1034-1085
        self.batch.groups(o["max_matches"], o["ref_metacell_match_multiplier"])            # helpers.py:105-138

    def close(self):
        if self.batch is not None:
            self.batch.close()
        self.section.close()


def _solve_window(run: _Run, w: int, aligned_src: pd.DataFrame, ref_src: pd.DataFrame, optim, gurobi, outprefix, backend):
    """Model build + solve + post-analysis of window `w` (src/same.py:1112-1481).  `aligned_src` / `ref_src` are the
    frames whose `.iloc[section rows]` give the window's post-KNN frames."""
    b = run.batch
    if w in run.window_errors:
        raise run.window_errors[w]
    t_stage = {"t0": time.perf_counter()}

    def lap(name):
        now = time.perf_counter()
        t_stage[name] = now - t_stage.pop("t0")
        t_stage["t0"] = now
    m = b.window_model(w)
    lap("fetch_model_arrays_s")
    if len(m["pairs"]) == 0:
        raise ValueError("No valid_pairs after KNN filtering. Increase radius and/or knn.")       # same.py:1002-1003
    commonCT = run.commonCT
    cell_id_col = optim["cell_id_col"]
    aligned_df = aligned_src.iloc[m["keepA"]].reset_index(drop=True)
    ref_df = ref_src.iloc[m["keepR"]].reset_index(drop=True)
    pairs, tri = m["pairs"], m["tri"]
    n_aligned, n_ref, P, T = len(aligned_df), len(ref_df), len(pairs), len(tri)
    lazy = bool(optim["lazy_constraints"])
    if not lazy:
        raise NotImplementedError("lazy_constraints=False (the O(n*k^3) eager builder, src/helpers.py:444-573) is outside the "
                                  "GPU hot path; use lazy_constraints=True (the reference's default)")
    # MIP start (src/same.py:1199-1215, src/init_helpers.py): 'greedy' for every window of the batch in one device pass
    # (same_batch_mip_start), 'hungarian' through the reference's dense host assignment
    start = None
    init_method = gurobi.get("init_method")
    if init_method is not None:
        method = str(init_method).lower()
        if method not in {"greedy", "hungarian"}:
            raise ValueError(f"Unknown init_method={init_method!r}. Use 'greedy' or 'hungarian'.")
        if method == "hungarian" and int(optim["max_matches"]) != 1:
            raise ValueError("init_method='hungarian' requires max_matches == 1.")
        if method == "greedy":
            if getattr(run, "_start_penalty", None) != float(optim["no_match_penalty"]):
                b.mip_start(float(optim["no_match_penalty"]))
                run._start_penalty = float(optim["no_match_penalty"])
            x0 = b.get_window(L.START_X, w).astype(np.float64)
            nm0 = b.get_window(L.START_UNMATCHED, w).astype(np.float64)
            start = (x0, nm0)
            print(f"Initialized MIP start (greedy): {int(x0.sum())} matches, {int(nm0.sum())} unmatched")
        else:
            from .init_helpers import mip_start_vectors
            start = mip_start_vectors(valid_pairs=pairs, costs=m["cost"], n_aligned=n_aligned, n_ref=n_ref,
                                      aligned_sizes=aligned_df["size"].to_numpy(dtype=float), no_match_penalty=optim["no_match_penalty"],
                                      max_matches=int(optim["max_matches"]), init_method=method, init_big_m=gurobi.get("init_big_m", 1e9),
                                      init_hungarian_max_n=gurobi.get("init_hungarian_max_n", 5000), verbose=True)

    spec = ModelSpec(n_pairs=P, n_ref=n_ref, n_aligned=n_aligned, n_tri=T, cost=m["cost"], row_ptr=m["row_ptr"],
                     ref_group_node=m["ref_group_node"], ref_group_ptr=m["ref_group_ptr"], ref_group_idx=m["ref_group_idx"],
                     ref_group_limit=m["ref_group_limit"], aligned_size=aligned_df["size"].to_numpy(dtype=np.float64),
                     tri_weight=m["weight"], penalty_coeff=optim["penalty_coeff"], no_match_penalty=optim["no_match_penalty"],
                     delaunay_penalty=optim["delaunay_penalty"])
    allowed = gurobi["lazy_allowed_flip_fraction"]
    per_inc = gurobi["lazy_max_cuts_per_incumbent"]
    lazy_max = gurobi["lazy_max_cuts"]

    a_xy_kept = aligned_df[["X", "Y"]].to_numpy(dtype=np.float64)
    r_xy_kept = ref_df[["X", "Y"]].to_numpy(dtype=np.float64)
    # exact-predicate diagnostic: how often could the naive fp64 orientation the reference (and this path) evaluates differ in sign
    # from the exact determinant?  Counted, never acted upon (helpers.exact_predicate_check).
    epc = {"source_signs": H.exact_predicate_check(b, w, 0, a_xy_kept, r_xy_kept), "separation_calls": 0,
           "separation_uncertain": 0, "separation_naive_differs_from_exact": 0}

    def separate(x_vals, cuts_so_far):
        """same.py:631-703 with the triangle loop on the GPU; returns the cuts to add, in order."""
        cap = T if per_inc is None else min(int(per_inc), T)
        if lazy_max is not None:
            cap = min(cap, max(0, int(lazy_max) - cuts_so_far))
        nv, nc, cuts = b.separation(x_vals, w, w + 1, cap=max(cap, 0))
        epc["separation_calls"] += 1
        if b.uncertain(1, cap=0)[0]:                 # the count came back with the call's own counters: no extra round trip
            chk = H.exact_predicate_check(b, w, 1, a_xy_kept, r_xy_kept, b.get_window(L.MATCH_J, w))
            epc["separation_uncertain"] += chk["uncertain"]
            epc["separation_naive_differs_from_exact"] += chk["naive_differs_from_exact"]
        viol, checked = int(nv[0]), int(nc[0])
        if checked == 0 or viol == 0:
            return np.zeros((0, 4), np.int32)
        if allowed is not None and viol / float(checked) <= allowed:
            return np.zeros((0, 4), np.int32)
        return cuts[0, :min(viol, cap)].copy()

    lap("prepare_model_s")
    if start is not None:
        res = backend.solve(spec, separate, gurobi, outprefix=outprefix, env_options=_env_options(), start=start)
    else:
        res = backend.solve(spec, separate, gurobi, outprefix=outprefix, env_options=_env_options())
    lap("solve_s")
    time_limit_reached = res.status == "time_limit"
    if res.status not in ("optimal", "time_limit"):
        out_df, var_out = pd.DataFrame(), {}
        if outprefix:
            os.makedirs(outprefix, exist_ok=True)
            out_df.to_csv(os.path.join(outprefix, "matches_df.csv"), index=False)
        return out_df, var_out

    x = np.asarray(res.x, dtype=np.float64)
    sel = np.flatnonzero(x > 0.5)
    ai, rj = pairs[sel, 0].astype(np.int64), pairs[sel, 1].astype(np.int64)
    cols = {"aligned_idx": ai, "ref_idx": rj}                                                      # same.py:1264-1278, one frame construction
    for ct in list(commonCT) + ["X", "Y"]:
        cols[ct] = aligned_df[ct].to_numpy()[ai]
    for ct in ["X", "Y"]:
        cols[f"ref_{ct}"] = ref_df[ct].to_numpy()[rj]
    cols["size"] = aligned_df["size"].to_numpy()[ai]
    cols["ref_size"] = ref_df["size"].to_numpy()[rj]
    cols[f"Ref_{cell_id_col}"] = ref_df[cell_id_col].to_numpy()[rj]
    cols[f"Aligned_{cell_id_col}"] = aligned_df[cell_id_col].to_numpy()[ai]
    cols["time_limit_reached"] = np.full(len(ai), time_limit_reached)
    out_df = pd.DataFrame(cols)

    # ---- post-solve analysis on the GPU (violationhelper.py:1-134, same.py:1355-1408) ----
    b.postsolve(x, w, w + 1)
    mask = b.get_window(L.TRI_MASK, w)
    area_before, area_after = b.get_window(L.AREA_BEFORE, w), b.get_window(L.AREA_AFTER, w)
    flipped = np.flatnonzero(b.get_window(L.FLIPPED, w))
    match_j = b.get_window(L.MATCH_J, w)
    # node -> triangle incidence from the device CSR (same.py:1096-1099): a node's triangles arrive ascending, i.e. in the order the
    # reference's triangle loop adds them to the node's set, so `set(list)` reproduces the reference's sets including their iteration order
    ka0 = int(b.offsets(L.KEEP_A)[w])
    nt_ptr = b.get(L.NODE_TRI_PTR, ka0, ka0 + n_aligned + 1).astype(np.int64)
    nt_len = b.get_window(L.NODE_TRI_LEN, w).astype(np.int64)
    nt_idx = b.get_window(L.NODE_TRI_IDX, w)
    base = int(nt_ptr[0]) if n_aligned else 0

    def build_simplex_map():
        flat = nt_idx.tolist()
        return {i: set(flat[s0:s0 + ln]) for i, (s0, ln) in enumerate(zip((nt_ptr[:-1] - base).tolist(), nt_len.tolist()))}
    aligned_simplex_map = build_simplex_map() if eager_report() else H.LazyDict(n_aligned, build_simplex_map)
    aligned_delaunay = tri.astype(int)
    order_cache = []

    def tri_order():      # key order of the reference's triangle_info (walks every node's set): computed when something first needs it
        if not order_cache:
            order_cache.append(H.triangle_info_order(n_aligned, aligned_simplex_map))
        return order_cache[0]
    eager = eager_report()
    triangle_info = H.precompute_triangle_info(aligned_df, aligned_delaunay, aligned_simplex_map, bounds=m["bounds"], argv=m["argv"],
                                               order=tri_order() if eager else tri_order, lazy=not eager, n_entries=T)
    a_xy = aligned_df[["X", "Y"]].to_numpy(dtype=np.float64)
    r_xy = ref_df[["X", "Y"]].to_numpy(dtype=np.float64)
    violations = violations_from_mask(mask, tri, match_j, a_xy, r_xy, tri_order() if eager else tri_order, T)
    pts = violations["points_with_violations"]
    violation_points = set(pts.unordered.tolist()) if getattr(pts, "unordered", None) is not None else set(pts)
    penalty_points = set()
    for t in np.flatnonzero(np.asarray(res.q) > 1e-6):                                             # same.py:1325-1346
        penalty_points.update(int(v) for v in tri[t])
    points_both = violation_points & penalty_points
    matched_bits = (mask >> 8) & 7
    per_triangle = (lambda build: build()) if eager else (lambda build: H.LazyDict(T, build))   # dict[t] -> value, built when read
    var_out = {
        "x": x.tolist(), "no_match_vars": np.asarray(res.no_match).tolist(), "penalty_vars": np.asarray(res.penalty).tolist(),
        "area_penalty_vars": np.asarray(res.q).tolist(), "violations": violations,
        "violation_penalty_comparison": {"points_both": list(points_both), "points_only_violations": list(violation_points - penalty_points),
                                         "points_only_penalties": list(penalty_points - violation_points)},
        "triangle_data": {
            "triangles": aligned_delaunay, "triangle_info": triangle_info, "aligned_simplex_map": aligned_simplex_map,
            "areas_before": per_triangle(lambda: dict(zip(range(T), area_before))),
            "areas_after": per_triangle(lambda: {t: (None if nan else a) for t, (a, nan) in enumerate(zip(area_after, np.isnan(area_after).tolist()))}),
            "flipped_triangles": flipped.tolist(),
            "matched_vertices": per_triangle(lambda: dict(zip(range(T), (((matched_bits[:, None] >> np.arange(3)) & 1) != 0).tolist())))},
        "lazy_constraints": lazy, "lazy_cuts_added": res.cuts_added if lazy else 0,
        "exact_predicate_check": epc,      # (extra key: diagnostic only, see helpers.exact_predicate_check)
    }
    lap("post_solve_analysis_s")
    # machine-readable stage timers of this window (the reference only has prints and Gurobi's own Runtime)
    timings = {k: float(v) for k, v in t_stage.items() if k != "t0"}
    timings.update(solver_runtime_s=float(res.runtime), separation_calls=int(epc["separation_calls"]), lazy_cuts_added=int(res.cuts_added),
                   n_pairs=int(P), n_triangles=int(T), n_aligned=int(n_aligned), n_ref=int(n_ref), solver=getattr(backend, "name", type(backend).__name__))
    var_out["timings"] = timings
    if outprefix:                                                                                  # same.py:1455-1463
        os.makedirs(outprefix, exist_ok=True)
        np.save(os.path.join(outprefix, "var_out.npy"), var_out, allow_pickle=True)
        aligned_df.to_csv(os.path.join(outprefix, "aligned_df.csv"), index=False)
        ref_df.to_csv(os.path.join(outprefix, "ref_df.csv"), index=False)
        import json
        with open(os.path.join(outprefix, "timings.json"), "w") as f:
            json.dump(timings, f, indent=1)
    flipped_nodes = np.unique(tri[flipped].reshape(-1)) if len(flipped) else np.zeros(0, np.int64)  # same.py:1466-1472
    out_df["triangle_violation"] = np.isin(ai, flipped_nodes)
    out_df["filtered_violation"] = np.isin(ai, np.fromiter(points_both, dtype=np.int64, count=len(points_both)))
    out_df["run_time"] = res.runtime
    if outprefix:
        out_df.to_csv(os.path.join(outprefix, "matches_df.csv"), index=False)
    return out_df, var_out


def _prepare_frames(ref_df, aligned_df, aligned_delaunay_vertex_col):
    """`size`, `__orig_idx`, `__tri_vid` helper columns on copies (src/same.py:934-970)."""
    if "size" not in aligned_df.columns:
        aligned_df = aligned_df.copy()
        aligned_df["size"] = 1
    if "size" not in ref_df.columns:
        ref_df = ref_df.copy()
        ref_df["size"] = 1
    aligned_df, ref_df = aligned_df.copy(), ref_df.copy()
    if "__orig_idx" not in aligned_df.columns:
        aligned_df["__orig_idx"] = aligned_df.index.to_numpy()
    if "__orig_idx" not in ref_df.columns:
        ref_df["__orig_idx"] = ref_df.index.to_numpy()
    if aligned_delaunay_vertex_col is None:
        aligned_df["__tri_vid"] = aligned_df.index.to_numpy()
    else:
        if aligned_delaunay_vertex_col not in aligned_df.columns:
            raise ValueError(f"aligned_delaunay_vertex_col='{aligned_delaunay_vertex_col}' not in aligned_df")
        aligned_df["__tri_vid"] = aligned_df[aligned_delaunay_vertex_col].to_numpy()
    return ref_df, aligned_df


def run_same(ref_df, aligned_df, commonCT, outprefix=None, aligned_delaunay=None, aligned_delaunay_vertex_col=None,
             optim_params: Optional[Dict[str, Any]] = None, gurobi_params: Optional[Dict[str, Any]] = None,
             ignore_precomputed_triangulation: bool = False, solver=None):
    """Optimal spatial matches between aligned and reference cells (reference `run_same`, src/same.py:706-1489).

    Same arguments, DataFrame contract, return value `(matches_df, var_out)` and error behaviour as the reference;
    `solver` (extra, optional) selects the host MIP back-end (`'gurobi'` default, `'highs'`, or an object)."""
    if gurobi_params is None:
        gurobi_params = {}
    if optim_params is None:
        optim_params = {}
    if hasattr(aligned_df, "metacell_df") and hasattr(aligned_df, "metacell_delaunay"):          # same.py:891-900
        mc = aligned_df
        aligned_df = mc.metacell_df
        if aligned_delaunay is None and not ignore_precomputed_triangulation:
            aligned_delaunay = mc.metacell_delaunay
        if aligned_delaunay_vertex_col is None and hasattr(mc, "metacell_idx_col"):
            aligned_delaunay_vertex_col = mc.metacell_idx_col
        if (optim_params.get("cell_id_col") is None) and hasattr(mc, "metacell_idx_col"):
            optim_params["cell_id_col"] = mc.metacell_idx_col
    if hasattr(ref_df, "metacell_df"):
        ref_df = ref_df.metacell_df
    optim = init_optim_params(**(optim_params or {}))
    gurobi = init_gurobi_params(**gurobi_params)
    ref_df, aligned_df = _prepare_frames(ref_df, aligned_df, aligned_delaunay_vertex_col)
    vid = aligned_df["__tri_vid"].to_numpy()
    run = _Run(aligned_df, ref_df, commonCT, optim, rects=None, aligned_delaunay=aligned_delaunay,
               vertex_ids=None if aligned_delaunay is None else vid.astype(np.int64),
               ignore_precomputed=ignore_precomputed_triangulation)
    try:
        return _solve_window(run, 0, aligned_df, ref_df, optim, gurobi, outprefix, get_backend(solver))
    finally:
        run.close()


def sliding_window_matching(ref, moving, commonCT=None, outprefix=None, moving_delaunay=None, moving_delaunay_vertex_col=None,
                            optim_params: Optional[Dict[str, Any]] = None, gurobi_params: Optional[Dict[str, Any]] = None,
                            ignore_precomputed_triangulation: bool = False, solver=None, window_shard=None):
    """Sliding-window matching (reference `sliding_window_matching`, src/same.py:297-595): same window grid, merge
    rule, central-region ownership, `window_id` and CSV checkpoint/resume.  All runnable windows are cut, searched
    and tabulated on the GPU as ONE batch; the per-window MIPs then run on the host in the reference's order.

    `window_shard=(rank, world_size)` (extra, optional) restricts this process to its contiguous block of the
    window list (one process per GPU; results are concatenated by the caller)."""
    ref_cell_type_col = moving_cell_type_col = "cell_type"
    if optim_params is None:
        optim_params = {}
    if gurobi_params is None:
        gurobi_params = {}
    if hasattr(ref, "metacell_df"):                                                                # same.py:415-421
        mc_ref = ref
        ref = mc_ref.metacell_df
        if hasattr(mc_ref, "cell_type_col"):
            ref_cell_type_col = mc_ref.cell_type_col
        if (optim_params.get("cell_id_col") is None) and hasattr(mc_ref, "metacell_idx_col"):
            optim_params["cell_id_col"] = mc_ref.metacell_idx_col
    if hasattr(moving, "metacell_df") and hasattr(moving, "metacell_delaunay"):                    # same.py:422-432
        mc = moving
        moving = mc.metacell_df
        if moving_delaunay is None and not ignore_precomputed_triangulation:
            moving_delaunay = mc.metacell_delaunay
        if moving_delaunay_vertex_col is None and hasattr(mc, "metacell_idx_col"):
            moving_delaunay_vertex_col = mc.metacell_idx_col
        if hasattr(mc, "cell_type_col"):
            moving_cell_type_col = mc.cell_type_col
        if (optim_params.get("cell_id_col") is None) and hasattr(mc, "metacell_idx_col"):
            optim_params["cell_id_col"] = mc.metacell_idx_col
    optim = init_optim_params(**(optim_params or {}))
    gurobi = init_gurobi_params(**(gurobi_params or {}))
    window_size, overlap = optim["window_size"], optim["overlap"]
    min_cells, cell_id_col = optim["min_cells_per_window"], optim["cell_id_col"]

    ref_types = mov_types = None                                                                   # same.py:445-458
    if ref_cell_type_col in ref.columns and moving_cell_type_col in moving.columns:
        ref_types = set(pd.Series(ref[ref_cell_type_col]).dropna().unique().tolist())
        mov_types = set(pd.Series(moving[moving_cell_type_col]).dropna().unique().tolist())
        if ref_types != mov_types:
            raise ValueError(
                f"Cell type categories differ between ref and moving.\n"
                f"ref ({ref_cell_type_col}) has {len(ref_types)} types, moving ({moving_cell_type_col}) has {len(mov_types)} types.\n"
                f"Only-in-ref: {sorted(ref_types - mov_types)[:20]}\n"
                f"Only-in-moving: {sorted(mov_types - ref_types)[:20]}")
    if commonCT is None:                                                                           # same.py:462-478
        if ref_types is None:
            raise ValueError("commonCT is None, but cell_type columns were not found to infer it. Pass commonCT explicitly "
                             f"(list of probability/one-hot columns), or ensure both dataframes have "
                             f"'{ref_cell_type_col}'/'{moving_cell_type_col}'.")
        commonCT = sorted(ref_types)
        missing_ref = [c for c in commonCT if c not in ref.columns]
        missing_mov = [c for c in commonCT if c not in moving.columns]
        if missing_ref or missing_mov:
            raise ValueError("commonCT is None so it was inferred as the unique values of the cell_type column, but those names are "
                             f"not present as probability/one-hot columns.\nMissing in ref columns (first 20): {missing_ref[:20]}\n"
                             f"Missing in moving columns (first 20): {missing_mov[:20]}\nEither rename your probability columns to "
                             "match cell type names, or pass commonCT explicitly.")

    x_min, x_max = min(ref["X"].min(), moving["X"].min()), max(ref["X"].max(), moving["X"].max())  # same.py:481-488
    y_min, y_max = min(ref["Y"].min(), moving["Y"].min()), max(ref["Y"].max(), moving["Y"].max())
    x_windows, y_windows = WN.window_grid(x_min, x_max, y_min, y_max, window_size, overlap)
    all_matches: List[pd.DataFrame] = []
    output_file = None
    if outprefix:
        os.makedirs(outprefix, exist_ok=True)
        output_file = os.path.join(outprefix, "matchedDF.csv")

    ref_p, mov_p = _prepare_frames(ref, moving, moving_delaunay_vertex_col)
    # cell counts of every rectangle the merge rule may ask for: one GPU launch (same.py:523-542)
    section = build_section(mov_p, ref_p, commonCT)
    rects, keys = WN.candidate_rects(x_windows, y_windows, window_size)
    cnt_mov, cnt_ref = np.zeros(len(rects), np.int64), np.zeros(len(rects), np.int64)
    for lo in range(0, len(rects), 60000):
        a, r = section.count_rects(rects[lo:lo + 60000])
        cnt_mov[lo:lo + 60000], cnt_ref[lo:lo + 60000] = a, r
    counter = WN.table_counter(keys, cnt_ref, cnt_mov)
    windows_to_process = None
    if outprefix:                                                                                  # same.py:497-501
        base_counts = {(i, j): int(cnt_mov[keys[(x, x + window_size, y, y + window_size)]])
                       for i, x in enumerate(x_windows) for j, y in enumerate(y_windows)}
        windows_to_process, existing = H.get_unprocessed_windows(moving, output_file, x_windows, y_windows, window_size, overlap,
                                                                 cell_id_col=cell_id_col, counts=base_counts)
        if existing is not None:
            all_matches.append(existing)
    wins = WN.enumerate_windows(x_windows, y_windows, window_size, overlap, min_cells, counter,
                                (int(x_min), int(x_max), int(y_min), int(y_max)), windows_to_process)
    runnable = [wd for wd in wins if wd.run]
    if window_shard is not None:
        lo, hi = WN.shard_windows(len(runnable), window_shard[1], window_shard[0])
        runnable = runnable[lo:hi]
    if not runnable:
        section.close()
        return pd.concat(all_matches, ignore_index=True) if all_matches else pd.DataFrame()

    vid = mov_p["__tri_vid"].to_numpy()      # (_Run closes the section itself when one of its stages fails)
    run = _Run(mov_p, ref_p, commonCT, optim, rects=np.asarray([wd.rect for wd in runnable], dtype=np.float64),
               aligned_delaunay=moving_delaunay, vertex_ids=None if moving_delaunay is None else vid.astype(np.int64),
               ignore_precomputed=ignore_precomputed_triangulation, section=section)
    backend = get_backend(solver)
    try:
        for w, wd in enumerate(runnable):
            window_outprefix = os.path.join(outprefix, f"window_{wd.window_id}") if outprefix else None
            window_matches, _ = _solve_window(run, w, mov_p, ref_p, optim, gurobi, window_outprefix, backend)
            if window_matches.shape[0] > 0:                                                        # same.py:564-590
                cx0, cx1, cy0, cy1 = wd.central
                central = window_matches[(window_matches["X"] >= cx0) & (window_matches["X"] < cx1) &
                                         (window_matches["Y"] >= cy0) & (window_matches["Y"] < cy1)].copy()
                central["window_id"] = wd.window_id
                if len(central) > 0:
                    all_matches.append(central)
                    if outprefix:
                        pd.concat(all_matches, ignore_index=True).to_csv(output_file, index=False)
    finally:
        run.close()
    return pd.concat(all_matches, ignore_index=True) if all_matches else pd.DataFrame()
