"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz

Runs only where `/root/reference` exists (the build container).  It imports the
reference under stub modules (`oracle/ref_loader.py`), drives every function of
the hot path on small inputs and freezes inputs +
outputs as `.npz`.  The committed `.npz` files are what the CPU and GPU parity
tests read; nothing at test/bench time touches `/root/reference`.

Inputs come from (i) the reference's own shipped example data (Fig-2 synthetic
411/372 cells, `examples/simulated_st` 144/144 cells; attribution in README.md
next to this file) and (ii) `same_b200.datagen` seeded sections.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from same_b200 import datagen  # noqa: E402

REF = ref_loader.load_reference()
import src.same as rsame  # noqa: E402
import src.helpers as rhelpers  # noqa: E402
import src.utils as rutils  # noqa: E402
import src.knn_utils as rknn  # noqa: E402
import src.violationhelper as rviol  # noqa: E402
from scipy.spatial import Delaunay  # noqa: E402


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def frame_arrays(df, commonCT, id_col, prefix):
    """Numeric restatement of a frame so tests can rebuild it without CSVs."""
    types, codes = np.unique(df["cell_type"].astype(str).to_numpy(), return_inverse=True)
    out = {
        f"{prefix}_xy": df[["X", "Y"]].to_numpy(np.float64),
        f"{prefix}_prob": df[list(commonCT)].to_numpy(np.float64),
        f"{prefix}_type_names": types.astype("U"),
        f"{prefix}_type_code": codes.astype(np.int32),
        f"{prefix}_id": (lambda v: np.asarray(v.tolist(), dtype="U") if v.dtype.kind in "OUT" else v)(df[id_col].to_numpy()),
        f"{prefix}_index": df.index.to_numpy(),
    }
    if "size" in df.columns:
        out[f"{prefix}_size"] = df["size"].to_numpy(np.float64)
    return out


def incumbent_rule(pairs, seed, p_match=0.9):
    """Deterministic pseudo-incumbent: each aligned row takes one of its pairs
    (uniform) with probability p_match.  Restated in tests/util.py."""
    pairs = np.asarray(pairs).reshape(-1, 2)
    rng = np.random.default_rng(seed)
    x = np.zeros(len(pairs))
    if len(pairs) == 0:
        return x
    i = pairs[:, 0]
    starts = np.flatnonzero(np.r_[True, i[1:] != i[:-1]])
    counts = np.diff(np.r_[starts, len(i)])
    u = rng.uniform(size=len(starts))
    pick = rng.integers(0, 1 << 30, size=len(starts)) % counts
    sel = starts + pick
    x[sel[u < p_match]] = 1.0
    return x


def make_incumbent_fn(seed):
    def fn(model):
        return incumbent_rule(np.asarray(model._valid_pairs).reshape(-1, 2), seed)
    return fn


def record_model(model, commonCT):
    """Arrays out of the recording fake (oracle/ref_loader.py)."""
    names = [v.VarName for v in model.vars]
    xs = [v.index for v in model.vars if v.VarName.startswith("x[")]
    qs = [v.index for v in model.vars if v.VarName.startswith("q_tri[")]
    pen = [v.index for v in model.vars if v.VarName.startswith("penalty[")]
    nom = [v.index for v in model.vars if v.VarName.startswith("no_match[")]
    obj = model.objective[0].terms
    out = {}
    out["pairs"] = np.asarray(model._valid_pairs, dtype=np.int64).reshape(-1, 2)
    out["cost"] = np.array([obj.get(i, 0.0) for i in xs], dtype=np.float64)
    out["obj_q"] = np.array([obj.get(i, 0.0) for i in qs], dtype=np.float64)
    out["obj_penalty"] = np.array([obj.get(i, 0.0) for i in pen], dtype=np.float64)
    out["obj_no_match"] = np.array([obj.get(i, 0.0) for i in nom], dtype=np.float64)
    out["n_vars"] = np.int64(len(names))
    tri = np.asarray(model._aligned_delaunay, dtype=np.int64).reshape(-1, 3)
    out["tri"] = tri
    out["source_signs"] = np.asarray(model._source_signs, dtype=np.float64)
    # constraints, in creation order, flattened CSR
    cname, csense, crhs, cptr, cidx, cval = [], [], [], [0], [], []
    for (nm, sense, terms, rhs) in model.constrs:
        cname.append(nm)
        csense.append(sense)
        crhs.append(rhs)
        for k, v in terms.items():
            cidx.append(k)
            cval.append(v)
        cptr.append(len(cidx))
    out["con_name"] = np.asarray(cname, dtype="U")
    out["con_sense"] = np.asarray(csense, dtype="U")
    out["con_rhs"] = np.asarray(crhs, dtype=np.float64)
    out["con_ptr"] = np.asarray(cptr, dtype=np.int64)
    out["con_idx"] = np.asarray(cidx, dtype=np.int64)
    out["con_val"] = np.asarray(cval, dtype=np.float64)
    # lazy cuts: x-var indices (pair indices) + triangle index
    cuts = []
    for terms, sense, rhs in model.lazy:
        p = [k for k, v in terms.items() if v > 0]
        q = [k for k, v in terms.items() if v < 0]
        cuts.append([xs.index(p[0]), xs.index(p[1]), xs.index(p[2]), qs.index(q[0])])
    out["cuts"] = np.asarray(cuts, dtype=np.int64).reshape(-1, 4)
    out["x_sol"] = np.array([model._sol[i] for i in xs])
    return out


def patched_optimize(self, callback=None):
    """Fake solve: seeded incumbent, ONE MIPSOL callback, and q_tri=1 on cut triangles
    (so that `filtered_violation`, same.py:1325-1346, is exercised)."""
    ref_loader.Model._orig_optimize(self, callback)
    for terms, sense, rhs in self.lazy:
        for k, v in terms.items():
            if v < 0:
                self.vars[k].x = 1.0


ref_loader.Model._orig_optimize = ref_loader.Model.optimize
ref_loader.Model.optimize = patched_optimize


def stage_records(ref_df, aligned_df, commonCT, radius, knn, min_angle_deg, seed):
    """Direct calls of the hot-path helper functions (a1, a2, a5, a6, a8, a11)."""
    out = {}
    a = aligned_df.copy()
    r = ref_df.copy()
    a["__orig_idx"] = np.arange(len(a))
    r["__orig_idx"] = np.arange(len(r))
    with quiet():
        a1, r1, pairs = rutils.find_knn_within_radius(a, r, radius, knn=knn)
    out["knn_keepA"] = a1["__orig_idx"].to_numpy(np.int64)
    out["knn_keepR"] = r1["__orig_idx"].to_numpy(np.int64)
    out["knn_pairs"] = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    with quiet():
        a2, r2, pairs2 = rknn.find_knn_with_cell_type_priority(a, r, radius, knn=knn)
    out["prio_pairs"] = np.asarray(pairs2, dtype=np.int64).reshape(-1, 2)
    # triangles on the post-KNN aligned frame
    pts = a1[["X", "Y"]].values
    tri = Delaunay(pts).simplices
    out["delaunay"] = tri.astype(np.int64)
    for flag in (False, True):
        with quiet():
            kept, unc = rhelpers.filter_triangles_by_radius(
                pts, tri, radius, aligned_df=a1, ignore_same_type_triangles=flag,
                remove_unconstrained_nodes=True, min_angle_deg=min_angle_deg)
        out[f"filt_same{int(flag)}"] = np.asarray(kept, dtype=np.int64).reshape(-1, 3)
        out[f"filt_same{int(flag)}_unc"] = np.asarray(sorted(unc), dtype=np.int64)
    # a tighter radius so that the radius filter and unconstrained nodes actually fire
    side = np.linalg.norm(pts[tri] - pts[np.roll(tri, 1, axis=1)], axis=2).max(axis=1)
    r_tight = float(np.quantile(side, 0.7))
    out["filt_tight_radius"] = np.float64(r_tight)
    with quiet():
        kept, unc = rhelpers.filter_triangles_by_radius(
            pts, tri, r_tight, aligned_df=a1, ignore_same_type_triangles=True,
            remove_unconstrained_nodes=True, min_angle_deg=min_angle_deg)
    out["filt_tight"] = np.asarray(kept, dtype=np.int64).reshape(-1, 3)
    out["filt_tight_unc"] = np.asarray(sorted(unc), dtype=np.int64)
    # a5 remap: global triangles in an id space, window = random 60% of rows
    rng = np.random.default_rng(seed + 17)
    vid = rng.permutation(len(a1)) * 3 + 7
    sub = np.sort(rng.choice(len(a1), size=int(0.6 * len(a1)), replace=False))
    out["remap_vid_all"] = vid.astype(np.int64)
    out["remap_rows"] = sub.astype(np.int64)
    out["remap_tri_global"] = vid[tri].astype(np.int64)
    out["remap_out"] = rsame._remap_triangles_by_vertex_ids(vid[tri], vid[sub]).astype(np.int64)
    return out


def run_case(name, ref_df, aligned_df, commonCT, optim, gurobi, id_col, seed,
             use_metacell=False, sliding=False, stage=True, mc_params=None):
    print(f"[golden] {name}: ref={len(ref_df)} aligned={len(aligned_df)} K={len(commonCT)}", flush=True)
    rec = {}
    rec.update(frame_arrays(ref_df, commonCT, id_col, "ref"))
    rec.update(frame_arrays(aligned_df, commonCT, id_col, "aligned"))
    rec["commonCT"] = np.asarray(commonCT, dtype="U")
    rec["id_col"] = np.asarray(id_col)
    rec["seed"] = np.int64(seed)
    for k, v in optim.items():
        rec[f"optim_{k}"] = np.asarray(np.nan if v is None else v)
    for k, v in gurobi.items():
        rec[f"gurobi_{k}"] = np.asarray(np.nan if v is None else v)
    if stage:
        rec.update(stage_records(ref_df, aligned_df, commonCT, optim["radius"], optim["knn"],
                                 optim.get("min_angle_deg", 15), seed))

    ref_in, al_in = ref_df, aligned_df
    tri_in, vcol = None, None
    if use_metacell:
        with quiet():
            mc_al = REF.greedy_triangle_collapse(aligned_df, cell_type_col="cell_type", original_idx_col=id_col,
                                                 return_object=True, **mc_params)
            mc_rf = REF.greedy_triangle_collapse(ref_df, cell_type_col="cell_type", original_idx_col=id_col,
                                                 return_object=True, **mc_params)
        rec["mc_aligned_delaunay"] = np.asarray(mc_al.metacell_delaunay, dtype=np.int64).reshape(-1, 3)
        rec["mc_aligned_metacell_id"] = mc_al.metacell_df["metacell_id"].to_numpy(np.int64)
        rec["mc_ref_metacell_id"] = mc_rf.metacell_df["metacell_id"].to_numpy(np.int64)
        assert np.allclose(mc_al.metacell_df[["X", "Y"]].to_numpy(), aligned_df[["X", "Y"]].to_numpy())
        ref_in, al_in = mc_rf, mc_al

    ref_loader.MODELS.clear()
    ref_loader.INCUMBENT_FN = make_incumbent_fn(seed)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            with quiet():
                if sliding:
                    matches = REF.sliding_window_matching(
                        ref_in, al_in, commonCT=list(commonCT), outprefix=os.path.join(td, "out"),
                        optim_params=dict(optim), gurobi_params=dict(gurobi))
                    var_out = None
                else:
                    matches, var_out = REF.run_same(ref_in, al_in, list(commonCT), outprefix=None,
                                                    optim_params=dict(optim), gurobi_params=dict(gurobi))
        finally:
            os.chdir(cwd)
    rec["n_models"] = np.int64(len(ref_loader.MODELS))
    for w, m in enumerate(ref_loader.MODELS):
        for k, v in record_model(m, commonCT).items():
            rec[f"w{w}_{k}"] = v
    # matches frame
    rec["matches_columns"] = np.asarray(list(matches.columns), dtype="U")
    for col in matches.columns:
        v = matches[col].to_numpy()
        if v.dtype == object:
            v = v.astype("U")
        rec[f"matches__{col}"] = v
    if var_out:
        td_ = var_out["triangle_data"]
        T = len(td_["triangles"])
        rec["vo_areas_before"] = np.array([td_["areas_before"][t] for t in range(T)], dtype=np.float64)
        rec["vo_areas_after"] = np.array(
            [np.nan if td_["areas_after"][t] is None else td_["areas_after"][t] for t in range(T)], dtype=np.float64)
        rec["vo_flipped"] = np.asarray(td_["flipped_triangles"], dtype=np.int64)
        rec["vo_matched_vertices"] = np.array([td_["matched_vertices"][t] for t in range(T)], dtype=bool).reshape(-1, 3)
        vio = var_out["violations"]
        s = vio["violation_summary"]
        rec["vo_summary"] = np.array([s["total_triangles"], s["violated_triangles"], s["total_comparisons"],
                                      s["total_violations"]], dtype=np.int64)
        rec["vo_percent"] = np.array([s["percent_triangles_violated"], s["percent_violations"]])
        rec["vo_tri_with_viol"] = np.asarray(sorted(vio["triangles_with_violations"]), dtype=np.int64)
        rec["vo_pts_with_viol"] = np.asarray(sorted(int(p) for p in vio["points_with_violations"]), dtype=np.int64)
        rec["vo_xviol"] = np.asarray([[d["triangle_idx"], d["point1"]["aligned_idx"], d["point2"]["aligned_idx"]]
                                      for d in vio["x_order_violations"]], dtype=np.int64).reshape(-1, 3)
        rec["vo_yviol"] = np.asarray([[d["triangle_idx"], d["point1"]["aligned_idx"], d["point2"]["aligned_idx"]]
                                      for d in vio["y_order_violations"]], dtype=np.int64).reshape(-1, 3)
        rec["vo_tri_info_order"] = np.asarray(list(td_["triangle_info"].keys()), dtype=np.int64)
        ti = td_["triangle_info"]
        rec["vo_tri_info"] = np.asarray(
            [[ti[t]["max_x_vertex"], ti[t]["min_x_vertex"], ti[t]["max_y_vertex"], ti[t]["min_y_vertex"]]
             for t in range(T) if t in ti], dtype=np.int64).reshape(-1, 4)
        rec["vo_tri_bounds"] = np.asarray(
            [[ti[t]["bounds"]["min_x"], ti[t]["bounds"]["max_x"], ti[t]["bounds"]["min_y"], ti[t]["bounds"]["max_y"]]
             for t in range(T) if t in ti], dtype=np.float64).reshape(-1, 4)
        rec["vo_lazy_cuts_added"] = np.int64(var_out["lazy_cuts_added"])
        vpc = var_out["violation_penalty_comparison"]
        rec["vo_points_both"] = np.asarray(sorted(int(p) for p in vpc["points_both"]), dtype=np.int64)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden]   -> {path} ({os.path.getsize(path) / 1024:.0f} KiB, {len(matches)} matches, "
          f"{len(ref_loader.MODELS)} window model(s))", flush=True)


def main(only=None):
    ex = os.path.join(ref_loader.REFERENCE_ROOT, "examples")
    cases = []

    # ---- Fig-2 data, the script's parameters (examples/synthetic/run_same.sh:34-133) ----
    ref = pd.read_csv(os.path.join(ex, "synthetic/data/ref.csv"), index_col=0)
    qry = pd.read_csv(os.path.join(ex, "synthetic/data/query.csv"), index_col=0)
    ct = ["c1", "c2", "c3"]
    g_fig2 = dict(mip_gap=0.025, lazy_allowed_flip_fraction=0.0, time_limit=7200, mip_focus=2)
    o_fig2 = dict(window_size=100, overlap=0, min_cells_per_window=30, max_matches=2, radius=5, knn=8,
                  no_match_penalty=10000, dist_ct_coeff=1, min_angle_deg=5, penalty_coeff=100,
                  delaunay_penalty=10, cell_id_col="metacell_id", ref_metacell_match_multiplier=1,
                  ignore_same_type_triangles=False, lazy_constraints=True)
    cases.append(("fig2_script", lambda: run_case(
        "fig2_script", ref, qry, ct, o_fig2, g_fig2, "cell_idx", seed=11, use_metacell=True, sliding=True,
        mc_params=dict(max_metacell_size=1, r_max=5, min_angle_deg=5, use_alpha_shape=False, alpha=None))))
    # ---- Fig-2 data, direct run_same with defaults-ish (fresh Delaunay, same-type filter on) ----
    o2 = dict(radius=1.2, knn=6, max_matches=1, min_angle_deg=15, cell_id_col="cell_idx", dist_ct_coeff=1,
              ignore_same_type_triangles=True, delaunay_penalty=5)
    cases.append(("fig2_direct", lambda: run_case(
        "fig2_direct", ref, qry, ct, o2, dict(lazy_allowed_flip_fraction=0.0), "cell_idx", seed=12)))
    # ---- priority KNN through run_same ----
    o3 = dict(o2, ignore_knn_if_matched=True, radius=1.5, knn=8, dist_ct_coeff=2.5)
    cases.append(("fig2_priority", lambda: run_case(
        "fig2_priority", ref, qry, ct, o3, dict(lazy_allowed_flip_fraction=0.02, lazy_max_cuts_per_incumbent=5),
        "cell_idx", seed=13, stage=False)))

    # ---- simulated_st known-answer: post-KNN frames shipped by the reference ----
    sa = pd.read_csv(os.path.join(ex, "simulated_st/aligned_df.csv"))
    sr = pd.read_csv(os.path.join(ex, "simulated_st/ref_df.csv"))
    for d in (sa, sr):
        if "cell_type" not in d.columns and "Cell Type" in d.columns:
            d["cell_type"] = d["Cell Type"].astype(str)
    st_ct = sorted(set(sa["cell_type"]))
    if all(c in sa.columns for c in st_ct):
        idc = "Cell_Num_Old" if "Cell_Num_Old" in sa.columns else sa.columns[0]
        o4 = dict(radius=3, knn=8, cell_id_col=idc, min_angle_deg=15, ignore_same_type_triangles=True)
        cases.append(("simulated_st", lambda: run_case(
            "simulated_st", sr, sa, st_ct, o4, dict(lazy_allowed_flip_fraction=0.0), idc, seed=14)))
    else:
        print("[golden] simulated_st: probability columns not found, skipped:", list(sa.columns)[:12])

    # ---- simulated_elastic: the reference's second saved run (elastically deformed copy of the same 144 cells) ----
    ea = pd.read_csv(os.path.join(ex, "simulated_elastic/aligned_df.csv"))
    er = pd.read_csv(os.path.join(ex, "simulated_elastic/ref_df.csv"))
    for d in (ea, er):
        if "cell_type" not in d.columns and "Cell Type" in d.columns:
            d["cell_type"] = d["Cell Type"].astype(str)
    el_ct = sorted(set(ea["cell_type"]))
    if all(c in ea.columns for c in el_ct):
        idc_e = "Cell_Num_Old" if "Cell_Num_Old" in ea.columns else ea.columns[0]
        o4e = dict(radius=3, knn=8, cell_id_col=idc_e, min_angle_deg=15, ignore_same_type_triangles=True)
        cases.append(("simulated_elastic", lambda: run_case(
            "simulated_elastic", er, ea, el_ct, o4e, dict(lazy_allowed_flip_fraction=0.0), idc_e, seed=18)))
    else:
        print("[golden] simulated_elastic: probability columns not found, skipped:", list(ea.columns)[:12])

    # ---- seeded datagen section, K=3, sliding window 3x3 with overlap (driver a13) ----
    r5, q5, ct5 = datagen.make_section_pair(n_tiles=4, n_types=3, seed=5)
    o5 = dict(window_size=12, overlap=3, min_cells_per_window=10, radius=1.0, knn=5, max_matches=1,
              min_angle_deg=15, ignore_same_type_triangles=True, cell_id_col="Cell_Num_Old")
    cases.append(("tiles4_sliding", lambda: run_case(
        "tiles4_sliding", r5, q5, ct5, o5, dict(lazy_allowed_flip_fraction=0.01), "Cell_Num_Old", seed=15,
        sliding=True)))
    # ---- K=8 uniform section, direct, knn=12 ----
    r6, q6, ct6 = datagen.make_uniform_pair(700, 650, extent=100.0, n_types=8, seed=6)
    o6 = dict(radius=9.0, knn=12, max_matches=2, min_angle_deg=20, ignore_same_type_triangles=True,
              cell_id_col="Cell_Num_Old", dist_ct_coeff=0.7)
    cases.append(("uniform_k8", lambda: run_case(
        "uniform_k8", r6, q6, ct6, o6, dict(lazy_allowed_flip_fraction=0.05), "Cell_Num_Old", seed=16)))
    # ---- sparse section: small-window merge quirk (same.py:527-542) + sizes > 1 on the ref side ----
    r7, q7, ct7 = datagen.make_uniform_pair(260, 240, extent=60.0, n_types=3, seed=7)
    # thin out the right-hand third so some windows fall under min_cells_per_window
    r7 = r7[(r7["X"] < 38) | (r7["Y"] < 20)].reset_index(drop=True)
    q7 = q7[(q7["X"] < 38) | (q7["Y"] < 20)].reset_index(drop=True)
    r7["size"] = np.where(np.arange(len(r7)) % 7 == 0, 3, 1)
    q7["size"] = np.where(np.arange(len(q7)) % 5 == 0, 2, 1)
    o7 = dict(window_size=20, overlap=4, min_cells_per_window=12, radius=6.0, knn=4, max_matches=1,
              min_angle_deg=10, ignore_same_type_triangles=False, cell_id_col="Cell_Num_Old")
    cases.append(("sparse_merge", lambda: run_case(
        "sparse_merge", r7, q7, ct7, o7, dict(lazy_allowed_flip_fraction=0.0), "Cell_Num_Old", seed=17,
        sliding=True, stage=False)))

    for nm, fn in cases:
        if only and nm not in only:
            continue
        fn()


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
